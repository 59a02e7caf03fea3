"""Randomised stress of the emulated kernels (tests/emu): random grid sizes, flats, nodata holes, strip counts, the warp-per-tile
sweep (single strip and exchange rounds) and strip flats against the oracle.  python scripts/emu_stress.py [first_seed] [n_seeds] [seconds]"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, test_emu
from taudem_b200 import synth
from oracle import port
lib=test_emu._build()
t0=time.time(); bad=0; n=0
first = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
count = int(sys.argv[2]) if len(sys.argv) > 2 else 100
budget = float(sys.argv[3]) if len(sys.argv) > 3 else 1500.0
for seed in range(first, first + count):
    rng=np.random.default_rng(seed)
    ny=int(rng.integers(8,200)); nx=int(rng.integers(5,300))
    dem=synth.punch_holes(synth.gen_dem(ny,nx,hurst=float(rng.choice([0.6,0.8])),tilt=float(rng.choice([0.0,1.0,4.0])),seed=seed),seed=seed)
    if rng.random()<0.4:
        q=(dem.max()-dem.min())/8; m=dem!=-9999.0; dem=np.where(m,(np.round(dem/q)*q),dem).astype(np.float32)
    fel=port.pitremove(dem); p,_=port.d8flowdir(fel); ang,_=port.dinfflowdir(fel)
    w=synth.gen_weights(ny,nx,seed=seed)
    ad8=port.aread8(p); sca=port.areadinf(ang); ad8w=port.aread8(p,weights=w,contcheck=False); scaw=port.areadinf(ang,weights=w,contcheck=False)
    strips=int(rng.integers(1,4))
    checks=[('ad8',test_emu._run(lib,False,0,0,p,None,True,seed,strips),ad8),
            ('sca',test_emu._run(lib,True,0,0,ang,None,True,seed+1,strips),sca),
            ('ad8w',test_emu._run(lib,False,0,0,p,w,False,seed+2,strips),ad8w),
            ('scaw',test_emu._run(lib,True,0,0,ang,w,False,seed+3,strips),scaw),
            ('ad8 1 strip',test_emu._run(lib,False,0,0,p,None,True,seed+4),ad8),
            ('sca 1 strip',test_emu._run(lib,True,0,0,ang,None,True,seed+5),sca)]
    # flats over strips
    p0,_=port.d8flowdir(fel,flats=False); a0,_=port.dinfflowdir(fel,flats=False)
    fs=int(rng.integers(1,5))
    if ny//fs>=1:
        checks+= [('p strips',test_emu._flats(lib,False,fel,p0,fs,seed+6)[0],p),('ang strips',test_emu._flats(lib,True,fel,a0,fs,seed+7)[0],ang)]
    for name,a,b in checks:
        n+=1
        if not np.array_equal(np.ascontiguousarray(a).view(np.int32 if a.dtype==np.float32 else a.dtype), np.ascontiguousarray(b).view(np.int32 if b.dtype==np.float32 else b.dtype)):
            bad+=1; print('MISMATCH',seed,name,ny,nx,strips,flush=True)
    if time.time()-t0>budget: break
print('checks',n,'bad',bad,'seeds up to',seed,'time',round(time.time()-t0))
sys.exit(1 if bad else 0)
