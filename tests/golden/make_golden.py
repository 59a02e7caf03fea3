"""Generates tests/golden/*.npz: inputs and the outputs of the reference's own tools
(oracle/_ref, built unchanged from /root/reference by oracle/Makefile) for small cases.
Run in the build container (needs oracle/_ref):  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import refrun
from taudem_b200 import synth


def cases():
    yield "hills_holes", synth.punch_holes(synth.gen_dem(96, 128, hurst=0.8, tilt=1.0)), 30.0, 20.0
    yield "rough", synth.gen_dem(128, 128, family="rough", seed=99), 30.0, 30.0
    plateau = np.full((40, 50), 250.0, np.float32)
    plateau[10:20, 5:25] = 260.0          # a mesa on the plateau
    plateau[30, 44] = 240.0               # and one pit
    yield "plateau", plateau, 10.0, 10.0
    yy, xx = np.mgrid[0:64, 0:80].astype(np.float32)
    bowl = (((yy - 30) ** 2 + (xx - 35) ** 2) * np.float32(0.05) + synth.gen_dem(64, 80, seed=5, hurst=0.9, tilt=0.0) * np.float32(0.02)).astype(np.float32)
    bowl[:, 60:] += np.float32(40.0)      # a dam: the bowl fills to its lowest pass
    yield "lake", bowl, 30.0, 30.0
    yield "tiny", synth.gen_dem(5, 7, seed=3, hurst=0.8, tilt=1.0), 30.0, 30.0


def main():
    assert refrun.available(), "build oracle/_ref first (make -C oracle ref)"
    for name, dem, dx, dy in cases():
        R = refrun.RefPipeline(dx=dx, dy=dy)
        w = synth.gen_weights(*dem.shape)
        fel = R.pitremove(dem)
        fel4 = R.pitremove(dem, four_way=True)
        p, sd8 = R.d8flowdir(fel)
        ang, slp = R.dinfflowdir(fel)
        out = dict(dem=dem, dx=np.float64(dx), dy=np.float64(dy), w=w, fel=fel, fel4=fel4, p=p, sd8=sd8, ang=ang, slp=slp,
                   ad8=R.aread8(p), ad8_w=R.aread8(p, weights=w), ad8_nc=R.aread8(p, contcheck=False),
                   sca=R.areadinf(ang), sca_w=R.areadinf(ang, weights=w), sca_nc=R.areadinf(ang, contcheck=False))
        if name == "lake":
            # depression mask (-depmask): the cells marked 1 are real depressions and keep their elevation
            mask = np.zeros(dem.shape, np.int16)
            mask[24:36, 28:42] = 1
            out["depmask"] = mask
            out["fel_mask"] = R.pitremove(dem, depmask=mask)
            out["fel_mask4"] = R.pitremove(dem, depmask=mask, four_way=True)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        flats = int(((sd8 == 0) & (p != -32768)).sum())
        print(f"{name}: {dem.shape} filled {(fel != dem).sum()} flats {flats} ad8 max {out['ad8'].max()} sca nodata {(out['sca'] == -1).sum()}")


def sibling_inputs(shape, seed=77):
    """the extra grids of the sibling sweep tools and the point-wise consumers"""
    rng = np.random.default_rng(seed)
    q = rng.uniform(0.5, 3.0, shape).astype(np.float32)
    q[rng.random(shape) < 0.004] = -9999.0
    q[rng.random(shape) < 0.004] = 0.0
    dm = rng.uniform(0.2, 1.0, shape).astype(np.float32)
    dm[rng.random(shape) < 0.003] = -9999.0
    dg = (rng.random(shape) < 0.02).astype(np.int16)
    tc = rng.uniform(0.0, 8.0, shape).astype(np.float32)
    tc[rng.random(shape) < 0.003] = -9999.0
    cs = rng.uniform(0.0, 2.0, shape).astype(np.float32)
    cs[rng.random(shape) < 0.003] = -9999.0
    sa = (rng.random(shape) * 100.0 - 20.0).astype(np.float32)
    return dict(q=q, dm=dm, dg=dg, tc=tc, cs=cs, sa=sa)


def siblings():
    """tests/golden/siblings.npz: the nine tools of SURVEY.md 8(f) ranks 3 and 4 on the rasters of the hills_holes case"""
    g = np.load(os.path.join(HERE, "hills_holes.npz"))
    dx, dy = float(g["dx"]), float(g["dy"])
    p, ang, slp, ad8, sca = g["p"], g["ang"], g["slp"], g["ad8"], g["sca"]
    x = sibling_inputs(p.shape)
    R = refrun.RefPipeline(dx=dx, dy=dy)
    out = dict(x)
    out["ssa_max"] = R.d8flowpathextremeup(p, x["sa"], usemax=True)
    out["ssa_min_nc"] = R.d8flowpathextremeup(p, x["sa"], usemax=False, contcheck=False)
    out["plen"], out["tlen"], out["gord"] = R.gridnet(p)
    mask = np.where(ad8 >= 0, ad8, 0).astype(np.int32)
    out["gn_mask"] = mask
    out["plen_m"], out["tlen_m"], out["gord_m"] = R.gridnet(p, mask=mask, thresh=5)
    out["dsca"] = R.dinfdecayaccum(ang, x["dm"])
    out["dsca_w_nc"] = R.dinfdecayaccum(ang, x["dm"], weights=g["w"], contcheck=False)
    out["ctpt"] = R.dinfconclimaccum(ang, x["dm"], x["q"], x["dg"], csol=2.5)
    out["ctpt_nc"] = R.dinfconclimaccum(ang, x["dm"], x["q"], x["dg"], csol=2.5, contcheck=False)
    out["tla"], out["tdep"], _ = R.dinftranslimaccum(ang, x["q"], x["tc"])
    out["tla_c"], out["tdep_c"], out["ctpt_c"] = R.dinftranslimaccum(ang, x["q"], x["tc"], cs=x["cs"], contcheck=False)
    out["src"] = R.threshold(ad8, 50.0)
    out["twi"] = R.twi(slp, sca)
    out["sa_default"] = R.slopearea(slp, sca)
    out["sar"] = R.slopearearatio(slp, sca)
    np.savez_compressed(os.path.join(HERE, "siblings.npz"), **out)
    print("siblings:", {k: (v.dtype.name, int((v > -1e38).sum()) if v.dtype.kind == "f" else int(v.max())) for k, v in out.items() if k not in x})


if __name__ == "__main__":
    main()
    siblings()
