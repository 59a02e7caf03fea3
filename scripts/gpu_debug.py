"""First-light check on a GPU box: every tool against the reference binaries, mismatch statistics."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import refrun
import taudem_b200 as td
from taudem_b200 import synth


def cmp(name, a, b, tol=None):
    if tol is None:
        bad = a.view(np.uint32 if a.dtype == np.float32 else a.dtype) != b.view(np.uint32 if b.dtype == np.float32 else b.dtype)
        # +0/-0 and identical NaNs aside, bit equality
        print(f"  {name}: {int(bad.sum())} / {a.size} cells differ (bit-exact check)")
        if bad.any():
            idx = np.argwhere(bad)[:5]
            for y, x in idx:
                print(f"     ({y},{x}) gpu={a[y, x]!r} ref={b[y, x]!r}")
    else:
        nd_a, nd_b = a <= -1, b <= -1
        print(f"  {name}: nodata/flat masks equal: {bool((nd_a == nd_b).all())}", end="")
        ok = ~nd_a & ~nd_b
        rel = np.abs(a[ok].astype(np.float64) - b[ok]) / np.maximum(np.abs(b[ok]), 1e-30)
        print(f"; max rel err {rel.max() if rel.size else 0:.3e}; bit-different {int((a[ok] != b[ok]).sum())}")
    sys.stdout.flush()


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    print("devices:", td.device_count())
    cases = [("rough", synth.gen_dem(n, family="rough"), 30.0, 30.0),
             ("hills+holes dx!=dy", synth.punch_holes(synth.gen_dem(n * 3 // 4, n, hurst=0.8, tilt=1.0)), 30.0, 20.0)]
    for name, dem, dx, dy in cases:
        print(f"== case {name} {dem.shape}")
        R = refrun.RefPipeline(dx=dx, dy=dy)
        fel_r = R.pitremove(dem)
        p_r, sd8_r = R.d8flowdir(fel_r)
        ad8_r = R.aread8(p_r)
        ang_r, slp_r = R.dinfflowdir(fel_r)
        sca_r = R.areadinf(ang_r)
        w = synth.gen_weights(*dem.shape)
        ad8w_r = R.aread8(p_r, weights=w)
        scaw_r = R.areadinf(ang_r, weights=w, contcheck=False)
        t = time.time(); fel = td.pitremove_grid(dem); print(" pitremove", time.time() - t, td.last_compute_seconds()); cmp("fel", fel, fel_r)
        t = time.time(); p, sd8 = td.d8flowdir_grid(fel_r, dx=dx, dy=dy); print(" d8flowdir", time.time() - t, td.last_compute_seconds()); cmp("sd8", sd8, sd8_r); cmp("p", p, p_r)
        t = time.time(); ad8 = td.aread8_grid(p_r); print(" aread8", time.time() - t, td.last_compute_seconds()); cmp("ad8", ad8, ad8_r)
        ad8w = td.aread8_grid(p_r, weights=w); cmp("ad8 -wg", ad8w, ad8w_r)
        t = time.time(); ang, slp = td.dinfflowdir_grid(fel_r, dx=dx, dy=dy); print(" dinfflowdir", time.time() - t, td.last_compute_seconds()); cmp("slp", slp, slp_r); cmp("ang", ang, ang_r, tol=1e-5); cmp("ang bits", ang, ang_r)
        t = time.time(); sca = td.areadinf_grid(ang_r, dx=dx, dy=dy); print(" areadinf", time.time() - t, td.last_compute_seconds()); cmp("sca", sca, sca_r, tol=1e-5); cmp("sca bits", sca, sca_r)
        scaw = td.areadinf_grid(ang_r, weights=w, dx=dx, dy=dy, contcheck=False); cmp("sca -wg -nc bits", scaw, scaw_r)


if __name__ == "__main__":
    main()
