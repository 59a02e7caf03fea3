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


if __name__ == "__main__":
    main()
