"""A/B of the single-strip sweep implementations (TAUDEM_B200_SWEEP = tiles | levels | hybrid | walk | chain):
bit-identity of ad8 / sca against the first mode listed and CUDA-event timings of the sweep alone.

  python scripts/sweep_modes.py [n=4096] [modes=tiles,levels,levels:48,levels+river:64,hybrid,walk] [reps=3]

A mode is NAME[:passes][+river:hops]: `levels:48` = 48 level passes (TAUDEM_B200_LEVELS; `levels:auto` = until a pass stops paying), `+river:64` = D8 chains
longer than 64 cells go to the look-ahead river kernel (TAUDEM_B200_RIVER; with RIVER_DINF=1 in the environment also
TAUDEM_B200_RIVER_DINF for areadinf, where the look-ahead rarely pays: braided strands).

Every mode runs under the caller's own `timeout`; a mode that differs prints DIFFERENT and the script
exits 1 at the end."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from taudem_b200.device import DeviceStrip, Tools


def timed(fn):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record(); r = fn(); b.record(); torch.cuda.synchronize()
    return r, a.elapsed_time(b)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    modes = (sys.argv[2] if len(sys.argv) > 2 else "tiles,levels,levels:48,levels+river:64,hybrid,walk").split(",")
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    T = Tools()
    s = DeviceStrip(n, n)
    dxc, dyc = s.rows(30.0), s.rows(30.0)
    dem = T.gen_dem(s, hurst=0.8, tilt=1.0)
    fel = T.pitremove(s, dem); del dem
    p, sd8, _ = T.d8_slopes(s, fel, dxc, dyc); del sd8
    felc = fel.clone(); T.d8_flats(s, felc, p, dxc, dyc)
    ang, slp, _ = T.dinf_slopes(s, fel, dxc, dyc); del slp
    felc.copy_(fel); T.dinf_flats(s, felc, ang, dxc, dyc); del felc, fel
    ref, bad = {}, 0
    for mode in modes:
        base, _, river = mode.partition("+river:")
        name, _, passes = base.partition(":")
        os.environ["TAUDEM_B200_SWEEP"] = name
        for key, val in (("TAUDEM_B200_LEVELS", passes), ("TAUDEM_B200_RIVER", river), ("TAUDEM_B200_RIVER_DINF", river if os.environ.get("RIVER_DINF") else "")):
            if val: os.environ[key] = val
            else: os.environ.pop(key, None)
        for tool in ("aread8", "areadinf"):
            out = s.empty(torch.float32)
            best = 1e30
            for _ in range(reps):
                if tool == "aread8":
                    T.aread8_deps(s, p, out); _, t = timed(lambda: T.aread8_sweep(s, out))
                else:
                    T.areadinf_deps(s, ang, out, dxc, dyc); _, t = timed(lambda: T.areadinf_sweep(s, ang, out, dxc))
                best = min(best, t)
            own = s.owned(out)
            if tool not in ref: ref[tool] = own.clone(); verdict = "reference"
            else:
                same = torch.equal(own.view(torch.int32), ref[tool].view(torch.int32))
                verdict = "identical" if same else f"DIFFERENT ({int((own.view(torch.int32) != ref[tool].view(torch.int32)).sum())} cells)"
                bad += 0 if same else 1
            st = ""
            if os.environ.get("TAUDEM_B200_TIMING") and name in ("warp", "tiles"):
                c = [T.l.td_ctx_counter(T.ctx, 24 + i) for i in range(8)]
                v = max(c[3], 1)
                st = f"  [visits {c[3]} cycles/visit: wait {c[4]//v} load {c[5]//v} walk {c[6]//v} wb {c[7]//v}]"
                if name == "warp":
                    hh = [T.l.td_ctx_sweep_hist(T.ctx, i) for i in range(20)]
                    st += "\n      by cells/visit (<8,<32,<128,more): " + "  ".join(
                        f"[{hh[4*b]} visits, {hh[4*b+1]/max(hh[4*b],1):.0f} cells, {hh[4*b+2]/max(hh[4*b],1):.0f} iters, {hh[4*b+3]/max(hh[4*b],1):.0f} cyc]" for b in range(4))
            ph = [T.l.td_ctx_phase_ms(T.ctx, i) for i in range(4)]
            phases = "" if not any(ph) else "  [levels %.1f ready %.1f walk %.1f river %.1f ms]" % tuple(ph)
            print(f"{mode:18s} {tool:9s} sweep {best:9.2f} ms  {n * n / best / 1e3:9.1f} Mcells/s  max {float(own.max()):.6g}  {verdict}{phases}{st}", flush=True)
            del out
    for key in ("TAUDEM_B200_SWEEP", "TAUDEM_B200_LEVELS", "TAUDEM_B200_RIVER", "TAUDEM_B200_RIVER_DINF"): os.environ.pop(key, None)
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
