"""Per-tool device timings on a synthetic DEM (CUDA events)."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import taudem_b200 as td
from taudem_b200.device import DeviceStrip, Tools


def timed(fn):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record(); r = fn(); b.record(); torch.cuda.synchronize()
    return r, a.elapsed_time(b) * 1e-3


def stats(T):
    c = [T.l.td_ctx_counter(T.ctx, 24 + i) for i in range(8)]
    v = max(c[3], 1)
    return f"visits {c[3]} per-visit cycles: wait {c[4]//v} load {c[5]//v} wavefront {c[6]//v} write-back {c[7]//v}" if c[3] else "(TAUDEM_B200_TIMING=1 for per-visit statistics)"


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    hurst = float(sys.argv[2]) if len(sys.argv) > 2 else 0.8
    tilt = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
    T = Tools()
    s = DeviceStrip(n, n)
    # box calibration: plain copy bandwidth + SM clock (timings differ between boxes of the pool)
    a = torch.empty(1 << 28, dtype=torch.float32, device="cuda"); b = torch.empty_like(a)
    best = min(timed(lambda: b.copy_(a))[1] for _ in range(5))
    try:
        import pynvml
        pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
        clk = f"sm {pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)} MHz (max {pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)})"
    except Exception as e:
        clk = str(e)
    print(f"calibration: copy {2 * a.numel() * 4 / best / 1e9:.0f} GB/s, {clk}")
    del a, b
    mc = n * n / 1e6
    dxc, dyc = s.rows(30.0), s.rows(30.0)
    dem, t = timed(lambda: T.gen_dem(s, hurst=hurst, tilt=tilt)); print(f"gen_dem      {t*1e3:9.2f} ms")
    for rep in range(2):
        fel, t = timed(lambda: T.pitremove(s, dem)); print(f"pitremove    {t*1e3:9.2f} ms  {mc/t:10.1f} Mcells/s  {8*mc/t/1e6*1e3/1e3:8.3f} GB/s")
    del dem
    for rep in range(4):
        (p, sd8, nflat), t = timed(lambda: T.d8_slopes(s, fel, dxc, dyc)); print(f"d8 stencil   {t*1e3:9.2f} ms  {mc/t:10.1f} Mcells/s  {10*mc/t/1e3:8.1f} GB/s  flats {nflat}")
    felc = fel.clone()
    left, t = timed(lambda: T.d8_flats(s, felc, p, dxc, dyc)); print(f"d8 flats     {t*1e3:9.2f} ms  left {left}")
    del sd8
    for rep in range(2):
        ad8, t = timed(lambda: T.aread8(s, p)); print(f"aread8       {t*1e3:9.2f} ms  {mc/t:10.1f} Mcells/s  {6*mc/t/1e3:8.1f} GB/s")
        ad8 = None
    ad8 = s.empty(torch.float32)
    _, t1 = timed(lambda: T.aread8_deps(s, p, ad8)); _, t2 = timed(lambda: T.aread8_sweep(s, ad8)); print(f"  deps {t1*1e3:.2f} ms  sweep {t2*1e3:.2f} ms  max area {float(s.owned(ad8).max())}  {stats(T)} of {((n+31)//32)**2} tiles")
    del ad8, p
    for rep in range(4):
        (ang, slp, nflat), t = timed(lambda: T.dinf_slopes(s, fel, dxc, dyc)); print(f"dinf stencil {t*1e3:9.2f} ms  {mc/t:10.1f} Mcells/s  {12*mc/t/1e3:8.1f} GB/s  flats {nflat}")
    del slp
    felc.copy_(fel)
    left, t = timed(lambda: T.dinf_flats(s, felc, ang, dxc, dyc)); print(f"dinf flats   {t*1e3:9.2f} ms  left {left}")
    del felc, fel
    for rep in range(2):
        sca, t = timed(lambda: T.areadinf(s, ang, dxc, dyc)); print(f"areadinf     {t*1e3:9.2f} ms  {mc/t:10.1f} Mcells/s  {8*mc/t/1e3:8.1f} GB/s")
        sca = None
    sca = s.empty(torch.float32)
    _, t1 = timed(lambda: T.areadinf_deps(s, ang, sca, dxc, dyc)); _, t2 = timed(lambda: T.areadinf_sweep(s, ang, sca, dxc)); print(f"  deps {t1*1e3:.2f} ms  sweep {t2*1e3:.2f} ms  max sca {float(s.owned(sca).max())}  {stats(T)}")
    print("launches", td.launch_count(), "mem GB", torch.cuda.max_memory_allocated() / 1e9)


if __name__ == "__main__":
    main()
