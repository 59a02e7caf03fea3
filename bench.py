#!/usr/bin/env python
"""bench.py — Mcells/s for aread8 + areadinf on a synthetic fractal DEM (BASELINE.json metric).

One step = one pass of the hot path over the whole DEM: aread8 (dependency stencil + evaluation
sweep) followed by areadinf (same, D-infinity) — `value` = DEM cells / step time with the
direction rasters already resident in HBM; `e2e` = the same through the host-grid C-ABI calls
(td_aread8_host / td_area_host) with pinned HOST buffers, copies inside the timed region.

  python bench.py --gpus N --steps K --warmup W            # our arm
  python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU tools (oracle/_ref)

Inputs are produced on the device by this library's own pitremove -> d8flowdir / dinfflowdir on a
generated DEM (family "hills": H = 0.8, tilt = 1 x relief, seed 1234, 30 m cells); none of that is
inside the timed region.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

T0 = time.time()
METRIC = "Mcells/s for aread8+areadinf on synthetic fractal DEM"
HURST, TILT, SEED = 0.8, 1.0, 1234
# dram__bytes_read + dram__bytes_write per cell from the ncu --set full captures (sweeps at 16384^2, dependency stencils at 8192^2: profiles/r02_ncu_summary.md); None = not captured
NCU_TRAFFIC_PER_CELL = {"aread8_sweep": None, "areadinf_sweep": None, "aread8_deps": None, "areadinf_deps": None}
try:
    with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as _f:
        NCU_TRAFFIC_PER_CELL.update(json.load(_f))
except Exception:
    pass
ALG_BYTES = {"aread8_deps": 2 + 0, "aread8_sweep": 2 + 4, "areadinf_deps": 4 + 0, "areadinf_sweep": 4 + 4}
KERNEL = {"aread8_deps": "k_deps_d8", "aread8_sweep": "k_sweep_warp<d8>", "areadinf_deps": "k_deps_dinf", "areadinf_sweep": "k_sweep_warp<dinf>"}
# algorithmic bytes per cell of the flow-direction pipeline: read + written rasters
PIPE_BYTES = {"pitremove_init": 4 + 4, "d8_stencil": 4 + 2 + 4, "dinf_stencil": 4 + 4 + 4}
REF_SIZE = 16384      # the configuration both arms run in full (BASELINE.json configs[1] / [2])


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock + throttle reasons sampled through NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "hw_power_brake": 0x80}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        if self.nv:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._t:
            self._t.join()

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def build_inputs(T, n, torch):
    """DEM -> fel -> (p, ang) on the device (outside the timed steps); every tool is timed with CUDA events on the way."""
    from taudem_b200.device import DeviceStrip
    s = DeviceStrip(n, n)
    dxc, dyc = s.rows(30.0), s.rows(30.0)
    t0 = time.time()
    ms = {}

    def timed(name, fn):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(); r = fn(); e1.record(); torch.cuda.synchronize()
        ms[name] = round(e0.elapsed_time(e1), 3)
        return r

    dem = T.gen_dem(s, seed=SEED, hurst=HURST, tilt=TILT)
    T.flood_init(s, dem)              # first launch of the kernel: module load, tensor map, occupancy query (like the stencils below)
    fel = timed("pitremove_init", lambda: T.flood_init(s, dem))
    timed("pitremove_relax", lambda: T.flood_relax(s, dem, fel))
    del dem
    T.d8_slopes(s, fel, dxc, dyc)     # first launch of the kernel: module load, tensor map, occupancy query
    p, sd8, nflat = timed("d8_stencil", lambda: T.d8_slopes(s, fel, dxc, dyc))
    del sd8
    felw = fel.clone()
    left = timed("d8_flats", lambda: T.d8_flats(s, felw, p, dxc, dyc)) if nflat else 0
    T.dinf_slopes(s, fel, dxc, dyc)
    ang, slp, nflat2 = timed("dinf_stencil", lambda: T.dinf_slopes(s, fel, dxc, dyc))
    del slp
    felw.copy_(fel)
    left2 = timed("dinf_flats", lambda: T.dinf_flats(s, felw, ang, dxc, dyc)) if nflat2 else 0
    del felw, fel
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    info = {"flat_cells": int(nflat), "flats_left_d8": int(left), "flats_left_dinf": int(left2), "setup_s": round(time.time() - t0, 2)}
    return s, dxc, dyc, p, ang, info, ms


HASH_K = 0x9E3779B97F4A7C15 - (1 << 64)     # odd 64-bit constant as a signed int64


def raster_hash(torch, owned, row0, nx_total):
    """64-bit position-weighted sum of the raw float32 bits of a raster (mod 2^64): equal rasters <=> equal hashes for
    all practical purposes, and the hashes of row strips add up to the hash of the whole grid."""
    h = 0
    ny, nx = owned.shape
    step = max(1, (1 << 26) // max(nx, 1))
    for r in range(0, ny, step):
        blk = owned[r:r + step].contiguous()
        v = blk.view(torch.int32).to(torch.int64) & 0xFFFFFFFF
        rows = torch.arange(row0 + r, row0 + r + blk.shape[0], device=blk.device, dtype=torch.int64)[:, None]
        cols = torch.arange(nx, device=blk.device, dtype=torch.int64)[None, :]
        idx = rows * nx_total + cols
        h = (h + int((v * ((2 * idx + 1) * HASH_K)).sum().item())) & ((1 << 64) - 1)
    return h


def pick_size(torch, want):
    if want:
        return want
    free, _ = torch.cuda.mem_get_info()
    # 65536^2 needs ~90 GB during input preparation (SURVEY.md section 8 budget)
    return 65536 if free > 120e9 else 16384


def sweep_name():
    return "warp-per-tile dataflow (sweep_warp.cu)"


def run_size(args, torch, td, T, n, steps, warmup, e2e_steps, log, local):
    """The whole measurement at one DEM size: inputs, per-tool pipeline timings, timed steps, hashes, end to end."""
    cells = n * n
    s, dxc, dyc, p, ang, info, pipe_ms = build_inputs(T, n, torch)
    log(f"{n}^2 inputs ready", info, pipe_ms)
    ad8, sca = s.empty(torch.float32), s.empty(torch.float32)
    parts = ("aread8_deps", "aread8_sweep", "areadinf_deps", "areadinf_sweep")

    def step(ev=None):
        def mark(i):
            if ev is not None:
                ev[i].record()
        mark(0); T.aread8_deps(s, p, ad8)
        mark(1); T.aread8_sweep(s, ad8)
        mark(2); T.areadinf_deps(s, ang, sca, dxc, dyc)
        mark(3); T.areadinf_sweep(s, ang, sca, dxc)
        mark(4)

    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    td.reset_launch_count()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(steps)]
    with ClockSampler(local) as clk:
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for k in range(steps):
            step(evs[k])
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
    launches = td.launch_count()
    total_ms = sum(e[0].elapsed_time(e[4]) for e in evs)
    part_ms = {name: sum(e[i].elapsed_time(e[i + 1]) for e in evs) / steps for i, name in enumerate(parts)}
    ms_per_step = total_ms / steps
    res = {"n": n, "cells": cells, "ms_per_step": ms_per_step, "value": cells / 1e6 / (ms_per_step * 1e-3), "part_ms": part_ms, "wall": wall,
           "launches": int(launches), "clocks": clk.summary(), "info": info, "pipe_ms": pipe_ms,
           "max_ad8": float(s.owned(ad8).max()), "max_sca": float(s.owned(sca).max()),
           "hash_ad8": "%016x" % raster_hash(torch, s.owned(ad8), 0, n), "hash_sca": "%016x" % raster_hash(torch, s.owned(sca), 0, n)}
    # sweep statistics of one extra (untimed) step: visits per tile and where a visit's time goes
    os.environ["TAUDEM_B200_TIMING"] = "1"
    stats = {}
    for tool, run in (("aread8", lambda: (T.aread8_deps(s, p, ad8), T.aread8_sweep(s, ad8))),
                      ("areadinf", lambda: (T.areadinf_deps(s, ang, sca, dxc, dyc), T.areadinf_sweep(s, ang, sca, dxc)))):
        run(); torch.cuda.synchronize()
        c = [T.l.td_ctx_counter(T.ctx, 24 + i) for i in range(8)]
        v = max(c[3], 1)
        stats[tool] = {"tile_visits": c[3], "tiles": ((n + 31) // 32) ** 2, "cycles_per_visit": {"queue_wait": c[4] // v, "load": c[5] // v, "wavefront": c[6] // v, "write_back": c[7] // v}}
    os.environ.pop("TAUDEM_B200_TIMING", None)
    res["sweep_stats"] = stats

    # ---- end to end through the host-grid C ABI with pinned host buffers
    del ad8, sca
    hp = torch.empty((n, n), dtype=torch.int16, pin_memory=True); hp.copy_(s.owned(p))
    ha = torch.empty((n, n), dtype=torch.float32, pin_memory=True); ha.copy_(s.owned(ang))
    del p, ang
    torch.cuda.synchronize(); torch.cuda.empty_cache()
    o1 = torch.empty((n, n), dtype=torch.float32, pin_memory=True)
    o2 = torch.empty((n, n), dtype=torch.float32, pin_memory=True)
    hpn, han, o1n, o2n = hp.numpy(), ha.numpy(), o1.numpy(), o2.numpy()

    def e2e_pair():          # both tools in one call: copies overlapped with the kernels (td_contributing_areas_host)
        td.contributing_areas_grid(hpn, han, dx=30.0, dy=30.0, out_ad8=o1n, out_sca=o2n)

    def e2e_sequential():    # the two reference-shaped calls one after the other: copy in, compute, copy out, twice
        td.aread8_grid(hpn, out=o1n)
        td.areadinf_grid(han, dx=30.0, dy=30.0, out=o2n)

    def run(fn):
        fn()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            fn()
        return (time.perf_counter() - t0) / e2e_steps

    seq_s = run(e2e_sequential)
    o1.zero_(); o2.zero_()
    e2e_s = run(e2e_pair)
    log(f"{n}^2 e2e done", e2e_s, "sequential", seq_s)
    res["e2e"] = {"value": round(cells / 1e6 / e2e_s, 2), "unit": "Mcells/s", "h2d_bytes_per_step": hpn.nbytes + han.nbytes,
                  "d2h_bytes_per_step": o1n.nbytes + o2n.nbytes, "steps": e2e_steps, "ms_per_step": round(e2e_s * 1e3, 2),
                  "api": "td_contributing_areas_host (pinned host p + ang in, pinned host ad8 + sca out; three streams: copies overlap the kernels)",
                  "sequential_calls": {"value": round(cells / 1e6 / seq_s, 2), "ms_per_step": round(seq_s * 1e3, 2),
                                       "api": "td_aread8_host then td_area_host (each: copy in, compute, copy out)"}}
    assert float(o1.max()) == res["max_ad8"] and float(o2.max()) == res["max_sca"], "e2e result differs from the device-resident run"
    res["host"] = (hpn, han)
    return res


def ours(args):
    import torch
    import torch.distributed as dist
    import taudem_b200 as td
    from taudem_b200.device import Tools

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if world != args.gpus:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE {world}", file=sys.stderr)
    if world > 1:
        from taudem_b200 import dist as tdist
        return tdist.bench_main(args, rank, world, local)

    n = pick_size(torch, args.size)
    cells = n * n
    T = Tools()
    log = lambda *a: print("[bench %.1fs]" % (time.time() - T0), *a, file=sys.stderr, flush=True)
    R = run_size(args, torch, td, T, n, args.steps, args.warmup, max(1, min(args.steps, args.e2e_steps)), log, local)
    part_ms, pipe_ms = R["part_ms"], R["pipe_ms"]
    peak, peak_src = peaks()
    dom = max(part_ms, key=part_ms.get)
    achieved = ALG_BYTES[dom] * cells / (part_ms[dom] * 1e-3) / 1e9
    traffic = NCU_TRAFFIC_PER_CELL.get(dom)
    gbs = lambda b, ms: b * cells / (ms * 1e-3) / 1e9
    pipeline = {k: {"ms": v, "Mcells_per_s": round(cells / 1e6 / (v * 1e-3), 1)} for k, v in pipe_ms.items()}
    for k, b in PIPE_BYTES.items():
        if k in pipe_ms:
            pipeline[k].update({"algorithmic_bytes_per_cell": b, "GB_per_s": round(gbs(b, pipe_ms[k]), 1), "frac_of_hbm_peak": round(gbs(b, pipe_ms[k]) / peak, 4)})
    roofline = {"bound": "hbm", "kernel": KERNEL[dom], "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 5), "traffic": round(traffic * cells) if traffic else None,
                "traffic_source": "ncu --set full capture at 16384^2 (dram__bytes_read + write of one launch), scaled per cell to this size (profiles/r02_ncu_summary.md)" if traffic else "not captured",
                "peak_source": peak_src, "sweep": sweep_name(), "sweep_stats": R["sweep_stats"],
                "algorithmic_bytes_per_cell": ALG_BYTES[dom], "ms_per_launch": round(part_ms[dom], 3),
                "note": "the contributing-area sweep is bound by instruction issue (aread8) / dependent-issue latency x tiles in flight (areadinf), not by bandwidth (DESIGN.md section 4.1)",
                "per_kernel_ms": {KERNEL[k]: round(v, 3) for k, v in part_ms.items()},
                "per_kernel_frac": {KERNEL[k]: round(ALG_BYTES[k] * cells / (v * 1e-3) / 1e9 / peak, 5) for k, v in part_ms.items()},
                "pipeline": pipeline}
    hpn, han = R.pop("host")
    same = None
    if n != REF_SIZE and not args.no_same_config:
        # the configuration the reference arm runs in full: one more, smaller measurement so that one pair of lines is same-config
        del hpn, han
        torch.cuda.empty_cache()
        R2 = run_size(args, torch, td, T, REF_SIZE, max(3, args.steps), 3, 2, log, local)
        hpn, han = R2.pop("host")
        same = {"workload": workload_name(REF_SIZE), "value": round(R2["value"], 2), "unit": "Mcells/s", "ms_per_step": round(R2["ms_per_step"], 3),
                "e2e": R2["e2e"], "hash_ad8": R2["hash_ad8"], "hash_sca": R2["hash_sca"], "per_kernel_ms": {KERNEL[k]: round(v, 3) for k, v in R2["part_ms"].items()},
                "pipeline_ms": R2["pipe_ms"]}
    T.close()
    cpu = cpu_reference_sample(hpn, han, args.cpu_sample, args.cpu_ranks) if not args.no_cpu else None
    line = {"metric": METRIC, "value": round(R["value"], 2), "unit": "Mcells/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(R["ms_per_step"], 3), "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": workload_name(n), "cells": cells,
                                             "l2": "inputs (>= 1.5 GiB) exceed the 126 MB L2; no explicit flush", "timed": "CUDA events on the launching stream, wall %.3f s for %d steps" % (R["wall"], args.steps),
                                             "max_ad8": R["max_ad8"], "max_sca": R["max_sca"], "hash_ad8": R["hash_ad8"], "hash_sca": R["hash_sca"], "sweep": sweep_name(), **R["info"]},
            "clocks": R["clocks"], "e2e": R["e2e"], "gpu_launches": R["launches"], "roofline": roofline, "cpu_baseline": cpu,
            "same_config_as_reference_arm": same}
    print(json.dumps(line))


def workload_name(n):
    return (f"aread8 + areadinf on {n}x{n} float32 synthetic fractal DEM (hills: H={HURST}, tilt={TILT}, seed={SEED}, 30 m cells), "
            "contamination check on, no weights")


def host_cores():
    """CPUs this process may really use: the affinity mask, capped by the cgroup CPU quota (a container can see 128 CPUs and own 8;
    48 pinned ranks on 8 cores made the reference arm 12x slower on one of the pool's hosts)."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            txt = open(path).read().split()
            if path.endswith("cpu.max"):
                if txt[0] != "max":
                    n = min(n, max(1, int(int(txt[0]) / int(txt[1]))))
            else:
                q = int(txt[0]); per = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
                if q > 0:
                    n = min(n, max(1, q // per))
            break
        except Exception:
            continue
    return max(1, n)


def cpu_reference_sample(p_host, ang_host, sample, ranks):
    """Times the reference's own aread8 + areadinf (oracle/_ref, sources compiled unchanged) on a
    window of the same rasters, on this box's host cores."""
    import numpy as np
    import refrun
    if not refrun.available():
        return {"value": None, "unit": "Mcells/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref not built"}
    m = min(sample, p_host.shape[0])
    ranks = max(1, min(ranks, host_cores()))
    os.environ["MINIMPI_PIN"] = "1"
    R = refrun.RefPipeline(np_ranks=ranks)
    pw = np.ascontiguousarray(p_host[:m, :m]); aw = np.ascontiguousarray(ang_host[:m, :m])
    R.aread8(pw); R.areadinf(aw)
    t = R.times["aread8"]["Compute time"] + R.times["areadinf"]["Compute time"]
    return {"value": round(m * m / 1e6 / t, 3), "unit": "Mcells/s", "cores": ranks, "kind": "reference",
            "sample": f"top-left {m}x{m} window of the {p_host.shape[0]}^2 rasters; the reference tools' own 'Compute time' lines (aread8 {R.times['aread8']['Compute time']:.2f} s + areadinf {R.times['areadinf']['Compute time']:.2f} s), {ranks} ranks pinned one per core over the fork/socketpair MPI shim ({host_cores()} usable cores of {os.cpu_count()})"}


def prep(args):
    """Writes the direction rasters of the REF_SIZE configuration as TIFF files (untimed input preparation for the
    reference arm, run as a separate process: the reference arm itself never loads this library)."""
    import torch
    import taudem_b200 as td
    from taudem_b200.device import Tools
    T = Tools()
    s, dxc, dyc, p, ang, info, _ = build_inputs(T, args.size or REF_SIZE, torch)
    os.makedirs(args.prep, exist_ok=True)
    td.write_raster(os.path.join(args.prep, "p.tif"), s.owned(p).contiguous().cpu().numpy(), -32768, dx=30.0, dy=30.0, compression=1)
    td.write_raster(os.path.join(args.prep, "ang.tif"), s.owned(ang).contiguous().cpu().numpy(), -3.4028234663852886e38, dx=30.0, dy=30.0, compression=1)
    T.close()


def reference(args):
    """Reference arm: the reference's own aread8 + areadinf (oracle/_ref: its sources compiled unchanged) on the full
    REF_SIZE^2 rasters, every step, on this box's host cores.  This process loads neither the product library nor torch:
    the input rasters are files written by a separate preparation process (`bench.py --prep DIR`, the same generator and
    flow-direction pipeline as our arm) or, without a GPU, by the reference's own tools on a smaller DEM."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import subprocess
    import tempfile
    import refrun
    if not refrun.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref is not built on this box"}))
        return
    ranks = max(1, min(args.cpu_ranks, host_cores()))
    os.environ["MINIMPI_PIN"] = "1"
    n = args.size or REF_SIZE
    work = tempfile.mkdtemp(prefix="tdbench_ref_")
    note = f"{n}x{n} (full grid) of the hills DEM"
    r = subprocess.run([sys.executable, os.path.abspath(__file__), "--prep", work, "--size", str(n)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    if r.returncode != 0 or not os.path.exists(os.path.join(work, "ang.tif")):
        # no usable GPU for the preparation step: the reference's own pitremove / d8flowdir / dinfflowdir on a small DEM
        import numpy as np
        from taudem_b200 import synth
        n = min(n, 2048)
        R0 = refrun.RefPipeline(workdir=work, np_ranks=ranks)
        fel = R0.pitremove(synth.gen_dem(n, hurst=HURST, tilt=TILT, seed=SEED))
        p, _ = R0.d8flowdir(fel); ang, _ = R0.dinfflowdir(fel)
        R0.put("p.tif", p.astype(np.int16), -32768); R0.put("ang.tif", ang, -3.4028234663852886e38)
        note = f"{n}x{n} DEM prepared by the reference tools (GPU preparation failed: {r.stderr.strip()[-200:]})"
    pf, af = os.path.join(work, "p.tif"), os.path.join(work, "ang.tif")
    steps = max(1, min(args.steps, 2))
    times = []
    for k in range((1 if args.warmup else 0) + steps):
        _, t1 = refrun.run_tool("aread8", ["-p", pf, "-ad8", os.path.join(work, "ad8.tif")], ranks)
        _, t2 = refrun.run_tool("areadinf", ["-ang", af, "-sca", os.path.join(work, "sca.tif")], ranks)
        times.append((t1["Compute time"], t2["Compute time"]))
    times = times[-steps:]
    t = sum(a + b for a, b in times) / len(times)
    v = round(n * n / 1e6 / t, 3)
    sample = f"{note}; the reference tools' own 'Compute time' lines (aread8 {times[-1][0]:.2f} s + areadinf {times[-1][1]:.2f} s), {ranks} ranks pinned one per core (fork/socketpair MPI shim; {host_cores()} usable cores of {os.cpu_count()})"
    import shutil
    shutil.rmtree(work, ignore_errors=True)
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": "Mcells/s", "n_gpus": args.gpus, "steps": steps, "warmup": 1 if args.warmup else 0,
                      "ms_per_step": round(t * 1e3, 2), "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                      "data": "synthetic", "config": {"workload": workload_name(n), "cells": n * n},
                      "cpu_baseline": {"value": v, "unit": "Mcells/s", "cores": ranks, "kind": "reference", "sample": sample},
                      "e2e": {"value": v, "unit": "Mcells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def files_mode(args):
    """File-level wall clock of the executables (not a bench line): `aread8` / `areadinf` on GeoTIFF files of the REF_SIZE
    configuration, ours with TAUDEM_B200_GPUS = 1 .. --gpus (the reference's `mpiexec -n N`), the reference executables
    (oracle/_ref, all host cores) beside them.  Wall time covers process start, reading, computing and writing (LZW)."""
    import filecmp
    import re
    import subprocess
    import tempfile
    import time
    n = args.size or REF_SIZE
    work = tempfile.mkdtemp(prefix="tdbench_files_")
    r = subprocess.run([sys.executable, os.path.abspath(__file__), "--prep", work, "--size", str(n)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    if r.returncode != 0:
        print(json.dumps({"files": "unavailable", "why": r.stderr.strip()[-300:]}))
        return
    bindir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "taudem_b200", "bin")
    rows = []
    ns = [k for k in (1, 2, 4, 8) if k <= max(1, args.gpus)]
    for tool, flag_in, fin, flag_out in (("aread8", "-p", "p.tif", "-ad8"), ("areadinf", "-ang", "ang.tif", "-sca")):
        first = None
        for k in ns:
            out = os.path.join(work, f"{tool}_{k}.tif")
            env = dict(os.environ, TAUDEM_B200_GPUS=str(k))
            t0 = time.time()
            rr = subprocess.run([os.path.join(bindir, tool), flag_in, os.path.join(work, fin), flag_out, out], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)
            wall = time.time() - t0
            tm = {m.group(1).strip(): float(m.group(2)) for m in re.finditer(r"^([A-Za-z ]+time): ([0-9.eE+-]+)$", rr.stdout, re.M)}
            same = None
            if first is None:
                first = out
            else:
                same = filecmp.cmp(first, out, shallow=False)
            row = {"tool": tool, "impl": "ours", "ranks": k, "rc": rr.returncode, "wall_s": round(wall, 3), "tool_times_s": tm, "file_identical_to_1_rank": same}
            trace = [(round(float(m.group(1)) - t0, 3), m.group(2)) for m in re.finditer(r"^\[td trace\] ([0-9.]+) (.*)$", rr.stdout, re.M)]
            if trace:
                row["trace_s_after_launch"] = trace
            rows.append(row)
        if not args.no_cpu:
            sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle"))
            import refrun
            if refrun.available():
                ranks = max(1, min(args.cpu_ranks, host_cores()))
                os.environ["MINIMPI_PIN"] = "1"
                t0 = time.time()
                _, tm = refrun.run_tool(tool, [flag_in, os.path.join(work, fin), flag_out, os.path.join(work, f"{tool}_ref.tif")], ranks)
                rows.append({"tool": tool, "impl": "reference", "ranks": ranks, "rc": 0, "wall_s": round(time.time() - t0, 3), "tool_times_s": tm})
    import shutil
    shutil.rmtree(work, ignore_errors=True)
    print(json.dumps({"files": f"{n}x{n}", "rows": rows}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=0, help="DEM edge (default 65536 when it fits, else 16384; reference arm: 16384)")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--cpu-sample", type=int, default=6144)
    ap.add_argument("--cpu-ranks", type=int, default=48, help="MPI ranks of the CPU reference (48 measured 1.9x faster than 16 on the 128-core bench host)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-same-config", action="store_true", help="skip the additional 16384^2 measurement (the reference arm's configuration)")
    ap.add_argument("--prep", default="", help="write the REF_SIZE direction rasters into this directory and exit (input preparation of the reference arm)")
    ap.add_argument("--files", action="store_true", help="file-level wall clock of the executables at 1 .. --gpus ranks next to the reference executables")
    args = ap.parse_args()
    if args.files:
        files_mode(args)
    elif args.prep:
        prep(args)
    elif args.impl == "reference":
        reference(args)
    else:
        ours(args)


if __name__ == "__main__":
    main()
