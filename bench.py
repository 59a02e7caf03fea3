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
# dram__bytes_read+write per cell of the tile sweeps from the round-1 ncu capture (profiles/r01_ncu_summary.md)
NCU_TRAFFIC_PER_CELL = {"aread8_sweep": 70.3, "areadinf_sweep": 132.0, "aread8_deps": 8.1, "areadinf_deps": 10.2}
ALG_BYTES = {"aread8_deps": 2 + 0, "aread8_sweep": 2 + 4, "areadinf_deps": 4 + 0, "areadinf_sweep": 4 + 4}
KERNEL = {"aread8_deps": "k_deps_d8", "aread8_sweep": "k_sweep_tiles<d8,64x32>", "areadinf_deps": "k_deps_dinf", "areadinf_sweep": "k_sweep_tiles<dinf,64x16>"}


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
    """DEM -> fel -> (p, ang) on the device (untimed)."""
    from taudem_b200.device import DeviceStrip
    s = DeviceStrip(n, n)
    dxc, dyc = s.rows(30.0), s.rows(30.0)
    t0 = time.time()
    dem = T.gen_dem(s, seed=SEED, hurst=HURST, tilt=TILT)
    fel = T.pitremove(s, dem)
    del dem
    p, sd8, nflat = T.d8_slopes(s, fel, dxc, dyc)
    del sd8
    felw = fel.clone()
    left = T.d8_flats(s, felw, p, dxc, dyc) if nflat else 0
    ang, slp, nflat2 = T.dinf_slopes(s, fel, dxc, dyc)
    del slp
    felw.copy_(fel)
    left2 = T.dinf_flats(s, felw, ang, dxc, dyc) if nflat2 else 0
    del felw, fel
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    info = {"flat_cells": int(nflat), "flats_left_d8": int(left), "flats_left_dinf": int(left2), "setup_s": round(time.time() - t0, 2)}
    return s, dxc, dyc, p, ang, info


def pick_size(torch, want):
    if want:
        return want
    free, _ = torch.cuda.mem_get_info()
    # 65536^2 needs ~90 GB during input preparation (SURVEY.md section 8 budget)
    return 65536 if free > 120e9 else 16384


def apply_sweep_spec(spec):
    """NAME[:passes][+river:cells] -> the environment variables the library reads (capi.cu, sweep_walk.cu)."""
    if not spec:
        return
    base, _, river = spec.partition("+river:")
    name, _, passes = base.partition(":")
    os.environ["TAUDEM_B200_SWEEP"] = name
    for key, val in (("TAUDEM_B200_LEVELS", passes), ("TAUDEM_B200_RIVER", river)):
        if val:
            os.environ[key] = val
        else:
            os.environ.pop(key, None)


def sweep_name():
    m = os.environ.get("TAUDEM_B200_SWEEP", "") or "tiles"
    if m == "levels":
        m += ":" + os.environ.get("TAUDEM_B200_LEVELS", "24")
    if os.environ.get("TAUDEM_B200_RIVER"):
        m += "+river:" + os.environ["TAUDEM_B200_RIVER"]
    return m


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
    s, dxc, dyc, p, ang, info = build_inputs(T, n, torch)
    log("inputs ready", info)
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

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    log("warm-up done")
    td.reset_launch_count()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(args.steps)]
    with ClockSampler(local) as clk:
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for k in range(args.steps):
            step(evs[k])
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
    launches = td.launch_count()
    log("timed steps done, wall", wall)
    step_ms = [e[0].elapsed_time(e[4]) for e in evs]
    total_ms = sum(step_ms)
    part_ms = {name: sum(e[i].elapsed_time(e[i + 1]) for e in evs) / args.steps for i, name in enumerate(parts)}
    ms_per_step = total_ms / args.steps
    value = cells / 1e6 / (ms_per_step * 1e-3)
    # sanity of the result that was timed
    max_ad8 = float(s.owned(ad8).max()); max_sca = float(s.owned(sca).max())
    peak, peak_src = peaks()
    dom = max(part_ms, key=part_ms.get)
    achieved = ALG_BYTES[dom] * cells / (part_ms[dom] * 1e-3) / 1e9
    tiles = sweep_name() == "tiles"
    kname = dict(KERNEL)
    phases = None
    if not tiles:
        # level / walk schedules: the sweep is several kernels; one extra untimed step with phase timers says which one dominates
        names = ("k_level", "k_ready", "k_walk", "k_river")
        os.environ["TAUDEM_B200_TIMING"] = "1"
        phases = {}
        for tool, run in (("aread8", lambda: (T.aread8_deps(s, p, ad8), T.aread8_sweep(s, ad8))),
                          ("areadinf", lambda: (T.areadinf_deps(s, ang, sca, dxc, dyc), T.areadinf_sweep(s, ang, sca, dxc)))):
            run(); torch.cuda.synchronize()
            phases[tool] = {names[i]: round(T.l.td_ctx_phase_ms(T.ctx, i), 3) for i in range(4)}
        os.environ.pop("TAUDEM_B200_TIMING", None)
        for tool in ("aread8", "areadinf"):
            top = max(phases[tool], key=phases[tool].get)
            kname[tool + "_sweep"] = f"{top}<{'d8' if tool == 'aread8' else 'dinf'}> (+ the other phases of sweep_walk.cu)"
    roofline = {"bound": "hbm", "kernel": kname[dom], "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 5), "traffic": round(NCU_TRAFFIC_PER_CELL[dom] * cells) if tiles or dom.endswith("deps") else None,
                "traffic_source": "ncu --set full capture at 8192^2 scaled per cell (profiles/r01_ncu_summary.md)" if tiles or dom.endswith("deps") else "not captured for this schedule yet",
                "peak_source": peak_src, "sweep": sweep_name(), "sweep_phases_ms": phases,
                "algorithmic_bytes_per_cell": ALG_BYTES[dom], "ms_per_launch": round(part_ms[dom], 3),
                "per_kernel_ms": {kname[k]: round(v, 3) for k, v in part_ms.items()},
                "per_kernel_frac": {kname[k]: round(ALG_BYTES[k] * cells / (v * 1e-3) / 1e9 / peak, 5) for k, v in part_ms.items()}}

    # ---- end to end through the host-grid C ABI with pinned host buffers
    T.close(); del ad8, sca
    hp = torch.empty((n, n), dtype=torch.int16, pin_memory=True); hp.copy_(s.owned(p))
    ha = torch.empty((n, n), dtype=torch.float32, pin_memory=True); ha.copy_(s.owned(ang))
    del p, ang
    torch.cuda.synchronize(); torch.cuda.empty_cache()
    o1 = torch.empty((n, n), dtype=torch.float32, pin_memory=True)
    o2 = torch.empty((n, n), dtype=torch.float32, pin_memory=True)
    hpn, han, o1n, o2n = hp.numpy(), ha.numpy(), o1.numpy(), o2.numpy()
    log("pinned host buffers ready")
    e2e_steps = max(1, min(args.steps, args.e2e_steps))

    def e2e_step():
        td.aread8_grid(hpn, out=o1n)
        td.areadinf_grid(han, dx=30.0, dy=30.0, out=o2n)

    for _ in range(1 if n > 20000 else min(args.warmup, 3)):
        e2e_step()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    log("e2e done", e2e_s)
    e2e = {"value": round(cells / 1e6 / e2e_s, 2), "unit": "Mcells/s", "h2d_bytes_per_step": hpn.nbytes + han.nbytes,
           "d2h_bytes_per_step": o1n.nbytes + o2n.nbytes, "steps": e2e_steps, "ms_per_step": round(e2e_s * 1e3, 2),
           "api": "td_aread8_host + td_area_host (pinned host rasters in, pinned host rasters out)"}
    assert float(o1.max()) == max_ad8 and float(o2.max()) == max_sca, "e2e result differs from the device-resident run"

    cpu = cpu_reference_sample(hpn, han, args.cpu_sample, args.cpu_ranks) if not args.no_cpu else None
    line = {"metric": METRIC, "value": round(value, 2), "unit": "Mcells/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms_per_step, 3), "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": f"aread8 + areadinf on {n}x{n} float32 synthetic fractal DEM (hills: H={HURST}, tilt={TILT}, seed={SEED}, 30 m cells), contamination check on, no weights",
                                             "cells": cells, "l2": "inputs (>= 1.5 GiB) exceed the 126 MB L2; no explicit flush", "timed": "CUDA events on the launching stream, wall %.3f s for %d steps" % (wall, args.steps),
                                             "max_ad8": max_ad8, "max_sca": max_sca, "sweep": sweep_name(), **info},
            "clocks": clk.summary(), "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu}
    print(json.dumps(line))


def cpu_reference_sample(p_host, ang_host, sample, ranks):
    """Times the reference's own aread8 + areadinf (oracle/_ref, sources compiled unchanged) on a
    window of the same rasters, on this box's host cores."""
    import numpy as np
    import refrun
    if not refrun.available():
        return {"value": None, "unit": "Mcells/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref not built"}
    m = min(sample, p_host.shape[0])
    ranks = max(1, min(ranks, os.cpu_count() or 1))
    R = refrun.RefPipeline(np_ranks=ranks)
    pw = np.ascontiguousarray(p_host[:m, :m]); aw = np.ascontiguousarray(ang_host[:m, :m])
    R.aread8(pw); R.areadinf(aw)
    t = R.times["aread8"]["Compute time"] + R.times["areadinf"]["Compute time"]
    return {"value": round(m * m / 1e6 / t, 3), "unit": "Mcells/s", "cores": ranks, "kind": "reference",
            "sample": f"top-left {m}x{m} window of the bench rasters; the reference tools' own 'Compute time' lines (aread8 {R.times['aread8']['Compute time']:.2f} s + areadinf {R.times['areadinf']['Compute time']:.2f} s), {ranks} ranks over the fork/socketpair MPI shim"}


def reference(args):
    """Reference arm: the reference's CPU implementation (oracle/_ref) on a bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    import refrun
    if not refrun.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref is not built on this box"}))
        return
    m = args.cpu_sample
    ranks = max(1, min(args.cpu_ranks, os.cpu_count() or 1))
    n = args.size or 65536
    # inputs for the sample window: prepared on the GPU when there is one (untimed), else by the reference tools
    p = ang = None
    try:
        import torch
        if torch.cuda.is_available():
            from taudem_b200.device import Tools
            T = Tools()
            s, dxc, dyc, dp, dang, _ = build_inputs(T, min(n, 16384), torch)
            p = s.owned(dp)[:m, :m].contiguous().cpu().numpy(); ang = s.owned(dang)[:m, :m].contiguous().cpu().numpy()
            T.close()
    except Exception as e:  # pragma: no cover
        print("reference arm: GPU input preparation failed, using the reference tools:", e, file=sys.stderr)
    if p is None:
        from taudem_b200 import synth
        m = min(m, 2048)
        R0 = refrun.RefPipeline(np_ranks=ranks)
        fel = R0.pitremove(synth.gen_dem(m, hurst=HURST, tilt=TILT, seed=SEED))
        p, _ = R0.d8flowdir(fel); ang, _ = R0.dinfflowdir(fel)
    R = refrun.RefPipeline(np_ranks=ranks)
    times = []
    for k in range(args.warmup and 1 or 0):
        R.aread8(p); R.areadinf(ang)
    steps = max(1, min(args.steps, 3))
    for k in range(steps):
        R.aread8(p); R.areadinf(ang)
        times.append(R.times["aread8"]["Compute time"] + R.times["areadinf"]["Compute time"])
    t = sum(times) / len(times)
    v = round(p.size / 1e6 / t, 3)
    sample = f"{p.shape[0]}x{p.shape[1]} window of the hills DEM rasters per step, reference 'Compute time' lines, {ranks} ranks (fork/socketpair MPI shim)"
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": "Mcells/s", "n_gpus": args.gpus, "steps": steps, "warmup": 1,
                      "ms_per_step": round(t * 1e3, 2), "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                      "data": "synthetic", "config": {"workload": f"aread8 + areadinf, {sample}"},
                      "cpu_baseline": {"value": v, "unit": "Mcells/s", "cores": ranks, "kind": "reference", "sample": sample},
                      "e2e": {"value": v, "unit": "Mcells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=0, help="DEM edge (default 65536 when it fits, else 16384)")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--cpu-sample", type=int, default=6144)
    ap.add_argument("--cpu-ranks", type=int, default=48, help="MPI ranks of the CPU reference (48 measured 1.9x faster than 16 on the 128-core bench host)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--sweep", default=os.environ.get("TAUDEM_B200_SWEEP_SPEC", ""),
                    help="sweep schedule: tiles (default) | levels[:passes][+river:cells] | hybrid | walk  (sets TAUDEM_B200_SWEEP / _LEVELS / _RIVER)")
    args = ap.parse_args()
    apply_sweep_spec(args.sweep)
    if args.impl == "reference":
        reference(args)
    else:
        ours(args)


if __name__ == "__main__":
    main()
