"""The queue / count protocols of the level + walk sweep kernels (taudem_b200/csrc/sweep_walk.cu), executed on a
CPU emulation of the CUDA thread model (tests/emu: fibers, warp collectives, randomised interleavings) and
compared bit for bit with the oracle.  This checks protocol logic and arithmetic, not the GPU memory model —
the GPU parity tests do that (TAUDEM_B200_SWEEP=levels|walk|hybrid in tests/test_gpu_parity.py)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from taudem_b200 import synth
from util import assert_bits

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu")
CSRC = os.path.join(ROOT, "taudem_b200", "csrc")
BUILD = os.path.join(EMU, "_build")


def _build(tag="", defines=()):
    os.makedirs(BUILD, exist_ok=True)
    src = open(os.path.join(CSRC, "sweep_walk.cu")).read()
    # kernel<<<grid, block, smem, stream>>>(args);  ->  emu_launch(grid, block, [&]{ kernel(args); });
    src, n = re.subn(r"(k_\w+(?:<\w+>)?)<<<([^,]+),\s*([^,]+),[^>]*>>>\(([^;]*)\);",
                     r"emu_launch(dim3(\2), dim3(\3), [&] { \1(\4); });", src)
    assert n >= 6, n
    inc = os.path.join(BUILD, "sweep_walk_emu.inc")
    if not os.path.exists(inc) or open(inc).read() != src:
        open(inc, "w").write(src)
    so = os.path.join(BUILD, f"libemu{tag}.so")
    deps = [inc, os.path.join(EMU, "driver.cpp"), os.path.join(EMU, "emu.cpp"), os.path.join(EMU, "cuda_runtime.h"),
            os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "dinf_common.cuh"), os.path.join(CSRC, "ctx.h")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-I", EMU, "-I", BUILD, "-I", CSRC,
                               *[f"-D{d}" for d in defines], "-o", so, os.path.join(EMU, "driver.cpp"), os.path.join(EMU, "emu.cpp")])
    lib = C.CDLL(so)
    lib.emu_sweep.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int,
                              C.c_float, C.c_double, C.c_double, C.c_ulonglong, C.c_int, C.c_void_p]
    return lib


@pytest.fixture(scope="module")
def emu():
    return _build()


@pytest.fixture(scope="module")
def fields():
    from oracle import port
    dem = synth.punch_holes(synth.gen_dem(150, 190, hurst=0.8, tilt=1.0, seed=5))
    fel = port.pitremove(dem)
    p, _ = port.d8flowdir(fel)
    ang, _ = port.dinfflowdir(fel)
    w = synth.gen_weights(*dem.shape)
    return port, p, ang, w


def _run(lib, dinf, mode, passes, direction, w, contcheck, seed, nstrips=1, rounds=None):
    ny, nx = direction.shape
    out = np.empty((ny, nx), np.float32)
    d = np.ascontiguousarray(direction)
    wp = None if w is None else np.ascontiguousarray(w, np.float32)
    nodata = -3.4028234663852886e38 if dinf else -32768.0
    rc = lib.emu_sweep(int(dinf), mode, passes, d.ctypes.data, out.ctypes.data, None if wp is None else wp.ctypes.data, nx, ny, nodata,
                       int(w is not None), int(contcheck), -9999.0, 30.0, 30.0, seed, nstrips, None if rounds is None else rounds.ctypes.data)
    assert rc == 0
    return out


@pytest.mark.parametrize("mode,passes", [(0, 0), (1, 1), (1, 4), (1, 40)])
def test_emulated_d8_sweeps_match_the_oracle(emu, fields, mode, passes):
    port, p, _, w = fields
    for seed in (1, 2):
        assert_bits(_run(emu, False, mode, passes, p, None, True, seed), port.aread8(p), f"ad8 mode {mode}/{passes} seed {seed}")
    assert_bits(_run(emu, False, mode, passes, p, w, False, 3), port.aread8(p, weights=w, contcheck=False), "ad8 -wg -nc")


@pytest.mark.parametrize("mode,passes", [(0, 0), (1, 1), (1, 4), (1, 40)])
def test_emulated_dinf_sweeps_match_the_oracle(emu, fields, mode, passes):
    port, _, ang, w = fields
    for seed in (1, 2):
        assert_bits(_run(emu, True, mode, passes, ang, None, True, seed), port.areadinf(ang), f"sca mode {mode}/{passes} seed {seed}")
    assert_bits(_run(emu, True, mode, passes, ang, w, False, 3), port.areadinf(ang, weights=w, contcheck=False), "sca -wg -nc")


def test_emulated_dinf_fork_stack_spills_to_the_global_list(fields):
    """With a two-entry fork stack per warp nearly every fork spills: the host loop must drain the spill lists."""
    port, _, ang, _ = fields
    lib = _build("_wq2", ["TD_WALK_WQ=2"])
    for mode, passes, seed in ((0, 0, 7), (1, 3, 8)):
        assert_bits(_run(lib, True, mode, passes, ang, None, True, seed), port.areadinf(ang), f"sca, spilling, mode {mode}")


@pytest.mark.parametrize("nstrips", [2, 3])
def test_emulated_row_strips_with_exchange_rounds(emu, fields, nstrips):
    """The multi-strip protocol of the level / walk sweeps (halo decrement counts, area rows, plain apply, re-collection
    of the ready cells each round) reproduces the single-strip rasters."""
    port, p, ang, w = fields
    rounds = np.zeros(1, np.int32)
    assert_bits(_run(emu, False, 1, 3, p, None, True, 11, nstrips, rounds), port.aread8(p), "ad8 strips")
    assert rounds[0] > 1
    assert_bits(_run(emu, False, 0, 0, p, w, False, 12, nstrips), port.aread8(p, weights=w, contcheck=False), "ad8 -wg -nc strips")
    assert_bits(_run(emu, True, 1, 3, ang, None, True, 13, nstrips, rounds), port.areadinf(ang), "sca strips")
    assert rounds[0] > 1
    assert_bits(_run(emu, True, 0, 0, ang, w, False, 14, nstrips), port.areadinf(ang, weights=w, contcheck=False), "sca -wg -nc strips")
