"""Driver of the sweep scheduling model (sim.c): flow field -> downslope lists -> policies / tile shapes.
Research tooling only (see sim.c)."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libsweepsim.so")
DR = np.array([0, 0, -1, -1, -1, 0, 1, 1, 1])   # d8 code k -> row step (1 = east, counter-clockwise; src/commonLib.h d1/d2)
DC = np.array([0, 1, 1, 0, -1, -1, -1, 0, 1])


def lib():
    if not os.path.exists(SO) or os.path.getmtime(SO) < os.path.getmtime(os.path.join(HERE, "sim.c")):
        subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", SO, os.path.join(HERE, "sim.c")])
    l = C.CDLL(SO)
    l.sim_run.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                          C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
    return l


def _target(k, ny, nx):
    r, c = np.meshgrid(np.arange(ny, dtype=np.int64), np.arange(nx, dtype=np.int64), indexing="ij")
    rr, cc = r + DR[k], c + DC[k]
    ok = (k >= 1) & (k <= 8) & (rr >= 0) & (rr < ny) & (cc >= 0) & (cc < nx)
    return np.where(ok, rr * nx + cc, -1).astype(np.int32)


def graph_d8(p):
    ny, nx = p.shape
    k = np.where((p >= 1) & (p <= 8), p, 0).astype(np.int64)
    return _target(k, ny, nx), None


def graph_dinf(ang):
    """square-cell sector rule: flow splits between direction j+1 and j+2 of sector j = floor(ang / 45 deg)"""
    ny, nx = ang.shape
    ok = (ang >= 0) & (ang < 7.0)
    a = np.where(ok, ang, 0).astype(np.float64)
    sec = np.floor(a / (np.pi / 4)).astype(np.int64) % 8
    frac = a / (np.pi / 4) - np.floor(a / (np.pi / 4))
    k1 = np.where(ok, sec + 1, 0)
    k2 = np.where(ok & (frac >= 1e-5), (sec + 1) % 8 + 1, 0)
    k1 = np.where(ok & (1 - frac < 1e-5), 0, k1)
    return _target(k1, ny, nx), _target(k2, ny, nx)


def counts(d1, d2):
    n = d1.size
    cnt = np.bincount(d1.ravel()[d1.ravel() >= 0], minlength=n)
    if d2 is not None:
        cnt = cnt + np.bincount(d2.ravel()[d2.ravel() >= 0], minlength=n)
    return cnt.astype(np.uint8)


def run(d1, d2, cnt, twx, twy, workers, policy=0, key=None, c_fixed=10.0, c_cell=0.01, c_hop=0.1, cnt_out=None):
    ny, nx = d1.shape
    ntx, nty = -(-nx // twx), -(-ny // twy)
    res = np.zeros(8)
    vpt = np.zeros(ntx * nty, np.int32)
    keyarr = None if key is None else np.ascontiguousarray(key, np.float64)
    kp = None if keyarr is None else keyarr.ctypes.data
    lib().sim_run(nx, ny, d1.ctypes.data, None if d2 is None else d2.ctypes.data, cnt.ctypes.data, twx, twy, workers, policy,
                  kp, c_fixed, c_cell, c_hop, res.ctypes.data, vpt.ctypes.data, None if cnt_out is None else cnt_out.ctypes.data)
    return dict(visits=res[0], per_tile=res[0] / (ntx * nty), makespan_ms=res[1] / 1e3, busy_ms=res[2] / 1e3 / workers,
                cells=res[3], max_visits=res[4], empty=res[5], left=res[6]), vpt.reshape(nty, ntx)


def tile_key_elev(fel, twx, twy):
    ny, nx = fel.shape
    ntx, nty = -(-nx // twx), -(-ny // twy)
    pad = np.full((nty * twy, ntx * twx), np.nan, np.float32)
    pad[:ny, :nx] = fel
    return -np.nanmean(pad.reshape(nty, twy, ntx, twx), axis=(1, 3)).ravel()


def longest(d1, d2, cnt, weight=None):
    ny, nx = d1.shape
    l = lib()
    l.sim_longest.restype = C.c_long
    l.sim_longest.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    depth = np.zeros(d1.size, np.int32)
    w = None if weight is None else np.ascontiguousarray(weight, np.uint8)
    best = l.sim_longest(nx, ny, d1.ctypes.data, None if d2 is None else d2.ctypes.data, cnt.ctypes.data,
                         None if w is None else w.ctypes.data, depth.ctypes.data)
    return best, depth.reshape(ny, nx)


def local_cells(d1, d2, cnt, twx, twy):
    ny, nx = d1.shape
    l = lib()
    l.sim_local.restype = C.c_long
    l.sim_local.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    loc = np.zeros(d1.size, np.uint8)
    l.sim_local(nx, ny, d1.ctypes.data, None if d2 is None else d2.ctypes.data, cnt.ctypes.data, twx, twy, loc.ctypes.data)
    return loc.reshape(ny, nx)


def level_passes(d1, d2, cnt, band, passes, direction=1):
    """Model of k_level: `passes` streaming passes, rows visited in bands of `band` rows (direction +1 forward, -1 reverse,
    0 alternating).  Returns (cells evaluated after each pass, done mask)."""
    ny, nx = d1.shape
    l = lib()
    l.sim_levelpasses.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    out = np.zeros(passes, np.int64); done = np.zeros(d1.size, np.uint8)
    l.sim_levelpasses(nx, ny, d1.ctypes.data, None if d2 is None else d2.ctypes.data, cnt.ctypes.data, band, passes, direction,
                      out.ctypes.data, done.ctypes.data)
    return out, done.reshape(ny, nx)
