"""Synthetic fractal DEMs (numpy twin of csrc/gen.cu).

Multi-octave lattice value noise with an integer hash, amplitude 2^(-H*octave),
scaled to 1000 m relief + 100 m, plus the plane tilt*1000*(y + 0.37 x)/n.
Two families: ``rough`` (H=0.6, tilt 0: about a third of the cells end up in flats
after pit filling) and ``tilted`` (H=0.8, tilt 8: <1 % flats).
"""
import numpy as np

FAMILIES = {"rough": dict(hurst=0.6, tilt=0.0), "tilted": dict(hurst=0.8, tilt=8.0)}


def _hash3(x, y, z):
    x = x.astype(np.uint32)
    y = y.astype(np.uint32)
    z = np.uint32(z)
    with np.errstate(over="ignore"):
        h = (x * np.uint32(0x9E3779B1)) ^ (y * np.uint32(0x85EBCA77) + np.uint32(0xC2B2AE3D)) ^ (z * np.uint32(0x27D4EB2F) + np.uint32(0x165667B1))
        h ^= h >> np.uint32(15)
        h = h * np.uint32(0x2C1B3C6D)
        h ^= h >> np.uint32(12)
        h = h * np.uint32(0x297A2D39)
        h ^= h >> np.uint32(15)
    return h


def _u01(h):
    return (h >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)


def gen_dem(ny, nx=None, seed=1234, hurst=0.8, tilt=8.0, family=None):
    if nx is None:
        nx = ny
    if family is not None:
        hurst, tilt = FAMILIES[family]["hurst"], FAMILIES[family]["tilt"]
    n = max(nx, ny)
    noct = 0
    while (n >> (noct + 1)) >= 2:
        noct += 1
    yy, xx = np.meshgrid(np.arange(ny, dtype=np.float32), np.arange(nx, dtype=np.float32), indexing="ij")
    total = np.zeros((ny, nx), np.float32)
    norm = np.float32(0)
    for o in range(noct):
        L = np.float32(n) / np.float32(2 << o)
        gx, gy = xx / L, yy / L
        fx0, fy0 = np.floor(gx), np.floor(gy)
        ix, iy = fx0.astype(np.uint32), fy0.astype(np.uint32)
        fx, fy = gx - fx0, gy - fy0
        fx = fx * fx * (np.float32(3) - np.float32(2) * fx)
        fy = fy * fy * (np.float32(3) - np.float32(2) * fy)
        so = np.uint32((seed + 7919 * o) & 0xFFFFFFFF)
        v00, v10 = _u01(_hash3(ix, iy, so)), _u01(_hash3(ix + np.uint32(1), iy, so))
        v01, v11 = _u01(_hash3(ix, iy + np.uint32(1), so)), _u01(_hash3(ix + np.uint32(1), iy + np.uint32(1), so))
        a = v00 + (v10 - v00) * fx
        b = v01 + (v11 - v01) * fx
        amp = np.float32(2.0 ** (-hurst * o))
        total += amp * (a + (b - a) * fy)
        norm += amp
    v = total / norm
    return (np.float32(100) + np.float32(1000) * v + np.float32(tilt) * np.float32(1000) * (yy + np.float32(0.37) * xx) / np.float32(n)).astype(np.float32)


def gen_weights(ny, nx=None, seed=4321):
    if nx is None:
        nx = ny
    yy, xx = np.meshgrid(np.arange(ny, dtype=np.uint32), np.arange(nx, dtype=np.uint32), indexing="ij")
    return _u01(_hash3(xx, yy, np.uint32(seed)))


def punch_holes(dem, nodata=-9999.0, seed=7):
    """Rectangular + blob nodata holes for parity tests."""
    rng = np.random.default_rng(seed)
    out = dem.copy()
    ny, nx = dem.shape
    y0, x0 = ny // 5, nx // 3
    out[y0:y0 + max(2, ny // 12), x0:x0 + max(3, nx // 9)] = nodata
    cy, cx, rad = int(ny * 0.7), int(nx * 0.6), max(2, min(ny, nx) // 14)
    yy, xx = np.ogrid[:ny, :nx]
    out[(yy - cy) ** 2 + (xx - cx) ** 2 <= rad * rad] = nodata
    for _ in range(3):
        out[rng.integers(2, ny - 2), rng.integers(2, nx - 2)] = nodata
    return out
