"""Host-side mirror of the reference's interface for the hot path.

File level: ``flood``, ``setdird8``, ``setdir``, ``aread8``, ``area`` take the same
arguments, in the same order and with the same meaning, as the reference functions
(src/flood.cpp:50, src/d8.cpp:181, src/dinf.cpp:109, src/aread8.cpp:56,
src/areadinf.cpp:53) and return 0 on success like they do.

Grid level: ``*_grid`` functions run the same device path on numpy arrays (row 0 =
north).  dx/dy may be scalars (projected grids) or per-row arrays (geographic grids,
tiffIO::getdxc/getdyc).
"""
import ctypes as C

import numpy as np

from ._lib import check, lib

_DT = {np.dtype(np.int16): 0, np.dtype(np.int32): 1, np.dtype(np.float32): 2}

FEL_NODATA = np.float32(-3.0e38)            # src/flood.cpp:136
MISSINGFLOAT = np.float32(-3.4028234663852886e38)   # src/commonLib.h:80
MISSINGSHORT = np.int16(-32768)             # src/commonLib.h:77


def _b(s):
    return (s or "").encode()


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _rows(v, ny):
    a = np.asarray(v, dtype=np.float64)
    if a.ndim == 0:
        a = np.full(ny, float(a), dtype=np.float64)
    if a.shape != (ny,):
        raise ValueError("per-row cell sizes must have one value per row")
    return np.ascontiguousarray(a)


def _grid(a, dtype):
    a = np.ascontiguousarray(a, dtype=dtype)
    if a.ndim != 2:
        raise ValueError("rasters are 2-D arrays")
    return a


# --------------------------------------------------------------------- file level
def flood(demfile, felfile, sfdrfile="", usesfdr=0, verbose=False, is_4Point=False, use_mask=False, maskfile=""):
    return lib().td_flood(_b(demfile), _b(felfile), _b(sfdrfile), int(usesfdr), int(verbose), int(is_4Point), int(use_mask), _b(maskfile))


def setdird8(demfile, pointfile, slopefile, flowfile="", useflowfile=0):
    return lib().td_setdird8(_b(demfile), _b(pointfile), _b(slopefile), _b(flowfile), int(useflowfile))


def setdir(demfile, angfile, slopefile, flowfile="", useflowfile=0):
    return lib().td_setdir(_b(demfile), _b(angfile), _b(slopefile), _b(flowfile), int(useflowfile))


def aread8(pfile, afile, datasrc="", lyrname="", uselyrname=0, lyrno=0, wfile="", useOutlets=0, usew=0, contcheck=1):
    return lib().td_aread8(_b(pfile), _b(afile), _b(datasrc), _b(lyrname), int(uselyrname), int(lyrno), _b(wfile), int(useOutlets), int(usew), int(contcheck))


def area(angfile, scafile, datasrc="", lyrname="", uselyrname=0, lyrno=0, wfile="", useOutlets=0, usew=0, contcheck=1):
    return lib().td_area(_b(angfile), _b(scafile), _b(datasrc), _b(lyrname), int(uselyrname), int(lyrno), _b(wfile), int(useOutlets), int(usew), int(contcheck))


def nameadd(arg, suff):
    buf = C.create_string_buffer(4096)
    lib().td_nameadd(buf, _b(arg), _b(suff))
    return buf.value.decode()


# --------------------------------------------------------------------- raster files
def raster_info(path):
    nx, ny, hn, geo, bits, fmt = C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_int()
    nd, dx, dy = C.c_double(), C.c_double(), C.c_double()
    check(lib().td_raster_info(_b(path), C.byref(nx), C.byref(ny), C.byref(nd), C.byref(hn), C.byref(dx), C.byref(dy),
                               C.byref(geo), C.byref(bits), C.byref(fmt)))
    return dict(nx=nx.value, ny=ny.value, nodata=nd.value, has_nodata=bool(hn.value), dx=dx.value, dy=dy.value,
                is_geographic=bool(geo.value), bits=bits.value, sample_format=fmt.value)


def read_raster(path, dtype=np.float32):
    info = raster_info(path)
    out = np.empty((info["ny"], info["nx"]), dtype=dtype)
    check(lib().td_raster_read(_b(path), _DT[np.dtype(dtype)], _ptr(out), info["nx"], info["ny"]))
    return out


def write_raster(path, arr, nodata, like=None, dx=30.0, dy=30.0, compression=1):
    arr = np.ascontiguousarray(arr)
    check(lib().td_raster_write(_b(path), _DT[arr.dtype], _ptr(arr), arr.shape[1], arr.shape[0], float(nodata),
                                _b(like) if like else None, float(dx), float(dy), int(compression)))


# --------------------------------------------------------------------- grid level
def pitremove_grid(dem, nodata=-9999.0, depmask=None, is_4Point=False, out=None):
    dem = _grid(dem, np.float32)
    ny, nx = dem.shape
    fel = out if out is not None else np.empty_like(dem)
    m = None if depmask is None else _grid(depmask, np.int16)
    check(lib().td_flood_host(_ptr(dem), _ptr(fel), _ptr(m), nx, ny, np.float32(nodata), int(is_4Point)))
    return fel


def d8flowdir_grid(fel, nodata=float(FEL_NODATA), dx=30.0, dy=30.0, out=None):
    fel = _grid(fel, np.float32)
    ny, nx = fel.shape
    p, sd8 = out if out is not None else (np.empty((ny, nx), np.int16), np.empty((ny, nx), np.float32))
    dxc, dyc = _rows(dx, ny), _rows(dy, ny)
    check(lib().td_setdird8_host(_ptr(fel), _ptr(p), _ptr(sd8), nx, ny, np.float32(nodata), _ptr(dxc), _ptr(dyc)))
    return p, sd8


def dinfflowdir_grid(fel, nodata=float(FEL_NODATA), dx=30.0, dy=30.0, out=None):
    fel = _grid(fel, np.float32)
    ny, nx = fel.shape
    ang, slp = out if out is not None else (np.empty((ny, nx), np.float32), np.empty((ny, nx), np.float32))
    dxc, dyc = _rows(dx, ny), _rows(dy, ny)
    check(lib().td_setdir_host(_ptr(fel), _ptr(ang), _ptr(slp), nx, ny, np.float32(nodata), _ptr(dxc), _ptr(dyc)))
    return ang, slp


def _outlet_args(outlets):
    """outlets = (cols, rows) of the outlet cells -> (cols array, rows array, n); None -> no outlets (n = -1)."""
    if outlets is None:
        return None, None, -1
    c = np.ascontiguousarray(outlets[0], np.int32); r = np.ascontiguousarray(outlets[1], np.int32)
    assert c.shape == r.shape and c.ndim == 1
    return c, r, len(c)


def read_outlets(datasrc, lyrname="", uselyrname=0, lyrno=0):
    """Outlet points (x, y) of a shapefile / GeoJSON data source (readoutlets, src/ReadOutlets.cpp)."""
    n = C.c_int(0)
    check(lib().td_outlets_read(_b(datasrc), _b(lyrname), int(uselyrname), int(lyrno), None, None, 0, C.byref(n)))
    x = np.empty(max(n.value, 1), np.float64); y = np.empty(max(n.value, 1), np.float64)
    check(lib().td_outlets_read(_b(datasrc), _b(lyrname), int(uselyrname), int(lyrno), _ptr(x), _ptr(y), n.value, C.byref(n)))
    return x[:n.value], y[:n.value]


def aread8_grid(p, nodata=int(MISSINGSHORT), weights=None, w_nodata=-9999.0, contcheck=True, out=None, outlets=None):
    p = _grid(p, np.int16)
    ny, nx = p.shape
    ad8 = out if out is not None else np.empty((ny, nx), np.float32)
    w = None if weights is None else _grid(weights, np.float32)
    oc, orow, nout = _outlet_args(outlets)
    check(lib().td_aread8_outlets_host(_ptr(p), _ptr(w), _ptr(ad8), nx, ny, int(nodata), np.float32(w_nodata), int(contcheck),
                                       _ptr(oc), _ptr(orow), nout))
    return ad8


def d8flowpathextremeup_grid(p, sa, usemax=True, nodata=int(MISSINGSHORT), contcheck=True, outlets=None):
    """The largest (smallest) value of `sa` on the D8 flow paths above each cell (td_d8flowpathextremeup_host;
    src/D8flowpathextremeup.cpp:182-215).  nodata = -FLT_MAX."""
    p = _grid(p, np.int16); sa = _grid(sa, np.float32)
    ny, nx = p.shape
    assert sa.shape == p.shape
    ssa = np.empty((ny, nx), np.float32)
    oc, orow, nout = _outlet_args(outlets)
    check(lib().td_d8flowpathextremeup_host(_ptr(p), _ptr(sa), _ptr(ssa), nx, ny, int(nodata), int(usemax), int(contcheck), _ptr(oc), _ptr(orow), nout))
    return ssa


def gridnet_grid(p, mask=None, thresh=0, dx=30.0, dy=30.0, nodata=int(MISSINGSHORT), outlets=None, dxc=None, dyc=None):
    """(plen, tlen, gord): longest and total upstream path length and Strahler order of the D8 flow field (td_gridnet_host;
    src/gridnet.cpp:383-420).  mask (int32): only cells with mask >= thresh are evaluated and contribute.  nodata = -1."""
    p = _grid(p, np.int16)
    ny, nx = p.shape
    m = None if mask is None else _grid(mask, np.int32)
    dxc = _rows(dx, ny) if dxc is None else np.ascontiguousarray(dxc, np.float64)
    dyc = _rows(dy, ny) if dyc is None else np.ascontiguousarray(dyc, np.float64)
    plen = np.empty((ny, nx), np.float32); tlen = np.empty((ny, nx), np.float32); gord = np.empty((ny, nx), np.int16)
    oc, orow, nout = _outlet_args(outlets)
    check(lib().td_gridnet_host(_ptr(p), _ptr(m), int(thresh), _ptr(plen), _ptr(tlen), _ptr(gord), nx, ny, int(nodata), _ptr(dxc), _ptr(dyc),
                                _ptr(oc), _ptr(orow), nout))
    return plen, tlen, gord


def dinfdecayaccum_grid(ang, dm, weights=None, dx=30.0, dy=30.0, nodata=float(MISSINGFLOAT), dm_nodata=-9999.0, contcheck=True, outlets=None,
                        dxc=None, dyc=None):
    """Decaying accumulation on the D-infinity flow field (td_dinfdecayaccum_host; src/dinfdecayaccum.cpp:205-235): a cell starts from
    its weight (or dx) and receives (float)(dm * area * p) from every contributor.  nodata = -FLT_MAX."""
    ang = _grid(ang, np.float32); dm = _grid(dm, np.float32)
    ny, nx = ang.shape
    assert dm.shape == ang.shape
    w = None if weights is None else _grid(weights, np.float32)
    dxc = _rows(dx, ny) if dxc is None else np.ascontiguousarray(dxc, np.float64)
    dyc = _rows(dy, ny) if dyc is None else np.ascontiguousarray(dyc, np.float64)
    out = np.empty((ny, nx), np.float32)
    oc, orow, nout = _outlet_args(outlets)
    check(lib().td_dinfdecayaccum_host(_ptr(ang), _ptr(dm), _ptr(w), _ptr(out), nx, ny, np.float32(nodata), np.float32(dm_nodata), _ptr(dxc), _ptr(dyc),
                                       int(contcheck), _ptr(oc), _ptr(orow), nout))
    return out


def dinfconclimaccum_grid(ang, dm, q, dg, csol=1.0, dx=30.0, dy=30.0, nodata=float(MISSINGFLOAT), dm_nodata=-9999.0, q_nodata=-9999.0, contcheck=True,
                          outlets=None):
    """Concentration limited accumulation on the D-infinity flow field (td_dinfconclimaccum_host; src/DinfConcLimAccum.cpp:242-270):
    cells with q > 0 only; an indicator cell (dg > 0) has the concentration csol, any other the float sum of p * ctpt * q * dm over its
    contributors divided by its own q.  nodata = -FLT_MAX."""
    ang = _grid(ang, np.float32); dm = _grid(dm, np.float32); q = _grid(q, np.float32); dg = _grid(dg, np.int16)
    ny, nx = ang.shape
    assert dm.shape == ang.shape and q.shape == ang.shape and dg.shape == ang.shape
    dxc, dyc = _rows(dx, ny), _rows(dy, ny)
    out = np.empty((ny, nx), np.float32)
    oc, orow, nout = _outlet_args(outlets)
    check(lib().td_dinfconclimaccum_host(_ptr(ang), _ptr(dm), _ptr(q), _ptr(dg), _ptr(out), nx, ny, np.float32(nodata), np.float32(dm_nodata), np.float32(q_nodata),
                                         np.float32(csol), _ptr(dxc), _ptr(dyc), int(contcheck), _ptr(oc), _ptr(orow), nout))
    return out


def dinftranslimaccum_grid(ang, tsup, tc, cs=None, dx=30.0, dy=30.0, nodata=float(MISSINGFLOAT), tsup_nodata=-9999.0, tc_nodata=-9999.0, cs_nodata=-9999.0,
                           contcheck=True, outlets=None):
    """Transport limited accumulation on the D-infinity flow field (td_dinftranslimaccum_host; src/DinfTransLimAccum.cpp:237-302):
    returns (tla, tdep, ctpt) — ctpt is None without a supply concentration grid `cs`.  nodata = -FLT_MAX."""
    ang = _grid(ang, np.float32); tsup = _grid(tsup, np.float32); tc = _grid(tc, np.float32)
    ny, nx = ang.shape
    assert tsup.shape == ang.shape and tc.shape == ang.shape
    c = None if cs is None else _grid(cs, np.float32)
    dxc, dyc = _rows(dx, ny), _rows(dy, ny)
    tla = np.empty((ny, nx), np.float32); dep = np.empty((ny, nx), np.float32)
    cout = None if cs is None else np.empty((ny, nx), np.float32)
    oc, orow, nout = _outlet_args(outlets)
    check(lib().td_dinftranslimaccum_host(_ptr(ang), _ptr(tsup), _ptr(tc), _ptr(c), _ptr(tla), _ptr(dep), _ptr(cout), nx, ny, np.float32(nodata),
                                          np.float32(tsup_nodata), np.float32(tc_nodata), np.float32(cs_nodata), _ptr(dxc), _ptr(dyc), int(contcheck),
                                          _ptr(oc), _ptr(orow), nout))
    return tla, dep, cout


def threshold_grid(ssa, thresh=100.0, mask=None, nodata=-1.0):
    """src = (ssa >= thresh [& mask >= 0]) ? 1 : 0, -32768 where ssa is nodata (td_threshold_host; src/Threshold.cpp:109-131)."""
    ssa = _grid(ssa, np.float32)
    ny, nx = ssa.shape
    m = None if mask is None else _grid(mask, np.float32)
    src = np.empty((ny, nx), np.int16)
    check(lib().td_threshold_host(_ptr(ssa), _ptr(m), _ptr(src), nx, ny, np.float32(thresh), np.float32(nodata)))
    return src


def twi_grid(slp, sca, slp_nodata=-1.0, sca_nodata=-1.0):
    """twi = ln(sca / slp) where both are data and positive, else -1 (td_twi_host; src/TWI.cpp:108-124)."""
    slp = _grid(slp, np.float32); sca = _grid(sca, np.float32)
    ny, nx = slp.shape
    assert sca.shape == slp.shape
    twi = np.empty((ny, nx), np.float32)
    check(lib().td_twi_host(_ptr(slp), _ptr(sca), _ptr(twi), nx, ny, np.float32(slp_nodata), np.float32(sca_nodata)))
    return twi


def slopearea_grid(slp, sca, m=2.0, n=1.0):
    """sa = slp^m * sca^n where slp >= 0 and sca >= 0, else -1 (td_slopearea_host; src/SlopeArea.cpp:114-125)."""
    slp = _grid(slp, np.float32); sca = _grid(sca, np.float32)
    ny, nx = slp.shape
    assert sca.shape == slp.shape
    sa = np.empty((ny, nx), np.float32)
    check(lib().td_slopearea_host(_ptr(slp), _ptr(sca), _ptr(sa), nx, ny, np.float32(m), np.float32(n)))
    return sa


def slopearearatio_grid(slp, sca, sca_nodata=-1.0):
    """sar = slp / sca where sca is data, else -1 (td_slopearearatio_host; src/SlopeAreaRatio.cpp:107-118)."""
    slp = _grid(slp, np.float32); sca = _grid(sca, np.float32)
    ny, nx = slp.shape
    assert sca.shape == slp.shape
    sar = np.empty((ny, nx), np.float32)
    check(lib().td_slopearearatio_host(_ptr(slp), _ptr(sca), _ptr(sar), nx, ny, np.float32(sca_nodata)))
    return sar


def contributing_areas_grid(p, ang, p_nodata=int(MISSINGSHORT), ang_nodata=float(MISSINGFLOAT), dx=30.0, dy=30.0, contcheck=True, out_ad8=None, out_sca=None):
    """aread8 + areadinf of one DEM in one call, copies overlapped with the kernels (td_contributing_areas_host)."""
    p = _grid(p, np.int16); ang = _grid(ang, np.float32)
    ny, nx = p.shape
    assert ang.shape == p.shape
    ad8 = out_ad8 if out_ad8 is not None else np.empty((ny, nx), np.float32)
    sca = out_sca if out_sca is not None else np.empty((ny, nx), np.float32)
    dxc, dyc = _rows(dx, ny), _rows(dy, ny)
    check(lib().td_contributing_areas_host(_ptr(p), _ptr(ang), _ptr(ad8), _ptr(sca), nx, ny, int(p_nodata), np.float32(ang_nodata), _ptr(dxc), _ptr(dyc), int(contcheck)))
    return ad8, sca


def areadinf_grid(ang, nodata=float(MISSINGFLOAT), weights=None, w_nodata=-9999.0, dx=30.0, dy=30.0, contcheck=True, out=None, outlets=None):
    ang = _grid(ang, np.float32)
    ny, nx = ang.shape
    sca = out if out is not None else np.empty((ny, nx), np.float32)
    w = None if weights is None else _grid(weights, np.float32)
    dxc, dyc = _rows(dx, ny), _rows(dy, ny)
    oc, orow, nout = _outlet_args(outlets)
    check(lib().td_area_outlets_host(_ptr(ang), _ptr(w), _ptr(sca), nx, ny, np.float32(nodata), np.float32(w_nodata), _ptr(dxc), _ptr(dyc),
                                     int(contcheck), _ptr(oc), _ptr(orow), nout))
    return sca
