"""Device-strip level driver: the five tools on device-resident row strips.

torch is used for what it is good at here — device memory, streams, events and (in
``dist.py``) the process group; every computation is a hand-written kernel reached
through the C ABI (include/taudem_b200.h, section 3).
"""
import ctypes as C

import numpy as np
import torch

from ._lib import Strip, check, lib

FEL_NODATA = -3.0e38
MISSINGFLOAT = -3.4028234663852886e38


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class DeviceStrip:
    """Geometry + allocation helper for one row strip (src/linearpart.h:125-166)."""

    def __init__(self, nx, ny, has_top=False, has_bot=False, row0=0, total_ny=None, device="cuda"):
        self.nx, self.ny = int(nx), int(ny)
        self.pitch = lib().td_pitch_for(self.nx)
        self.has_top, self.has_bot = int(has_top), int(has_bot)
        self.row0 = int(row0)
        self.total_ny = int(total_ny if total_ny is not None else ny)
        self.device = torch.device(device)
        self.c = Strip(self.nx, self.ny, self.pitch, self.has_top, self.has_bot)

    @property
    def cells(self):
        return self.nx * self.ny

    def empty(self, dtype):
        return torch.empty((self.ny + 2, self.pitch), dtype=dtype, device=self.device)

    def owned(self, t):
        return t[1:self.ny + 1, :self.nx]

    def rows(self, value):
        return torch.full((self.ny,), float(value), dtype=torch.float64, device=self.device)


class Tools:
    """One td_ctx (scratch queues/counters) bound to the current CUDA device."""

    def __init__(self):
        self.l = lib()
        self.ctx = self.l.td_ctx_create()
        if not self.ctx:
            raise RuntimeError("td_ctx_create failed: " + self.l.td_last_error().decode())

    def close(self):
        if self.ctx:
            self.l.td_ctx_destroy(self.ctx)
            self.ctx = None

    @staticmethod
    def _stream():
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    # synthetic inputs -------------------------------------------------------------
    def gen_dem(self, s, seed=1234, hurst=0.8, tilt=1.0):
        dem = s.empty(torch.float32)
        check(self.l.td_gen_dem_dev(_p(dem), s.c, s.row0, s.total_ny, seed, hurst, tilt, self._stream()))
        return dem

    def gen_weights(self, s, seed=4321):
        w = s.empty(torch.float32)
        check(self.l.td_gen_weights_dev(_p(w), s.c, s.row0, seed, self._stream()))
        return w

    # tools ------------------------------------------------------------------------
    def flood_init(self, s, dem, nodata=-9999.0, four_way=False, depmask=None):
        w = s.empty(torch.float32)
        check(self.l.td_flood_init_dev(self.ctx, _p(dem), _p(depmask), _p(w), s.c, nodata, int(four_way), self._stream()))
        return w

    def flood_relax(self, s, dem, w, four_way=False, edges_only=False):
        ch = C.c_int(0)
        fn = self.l.td_flood_relax_edges_dev if edges_only else self.l.td_flood_relax_dev
        check(fn(self.ctx, _p(dem), _p(w), s.c, int(four_way), C.byref(ch), self._stream()))
        return bool(ch.value)

    def pitremove(self, s, dem, nodata=-9999.0, four_way=False):
        w = self.flood_init(s, dem, nodata, four_way)
        self.flood_relax(s, dem, w, four_way)
        return w

    def d8_slopes(self, s, fel, dxc, dyc, nodata=FEL_NODATA, p=None, sd8=None):
        p = s.empty(torch.int16) if p is None else p
        sd8 = s.empty(torch.float32) if sd8 is None else sd8
        nflat = C.c_longlong(0)
        check(self.l.td_d8_slopes_dev(self.ctx, _p(fel), _p(p), _p(sd8), s.c, nodata, _p(dxc), _p(dyc), C.byref(nflat), self._stream()))
        return p, sd8, nflat.value

    def d8_flats(self, s, fel, p, dxc, dyc):
        left = C.c_longlong(0)
        check(self.l.td_d8_flats_dev(self.ctx, _p(fel), _p(p), s.c, _p(dxc), _p(dyc), C.byref(left), self._stream()))
        return left.value

    def dinf_slopes(self, s, fel, dxc, dyc, nodata=FEL_NODATA, ang=None, slp=None):
        ang = s.empty(torch.float32) if ang is None else ang
        slp = s.empty(torch.float32) if slp is None else slp
        nflat = C.c_longlong(0)
        check(self.l.td_dinf_slopes_dev(self.ctx, _p(fel), _p(ang), _p(slp), s.c, nodata, _p(dxc), _p(dyc), C.byref(nflat), self._stream()))
        return ang, slp, nflat.value

    def dinf_flats(self, s, fel, ang, dxc, dyc):
        left = C.c_longlong(0)
        check(self.l.td_dinf_flats_dev(self.ctx, _p(fel), _p(ang), s.c, _p(dxc), _p(dyc), C.byref(left), self._stream()))
        return left.value

    def aread8_deps(self, s, p, ad8, nodata=-32768):
        check(self.l.td_aread8_deps_dev(self.ctx, _p(p), _p(ad8), s.c, nodata, self._stream()))

    def aread8_sweep(self, s, ad8, w=None, w_nodata=-9999.0, contcheck=True):
        check(self.l.td_aread8_sweep_dev(self.ctx, _p(w), _p(ad8), s.c, w_nodata, int(w is not None), int(contcheck), self._stream()))

    def aread8(self, s, p, ad8=None, w=None, nodata=-32768, w_nodata=-9999.0, contcheck=True):
        ad8 = s.empty(torch.float32) if ad8 is None else ad8
        self.aread8_deps(s, p, ad8, nodata)
        self.aread8_sweep(s, ad8, w, w_nodata, contcheck)
        return ad8

    def areadinf_deps(self, s, ang, sca, dxc, dyc, nodata=MISSINGFLOAT):
        check(self.l.td_area_deps_dev(self.ctx, _p(ang), _p(sca), s.c, nodata, _p(dxc), _p(dyc), self._stream()))

    def areadinf_sweep(self, s, ang, sca, dxc, w=None, contcheck=True):
        check(self.l.td_area_sweep_dev(self.ctx, _p(ang), _p(w), _p(sca), s.c, int(w is not None), int(contcheck), _p(dxc), self._stream()))

    def areadinf(self, s, ang, dxc, dyc, sca=None, w=None, nodata=MISSINGFLOAT, contcheck=True):
        sca = s.empty(torch.float32) if sca is None else sca
        self.areadinf_deps(s, ang, sca, dxc, dyc, nodata)
        self.areadinf_sweep(s, ang, sca, dxc, w, contcheck)
        return sca
