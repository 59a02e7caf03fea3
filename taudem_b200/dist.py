"""Row-strip partitioned execution over several GPUs (one process per GPU).

Mirrors the reference's only parallelisation (src/linearpart.h): the grid is cut into
contiguous row strips, ``ny = total // size`` rows each with the remainder on the LAST rank
(linearpart::init, :125-166); neighbouring strips exchange one halo row (share(), :195-219)
and, for the contributing-area wavefront, the dependency decrements that crossed the strip
boundary (addBorders(), :314-328); a global reduction decides termination (ringTerm(), :344-384).
torch.distributed (NCCL send/recv + all_reduce on GPUs, gloo in the CPU tests) is the transport.
"""
import ctypes as C

import torch
import torch.distributed as dist

from ._lib import check, lib


def partition(total_ny, world):
    """[(row0, ny)] per rank, exactly like linearpart::init."""
    ny = total_ny // world
    out = []
    for r in range(world):
        n = ny + (total_ny % world if r == world - 1 else 0)
        out.append((r * ny, n))
    return out


def _staged():
    """gloo moves host memory only: device rows are staged through the host (used by the 2-rank
    single-GPU test; the production transport is NCCL over NVLink)."""
    return dist.get_backend() == "gloo"


def _p2p(sends, recvs, group=None):
    """sends: [(tensor, peer)], recvs: [(tensor, peer)] posted as one batch."""
    if not sends and not recvs:
        return
    stage = _staged()
    ops, back = [], []
    for t, peer in sends:
        ops.append(dist.P2POp(dist.isend, t.cpu() if (stage and t.is_cuda) else t, peer, group))
    for t, peer in recvs:
        if stage and t.is_cuda:
            h = torch.empty(t.shape, dtype=t.dtype)
            back.append((t, h))
            ops.append(dist.P2POp(dist.irecv, h, peer, group))
        else:
            ops.append(dist.P2POp(dist.irecv, t, peer, group))
    for w in dist.batch_isend_irecv(ops):
        w.wait()
    for t, h in back:
        t.copy_(h)


def all_reduce_scalar(value, op=None, device="cpu"):
    t = torch.tensor([value], dtype=torch.int64, device="cpu" if _staged() else device)
    dist.all_reduce(t, op=op or dist.ReduceOp.SUM)
    return int(t.item())


def exchange_rows(t, ny, rank, world, group=None):
    """share(): my first owned row -> rank-1's bottom halo, my last owned row -> rank+1's top halo.
    ``t`` has ny+2 rows (row 0 / ny+1 are the halos)."""
    if world == 1:
        return
    row = lambda i: t[i].view(torch.uint8)      # rows travel as bytes (NCCL has no int16)
    sends, recvs = [], []
    if rank > 0:
        sends.append((row(1), rank - 1)); recvs.append((row(0), rank - 1))
    if rank < world - 1:
        sends.append((row(ny), rank + 1)); recvs.append((row(ny + 1), rank + 1))
    _p2p(sends, recvs, group)


def exchange_counts(halo_out, pitch, rank, world, group=None):
    """Sends the decrements recorded for the strip above / below, returns (dec_top, dec_bot): the
    decrements the neighbours recorded for my first / last row (None at the grid edge)."""
    dec_top = torch.zeros(pitch, dtype=torch.int32, device=halo_out.device) if rank > 0 else None
    dec_bot = torch.zeros(pitch, dtype=torch.int32, device=halo_out.device) if rank < world - 1 else None
    sends, recvs = [], []
    if rank > 0:
        sends.append((halo_out[:pitch], rank - 1)); recvs.append((dec_top, rank - 1))
    if rank < world - 1:
        sends.append((halo_out[pitch:], rank + 1)); recvs.append((dec_bot, rank + 1))
    _p2p(sends, recvs, group)
    return dec_top, dec_bot


class DistTools:
    """The strip of this rank plus the exchange rounds around the device-strip level C ABI."""

    def __init__(self, nx, total_ny, rank, world, device="cuda", peer=None):
        import os
        from .device import DeviceStrip, Tools
        # peer mode: sweeps deliver across GPUs inside the kernel (CUDA IPC + NVLink atomics), no exchange
        # rounds.  Needs one GPU per rank and the NCCL backend; TAUDEM_B200_PEER=0/1 overrides.
        # Default: on for NCCL ranks (one GPU each); the round-based exchange remains for gloo / shared-GPU tests.
        if peer is None:
            env = os.environ.get("TAUDEM_B200_PEER")
            if env is not None:
                peer = env == "1"
            else:
                peer = dist.is_initialized() and dist.get_backend() == "nccl" and world >= 2
        self.peer = bool(peer) and world > 1
        self._peer_cache = None
        self.rank, self.world = rank, world
        self.row0, self.ny = partition(total_ny, world)[rank]
        self.s = DeviceStrip(nx, self.ny, has_top=rank > 0, has_bot=rank < world - 1, row0=self.row0, total_ny=total_ny, device=device)
        self.T = Tools()
        self.l = lib()
        self.rounds = 0

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def share(self, t):
        exchange_rows(t, self.ny, self.rank, self.world)

    def _peer_setup(self, dinf):
        """Exports this rank's IPC handles, gathers everybody's, opens the neighbours' (cached while unchanged)."""
        import numpy as np
        s = self.s
        handles = np.zeros(320, np.uint8); meta = np.zeros(8, np.int32)
        check(self.l.td_sweep_peer_export_dev(self.T.ctx, s.c, int(dinf), handles.ctypes.data_as(C.c_void_p), meta.ctypes.data_as(C.c_void_p), self._stream()))
        mine = torch.from_numpy(np.concatenate([handles, meta.view(np.uint8)])).to(s.device)
        allp = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(allp, mine)
        packs = [p.cpu().numpy() for p in allp]
        hkey = b"".join(p[:320].tobytes() for p in packs)      # the IPC handles: re-opened only when a buffer moved
        mkey = b"".join(p[320:].tobytes() for p in packs)      # tile geometry: cheap to refresh (D8 <-> D-infinity)
        def conn(which, r, handles=True):
            if r is None:
                check(self.l.td_sweep_peer_connect_dev(self.T.ctx, which, None, None))
            else:
                h = np.ascontiguousarray(packs[r][:320]); m = np.ascontiguousarray(packs[r][320:]).view(np.int32)
                check(self.l.td_sweep_peer_connect_dev(self.T.ctx, which, h.ctypes.data_as(C.c_void_p) if handles else None, m.ctypes.data_as(C.c_void_p)))
        up = self.rank - 1 if self.rank > 0 else None
        down = self.rank + 1 if self.rank < self.world - 1 else None
        if self._peer_cache is None or hkey != self._peer_cache[0]:
            conn(0, up); conn(1, down); conn(2, None if self.rank == 0 else 0)
        elif mkey != self._peer_cache[1]:
            if up is not None: conn(0, up, handles=False)
            if down is not None: conn(1, down, handles=False)
        self._peer_cache = (hkey, mkey)

    def _sweep_peer(self, run, out):
        """One kernel per rank: tiles deliver into the neighbour GPUs over NVLink themselves."""
        import os, sys, time
        dbg = os.environ.get("TD_DEBUG") == "1"
        s = self.s
        check(self.l.td_sweep_peer_begin_dev(self.T.ctx, s.c, self._stream()))
        torch.cuda.synchronize(); dist.barrier()          # every rank has announced its tiles in the global counter
        if dbg:
            print(f"[peer r{self.rank} {time.time():.3f}] begin done, launching", file=sys.stderr, flush=True)
        halo = torch.zeros(2 * s.pitch, dtype=torch.int32, device=s.device)
        run(halo)
        torch.cuda.synchronize()
        if dbg:
            print(f"[peer r{self.rank} {time.time():.3f}] kernel finished", file=sys.stderr, flush=True)
        dist.barrier()
        self.l.td_sweep_peer_off_dev(self.T.ctx)
        self.rounds = 1
        return out

    def _sweep(self, run, out):
        if self.peer:
            return self._sweep_peer(run, out)
        s = self.s
        check(self.l.td_sweep_begin_dev(self.T.ctx, s.c, self._stream()))
        halo = torch.zeros(2 * s.pitch, dtype=torch.int32, device=s.device)
        self.rounds = 0
        while True:
            halo.zero_()
            run(halo)
            self.rounds += 1
            if self.world == 1:
                break
            dec_top, dec_bot = exchange_counts(halo, s.pitch, self.rank, self.world)
            self.share(out)                                   # the area rows the decrements announce
            if all_reduce_scalar(int(halo.sum()), device=s.device) == 0:   # ringTerm: did anybody hand over work?
                break
            check(self.l.td_sweep_apply_halo_dev(self.T.ctx, s.c, None if dec_top is None else C.c_void_p(dec_top.data_ptr()),
                                                 None if dec_bot is None else C.c_void_p(dec_bot.data_ptr()), self._stream()))
        return out

    def _restrict(self, outlets):
        """-o over row strips: the upstream flood of the outlets in rounds (requests for the neighbours' edge rows travel like
        the sweeps' halo counts; the outlet branch of initNeighborD8up / initNeighborDinfup, src/commonLib.cpp:300-375)."""
        import numpy as np
        s = self.s
        cols = np.ascontiguousarray(outlets[0], np.int32)
        rows = np.ascontiguousarray(np.asarray(outlets[1], np.int64) - self.row0, np.int32)    # row 0 = my first owned row
        req = torch.zeros(2 * s.pitch, dtype=torch.int32, device=s.device)
        ptr = lambda t: None if t is None else C.c_void_p(t.data_ptr())
        in_top = in_bot = None
        nout = len(cols)
        while True:
            check(self.l.td_sweep_restrict_round_dev(self.T.ctx, s.c, cols.ctypes.data_as(C.c_void_p), rows.ctypes.data_as(C.c_void_p), nout,
                                                     ptr(in_top), ptr(in_bot), ptr(req), 0, self._stream()))
            nout = -1
            if self.world == 1:
                break
            in_top, in_bot = exchange_counts(req, s.pitch, self.rank, self.world)
            if all_reduce_scalar(int(req.sum()), device=s.device) == 0:
                break
        check(self.l.td_sweep_restrict_round_dev(self.T.ctx, s.c, None, None, -1, None, None, ptr(req), 1, self._stream()))

    def aread8(self, p, ad8=None, w=None, nodata=-32768, w_nodata=-9999.0, contcheck=True, shared=False, outlets=None):
        s = self.s
        ad8 = s.empty(torch.float32) if ad8 is None else ad8
        if not shared:
            self.share(p)
        if self.peer:
            self._peer_setup(False)
        self.T.aread8_deps(s, p, ad8, nodata)
        if outlets is not None:
            self._restrict(outlets)
        wp = None if w is None else C.c_void_p(w.data_ptr())
        return self._sweep(lambda halo: check(self.l.td_aread8_sweep_run_dev(self.T.ctx, wp, C.c_void_p(ad8.data_ptr()), s.c, w_nodata, int(w is not None),
                                                                              int(contcheck), C.c_void_p(halo.data_ptr()), self._stream())), ad8)

    def areadinf(self, ang, dxc, dyc, sca=None, w=None, nodata=-3.4028234663852886e38, contcheck=True, shared=False, outlets=None):
        s = self.s
        sca = s.empty(torch.float32) if sca is None else sca
        if not shared:
            self.share(ang)
        if self.peer:
            self._peer_setup(True)
        if self.world > 1:
            # the neighbour strips' edge rows keep their own cell sizes (geographic rasters: getdxdyc(jn), src/areadinf.cpp:199-201)
            cs = torch.zeros((self.ny + 2, 2), dtype=torch.float64, device=dxc.device)
            cs[1:self.ny + 1, 0] = dxc; cs[1:self.ny + 1, 1] = dyc
            exchange_rows(cs, self.ny, self.rank, self.world)
            top, bot = cs[0].tolist(), cs[self.ny + 1].tolist()
            self.l.td_set_halo_cell_sizes_dev(self.T.ctx, top[0], top[1], bot[0], bot[1])
        self.T.areadinf_deps(s, ang, sca, dxc, dyc, nodata)
        if outlets is not None:
            self._restrict(outlets)
        wp = None if w is None else C.c_void_p(w.data_ptr())
        return self._sweep(lambda halo: check(self.l.td_area_sweep_run_dev(self.T.ctx, C.c_void_p(ang.data_ptr()), wp, C.c_void_p(sca.data_ptr()), s.c,
                                                                            int(w is not None), int(contcheck), C.c_void_p(dxc.data_ptr()),
                                                                            C.c_void_p(halo.data_ptr()), self._stream())), sca)

    def d8_slopes(self, fel, dxc, dyc, nodata=-3.0e38):
        self.share(fel)
        return self.T.d8_slopes(self.s, fel, dxc, dyc, nodata)

    def dinf_slopes(self, fel, dxc, dyc, nodata=-3.0e38):
        self.share(fel)
        return self.T.dinf_slopes(self.s, fel, dxc, dyc, nodata)

    # ---- flow directions incl. flats.  The Garbrecht-Martz BFS runs on the row strips themselves (_flats_strips: one exchange
    # per level like the reference's share() + MPI_Allreduce per pass; checked against the oracle on the CPU emulation and
    # bit-identical to the single-strip run over NCCL on 2 GPUs, scripts/dist_check.py).  TAUDEM_B200_FLATS=replicated selects
    # the first-generation fallback: all-gather fel + directions, every rank resolves the whole grid and keeps its rows.
    def gather_full(self, t):
        """All-gather of the owned rows of a strip tensor -> full-grid strip tensor (halo rows unused)."""
        from .device import DeviceStrip
        parts = partition(self.s.total_ny, self.world)
        maxny = max(n for _, n in parts)
        mine = torch.zeros((maxny, self.s.pitch), dtype=t.dtype, device=t.device)
        mine[:self.ny].copy_(t[1:self.ny + 1])
        if self.world == 1:
            bufs = [mine]
        elif _staged():
            hm = mine.cpu().view(torch.uint8)                      # bytes: neither gloo nor NCCL moves int16
            hb = [torch.empty_like(hm) for _ in range(self.world)]
            dist.all_gather(hb, hm)
            bufs = [b.view(mine.dtype).to(t.device) for b in hb]
        else:
            flat = torch.empty((self.world,) + tuple(mine.shape), dtype=mine.dtype, device=t.device)
            dist.all_gather_into_tensor(flat.view(torch.uint8), mine.view(torch.uint8))
            bufs = [flat[r] for r in range(self.world)]
        sf = DeviceStrip(self.s.nx, self.s.total_ny, device=t.device)
        full = sf.empty(t.dtype)
        for (row0, n), b in zip(parts, bufs):
            full[1 + row0:1 + row0 + n].copy_(b[:n])
        return sf, full

    def _flats_strips(self, fel, d, dxc, dyc, dinf):
        """Flat resolution on the row strips themselves (TAUDEM_B200_FLATS=strips): every rank runs the BFS loop of
        flats.cu on its own flat cells; the td_strip_comm callbacks below do what resolveflats does with
        linearpart::share() and MPI_Allreduce — one collect + share + all-reduce per BFS level."""
        import ctypes as C
        s, dev = self.s, self.s.device
        rank, world, ny = self.rank, self.world, self.ny

        class _Dev:                                   # a device pointer as a byte tensor (CUDA array interface)
            def __init__(self, ptr, nbytes):
                self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}

        def rows(ptr, elem, nrows):
            return torch.as_tensor(_Dev(ptr, nrows * s.pitch * elem), device=dev).view(nrows, s.pitch * elem)

        def share(user, ptr, elem):
            try:
                exchange_rows(rows(ptr, elem, ny + 2), ny, rank, world)
                return 0
            except Exception as e:                    # an exception must not unwind through the C frames
                print("td_strip_comm.share:", e, flush=True)
                return 1

        def collect(user, ptr, elem, recv_top, recv_bot):
            try:
                t = rows(ptr, elem, ny + 2)
                sends, recvs = [], []
                if rank > 0:
                    sends.append((t[0], rank - 1)); recvs.append((rows(recv_top, elem, 1)[0], rank - 1))
                if rank < world - 1:
                    sends.append((t[ny + 1], rank + 1)); recvs.append((rows(recv_bot, elem, 1)[0], rank + 1))
                _p2p(sends, recvs)
                return 0
            except Exception as e:
                print("td_strip_comm.collect:", e, flush=True)
                return 1

        def allreduce(user, v, n):
            try:
                t = torch.tensor([int(v[i]) for i in range(n)], dtype=torch.int64, device="cpu" if _staged() else dev)
                dist.all_reduce(t)
                for i, x in enumerate(t.tolist()):
                    v[i] = x
                return 0
            except Exception as e:
                print("td_strip_comm.allreduce_sum:", e, flush=True)
                return 1

        SHARE = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int)
        COLLECT = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p)
        ALLRED = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_ulonglong), C.c_int)

        class Comm(C.Structure):
            _fields_ = [("user", C.c_void_p), ("share", SHARE), ("collect", COLLECT), ("allreduce_sum", ALLRED)]

        cbs = (SHARE(share), COLLECT(collect), ALLRED(allreduce))          # keep the thunks alive during the call
        comm = Comm(None, *cbs)
        felc = fel.clone()                                                  # the reference works on a copy of elevDEM too
        self.share(d)                                                       # halo rows of the directions the stencil produced
        left = C.c_longlong(0)
        fn = self.l.td_dinf_flats_strip_dev if dinf else self.l.td_d8_flats_strip_dev
        torch.cuda.synchronize()
        check(fn(self.T.ctx, C.c_void_p(felc.data_ptr()), C.c_void_p(d.data_ptr()), s.c, C.c_void_p(dxc.data_ptr()), C.c_void_p(dyc.data_ptr()),
                 C.byref(left), C.byref(comm) if world > 1 else None, self._stream()))
        return int(left.value)

    def _flats(self, fel, d, dxc, dyc, nflat, dinf):
        import os
        total = all_reduce_scalar(int(nflat), device=self.s.device) if self.world > 1 else int(nflat)
        if total == 0:
            return 0
        if os.environ.get("TAUDEM_B200_FLATS") != "replicated":
            return self._flats_strips(fel, d, dxc, dyc, dinf)
        sf, fel_full = self.gather_full(fel)
        _, d_full = self.gather_full(d)
        # per-row cell sizes of the whole grid
        rows = torch.zeros(2, max(n for _, n in partition(self.s.total_ny, self.world)), dtype=torch.float64, device=self.s.device)
        rows[0, :self.ny] = dxc; rows[1, :self.ny] = dyc
        if self.world > 1:
            allr = [torch.empty_like(rows) for _ in range(self.world)] if not _staged() else None
            if _staged():
                hb = [torch.empty(rows.shape, dtype=rows.dtype) for _ in range(self.world)]
                dist.all_gather(hb, rows.cpu()); allr = [b.to(self.s.device) for b in hb]
            else:
                dist.all_gather(allr, rows)
        else:
            allr = [rows]
        parts = partition(self.s.total_ny, self.world)
        dxf = torch.cat([a[0, :n] for a, (_, n) in zip(allr, parts)]).contiguous()
        dyf = torch.cat([a[1, :n] for a, (_, n) in zip(allr, parts)]).contiguous()
        left = (self.T.dinf_flats if dinf else self.T.d8_flats)(sf, fel_full, d_full, dxf, dyf)
        d[1:self.ny + 1].copy_(d_full[1 + self.row0:1 + self.row0 + self.ny])
        return left

    def d8flowdir(self, fel, dxc, dyc, nodata=-3.0e38):
        p, sd8, nflat = self.d8_slopes(fel, dxc, dyc, nodata)
        self._flats(fel, p, dxc, dyc, nflat, dinf=False)
        return p, sd8

    def dinfflowdir(self, fel, dxc, dyc, nodata=-3.0e38):
        ang, slp, nflat = self.dinf_slopes(fel, dxc, dyc, nodata)
        self._flats(fel, ang, dxc, dyc, nflat, dinf=True)
        return ang, slp

    def pitremove(self, dem, nodata=-9999.0, four_way=False):
        """flood(): local relaxation to convergence, halo exchange, repeat until no strip changes
        (src/flood.cpp:344-479 share()/ringTerm() structure)."""
        s = self.s
        self.share(dem)
        w = self.T.flood_init(s, dem, nodata, four_way)
        first = True
        while True:
            self.share(w)                                     # fresh halo rows from the neighbours
            # after the first pass only the tiles next to the halo rows can have something new to do
            moved = int(self.T.flood_relax(s, dem, w, four_way, edges_only=not first))
            first = False
            if self.world > 1:
                moved = all_reduce_scalar(moved, device=s.device)   # ringTerm: did any strip move?
            if moved == 0:
                break
        return w


def bench_main(args, rank, world, local):
    """N > 1 arm of bench.py: strong scaling of the same DEM over row strips."""
    import json
    import os
    import sys
    import time

    import bench as B
    import taudem_b200 as td
    dev = torch.device("cuda", local)
    n = B.pick_size(torch, args.size)
    cells = n * n
    log = lambda *a: rank == 0 and print("[bench %.1fs]" % (time.time() - B.T0), *a, file=sys.stderr, flush=True)
    # inputs (untimed): the distributed pipeline itself — generated DEM strip -> pitremove -> d8flowdir /
    # dinfflowdir over the row strips (flat resolution replicated after an all-gather, see DistTools._flats)
    t_setup = time.time()
    D = DistTools(n, n, rank, world, device=dev)
    s = D.s
    dxc, dyc = s.rows(30.0), s.rows(30.0)
    pipe_ms = {}

    def timed(name, fn):
        """one tool of the pipeline on all ranks: CUDA events per rank, max over ranks (BASELINE.json configs[3])"""
        torch.cuda.synchronize(); dist.barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(); r = fn(); a1.record(); torch.cuda.synchronize()
        t = torch.tensor([a0.elapsed_time(a1)], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        pipe_ms[name] = round(float(t.item()), 3)
        return r

    dem = D.T.gen_dem(s, seed=B.SEED, hurst=B.HURST, tilt=B.TILT)
    fel = timed("pitremove", lambda: D.pitremove(dem))
    del dem
    p, sd8 = timed("d8flowdir", lambda: D.d8flowdir(fel, dxc, dyc))
    del sd8
    ang, slp = timed("dinfflowdir", lambda: D.dinfflowdir(fel, dxc, dyc))
    del slp, fel
    torch.cuda.synchronize(); torch.cuda.empty_cache()
    info = {"setup_s": round(time.time() - t_setup, 2)}
    ad8, sca = s.empty(torch.float32), s.empty(torch.float32)
    log("inputs ready", info, pipe_ms)

    def step():
        D.aread8(p, ad8)
        r1 = D.rounds
        D.areadinf(ang, dxc, dyc, sca)
        return r1, D.rounds

    for _ in range(args.warmup):
        rounds = step()
    td.reset_launch_count()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with B.ClockSampler(local) as clk:
        e0.record()
        for _ in range(args.steps):
            rounds = step()
        e1.record()
        torch.cuda.synchronize(); dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    launches = torch.tensor([td.launch_count()], device=dev, dtype=torch.int64)
    dist.all_reduce(launches)
    # BASELINE.json configs[4]: areadinf with a weight grid (-wg), same strips
    w = D.T.gen_weights(s)
    scaw = s.empty(torch.float32)
    D.areadinf(ang, dxc, dyc, scaw, w=w)
    timed("areadinf_wg", lambda: D.areadinf(ang, dxc, dyc, scaw, w=w))
    timed("aread8", lambda: D.aread8(p, ad8))
    timed("areadinf", lambda: D.areadinf(ang, dxc, dyc, sca))
    del w, scaw
    mx = torch.stack([s.owned(ad8).max(), s.owned(sca).max()]).double()
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    # 64-bit hashes of the raw bits of the whole rasters: the strips' position-weighted sums add up (mod 2^64)
    to_i64 = lambda h: h - (1 << 64) if h >= (1 << 63) else h
    hs = torch.tensor([to_i64(B.raster_hash(torch, s.owned(ad8), D.row0, n)), to_i64(B.raster_hash(torch, s.owned(sca), D.row0, n))], device=dev, dtype=torch.int64)
    dist.all_reduce(hs)
    hashes = ["%016x" % (int(v) & ((1 << 64) - 1)) for v in hs.tolist()]
    ms_per_step = float(ms.item()) / args.steps
    value = cells / 1e6 / (ms_per_step * 1e-3)

    # end to end: pinned host strips in, pinned host strips out, copies inside the timed region
    hp = torch.empty((s.ny, n), dtype=torch.int16, pin_memory=True); hp.copy_(s.owned(p))
    ha = torch.empty((s.ny, n), dtype=torch.float32, pin_memory=True); ha.copy_(s.owned(ang))
    o1 = torch.empty((s.ny, n), dtype=torch.float32, pin_memory=True); o2 = torch.empty((s.ny, n), dtype=torch.float32, pin_memory=True)
    e2e_steps = max(1, min(args.steps, args.e2e_steps))

    def e2e_step():
        s.owned(p).copy_(hp, non_blocking=True); s.owned(ang).copy_(ha, non_blocking=True)
        step()
        o1.copy_(s.owned(ad8), non_blocking=True); o2.copy_(s.owned(sca), non_blocking=True)
        torch.cuda.synchronize()

    e2e_step()
    dist.barrier(); t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    dist.barrier()
    te = torch.tensor([(time.perf_counter() - t0) / e2e_steps], device=dev, dtype=torch.float64)
    dist.all_reduce(te, op=dist.ReduceOp.MAX)
    if rank == 0:
        peak, peak_src = B.peaks()
        line = {"metric": B.METRIC, "value": round(value, 2), "unit": "Mcells/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": round(ms_per_step, 3), "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"aread8 + areadinf on {n}x{n} float32 synthetic fractal DEM (hills: H={B.HURST}, tilt={B.TILT}, seed={B.SEED}, 30 m cells), contamination check on, no weights",
                           "cells": cells, "partition": f"{world} row strips (linearpart: total//size rows, remainder on the last rank); " + ("peer mode: the sweep kernels deliver across GPUs over NVLink (CUDA IPC, system-scope atomics), no exchange rounds" if D.peer else "NCCL send/recv halo + decrement exchange rounds, all_reduce termination"),
                           "exchange_rounds": {"aread8": rounds[0], "areadinf": rounds[1]}, "l2": "inputs exceed the 126 MB L2; no explicit flush",
                           "timed": "CUDA events per rank, max over ranks", "max_ad8": float(mx[0]), "max_sca": float(mx[1]),
                           "hash_ad8": hashes[0], "hash_sca": hashes[1], **info},
                "clocks": clk.summary(),
                "e2e": {"value": round(cells / 1e6 / float(te.item()), 2), "unit": "Mcells/s", "h2d_bytes_per_step": cells * 6, "d2h_bytes_per_step": cells * 8,
                        "steps": e2e_steps, "ms_per_step": round(float(te.item()) * 1e3, 2), "api": "per-rank pinned host strips -> device strips -> DistTools.aread8/areadinf -> pinned host strips"},
                "gpu_launches": int(launches.item()),
                "roofline": {"bound": "hbm", "kernel": "k_sweep_warp<dinf>", "achieved": round(8 * cells / (pipe_ms["areadinf"] * 1e-3) / 1e9 / world, 2), "peak": peak, "unit": "GB/s",
                             "frac": round(8 * cells / (pipe_ms["areadinf"] * 1e-3) / 1e9 / world / peak, 5), "traffic": None, "peak_source": peak_src,
                             "note": "per GPU: the strip's algorithmic bytes over the whole areadinf time (dependency stencil + sweep, max over ranks)",
                             "per_tool_ms": pipe_ms,
                             "per_tool_Mcells_per_s": {k: round(cells / 1e6 / (v * 1e-3), 1) for k, v in pipe_ms.items()}},
                "cpu_baseline": None}
        print(json.dumps(line))
    dist.barrier()
    dist.destroy_process_group()
