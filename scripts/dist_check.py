"""Multi-GPU parity check (launch with torchrun): the row-strip partitioned tools must reproduce the
single-strip rasters bit for bit (rank-count invariance)."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from taudem_b200.device import DeviceStrip, Tools
from taudem_b200.dist import DistTools


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    backend = os.environ.get("TD_BACKEND", "nccl")
    if backend == "gloo":
        local = 0                                   # every rank on cuda:0, host-staged transport
    torch.cuda.set_device(local)
    if backend == "nccl":
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        dist.init_process_group("gloo")
    ny, nx = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (3001, 2500)
    # single-strip truth on every rank
    T = Tools()
    sf = DeviceStrip(nx, ny)
    dxc, dyc = sf.rows(30.0), sf.rows(20.0)
    dem = T.gen_dem(sf, hurst=0.8, tilt=1.0)
    sf.owned(dem)[ny // 3: ny // 3 + 40, nx // 2: nx // 2 + 70] = -9999.0      # a nodata hole
    w = T.gen_weights(sf)
    fel = T.pitremove(sf, dem)
    p, sd8, nf = T.d8_slopes(sf, fel, dxc, dyc)
    felc = fel.clone(); T.d8_flats(sf, felc, p, dxc, dyc)
    ang, slp, nf2 = T.dinf_slopes(sf, fel, dxc, dyc)
    felc.copy_(fel); T.dinf_flats(sf, felc, ang, dxc, dyc)
    ad8 = T.aread8(sf, p); ad8w = T.aread8(sf, p, w=w, contcheck=False)
    sca = T.areadinf(sf, ang, dxc, dyc); scaw = T.areadinf(sf, ang, dxc, dyc, w=w)
    torch.cuda.synchronize()

    D = DistTools(nx, ny, rank, world)
    s = D.s
    rows = slice(1 + D.row0, 1 + D.row0 + s.ny)

    def strip_of(full, dtype):
        t = s.empty(dtype); t[1:s.ny + 1].copy_(full[rows]); return t

    def same(name, mine, full):
        ok = torch.equal(mine[1:s.ny + 1, :nx].view(torch.int32 if mine.dtype == torch.float32 else mine.dtype),
                         full[rows, :nx].view(torch.int32 if full.dtype == torch.float32 else full.dtype))
        if not ok:
            a_ = mine[1:s.ny + 1, :nx]; b_ = full[rows, :nx]
            bad = (a_ != b_)
            idx = bad.nonzero()
            print(f"   rank {rank}: {int(bad.sum())} cells differ; rows {int(idx[:,0].min())}..{int(idx[:,0].max())} of {s.ny}; "
                  f"first {[(int(y), int(x), float(a_[y, x]), float(b_[y, x])) for y, x in idx[:4].tolist()]}", flush=True)
        from taudem_b200.dist import all_reduce_scalar
        flag = all_reduce_scalar(int(ok), op=dist.ReduceOp.MIN, device=s.device)
        if rank == 0:
            print(f"{name}: {'identical' if flag else 'DIFFERENT'} (rounds {D.rounds})", flush=True)
        return bool(flag)

    ldx, ldy = s.rows(30.0), s.rows(20.0)
    ok = True
    ok &= same("fel", D.pitremove(strip_of(dem, torch.float32)), fel)
    p_d, sd8_d = D.d8flowdir(strip_of(fel, torch.float32), ldx, ldy)
    ok &= same("sd8", sd8_d, sd8)
    ok &= same("p (flats resolved)", p_d, p)
    a_d, slp_d = D.dinfflowdir(strip_of(fel, torch.float32), ldx, ldy)
    ok &= same("slp", slp_d, slp)
    ok &= same("ang (flats resolved)", a_d, ang)
    ok &= same("ad8", D.aread8(strip_of(p, torch.int16)), ad8)
    ok &= same("ad8 -wg -nc", D.aread8(strip_of(p, torch.int16), w=strip_of(w, torch.float32), contcheck=False), ad8w)
    ok &= same("sca", D.areadinf(strip_of(ang, torch.float32), ldx, ldy), sca)
    ok &= same("sca -wg", D.areadinf(strip_of(ang, torch.float32), ldx, ldy, w=strip_of(w, torch.float32)), scaw)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
