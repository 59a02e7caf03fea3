"""Host logic of the row-strip partition, world_size 2 and 3 over gloo on CPU."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from taudem_b200.dist import exchange_counts, exchange_rows, partition


def test_partition_matches_linearpart():
    # ny = total // size, remainder on the LAST rank (src/linearpart.h:132-134)
    assert partition(10, 3) == [(0, 3), (3, 3), (6, 4)]
    assert partition(512, 1) == [(0, 512)]
    assert partition(7, 7) == [(i, 1) for i in range(7)]
    for total, world in ((65536, 8), (1001, 4), (5, 2)):
        parts = partition(total, world)
        assert parts[0][0] == 0 and sum(n for _, n in parts) == total
        assert all(parts[i][0] + parts[i][1] == parts[i + 1][0] for i in range(world - 1))


def _worker(rank, world, port, total_ny, nx):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    row0, ny = partition(total_ny, world)[rank]
    full = torch.arange(total_ny * nx, dtype=torch.float32).reshape(total_ny, nx)
    t = torch.full((ny + 2, nx), -7.0)
    t[1:ny + 1] = full[row0:row0 + ny]
    exchange_rows(t, ny, rank, world)
    if rank > 0:
        assert torch.equal(t[0], full[row0 - 1])
    else:
        assert (t[0] == -7).all()
    if rank < world - 1:
        assert torch.equal(t[ny + 1], full[row0 + ny])
    else:
        assert (t[ny + 1] == -7).all()
    # decrement hand-over: what I recorded for the strip below arrives as that strip's dec_top
    halo = torch.zeros(2 * nx, dtype=torch.int32)
    halo[:nx] = 100 * rank + 1            # for the strip above
    halo[nx:] = 100 * rank + 2            # for the strip below
    top, bot = exchange_counts(halo, nx, rank, world)
    assert (top is None) == (rank == 0) and (bot is None) == (rank == world - 1)
    if top is not None:
        assert (top == 100 * (rank - 1) + 2).all()
    if bot is not None:
        assert (bot == 100 * (rank + 1) + 1).all()
    # the cell sizes of the neighbour strips' edge rows (DistTools.areadinf on geographic rasters): a (ny + 2) x 2 float64 strip
    dxf = 30.0 + 0.001 * torch.arange(total_ny, dtype=torch.float64); dyf = 25.0 - 0.002 * torch.arange(total_ny, dtype=torch.float64)
    cs = torch.zeros((ny + 2, 2), dtype=torch.float64)
    cs[1:ny + 1, 0] = dxf[row0:row0 + ny]; cs[1:ny + 1, 1] = dyf[row0:row0 + ny]
    exchange_rows(cs, ny, rank, world)
    assert cs[0].tolist() == ([float(dxf[row0 - 1]), float(dyf[row0 - 1])] if rank > 0 else [0.0, 0.0])            # 0 = none: the grid ends here
    assert cs[ny + 1].tolist() == ([float(dxf[row0 + ny]), float(dyf[row0 + ny])] if rank < world - 1 else [0.0, 0.0])
    total = torch.tensor([int(halo.sum())]); dist.all_reduce(total)
    assert int(total) == sum(nx * (200 * r + 3) for r in range(world))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_halo_exchange_gloo(world):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    mp.spawn(_worker, args=(world, port, 11, 8), nprocs=world, join=True)
