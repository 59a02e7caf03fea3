"""Timeline of td_contributing_areas_host (TAUDEM_B200_TRACE=1): when each copy and each tool starts / ends on its stream."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import taudem_b200 as td
import bench
from taudem_b200.device import Tools
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
T = Tools()
s, dxc, dyc, p, ang, info, pipe = bench.build_inputs(T, n, torch)
hp = torch.empty((n, n), dtype=torch.int16, pin_memory=True); hp.copy_(s.owned(p))
ha = torch.empty((n, n), dtype=torch.float32, pin_memory=True); ha.copy_(s.owned(ang))
del p, ang; T.close(); torch.cuda.synchronize(); torch.cuda.empty_cache()
o1 = torch.empty((n, n), dtype=torch.float32, pin_memory=True); o2 = torch.empty((n, n), dtype=torch.float32, pin_memory=True)
for k in range(2):
    t0 = time.perf_counter()
    td.contributing_areas_grid(hp.numpy(), ha.numpy(), dx=30.0, dy=30.0, out_ad8=o1.numpy(), out_sca=o2.numpy())
    print("call", k, "wall %.1f ms" % ((time.perf_counter() - t0) * 1e3), flush=True)
