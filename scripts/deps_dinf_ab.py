"""k_deps_dinf with TAUDEM_B200_DEPS_EDGE as set in the environment, on the angles of the bench DEM (flats resolved): CUDA-event timings.
   python scripts/deps_dinf_ab.py [n=65536] [reps=5]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from taudem_b200.device import DeviceStrip, Tools  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
T = Tools(); s = DeviceStrip(n, n); dxc, dyc = s.rows(30.0), s.rows(30.0)
dem = T.gen_dem(s, hurst=0.8, tilt=1.0)
fel = T.pitremove(s, dem)
del dem
ang, slp, nflat = T.dinf_slopes(s, fel, dxc, dyc)
del slp
torch.cuda.empty_cache()
T.dinf_flats(s, fel, ang, dxc, dyc)
del fel
torch.cuda.empty_cache()
sca = s.empty(torch.float32)
ts = []
for _ in range(reps + 1):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record(); T.areadinf_deps(s, ang, sca, dxc, dyc); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
ts = sorted(ts[1:])
print(f"n={n} TAUDEM_B200_DEPS_EDGE={os.environ.get('TAUDEM_B200_DEPS_EDGE', '(default)')}: areadinf_deps best {ts[0]:.3f} ms median {ts[len(ts) // 2]:.3f} ms")
