import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from taudem_b200.device import DeviceStrip, Tools
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
T = Tools(); s = DeviceStrip(n, n); dxc, dyc = s.rows(30.0), s.rows(30.0)
fel = T.pitremove(s, T.gen_dem(s, hurst=0.8, tilt=1.0))
for _ in range(3):
    p, sd8, nf = T.d8_slopes(s, fel, dxc, dyc)
    ang, slp, nf = T.dinf_slopes(s, fel, dxc, dyc)
torch.cuda.synchronize()
