"""GeoTIFF I/O throughput of the native reader / writer (taudem_b200/csrc/tiff_io.cpp): write + read of a float32 DEM,
uncompressed / LZW / Deflate.  CPU only.   python scripts/io_bench.py [n=8192]"""
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import taudem_b200 as td
from taudem_b200 import synth


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    dem = synth.gen_dem(1024, 1024, hurst=0.8, tilt=1.0)
    dem = np.ascontiguousarray(np.tile(dem, (n // 1024, n // 1024)) + np.linspace(0, 50, n, dtype=np.float32)[None, :])   # compressible but not trivial
    mb = dem.nbytes / 1e6
    d = tempfile.mkdtemp(dir=os.environ.get("TMPDIR", "/tmp"))
    print(f"{n}x{n} float32 = {mb:.0f} MB, {os.cpu_count()} host threads visible")
    for name, comp in (("none", 1), ("LZW", 5), ("Deflate", 8)):
        f = os.path.join(d, f"t_{name}.tif")
        t0 = time.perf_counter(); td.write_raster(f, dem, -9999.0, compression=comp); tw = time.perf_counter() - t0
        size = os.path.getsize(f) / 1e6
        t0 = time.perf_counter(); back = td.read_raster(f); tr = time.perf_counter() - t0
        assert np.array_equal(back, dem)
        print(f"{name:8s} file {size:8.0f} MB  write {tw:6.2f} s = {mb / tw:7.0f} MB/s   read {tr:6.2f} s = {mb / tr:7.0f} MB/s")
        os.remove(f)
    os.rmdir(d)


if __name__ == "__main__":
    main()
