"""TEST INFRASTRUCTURE ONLY — runs the reference's own tools (oracle/_ref/*, the
reference sources compiled unchanged against the MPI/GDAL shims, see oracle/Makefile)
on numpy arrays by going through TIFF files.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
import os
import re
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")
TOOLS = ("pitremove", "d8flowdir", "dinfflowdir", "aread8", "areadinf")


def available():
    return all(os.access(os.path.join(REF, t), os.X_OK) for t in TOOLS)


def run_tool(tool, args, np_ranks=1, timeout=None):
    """Runs one reference tool; returns (stdout, {label: seconds} of its timing block)."""
    env = dict(os.environ)
    env["MINIMPI_NP"] = str(np_ranks)
    r = subprocess.run([os.path.join(REF, tool)] + list(args), env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                       timeout=timeout, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"reference {tool} failed ({r.returncode}):\n{r.stdout}\n{r.stderr}")
    times = {m.group(1).strip(): float(m.group(2)) for m in re.finditer(r"^([A-Za-z ]+time): ([0-9.eE+-]+)$", r.stdout, re.M)}
    return r.stdout, times


class RefPipeline:
    """Convenience wrapper: arrays in, arrays out, through a scratch directory."""

    def __init__(self, workdir=None, dx=30.0, dy=30.0, np_ranks=1):
        import taudem_b200 as td   # only the raster file reader/writer is used here
        self.td = td
        self._tmp = None
        if workdir is None:
            self._tmp = tempfile.TemporaryDirectory(prefix="tdref_")
            workdir = self._tmp.name
        self.dir = workdir
        self.dx, self.dy, self.np_ranks = dx, dy, np_ranks
        self.times = {}

    def path(self, name):
        return os.path.join(self.dir, name)

    def put(self, name, arr, nodata):
        self.td.write_raster(self.path(name), arr, nodata, dx=self.dx, dy=self.dy)
        return self.path(name)

    def get(self, name, dtype):
        return self.td.read_raster(self.path(name), dtype)

    def pitremove(self, dem, nodata=-9999.0, four_way=False, depmask=None):
        self.put("dem.tif", dem, nodata)
        args = ["-z", self.path("dem.tif"), "-fel", self.path("fel.tif")]
        if four_way:
            args.append("-4way")
        if depmask is not None:
            self.put("mask.tif", depmask.astype(np.int16), -32768)
            args += ["-depmask", self.path("mask.tif")]
        _, self.times["pitremove"] = run_tool("pitremove", args, self.np_ranks)
        return self.get("fel.tif", np.float32)

    def d8flowdir(self, fel, nodata=-3.0e38):
        self.put("felin.tif", fel, nodata)
        _, self.times["d8flowdir"] = run_tool("d8flowdir", ["-fel", self.path("felin.tif"), "-p", self.path("p.tif"), "-sd8", self.path("sd8.tif")], self.np_ranks)
        return self.get("p.tif", np.int16), self.get("sd8.tif", np.float32)

    def dinfflowdir(self, fel, nodata=-3.0e38):
        self.put("felin.tif", fel, nodata)
        _, self.times["dinfflowdir"] = run_tool("dinfflowdir", ["-fel", self.path("felin.tif"), "-ang", self.path("ang.tif"), "-slp", self.path("slp.tif")], self.np_ranks)
        return self.get("ang.tif", np.float32), self.get("slp.tif", np.float32)

    def aread8(self, p, nodata=-32768, weights=None, w_nodata=-9999.0, contcheck=True, outlets=None):
        """outlets: path of a point shapefile (-o)."""
        self.put("pin.tif", p.astype(np.int16), nodata)
        args = ["-p", self.path("pin.tif"), "-ad8", self.path("ad8.tif")]
        if outlets is not None:
            args += ["-o", outlets]
        if weights is not None:
            self.put("w.tif", weights, w_nodata)
            args += ["-wg", self.path("w.tif")]
        if not contcheck:
            args.append("-nc")
        _, self.times["aread8"] = run_tool("aread8", args, self.np_ranks)
        return self.get("ad8.tif", np.float32)

    def areadinf(self, ang, nodata=-3.4028234663852886e38, weights=None, w_nodata=-9999.0, contcheck=True, outlets=None):
        self.put("angin.tif", ang, nodata)
        args = ["-ang", self.path("angin.tif"), "-sca", self.path("sca.tif")]
        if outlets is not None:
            args += ["-o", outlets]
        if weights is not None:
            self.put("w.tif", weights, w_nodata)
            args += ["-wg", self.path("w.tif")]
        if not contcheck:
            args.append("-nc")
        _, self.times["areadinf"] = run_tool("areadinf", args, self.np_ranks)
        return self.get("sca.tif", np.float32)

    def d8flowpathextremeup(self, p, sa, usemax=True, contcheck=True, outlets=None, nodata=-32768, sa_nodata=-9999.0):
        self.put("pin.tif", p.astype(np.int16), nodata); self.put("sa.tif", sa, sa_nodata)
        args = ["-p", self.path("pin.tif"), "-sa", self.path("sa.tif"), "-ssa", self.path("ssa_up.tif")]
        if outlets is not None:
            args += ["-o", outlets]
        if not usemax:
            args.append("-min")
        if not contcheck:
            args.append("-nc")
        _, self.times["d8flowpathextremeup"] = run_tool("d8flowpathextremeup", args, self.np_ranks)
        return self.get("ssa_up.tif", np.float32)

    def gridnet(self, p, mask=None, thresh=0, outlets=None, nodata=-32768):
        self.put("pin.tif", p.astype(np.int16), nodata)
        args = ["-p", self.path("pin.tif"), "-plen", self.path("plen.tif"), "-tlen", self.path("tlen.tif"), "-gord", self.path("gord.tif")]
        if outlets is not None:
            args += ["-o", outlets]
        if mask is not None:
            self.put("mask.tif", mask.astype(np.int32), -1)
            args += ["-mask", self.path("mask.tif"), "-thresh", str(int(thresh))]
        _, self.times["gridnet"] = run_tool("gridnet", args, self.np_ranks)
        return self.get("plen.tif", np.float32), self.get("tlen.tif", np.float32), self.get("gord.tif", np.int16)

    def dinfdecayaccum(self, ang, dm, weights=None, contcheck=True, outlets=None, nodata=-3.4028234663852886e38, dm_nodata=-9999.0, w_nodata=-9999.0):
        self.put("angin.tif", ang, nodata); self.put("dm.tif", dm, dm_nodata)
        args = ["-ang", self.path("angin.tif"), "-dm", self.path("dm.tif"), "-dsca", self.path("dsca.tif")]
        if outlets is not None:
            args += ["-o", outlets]
        if weights is not None:
            self.put("w.tif", weights, w_nodata)
            args += ["-wg", self.path("w.tif")]
        if not contcheck:
            args.append("-nc")
        _, self.times["dinfdecayaccum"] = run_tool("dinfdecayaccum", args, self.np_ranks)
        return self.get("dsca.tif", np.float32)

    def dinfconclimaccum(self, ang, dm, q, dg, csol=1.0, contcheck=True, outlets=None, nodata=-9999.0):
        self.put("angin.tif", ang, -3.4028234663852886e38); self.put("dm.tif", dm, nodata); self.put("q.tif", q, nodata)
        self.put("dg.tif", np.asarray(dg, np.int16), -32768)
        args = ["-ang", self.path("angin.tif"), "-dm", self.path("dm.tif"), "-q", self.path("q.tif"), "-dg", self.path("dg.tif"),
                "-ctpt", self.path("ctpt.tif"), "-csol", repr(float(csol))]
        if outlets is not None:
            args += ["-o", outlets]
        if not contcheck:
            args.append("-nc")
        _, self.times["dinfconclimaccum"] = run_tool("dinfconclimaccum", args, self.np_ranks)
        return self.get("ctpt.tif", np.float32)

    def dinftranslimaccum(self, ang, tsup, tc, cs=None, contcheck=True, outlets=None, nodata=-9999.0):
        self.put("angin.tif", ang, -3.4028234663852886e38); self.put("tsup.tif", tsup, nodata); self.put("tc.tif", tc, nodata)
        args = ["-ang", self.path("angin.tif"), "-tsup", self.path("tsup.tif"), "-tc", self.path("tc.tif"), "-tla", self.path("tla.tif"),
                "-tdep", self.path("tdep.tif")]
        if cs is not None:
            self.put("cs.tif", cs, nodata)
            args += ["-cs", self.path("cs.tif"), "-ctpt", self.path("ctptout.tif")]
        if outlets is not None:
            args += ["-o", outlets]
        if not contcheck:
            args.append("-nc")
        _, self.times["dinftranslimaccum"] = run_tool("dinftranslimaccum", args, self.np_ranks)
        return (self.get("tla.tif", np.float32), self.get("tdep.tif", np.float32), self.get("ctptout.tif", np.float32) if cs is not None else None)

    def threshold(self, ssa, thresh, mask=None, nodata=-1.0):
        self.put("ssa.tif", ssa, nodata)
        args = ["-ssa", self.path("ssa.tif"), "-src", self.path("src.tif"), "-thresh", repr(float(thresh))]
        if mask is not None:
            self.put("mask.tif", mask, -9999.0)
            args += ["-mask", self.path("mask.tif")]
        _, self.times["threshold"] = run_tool("threshold", args, self.np_ranks)
        return self.get("src.tif", np.int16)

    def slopearea(self, slp, sca, m=None, n=None, nodata=-1.0):
        self.put("slpin.tif", slp, nodata); self.put("scain.tif", sca, nodata)
        args = ["-slp", self.path("slpin.tif"), "-sca", self.path("scain.tif"), "-sa", self.path("sa.tif")]
        if m is not None:
            args += ["-par", repr(float(m)), repr(float(n))]
        _, self.times["slopearea"] = run_tool("slopearea", args, self.np_ranks)
        return self.get("sa.tif", np.float32)

    def slopearearatio(self, slp, sca, nodata=-1.0):
        self.put("slpin.tif", slp, nodata); self.put("scain.tif", sca, nodata)
        _, self.times["slopearearatio"] = run_tool("slopearearatio", ["-slp", self.path("slpin.tif"), "-sca", self.path("scain.tif"), "-sar", self.path("sar.tif")], self.np_ranks)
        return self.get("sar.tif", np.float32)

    def twi(self, slp, sca, nodata=-1.0):
        self.put("slpin.tif", slp, nodata); self.put("scain.tif", sca, nodata)
        _, self.times["twi"] = run_tool("twi", ["-slp", self.path("slpin.tif"), "-sca", self.path("scain.tif"), "-twi", self.path("twi.tif")], self.np_ranks)
        return self.get("twi.tif", np.float32)
