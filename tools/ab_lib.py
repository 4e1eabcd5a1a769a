"""Same-box A/B of two builds of the library on one preset: python tools/ab_lib.py PRESET GSCALE (BSG_LIB selects the build)."""
import os, sys, time
sys.path.insert(0, "tools"); sys.path.insert(0, ".")
import workloads as WL
import bamsignals_b200 as B
from bench import data_dir
preset, gs = sys.argv[1], float(sys.argv[2])
bam, info = WL.make_bam(preset, gs, data_dir())
gr, kw, fn = WL.regions(preset, gs)
opts = B.default_opts(devices=[0], inflate_threads=os.cpu_count(), gpu_inflate=1)
ts = []
for rep in range(6):
    t0 = time.perf_counter(); r = getattr(B, fn)(bam, gr, opts=opts, **kw); ts.append((time.perf_counter() - t0) * 1e3); del r
t = B.timings()
print(os.environ.get("BSG_LIB", "current").split("/")[-1], preset, gs, "median ms", round(sorted(ts[2:])[2], 1), "min", round(min(ts[2:]), 1), "inflate", round(t["ms_inflate_gpu"], 1), "fetch", round(t["ms_fetch"], 1))
