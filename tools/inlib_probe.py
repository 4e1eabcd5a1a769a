"""Per-call times of the in-library multi-device call (opts.devices = all visible GPUs) on C2: python tools/inlib_probe.py [n_calls]"""
import os, sys, time
sys.path.insert(0, "tools"); sys.path.insert(0, ".")
import workloads as WL
import bamsignals_b200 as B
from bench import data_dir
n_calls = int(sys.argv[1]) if len(sys.argv) > 1 else 10
bam, info = WL.make_bam("c2", 1.0, data_dir())
gr, kw, fn = WL.regions("c2", 1.0)
nd = B.lib().bsg_device_count()
one = B.default_opts(devices=[0], inflate_threads=os.cpu_count(), gpu_inflate=1)
for _ in range(3):
    r = getattr(B, fn)(bam, gr, opts=one, **kw); del r
opts = B.default_opts(devices=list(range(nd)), inflate_threads=os.cpu_count(), gpu_inflate=1)
for i in range(n_calls):
    t0 = time.perf_counter(); r = getattr(B, fn)(bam, gr, opts=opts, **kw); ms = (time.perf_counter() - t0) * 1e3; del r
    t = B.timings()
    print(f"call {i}: {ms:.1f} ms  plan {t['ms_plan']:.1f} fetch {t['ms_fetch']:.1f} h2d {t['ms_h2d']:.1f} inflate {t['ms_inflate_gpu']:.1f} d2h {t['ms_d2h']:.1f} total {t['ms_total']:.1f}")
