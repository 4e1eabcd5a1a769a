import sys, os
sys.path[:0]=['/root/repo','/root/repo/tests','/root/repo/tools']
import numpy as np, bamsignals_b200 as B, workloads as WL, oracle_api as O
bam, info = WL.make_bam("c2", 0.002, "/tmp/gen", unplaced=7)
gr, kw, fn = WL.regions("c2", 0.002)
got = getattr(B, fn)(bam, gr, opts=B.default_opts(gpu_inflate=1), **kw)
want = getattr(O, fn)(bam, gr, nthreads=8, **kw)
print("equal:", np.array_equal(WL.as_flat(got), WL.as_flat(want)), B.timings())
