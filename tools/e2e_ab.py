#!/usr/bin/env python
"""A/B timing of the end-to-end call on one box: python tools/e2e_ab.py [--preset c2] [--gscale 1] [--reps 5] NAME=ENV=VALUE ...
Every variant is "label:VAR=val,VAR=val" (environment switches the library reads per call, e.g. BSG_BATCH_MB);
variants are interleaved rep by rep so that box-to-box and warm-up effects cancel."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tools"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import bamsignals_b200 as B  # noqa: E402
import workloads as WL  # noqa: E402
from bench import data_dir  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--preset", default="c2")
    ap.add_argument("--gscale", type=float, default=1.0)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("variants", nargs="*", default=["default:"])
    a = ap.parse_args()
    bam, info = WL.make_bam(a.preset, a.gscale, data_dir())
    gr, kw, fn = WL.regions(a.preset, a.gscale)
    call = getattr(B, fn)
    opts = B.default_opts(devices=[0], inflate_threads=os.cpu_count() or 1, gpu_inflate=1)
    variants = []
    for v in a.variants:
        label, _, envs = v.partition(":")
        variants.append((label, dict(e.split("=", 1) for e in envs.split(",") if e)))
    keys = sorted({k for _, e in variants for k in e})
    res = {label: [] for label, _ in variants}
    first_result = None
    for rep in range(a.reps + 2):
        for label, env in variants:
            for k in keys:
                os.environ.pop(k, None)
            os.environ.update(env)
            t0 = time.perf_counter()
            got = call(bam, gr, opts=opts, **kw)
            wall = (time.perf_counter() - t0) * 1e3
            if rep == 0:                     # every variant must return what the first one returned, bit for bit
                import numpy as np
                flat = WL.as_flat(got)
                if first_result is None:
                    first_result = flat.copy()
                elif not np.array_equal(flat, first_result):
                    raise SystemExit(f"variant {label!r} changes the result")
            del got
            t = B.timings()
            if rep >= 2:
                res[label].append(dict(wall=wall, total=t["ms_total"], plan=t["ms_plan"], fetch=t["ms_fetch"],
                                       inflate=t["ms_inflate_gpu"], h2d=t["ms_h2d"], kernels=t["ms_kernels"], batches=t["n_batches"]))
    out = {}
    for label, rows in res.items():
        out[label] = {k: round(sorted(r[k] for r in rows)[len(rows) // 2], 2) for k in rows[0]} if rows else {}
    print(json.dumps({"preset": a.preset, "gscale": a.gscale, "reads": info["records"], "median_ms": out}))


if __name__ == "__main__":
    main()
