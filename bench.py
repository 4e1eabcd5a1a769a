#!/usr/bin/env python
"""bench.py — reads counted/s on BASELINE.json's config C2 (bamProfile binsize=1 ss=TRUE shift=75 over 100k 2 kb
windows on a synthetic 100M-read single-end BAM), B200 path vs the reference CPU path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--preset c2|c3|c4|c5] [--gscale g]

One JSON line on stdout (rank 0).  A "step" is one pass of the counting hot path over the whole workload:
  value  kernel-only: raw record bytes + record offsets already resident in HBM (bsg_stage), the step runs
         K1 decode+filter -> K3 join -> K4/K5 count on the device; timed with CUDA events on the library's
         compute stream (first event to last event of every step, summed), max over ranks.
  e2e    the same metric through the reference-facing C-ABI call (bsg_pileup / bsg_coverage) from the BAM *file*
         (page cache) to the result in host memory: BAI query, inflate, record walk, H2D, kernels, D2H all inside.
  roofline      dominant kernel: algorithmic bytes / its CUDA-event time vs MEASURED_PEAKS.json's hbm_gbs.
  cpu_baseline  the oracle (restated reference CPU path, single thread like the reference) on a bounded sample.
--impl reference times the restated reference CPU path (oracle/, zlib inflate) on all host threads.
For N > 1 (torchrun) the regions are sharded across ranks in genomic order (no collective on the data path; NCCL is
used only for the barrier and the max/sum of the timings) => "scaling": "strong" (one BAM, fixed total work).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tools"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import workloads as WL  # noqa: E402

FULL_READS = {"c2": 100e6, "c3": 2 * 24895642, "c4": 1e9, "c5": 500e6}
# algorithmic bytes per unit (SURVEY.md 8d / DESIGN.md): K1 (decode + filter fused: the 20-byte table row of SURVEY 8d
# never leaves registers) 4 + 36 + 4*n_cigar in, 16 out (tid, pos, c0, c1); n_cigar ~ 1.18 on the synthetic mix.
# K3 24 per tile, K4/K5 8*C + 4*B.
K1_BYTES_PER_READ = 4 + 36 + 4 * 1.18 + 16
K2_BYTES_PER_READ = 0


def data_dir():
    d = os.environ.get("BSG_BENCH_DIR") or os.path.join("/dev/shm" if os.path.isdir("/dev/shm") else "/tmp", "bsg_bench")
    os.makedirs(d, exist_ok=True)
    return d


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.rows = index, threading.Event(), []

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            return
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])
            if self.stop_flag.is_set():
                break
        self.proc.terminate()

    def summary(self):
        self.stop_flag.set()
        time.sleep(0.15)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


def run_ours(args, rank, world, local_rank):
    import bamsignals_b200 as B
    import torch
    if B.lib().bsg_device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device (the product path has no CPU fallback)")
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    preset, gs = args.preset, args.gscale
    d = data_dir()
    t_gen = time.time()
    if rank == 0:
        bam, info = WL.make_bam(preset, gs, d)
    barrier()
    if rank != 0:
        bam, info = WL.make_bam(preset, gs, d)
    t_gen = time.time() - t_gen
    gr_all, kw, fn = WL.regions(preset, gs)
    # bamCount regions own one counter each and cannot be cut; profile / coverage regions are cut into bin-aligned pieces
    # so that a few huge regions (C4) still balance over the ranks by reads
    gr, _ = WL.shard_regions(gr_all, rank, world, split_align=None if fn == "bamCount" else int(kw.get("binsize", 1)))
    ca = B.core_args(fn, **kw)
    gpu_inflate = 1 if args.gpu_inflate else -1
    opts = B.default_opts(devices=[local_rank], inflate_threads=max(1, (os.cpu_count() or 1) // world), gpu_inflate=gpu_inflate)
    is_cov = fn == "bamCoverage"
    ext = (ca["tlen_filter"][1] if (is_cov and ca["tspan"]) else 0) if is_cov else \
        abs(ca["shift"]) + (ca["tlen_filter"][1] if ca["pe_mid"] else 0)

    # ---- kernel-only: inputs resident in HBM -----------------------------------------------------------------
    st = B.Stage(bam, gr, ext_hint=ext, opts=opts)
    t_stage = B.timings()

    def step_resident():
        if is_cov:
            st.coverage(ca["tlen_filter"], ca["mapqual"], ca["requiredF"], ca["filteredF"], ca["tspan"], want_output=False)
        else:
            st.pileup(ca["tlen_filter"], ca["mapqual"], ca["binsize"], ca["shift"], ca["ss"], ca["requiredF"],
                      ca["filteredF"], ca["pe_mid"], want_output=False)
        return B.timings()

    sampler = ClockSampler(local_rank)
    sampler.start()
    t_wait = time.time()
    while not sampler.rows and time.time() - t_wait < 3.0:      # nvidia-smi needs a moment to produce its first row
        time.sleep(0.02)
    for _ in range(args.warmup):
        step_resident()
    barrier()
    dev_ms, launches, ksum = 0.0, 0, {k: 0.0 for k in ("ms_decode", "ms_filter", "ms_join", "ms_count")}
    wall0 = time.perf_counter()
    for _ in range(args.steps):
        t = step_resident()
        dev_ms += t["ms_device"]
        launches += t["n_launches"]
        for k in ksum:
            ksum[k] += t[k]
    barrier()
    wall_ms = (time.perf_counter() - wall0) * 1e3
    reads = t["records"]
    st.close()
    if args.profile:                     # ncu runs: only the resident steps matter
        sampler.summary()                # stops the nvidia-smi child
        if rank == 0:
            print(json.dumps({"profile_only": True, "ms_device_per_step": dev_ms / args.steps, "kernels_ms": ksum,
                              "launches": launches, "reads_decoded": reads}), flush=True)
        return

    # ---- end to end through the C ABI from the BAM file -----------------------------------------------------------
    call = getattr(B, fn)
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    for _ in range(min(args.warmup, 2)):
        call(bam, gr, opts=opts, **kw)
    barrier()
    e0 = time.perf_counter()
    for _ in range(e2e_steps):
        res = call(bam, gr, opts=opts, **kw)
        te = B.timings()
    barrier()
    e2e_ms = (time.perf_counter() - e0) * 1e3 / e2e_steps
    # clocks + throttle reasons sampled over BOTH timed regions (the resident steps alone last a few milliseconds, less
    # than one 100 ms nvidia-smi sample)
    clocks = sampler.summary()
    if gpu_inflate > 0:     # compressed bytes + block descriptors (16 B + 4 B CRC per <= 64 KiB block) + tiles
        h2d = te["bytes_compressed"] + 20 * (te["bytes_inflated"] // 65280 + 1) + 24 * te["n_tiles"]
    else:                   # inflated bytes + record offsets + tiles
        h2d = te["bytes_inflated"] + 4 * (te["records"] + te["n_batches"]) + 24 * te["n_tiles"]
    d2h = 4 * te["out_elems"]
    del res

    # ---- reduce over ranks ----------------------------------------------------------------------------------------
    vals = torch.tensor([dev_ms, wall_ms, e2e_ms] + [ksum[k] for k in ksum], dtype=torch.float64, device="cuda")
    sums = torch.tensor([float(reads), float(te["records"]), float(launches), float(h2d), float(d2h),
                         float(t["candidates"]), float(t["out_elems"]), float(t["n_tiles"])], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    vals, sums = vals.cpu().tolist(), sums.cpu().tolist()
    dev_ms, wall_ms, e2e_ms = vals[0], vals[1], vals[2]
    kms = dict(zip(ksum, vals[3:]))
    reads_all, reads_e2e, launches_all, h2d_all, d2h_all, cand, out_elems, n_tiles = sums
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    ms_per_step = dev_ms / args.steps
    # Unit of work for every arm (ours kernel-only, ours e2e, reference): the job = this BAM + these regions, counted
    # as the records of the BAM.  reads_all (records inside the fetched index ranges, what the kernels touch) is
    # reported beside it and is what the per-kernel roofline figures use.
    units = float(info["records"])
    value = units / (ms_per_step * 1e-3)

    # ---- roofline of the dominant kernel ---------------------------------------------------------------------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    else:
        peak, peak_src = 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"
    # per-rank work for per-launch figures: every rank runs the same kernels on 1/world of the data
    per = 1.0 / world
    alg = {"decode": K1_BYTES_PER_READ * reads_all * per, "filter": K2_BYTES_PER_READ * reads_all * per,
           "join": 24.0 * n_tiles * per, "count": (8.0 * cand + 4.0 * out_elems) * per}
    kern = {}
    for name, key in (("decode", "ms_decode"), ("filter", "ms_filter"), ("join", "ms_join"), ("count", "ms_count")):
        ms = kms[key] / args.steps
        kern[name] = {"ms": round(ms, 4), "alg_gb": round(alg[name] / 1e9, 4),
                      "gbs": round(alg[name] / 1e9 / (ms * 1e-3), 1) if ms > 0 else None}
    dom = max(kern, key=lambda k: kern[k]["ms"])
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        t_ent = json.load(open(tpath)).get(preset, {}).get(dom)
        if t_ent and "dram_bytes_per_read" in t_ent:       # ncu capture at a smaller scale, per record of the launch
            traffic = t_ent["dram_bytes_per_read"] * reads_all * per
    roofline = {"bound": "hbm", "kernel": dom, "achieved": kern[dom]["gbs"], "peak": peak, "unit": "GB/s",
                "frac": round(kern[dom]["gbs"] / peak, 4) if kern[dom]["gbs"] else None, "traffic": traffic,
                "peak_source": peak_src, "kernels": kern,
                "step_share": {k: round(kern[k]["ms"] / max(1e-9, sum(v["ms"] for v in kern.values())), 3) for k in kern}}

    # ---- CPU baseline: restated reference path, single thread, bounded sample ---------------------------------------
    cpu = cpu_baseline(bam, gr_all, fn, kw, preset, threads=1, budget_s=args.cpu_seconds, total_reads=units)

    line = {
        "metric": "reads counted/sec", "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": f"{preset}: {fn} {kw} on {int(FULL_READS[preset] * gs):,}-read synthetic BAM (gscale {gs:g}), "
                               f"{len(gr_all):,} regions", "preset": preset, "gscale": gs, "reads_in_bam": info["records"],
                   "reads_per_step": int(units), "reads_decoded_per_step": int(reads_all),
                   "regions": len(gr_all), "record_shape": "compact",
                   "l2": f"inputs {te['bytes_inflated'] * world / 1e9:.2f} GB raw + {4 * out_elems / 1e9:.2f} GB out per step >> 126 MB L2 (no flush needed)",
                   "parallelism": f"regions sharded over {world} GPU(s), no collective", "host_threads": os.cpu_count()},
        "e2e": {"value": units / (e2e_ms * 1e-3), "unit": "reads/s", "reads_decoded_per_step": int(reads_e2e), "ms_per_step": e2e_ms, "steps": e2e_steps,
                "h2d_bytes_per_step": int(h2d_all), "d2h_bytes_per_step": int(d2h_all),
                "breakdown_ms_rank0": {k: round(te[k], 2) for k in ("ms_plan", "ms_fetch", "ms_h2d", "ms_inflate_gpu", "ms_d2h", "ms_kernels", "ms_total")},
                "inflate": "gpu" if gpu_inflate > 0 else "host zlib",
                "note": "BAM file in page cache -> result in host memory; inflate (GPU kernel or host zlib pool) is inside"},
        "gpu_launches": int(launches_all),
        "wall_ms_per_step": wall_ms / args.steps,
        "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
        "setup": {"bam_generate_s": round(t_gen, 1), "stage_ms": round(t_stage["ms_total"], 1)},
    }
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def cpu_baseline(bam, gr_all, fn, kw, preset, threads, budget_s, total_reads):
    """The oracle in indexed mode (the reference's access pattern) on a genomic prefix of the regions sized for about
    `budget_s` seconds; reads/s = records streamed / wall time."""
    import oracle_api as O
    order = np.lexsort((gr_all.start, gr_all.seq_idx))
    n = len(order)
    # calibrate on a small prefix, then size the sample
    k0 = max(1, min(n, n // 200 if n > 400 else n))
    t0 = time.perf_counter()
    getattr(O, fn)(bam, gr_all[np.sort(order[:k0])], nthreads=threads, **kw)
    dt = max(1e-4, time.perf_counter() - t0)
    k = int(min(n, max(k0, k0 * budget_s / dt)))
    t0 = time.perf_counter()
    getattr(O, fn)(bam, gr_all[np.sort(order[:k])], nthreads=threads, **kw)
    dt = time.perf_counter() - t0
    s = O.stats()
    # job units: the sample is k/n of the regions in genomic order, i.e. about k/n of the BAM
    return {"value": total_reads * (k / n) / dt, "unit": "reads/s", "cores": threads, "kind": "port",
            "sample": f"first {k:,} of {n:,} regions in genomic order = {k / n:.3f} of the job ({s['records']:,} records "
                      f"streamed, {dt:.1f} s), restated reference CPU path (R/Rhtslib not installable offline), zlib "
                      f"inflate, indexed access, single thread like the reference",
            "seconds": dt}


def run_reference(args, rank, world):
    """The reference arm: the restated reference CPU path on all host threads (region shards, one reader each)."""
    if rank != 0:
        return
    import oracle_api as O
    preset, gs = args.preset, args.gscale
    bam, info = WL.make_bam(preset, gs, data_dir())
    gr, kw, fn = WL.regions(preset, gs)
    threads = os.cpu_count() or 1
    call = getattr(O, fn)
    for _ in range(min(args.warmup, 1)):
        call(bam, gr, nthreads=threads, **kw)
    t0 = time.perf_counter()
    recs = 0
    for _ in range(args.steps):
        call(bam, gr, nthreads=threads, **kw)
        recs = O.stats()["records"]
    dt = (time.perf_counter() - t0) / args.steps
    v = info["records"] / dt          # same unit of work as the GPU arm: the job's BAM records per second
    line = {"impl": "reference", "metric": "reads counted/sec", "value": v, "unit": "reads/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": f"{preset}: {fn} {kw} on {int(FULL_READS[preset] * gs):,}-read synthetic BAM (gscale {gs:g}), "
                                   f"{len(gr):,} regions", "preset": preset, "gscale": gs, "reads_in_bam": info["records"],
                       "reads_per_step": int(info["records"]), "reads_streamed_per_step": int(recs), "regions": len(gr)},
            "cpu_baseline": {"value": v, "unit": "reads/s", "cores": threads, "kind": "port",
                             "sample": "whole workload per step; restated reference CPU path (oracle/, zlib, indexed access) "
                                       f"on {threads} threads over disjoint region shards; R/Rhtslib not installable offline"},
            "e2e": {"value": v, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--preset", default="c2", choices=["c2", "c3", "c4", "c5"])
    ap.add_argument("--gscale", type=float, default=1.0, help="genome (and read-count) scale; 1.0 = the full configuration")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--gpu-inflate", type=int, default=1, help="1: inflate BGZF on the GPU (default), 0: host zlib pool")
    ap.add_argument("--profile", action="store_true", help="resident steps only (for runs under ncu)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
