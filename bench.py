#!/usr/bin/env python
"""bench.py — reads counted/s on BASELINE.json's configs, B200 path vs the reference CPU path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--preset c2|c3|c4|c5] [--gscale g]
                  [--also c1,c4|none]

One JSON line on stdout (rank 0).  The headline workload is C2 (BASELINE.json configs[1]: bamProfile binsize=1 ss=TRUE
shift=75 over 100 k 2 kb windows on a synthetic 100 M-read BAM); `--also` (default c1,c4 at full scale) adds, under
"configs", C4 (configs[3]: the config BASELINE.json quotes "at 1/2/4/8 GPUs") and the C1 call latency.
A "step" is one pass of the counting hot path over the whole workload:
  value   kernel-only, SURVEY 8d: records DECODED per second with raw record bytes + record offsets already resident in
          HBM (bsg_stage); a step runs K1 decode+filter -> K3 join -> K4/K5 count; CUDA events on the library's compute
          stream, max over ranks.  `value_job_units` is the same time in the unit the two arms share (below).
  e2e     through the reference-facing C-ABI call (bsg_pileup / bsg_coverage) from the BAM *file* (page cache) to the
          result in host memory: index query, H2D of the compressed bytes, inflate, record walk, kernels, D2H inside the
          timed region.  Unit shared with the reference arm: the job's BAM records per second (the reference's index
          queries touch fewer records than the GPU path decodes, and neither touches all of them: the job is the
          common unit).  The result of the last timed call is compared bit for bit with the CPU checker
          (oracle/_ref = the reference's own src/bamsignals.cpp, else the restated port) on a prefix of every rank's
          regions before anything is printed: "parity".  A mismatch exits 1.
  roofline      dominant kernel: algorithmic bytes / its CUDA-event time vs MEASURED_PEAKS.json's hbm_gbs; per kernel also
                frac_dram = ncu-measured DRAM bytes (profiles/traffic.json, per unit of work) / time / peak.
  cpu_baseline  the checker in the reference's access pattern, single thread like the reference, on a bounded sample.
--impl reference times the reference CPU path (oracle/_ref when built, else the port) on all host threads.
For N > 1 (torchrun) the regions are sharded across ranks in genomic order (no collective on the data path; NCCL only for
the barrier and the max/sum of the timings) => "scaling": "strong" (one BAM, fixed total work); rank 0 then also runs
ONE in-library multi-device call (opts.devices = all N GPUs) over the whole region set, timed and checked: "inlib".
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
# algorithmic bytes per unit (SURVEY.md 8d / DESIGN.md 4): K1 (decode + filter fused: the 20-byte table row of SURVEY 8d
# never leaves registers) 4 + 36 + 4*n_cigar in, 16 out (tid, pos, c0, c1); n_cigar ~ 1.18 on the synthetic mix.
# K3 24 per tile, K4/K5 8*C + 4*B.
K1_BYTES_PER_READ = 4 + 36 + 4 * 1.18 + 16


def data_dir():
    d = os.environ.get("BSG_BENCH_DIR") or os.path.join("/dev/shm" if os.path.isdir("/dev/shm") else "/tmp", "bsg_bench")
    os.makedirs(d, exist_ok=True)
    return d


def checker():
    """The CPU checker: the reference's own engine when oracle/_ref was built, else the restated port."""
    import oracle_api as O
    return ("ref" if O.ref_available() else "port"), O


def config_dict(preset, gs, info, n_regions, fn, kw):
    """Identical in both arms (the driver compares the two dicts)."""
    return {"workload": f"{preset}: {fn} {kw} on {int(FULL_READS[preset] * gs):,}-read synthetic BAM (gscale {gs:g}), "
                        f"{n_regions:,} regions", "preset": preset, "gscale": gs, "reads_in_bam": info["records"],
            "regions": n_regions, "record_shape": "compact", "unit_of_work": "the job's BAM records per step",
            "l2": "inputs and outputs per step >> 126 MB L2 (no flush needed)"}


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


def region_result(res, i):
    """Region i's slice of a bamCount array / CountSignals, flattened in bsg_output_layout() order."""
    if isinstance(res, np.ndarray):
        return (res[:, i] if res.ndim == 2 else res[i:i + 1]).ravel()
    return np.asarray(res[int(i)]).ravel(order="F")


def check_prefix(res, gr, fn, kw, bam, budget_s, threads, want=None, full=False):
    """Compare the first regions (genomic order) of `res` - the result of a timed call over `gr` - bit for bit with the CPU
    checker.  `want` = (k, checker result over the first k regions) when the caller already has one (cpu_baseline);
    full: every region of the job."""
    kind, O = checker()
    n = len(gr)
    if n == 0:
        return {"regions_checked": 0, "equal": True, "checker": kind}
    order = np.lexsort((gr.start, gr.seq_idx))
    if full:
        t0 = time.perf_counter()
        w = getattr(O, fn)(bam, gr, nthreads=threads, impl=kind, **kw)
        dt = time.perf_counter() - t0
        a, b = WL.as_flat(res), WL.as_flat(w)
        ok = a.shape == b.shape and bool(np.array_equal(a, b))
        return {"regions_checked": int(n), "ints_checked": int(a.size), "equal": ok, "checker": kind, "checker_threads": threads,
                "checker_seconds": round(dt, 2)}
    if want is None:
        k0 = max(1, min(n, n // 200 if n > 400 else max(1, n // 8)))
        t0 = time.perf_counter()
        w = getattr(O, fn)(bam, gr[np.sort(order[:k0])], nthreads=threads, impl=kind, **kw)
        dt = max(1e-4, time.perf_counter() - t0)
        k = int(min(n, max(k0, k0 * budget_s / dt)))
        if k > k0:
            w = getattr(O, fn)(bam, gr[np.sort(order[:k])], nthreads=threads, impl=kind, **kw)
    else:
        k, w = want
    idx = np.sort(order[:k])
    ok, ints = True, 0
    for j, i in enumerate(idx):
        a, b = region_result(res, i), region_result(w, j)
        ints += a.size
        if a.shape != b.shape or not np.array_equal(a, b):
            ok = False
            break
    return {"regions_checked": int(k), "ints_checked": int(ints), "equal": bool(ok), "checker": kind}


def cpu_baseline(bam, gr_all, fn, kw, threads, budget_s, total_reads):
    """The checker in indexed mode (the reference's access pattern) on a genomic prefix of the regions sized for about
    `budget_s` seconds; reads/s = job units (BAM records x share of the job) / wall time.  Also returns (k, result) so
    that the same numbers serve as the parity check of the timed GPU result."""
    kind, O = checker()
    order = np.lexsort((gr_all.start, gr_all.seq_idx))
    n = len(order)
    k0 = max(1, min(n, n // 200 if n > 400 else max(1, n // 8)))
    t0 = time.perf_counter()
    getattr(O, fn)(bam, gr_all[np.sort(order[:k0])], nthreads=threads, impl=kind, **kw)
    dt = max(1e-4, time.perf_counter() - t0)
    k = int(min(n, max(k0, k0 * budget_s / dt)))
    t0 = time.perf_counter()
    w = getattr(O, fn)(bam, gr_all[np.sort(order[:k])], nthreads=threads, impl=kind, **kw)
    dt = time.perf_counter() - t0
    s = O.stats()
    share = float(gr_all.width[order[:k]].astype(np.int64).sum()) / max(1.0, float(gr_all.width.astype(np.int64).sum()))
    what = ("the reference's own src/bamsignals.cpp compiled unchanged (oracle/_ref) over the repo's htslib stand-in (zlib)"
            if kind == "ref" else "restated reference CPU path (oracle/ port, zlib)")
    return {"value": total_reads * share / dt, "unit": "reads/s", "cores": threads, "kind": "reference" if kind == "ref" else "port",
            "sample": f"first {k:,} of {n:,} regions in genomic order = {share:.3f} of the job ({s['records']:,} records "
                      f"streamed, {dt:.1f} s); {what}, indexed access, {threads} thread(s)",
            "seconds": dt}, (k, w)


def measure_ours(args, preset, gs, rank, world, local_rank, dist, barrier, sampler, steps, warmup, e2e_steps):
    """Kernel-only + end-to-end + parity for one preset on this rank; returns the rank's partial figures."""
    import bamsignals_b200 as B
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
    host_threads = max(1, (os.cpu_count() or 1) // world)
    opts = B.default_opts(devices=[local_rank], inflate_threads=host_threads, gpu_inflate=1 if args.gpu_inflate else -1)
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

    for _ in range(warmup):
        step_resident()
    barrier()
    dev_ms, launches, ksum = 0.0, 0, {k: 0.0 for k in ("ms_decode", "ms_filter", "ms_join", "ms_count")}
    wall0 = time.perf_counter()
    for _ in range(steps):
        t = step_resident()
        dev_ms += t["ms_device"]
        launches += t["n_launches"]
        for k in ksum:
            ksum[k] += t[k]
    barrier()
    wall_ms = (time.perf_counter() - wall0) * 1e3
    st.close()
    out = dict(preset=preset, gs=gs, bam=bam, info=info, gr_all=gr_all, gr=gr, kw=kw, fn=fn, t_gen=t_gen, t_stage=t_stage,
               dev_ms=dev_ms, wall_ms=wall_ms, launches=launches, ksum=ksum, t_res=t, steps=steps, host_threads=host_threads)
    if args.profile:
        return out

    # ---- end to end through the C ABI from the BAM file -----------------------------------------------------------
    call = getattr(B, fn)
    for _ in range(min(warmup, 2)):
        call(bam, gr, opts=opts, **kw)
    barrier()
    per_call, e_launch = [], 0
    res = None
    e0 = time.perf_counter()
    for _ in range(e2e_steps):
        del res
        c0 = time.perf_counter()
        res = call(bam, gr, opts=opts, **kw)
        per_call.append((time.perf_counter() - c0) * 1e3)
        te = B.timings()
        e_launch += te["n_launches"]
    barrier()
    e2e_ms = (time.perf_counter() - e0) * 1e3 / e2e_steps
    if args.gpu_inflate:    # compressed bytes + block descriptors (16 B + 4 B CRC per <= 64 KiB block) + tiles
        h2d = te["bytes_compressed"] + 20 * (te["bytes_inflated"] // 65280 + 1) + 24 * te["n_tiles"]
    else:                   # inflated bytes + record offsets + tiles
        h2d = te["bytes_inflated"] + 4 * (te["records"] + te["n_batches"]) + 24 * te["n_tiles"]
    out.update(e2e_ms=e2e_ms, per_call=per_call, te=te, h2d=h2d, d2h=te["bytes_d2h"], e_launch=e_launch)

    # ---- parity of what was timed: the last timed call's result vs the CPU checker ----------------------------------
    units = float(info["records"])
    if rank == 0 and world == 1:
        cpu, want = cpu_baseline(bam, gr_all, fn, kw, threads=1, budget_s=args.cpu_seconds, total_reads=units)
        out["cpu"] = cpu
        out["parity"] = check_prefix(res, gr, fn, kw, bam, 0, 1, want=want)
        del want
        # the WHOLE timed result against the checker on all host threads, when that fits the budget (C2: ~1 s, C4: ~15 s)
        est = units / max(1.0, cpu["value"] * max(1.0, 0.6 * host_threads))
        if out["parity"]["equal"] and args.full_parity_seconds > 0 and est <= args.full_parity_seconds:
            out["parity"]["full"] = check_prefix(res, gr, fn, kw, bam, 0, host_threads, full=True)
            out["parity"]["equal"] = out["parity"]["equal"] and out["parity"]["full"]["equal"]
    else:
        out["parity"] = check_prefix(res, gr, fn, kw, bam, args.parity_seconds, host_threads)
        if rank == 0:
            out["cpu"], _ = cpu_baseline(bam, gr_all, fn, kw, threads=1, budget_s=args.cpu_seconds, total_reads=units)
    del res
    return out


def inlib_call(args, m, world, steps):
    """Rank 0, N > 1: ONE C-ABI call with opts.devices = all N GPUs over the whole region set (the library shards
    the regions itself, one pipeline per device, every device writes its slice of the caller's buffer), timed + checked."""
    import bamsignals_b200 as B
    opts = B.default_opts(devices=list(range(world)), inflate_threads=os.cpu_count() or 1, gpu_inflate=1 if args.gpu_inflate else -1)
    call = getattr(B, m["fn"])
    ms = []
    res = None
    warm = 5        # rank 0 has only seen its own shard so far: the other devices' buffers are allocated by the first of these
    for i in range(warm + steps):       # calls, and the rest of the BAM's mapping is page-locked in the background during the next few
        del res
        t0 = time.perf_counter()
        res = call(m["bam"], m["gr_all"], opts=opts, **m["kw"])
        if i >= warm:
            ms.append((time.perf_counter() - t0) * 1e3)
    t = B.timings()
    par = check_prefix(res, m["gr_all"], m["fn"], m["kw"], m["bam"], args.parity_seconds, os.cpu_count() or 1)
    del res
    return {"n_devices": int(t["n_devices"]), "ms_per_call_median": float(np.median(ms)), "ms_per_call_min": float(min(ms)),
            "calls": len(ms), "ms_per_call_all": [round(x, 1) for x in ms], "reads_per_s_job_units": m["info"]["records"] / (float(np.median(ms)) * 1e-3), "parity": par,
            "note": "one bsg_pileup/bsg_coverage call, opts.devices = all GPUs; regions sharded inside the library"}


def reduce_and_format(args, m, rank, world, dist, torch):
    """All-reduce one preset's figures over the ranks; rank 0 returns the dict of the JSON line (others None)."""
    ks = list(m["ksum"])
    t, te = m["t_res"], m.get("te")
    have_e2e = te is not None
    vals = [m["dev_ms"], m["wall_ms"], m.get("e2e_ms", 0.0), float(np.median(m["per_call"])) if have_e2e else 0.0,
            float(min(m["per_call"])) if have_e2e else 0.0] + [m["ksum"][k] for k in ks]
    sums = [float(t["records"]), float(te["records"]) if have_e2e else 0.0, float(m["launches"]), float(m.get("h2d", 0)),
            float(m.get("d2h", 0)), float(t["candidates"]), float(t["out_elems"]), float(t["n_tiles"]), float(m.get("e_launch", 0)),
            float(m["parity"]["regions_checked"]) if have_e2e else 0.0, float(m["parity"].get("ints_checked", 0)) if have_e2e else 0.0]
    mins = [1.0 if (not have_e2e or m["parity"]["equal"]) else 0.0]
    vals = torch.tensor(vals, dtype=torch.float64, device="cuda")
    sums = torch.tensor(sums, dtype=torch.float64, device="cuda")
    mins = torch.tensor(mins, dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
        dist.all_reduce(mins, op=dist.ReduceOp.MIN)
    vals, sums, mins = vals.cpu().tolist(), sums.cpu().tolist(), mins.cpu().tolist()
    if rank != 0:
        return None, mins[0] > 0
    dev_ms, wall_ms, e2e_ms, e2e_med, e2e_min = vals[:5]
    kms = dict(zip(ks, vals[5:]))
    reads_all, reads_e2e, launches_all, h2d_all, d2h_all, cand, out_elems, n_tiles, e_launch, p_regions, p_ints = sums
    steps = m["steps"]
    ms_per_step = dev_ms / steps
    preset, info = m["preset"], m["info"]
    units = float(info["records"])

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    else:
        peak, peak_src = 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"
    per = 1.0 / world       # per-launch figures: every rank runs the same kernels on 1/world of the data
    alg = {"decode": K1_BYTES_PER_READ * reads_all * per, "filter": 0.0,
           "join": 24.0 * n_tiles * per, "count": (8.0 * cand + 4.0 * out_elems) * per}
    traffic_tab = {}
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        traffic_tab = json.load(open(tpath)).get(preset, {})
    kern = {}
    for name, key in (("decode", "ms_decode"), ("join", "ms_join"), ("count", "ms_count")):
        ms = kms[key] / steps
        k = {"ms": round(ms, 4), "alg_gb": round(alg[name] / 1e9, 4),
             "gbs": round(alg[name] / 1e9 / (ms * 1e-3), 1) if ms > 0 else None}
        k["frac"] = round(k["gbs"] / peak, 4) if k["gbs"] else None
        ent = traffic_tab.get(name)
        if ent and ms > 0:      # ncu dram__bytes_read+write of one launch per unit of work (record / output int), scaled
            dram = ent["dram_bytes_per_read"] * reads_all * per if "dram_bytes_per_read" in ent else \
                (ent["dram_bytes_per_out_int"] * out_elems * per if "dram_bytes_per_out_int" in ent else None)
            if dram:
                k["dram_gb"] = round(dram / 1e9, 4)
                k["frac_dram"] = round(dram / 1e9 / (ms * 1e-3) / peak, 4)
        kern[name] = k
    dom = max(kern, key=lambda k: kern[k]["ms"])
    roofline = {"bound": "hbm", "kernel": dom, "achieved": kern[dom]["gbs"], "peak": peak, "unit": "GB/s",
                "frac": kern[dom]["frac"], "traffic": int(kern[dom]["dram_gb"] * 1e9) if "dram_gb" in kern[dom] else None,
                "frac_dram": kern[dom].get("frac_dram"), "peak_source": peak_src, "kernels": kern,
                "note": "achieved = algorithmic bytes (DESIGN 4) / CUDA-event time; frac_dram = ncu DRAM bytes "
                        "(profiles/traffic.json, scaled per unit) / time / peak - the figure that cannot exceed 1",
                "step_share": {k: round(kern[k]["ms"] / max(1e-9, sum(v["ms"] for v in kern.values())), 3) for k in kern}}
    line = {
        "metric": "reads counted/sec", "value": reads_all / (ms_per_step * 1e-3), "unit": "reads/s", "n_gpus": world,
        "steps": steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "value_definition": "kernel-only, SURVEY 8d: records decoded per step / CUDA-event time of K1+K3+K4/K5 with inputs resident in HBM",
        "value_job_units": units / (ms_per_step * 1e-3),
        "config": config_dict(preset, m["gs"], info, len(m["gr_all"]), m["fn"], m["kw"]),
        "details": {"reads_decoded_per_step": int(reads_all), "host_threads": os.cpu_count(),
                    "parallelism": f"regions sharded over {world} GPU(s), no collective",
                    "bytes_per_step": f"{(te or m['t_stage'])['bytes_inflated'] * world / 1e9:.2f} GB raw (rank 0 x N) + {4 * out_elems / 1e9:.2f} GB out"},
        "gpu_launches": int(launches_all + e_launch),
        "wall_ms_per_step": wall_ms / steps,
        "roofline": roofline,
        "setup": {"bam_generate_s": round(m["t_gen"], 1), "stage_ms": round(m["t_stage"]["ms_total"], 1)},
    }
    if have_e2e:
        line["e2e"] = {"value": units / (e2e_ms * 1e-3), "unit": "reads/s", "ms_per_step": e2e_ms,
                       "ms_per_call_median_maxrank": e2e_med, "ms_per_call_min_maxrank": e2e_min, "steps": len(m["per_call"]),
                       "reads_decoded_per_s": reads_e2e / (e2e_ms * 1e-3), "reads_decoded_per_step": int(reads_e2e),
                       "h2d_bytes_per_step": int(h2d_all), "d2h_bytes_per_step": int(d2h_all),
                       "breakdown_ms_rank0": {k: round(te[k], 2) for k in ("ms_plan", "ms_fetch", "ms_h2d", "ms_inflate_gpu", "ms_d2h", "ms_kernels", "ms_total")},
                       "inflate": "gpu" if args.gpu_inflate else "host zlib",
                       "inflate_kernel": ({"ms_rank0": round(te["ms_inflate_gpu"], 2), "out_gbs": round(te["bytes_inflated"] / max(1e-9, te["ms_inflate_gpu"]) / 1e6, 1),
                                           "hbm_frac_in_plus_out": round((te["bytes_inflated"] + te["bytes_compressed"]) / max(1e-9, te["ms_inflate_gpu"]) / 1e6 / peak, 4),
                                           "note": "CUDA-event time of the inflate launches of rank 0's last call (they share the SMs with the walk / CRC32 / K1 of the "
                                                   "neighbouring batches); Huffman decoding is bound by dependent-instruction latency, not by HBM"}
                                          if args.gpu_inflate and te["ms_inflate_gpu"] > 0 else None),
                       "note": "BAM file in page cache -> result in host memory; inflate (GPU kernel or host zlib pool) is inside; "
                               "value in job units (BAM records / s), the unit the reference arm uses"}
        line["parity"] = {"equal": mins[0] > 0, "regions_checked": int(p_regions), "ints_checked": int(p_ints),
                          "checker": m["parity"]["checker"],
                          "what": "result of the last timed e2e call vs the CPU checker, a genomic prefix of every rank's regions"}
        if "full" in m["parity"]:
            line["parity"]["full"] = m["parity"]["full"]
            line["parity"]["what"] += "; full = EVERY region of that result vs the checker on all host threads"
        if "cpu" in m:
            line["cpu_baseline"] = m["cpu"]
    return line, mins[0] > 0


def c1_latency(args):
    """BASELINE config C1: bamCount(randomBam.bam, 100 random 1 kb regions, mapqual=0, ss=FALSE) - call latency of the GPU
    path next to the single-threaded CPU checker (the reference 'runs on CPU today')."""
    import bamsignals_b200 as B
    import spec_r
    kind, O = checker()
    bam = os.path.join(ROOT, "tests", "golden", "randomBam.bam")
    g = spec_r.test_regions(seed=1, n=100)
    gr = B.GRanges([["chr1", "chr2", "chr3"][i] for i in g["rname"]], g["start"], [1000] * 100, g["strand"])
    want = O.bamCount(bam, gr, mapqual=0, ss=False, impl=kind)

    def med(f, n):
        ts = []
        for _ in range(n):
            t0 = time.perf_counter()
            r = f()
            ts.append((time.perf_counter() - t0) * 1e3)
        return float(np.median(ts)), float(min(ts)), r
    g_ms, g_min, got = med(lambda: B.bamCount(bam, gr, mapqual=0, ss=False), 25)
    t = B.timings()
    c_ms, c_min, _ = med(lambda: O.bamCount(bam, gr, mapqual=0, ss=False, impl=kind), 25)
    return {"workload": "c1: bamCount(randomBam.bam, 100 x 1 kb, mapqual=0, ss=FALSE), 99,000-read fixture of the reference",
            "gpu_ms_per_call_median": g_ms, "gpu_ms_per_call_min": g_min, "cpu_ms_per_call_median": c_ms, "cpu_ms_per_call_min": c_min,
            "cpu": f"{kind}, 1 thread", "gpu_over_cpu": c_ms / g_ms, "gpu_launches_per_call": int(t["n_launches"]),
            "gpu_breakdown_ms": {k: round(t[k], 3) for k in ("ms_plan", "ms_fetch", "ms_inflate_gpu", "ms_kernels", "ms_total")},
            "parity": {"equal": bool(np.array_equal(got, want)), "regions_checked": 100, "checker": kind}}


def cold_call(args, preset, gs):
    """First call of a fresh process (CUDA context, library buffers, index parse: nothing cached) next to its second."""
    r = subprocess.run([sys.executable, os.path.abspath(__file__), "--cold-child", "--preset", preset, "--gscale", repr(gs),
                        "--gpu-inflate", str(args.gpu_inflate)], capture_output=True, text=True, timeout=600)
    try:
        return json.loads(r.stdout.strip().splitlines()[-1])
    except Exception:
        return {"error": (r.stderr or r.stdout)[-300:]}


def cold_child(args):
    t_imp = time.perf_counter()
    import bamsignals_b200 as B
    bam, info = WL.make_bam(args.preset, args.gscale, data_dir())
    gr, kw, fn = WL.regions(args.preset, args.gscale)
    opts = B.default_opts(devices=[0], inflate_threads=os.cpu_count() or 1, gpu_inflate=1 if args.gpu_inflate else -1)
    B.lib()
    t_imp = time.perf_counter() - t_imp
    ms = []
    for _ in range(3):
        t0 = time.perf_counter()
        r = getattr(B, fn)(bam, gr, opts=opts, **kw)
        ms.append((time.perf_counter() - t0) * 1e3)
        del r
    print(json.dumps({"first_call_ms": ms[0], "second_call_ms": ms[1], "third_call_ms": ms[2],
                      "note": "fresh process: the first call creates the CUDA context, allocates the device / pinned buffers and parses the index"}))


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

    def wait_for_rank0(tag):
        """The other ranks wait on the CPU (TCP store), not in an NCCL kernel spinning on their GPUs, while rank 0 uses
        every GPU of the box for the in-library multi-device call."""
        if dist is None:
            return
        store = dist.distributed_c10d._get_default_store()
        if rank == 0:
            store.set(tag, "1")
        else:
            store.wait([tag])

    sampler = ClockSampler(local_rank)
    sampler.start()
    t_wait = time.time()
    while not sampler.rows and time.time() - t_wait < 3.0:      # nvidia-smi needs a moment to produce its first row
        time.sleep(0.02)
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    m = measure_ours(args, args.preset, args.gscale, rank, world, local_rank, dist, barrier, sampler, args.steps, args.warmup, e2e_steps)
    if args.profile:                     # ncu runs: only the resident steps matter
        sampler.summary()
        if rank == 0:
            print(json.dumps({"profile_only": True, "ms_device_per_step": m["dev_ms"] / args.steps, "kernels_ms": m["ksum"],
                              "launches": m["launches"], "reads_decoded": m["t_res"]["records"]}), flush=True)
        return
    # clocks + throttle reasons sampled over BOTH timed regions of the headline preset (the resident steps alone last a few
    # milliseconds, less than one 100 ms nvidia-smi sample)
    clocks = sampler.summary()
    line, ok = reduce_and_format(args, m, rank, world, dist, torch)
    all_ok = ok
    if rank == 0:
        line["clocks"] = clocks
    if world > 1:
        barrier()
        if rank == 0:
            line["inlib"] = inlib_call(args, m, world, steps=min(5, e2e_steps))
            all_ok = all_ok and line["inlib"]["parity"]["equal"]
        wait_for_rank0("inlib_" + args.preset)
        barrier()
    also = [] if args.also in ("", "none") else [x.strip() for x in args.also.split(",") if x.strip()]
    extra = {}
    for name in also:
        if name == "c1":
            if rank == 0:
                extra["c1"] = c1_latency(args)
                all_ok = all_ok and extra["c1"]["parity"]["equal"]
            barrier()
        elif name in FULL_READS and name != args.preset:
            m2 = measure_ours(args, name, args.gscale, rank, world, local_rank, dist, barrier, None, max(3, args.steps // 2),
                              min(args.warmup, 3), max(3, e2e_steps // 2))
            l2, ok2 = reduce_and_format(args, m2, rank, world, dist, torch)
            all_ok = all_ok and ok2
            if world > 1:
                barrier()
                if rank == 0:
                    l2["inlib"] = inlib_call(args, m2, world, steps=3)
                    all_ok = all_ok and l2["inlib"]["parity"]["equal"]
                wait_for_rank0("inlib_" + name)
                barrier()
            if rank == 0:
                extra[name] = l2
    if rank == 0:
        if world == 1 and not args.no_cold:
            line["e2e"]["cold"] = cold_call(args, args.preset, args.gscale)
        if extra:
            line["configs"] = extra
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()
    if not all_ok:
        sys.stderr.write("bench.py: PARITY MISMATCH between the timed GPU result and the CPU checker\n")
        sys.exit(1)


def reference_steps(call, bam, gr, kw, threads, kind, steps, warmup):
    import oracle_api as O
    for _ in range(warmup):
        call(bam, gr, nthreads=threads, impl=kind, **kw)
    t0 = time.perf_counter()
    recs = 0
    for _ in range(steps):
        call(bam, gr, nthreads=threads, impl=kind, **kw)
        recs = O.stats()["records"]
    return (time.perf_counter() - t0) / steps, recs


def run_reference(args, rank, world):
    """The reference arm: the reference CPU path on all host threads (region shards, one reader each) - oracle/_ref, the
    reference's own src/bamsignals.cpp compiled unchanged, when it was built; else the restated port."""
    if rank != 0:
        return
    kind, O = checker()
    preset, gs = args.preset, args.gscale
    bam, info = WL.make_bam(preset, gs, data_dir())
    gr, kw, fn = WL.regions(preset, gs)
    threads = os.cpu_count() or 1
    dt, recs = reference_steps(getattr(O, fn), bam, gr, kw, threads, kind, args.steps, min(args.warmup, 1))
    v = info["records"] / dt          # same unit of work as the GPU arm's e2e: the job's BAM records per second
    what = ("the reference's own src/bamsignals.cpp compiled unchanged (oracle/_ref) over the repo's htslib stand-in (zlib inflate)"
            if kind == "ref" else "restated reference CPU path (oracle/ port, zlib inflate; R/Rhtslib not installable offline)")
    line = {"impl": "reference", "metric": "reads counted/sec", "value": v, "unit": "reads/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": config_dict(preset, gs, info, len(gr), fn, kw),
            "details": {"reads_streamed_per_step": int(recs), "host_threads": threads},
            "cpu_baseline": {"value": v, "unit": "reads/s", "cores": threads, "kind": "reference" if kind == "ref" else "port",
                             "sample": f"whole workload per step; {what}, indexed access, {threads} threads over disjoint region shards"},
            "e2e": {"value": v, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    also = [] if args.also in ("", "none") else [x.strip() for x in args.also.split(",") if x.strip()]
    extra = {}
    for name in also:
        if name == "c1":
            import bamsignals_b200 as B
            import spec_r
            fb = os.path.join(ROOT, "tests", "golden", "randomBam.bam")
            g = spec_r.test_regions(seed=1, n=100)
            g1 = B.GRanges([["chr1", "chr2", "chr3"][i] for i in g["rname"]], g["start"], [1000] * 100, g["strand"])
            ts = []
            for _ in range(25):
                t0 = time.perf_counter()
                O.bamCount(fb, g1, mapqual=0, ss=False, impl=kind)
                ts.append((time.perf_counter() - t0) * 1e3)
            extra["c1"] = {"cpu_ms_per_call_median": float(np.median(ts)), "cpu": f"{kind}, 1 thread"}
        elif name in FULL_READS and name != preset:
            # a bounded sample: a genomic prefix of the (bin-aligned pieces of the) regions sized for ~cpu_seconds per step
            bam2, info2 = WL.make_bam(name, gs, data_dir())
            gr2, kw2, fn2 = WL.regions(name, gs)
            pieces = gr2 if fn2 == "bamCount" else WL.split_wide_regions(gr2, 64, int(kw2.get("binsize", 1)))
            order = np.lexsort((pieces.start, pieces.seq_idx))
            n = len(order)
            k0 = max(1, n // 400)
            t0 = time.perf_counter()
            getattr(O, fn2)(bam2, pieces[np.sort(order[:k0])], nthreads=threads, impl=kind, **kw2)
            d0 = max(1e-4, time.perf_counter() - t0)
            k = int(min(n, max(k0, k0 * args.cpu_seconds / d0)))
            sub = pieces[np.sort(order[:k])]
            dt2, recs2 = reference_steps(getattr(O, fn2), bam2, sub, kw2, threads, kind, 2, 0)
            share = float(sub.width.astype(np.int64).sum()) / float(pieces.width.astype(np.int64).sum())
            extra[name] = {"impl": "reference", "value": info2["records"] * share / dt2, "unit": "reads/s",
                           "config": config_dict(name, gs, info2, len(gr2), fn2, kw2),
                           "sample": f"first {k:,} of {n:,} bin-aligned region pieces in genomic order = {share:.3f} of the job "
                                     f"per step ({recs2:,} records streamed, {dt2:.1f} s), {threads} threads, extrapolated to the job"}
    if extra:
        line["configs"] = extra
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--preset", default="c2", choices=["c2", "c3", "c4", "c5"])
    ap.add_argument("--gscale", type=float, default=1.0, help="genome (and read-count) scale; 1.0 = the full configuration")
    ap.add_argument("--also", default=None, help="extra entries under 'configs': comma list of c1 (call latency on the "
                    "reference's fixture) and presets; default c1,c4 at gscale 1, c1 otherwise; 'none' to skip")
    ap.add_argument("--e2e-steps", type=int, default=12)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--full-parity-seconds", type=float, default=40.0, help="N = 1: also compare EVERY region of the timed result "
                    "with the checker on all host threads when that is estimated to take no longer than this (0 = off)")
    ap.add_argument("--parity-seconds", type=float, default=4.0, help="CPU-checker budget per rank for the parity check (N > 1)")
    ap.add_argument("--gpu-inflate", type=int, default=1, help="1: inflate BGZF on the GPU (default), 0: host zlib pool")
    ap.add_argument("--profile", action="store_true", help="resident steps only (for runs under ncu)")
    ap.add_argument("--no-cold", action="store_true", help="skip the fresh-process cold-call figure")
    ap.add_argument("--cold-child", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.also is None:
        args.also = "c1,c4" if args.gscale >= 1.0 and args.preset == "c2" else "c1"
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if args.cold_child:
        cold_child(args)
    elif args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
