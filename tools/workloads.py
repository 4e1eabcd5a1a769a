"""Synthetic workloads for BASELINE.json's configs C2..C5: BAM files come from tools/bamgen
(deterministic, seeded), region sets are drawn here with fixed seeds.  Used by tests/ and bench.py only.

`gscale` shrinks every contig (and the read count with it, so read density stays that of the full-size config);
gscale=1 is the full-size configuration."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bamsignals_b200.api import GRanges  # noqa: E402

BAMGEN = os.path.join(ROOT, "tools", "bamgen")
HG38 = [248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636, 138394717, 133797422,
        135086622, 133275309, 114364328, 107043718, 101991189, 90338345, 83257441, 80373285, 58617616, 64444167,
        46709983, 50818468, 156040895, 57227415]
NAMES = [f"chr{i}" for i in range(1, 23)] + ["chrX", "chrY"]


def contig_lens(preset, gscale):
    n = 1 if preset == "c3" else 24
    return [max(20000, int(l * gscale)) for l in HG38[:n]]


def ensure_bamgen():
    if not os.path.exists(BAMGEN) or os.path.getmtime(BAMGEN) < os.path.getmtime(BAMGEN + ".cpp"):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tools")])


def make_bam(preset, gscale, outdir, record="compact", threads=0, straddle=0.05, unplaced=0, level=1, density=1.0):
    """Generate (or reuse) the BAM of `preset` at genome scale `gscale`; returns (path, info dict)."""
    ensure_bamgen()
    os.makedirs(outdir, exist_ok=True)
    tag = f"{preset}_g{gscale:g}_d{density:g}_{record}_s{straddle:g}_u{unplaced}_l{level}"
    path = os.path.join(outdir, tag + ".bam")
    meta = path + ".json"
    if os.path.exists(path) and os.path.exists(path + ".bai") and os.path.exists(meta):
        return path, json.load(open(meta))
    cmd = [BAMGEN, "--out", path, "--preset", preset, "--scale", repr(gscale * density), "--genome-scale", repr(gscale),
           "--record", record, "--straddle", repr(straddle), "--unplaced", str(unplaced), "--level", str(level)]
    if threads:
        cmd += ["--threads", str(threads)]
    out = subprocess.check_output(cmd)
    info = json.loads(out.decode().strip().splitlines()[-1])
    json.dump(info, open(meta, "w"))
    return path, info


def regions(preset, gscale, seed=None):
    """The region set of the config at genome scale gscale -> (GRanges, call kwargs, api function name)."""
    lens = contig_lens(preset, gscale)
    rng = np.random.default_rng(seed if seed is not None else {"c2": 2, "c3": 3, "c4": 4, "c5": 5}[preset])
    if preset == "c2":      # bamProfile binsize=1 ss=TRUE shift=75 over 100k 2 kb promoter windows
        n = max(8, int(round(100000 * gscale)))
        w = np.asarray(lens, dtype=np.float64)
        ci = rng.choice(len(lens), n, p=w / w.sum())
        tss = (rng.random(n) * (np.asarray(lens)[ci] - 2000)).astype(np.int64) + 1000
        gr = GRanges.from_codes(NAMES[:len(lens)], ci, tss - 999, np.full(n, 2000), rng.choice(np.array([1, -1], np.int8), n))
        return gr, dict(binsize=1, ss=True, shift=75), "bamProfile"
    if preset == "c3":      # bamCoverage paired.end="extend" over 10k 100 kb windows on chr1
        n = max(4, int(round(10000 * gscale)))
        width = min(100000, lens[0] // 2)
        start = (rng.random(n) * (lens[0] - width)).astype(np.int64) + 1
        gr = GRanges.from_codes(NAMES[:1], np.zeros(n, np.int32), start, np.full(n, width), rng.choice(np.array([1, -1], np.int8), n))
        return gr, dict(paired_end="extend"), "bamCoverage"
    if preset == "c4":      # bamProfile binsize=200 midpoint tlenFilter=c(70,200), the 24 whole contigs, strand '*'
        gr = GRanges.from_codes(NAMES[:len(lens)], np.arange(len(lens)), np.ones(len(lens)), np.asarray(lens), np.zeros(len(lens), np.int8))
        return gr, dict(binsize=200, paired_end="midpoint", tlenFilter=(70, 200)), "bamProfile"
    if preset == "c5":      # bamCount over 5 kb windows every 3 kb, mapqual=30, filteredFlag=1024, ss=TRUE
        ci, st = [], []
        for c, l in enumerate(lens):
            s = np.arange(1, max(2, l - 5000 + 1), 3000, dtype=np.int64)
            ci.append(np.full(len(s), c, np.int32))
            st.append(s)
        ci, st = np.concatenate(ci), np.concatenate(st)
        strand = np.where(np.arange(len(st)) % 2 == 0, 1, -1).astype(np.int8)
        gr = GRanges.from_codes(NAMES[:len(lens)], ci, st, np.full(len(st), 5000), strand)
        return gr, dict(mapqual=30, filteredFlag=1024, ss=True), "bamCount"
    raise ValueError(preset)


def as_flat(result):
    """Flatten a bamCount array or a CountSignals into one int32 vector in bsg_output_layout() order."""
    if isinstance(result, np.ndarray):
        return result.ravel(order="F")
    sig = result.as_list()
    if not sig:
        return np.zeros(0, np.int32)
    return np.concatenate([s.ravel(order="F") for s in sig])


def split_wide_regions(gr, world, align=1):
    """Cut regions that are much wider than a rank's share into pieces of a multiple of `align` bp (the bin size),
    counted from the region's 5' end (start for '+' / '*', end for '-'): each piece behaves like a region of its own
    and the pieces' results, concatenated in 5' -> 3' order, are the region's result.  A config with a few
    huge regions (C4: 24 whole contigs) can then be balanced over the ranks by reads instead of by region count."""
    total = int(gr.width.astype(np.int64).sum())
    piece = max(align, -(-(total // (world * 64) + 1) // align) * align)
    ci, start, width, strand = [], [], [], []
    for i in range(len(gr)):
        w, s0, st, c = int(gr.width[i]), int(gr.start[i]), int(gr.strand[i]), int(gr.seq_idx[i])
        if w <= 2 * piece:
            ci.append(c); start.append(s0); width.append(w); strand.append(st)
            continue
        for lo in range(0, w, piece):
            pw = min(piece, w - lo)
            ci.append(c); width.append(pw); strand.append(st)
            start.append(s0 + w - lo - pw if st < 0 else s0 + lo)
    return GRanges.from_codes(gr.seqlevels, np.asarray(ci, np.int32), np.asarray(start, np.int64), np.asarray(width, np.int64),
                              np.asarray(strand, np.int8))


def shard_regions(gr, rank, world, split_align=None):
    """Contiguous slice (in (chromosome, start) order) of the region set for this rank -> (GRanges, original indices).
    Regions are independent: no exchange step, every output element has exactly one owner.
    split_align (a bin size): wide regions are first cut into bin-aligned pieces (split_wide_regions); the returned
    indices then refer to the pieces."""
    if world == 1:
        return gr, np.arange(len(gr))
    if split_align:
        gr = split_wide_regions(gr, world, int(split_align))
    order = np.lexsort((gr.start, gr.seq_idx))
    lo, hi = len(gr) * rank // world, len(gr) * (rank + 1) // world
    idx = np.sort(order[lo:hi])
    return gr[idx], idx


def merge_shards(parts, offsets):
    """parts: [(idx, flat)] from every rank, each flat in the shard's own bsg_output_layout() order; offsets: the
    full region set's layout (R+1).  Returns the full flat result."""
    out = np.zeros(int(offsets[-1]), dtype=np.int32)
    for idx, flat in parts:
        pos = 0
        for i in idx:
            n = int(offsets[i + 1] - offsets[i])
            out[offsets[i]:offsets[i + 1]] = flat[pos:pos + n]
            pos += n
        assert pos == len(flat)
    return out
