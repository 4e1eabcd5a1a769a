"""numpy restatement of the reference's OWN test oracle (the R functions df2gr / countR / profileR / coverageR in
/root/reference/tests/testthat/utils.R:178-311), fed with the reference's read table randomReads.RData (committed as
tests/golden/randomReads.npz).  It is independent of the BAM bytes and of src/bamsignals.cpp, which is what makes it a
pin for oracle/bsg_oracle.cpp: the reference's test-suite asserts bamCount/bamProfile/bamCoverage == these functions
(tests/testthat/test_methods.R:33-104).  Test infrastructure only.
"""
import itertools
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_reads():
    z = np.load(os.path.join(GOLDEN, "randomReads.npz"))
    return {k: z[k] for k in z.files}


def load_annot():
    return json.load(open(os.path.join(GOLDEN, "randomAnnot.json")))


def df2gr(reads, paired_end="ignore", shift=0, mapqual=0, tlenFilter=None):
    """utils.R:178-225 -> dict(rname, start, end, strand) with 1-based inclusive coordinates."""
    assert paired_end in ("ignore", "filter", "midpoint", "extend")
    keep = reads["mapq"] >= mapqual                                        # utils.R:184
    if paired_end != "ignore":
        keep &= reads["read1"]                                             # utils.R:188
        lo, hi = (0, 1000) if tlenFilter is None else tlenFilter           # utils.R:190-194
        a = np.abs(reads["isize"])
        keep &= (a >= lo) & (a <= hi)
    rname = reads["rname"][keep]
    pos = reads["pos"][keep].astype(np.int64)
    qwidth = reads["qwidth"][keep].astype(np.int64)
    strand = reads["strand"][keep].astype(np.int64)
    isize = reads["isize"][keep].astype(np.int64)
    if paired_end in ("extend", "midpoint"):                               # utils.R:198-204
        neg = strand < 0
        pos = np.where(neg, pos - np.abs(isize) + qwidth, pos)
        qwidth = np.abs(isize)
    start = pos
    end = pos + qwidth - 1
    if paired_end == "midpoint":                                           # utils.R:208-218
        mids2 = start + end                                                # 2 * mid
        mids = np.where(strand > 0, -((-mids2) // 2), mids2 // 2)          # ceil for +, floor for -
        start = mids
        end = mids
    sh = np.where(strand < 0, -shift, shift)                               # utils.R:221-224
    return dict(rname=rname, start=start + sh, end=end + sh, strand=strand)


def _genes(genes):
    g_start = np.asarray(genes["start"], dtype=np.int64)
    g_width = np.asarray(genes["width"], dtype=np.int64)
    return g_start, g_start + g_width - 1, list(genes["strand"]), list(genes["rname"])


def profileR(reads, genes, ss=False, **kw):
    """utils.R:255-291; genes = dict(rname(code), start(1-based), width, strand '+','-','*')."""
    gr = df2gr(reads, **kw)
    g_start, g_end, g_strand, g_rname = _genes(genes)
    out = []
    for g in range(len(g_start)):
        glen = int(g_end[g] - g_start[g] + 1)
        on = gr["rname"] == g_rname[g]
        posr = on & (gr["strand"] > 0)
        negr = on & (gr["strand"] < 0)
        p = gr["start"][posr] - g_start[g]
        n = gr["end"][negr] - g_start[g]
        mat = np.zeros((2, glen), dtype=np.int64)
        mat[0] = np.bincount(p[(p >= 0) & (p < glen)], minlength=glen)[:glen] if glen else 0
        mat[1] = np.bincount(n[(n >= 0) & (n < glen)], minlength=glen)[:glen] if glen else 0
        if g_strand[g] == "-":                                             # rev(mat): positions and rows flip
            mat = mat[::-1, ::-1]
        out.append(mat.astype(np.int32) if ss else mat.sum(axis=0).astype(np.int32))
    return out


def countR(reads, genes, ss=False, **kw):
    """utils.R:228-253: returns (2,R) [sense; antisense] or (R,)."""
    prof = profileR(reads, genes, ss=True, **kw)
    m = np.stack([p.sum(axis=1) for p in prof], axis=1).astype(np.int32) if prof else np.zeros((2, 0), np.int32)
    return m if ss else m.sum(axis=0).astype(np.int32)


def coverageR(reads, genes, **kw):
    """utils.R:293-311."""
    gr = df2gr(reads, **kw)
    g_start, g_end, g_strand, g_rname = _genes(genes)
    out = []
    for g in range(len(g_start)):
        glen = int(g_end[g] - g_start[g] + 1)
        on = gr["rname"] == g_rname[g]
        s = gr["start"][on]
        e = gr["end"][on]
        ok = e >= s                                                        # zero-width ranges cover nothing
        s, e = s[ok], e[ok]
        d = np.zeros(glen + 1, dtype=np.int64)
        hit = (s <= g_end[g]) & (e >= g_start[g])
        a = np.clip(s[hit] - g_start[g], 0, glen)
        b = np.clip(e[hit] + 1 - g_start[g], 0, glen)
        np.add.at(d, a, 1)
        np.add.at(d, b, -1)
        sig = np.cumsum(d[:glen])
        if g_strand[g] == "-":
            sig = sig[::-1]
        out.append(sig.astype(np.int32))
    return out


# ---- the parameter sweep of tests/testthat/test_methods.R:33-104 ------------------------------------------------
def sweep_pileup():
    for shift, mapq, ss, pe, tf in itertools.product((0, 100), (0, 100), (False, True),
                                                     ("ignore", "filter", "midpoint"), (None, (50, 200))):
        yield dict(shift=shift, mapqual=mapq, ss=ss, paired_end=pe, tlenFilter=tf)


def sweep_coverage():
    for mapq, pe, tf in itertools.product((0, 100), ("ignore", "extend"), (None, (50, 200))):
        yield dict(mapqual=mapq, paired_end=pe, tlenFilter=tf)


def case_key(kind, case):
    return kind + "|" + ",".join(f"{k}={case[k]}" for k in sorted(case))


def test_regions(seed=7, n=20):
    """Regions drawn like test_methods.R:11-20 (the reference draws them unseeded at test time)."""
    rng = np.random.default_rng(seed)
    return dict(rname=rng.integers(0, 3, n).tolist(), strand=rng.choice(["+", "-"], n).tolist(),
                start=rng.integers(1, 1001, n).tolist(), width=(1 + rng.poisson(199, n)).tolist())


def annot_regions():
    a = load_annot()
    reads = load_reads()
    lv = list(reads["rname_levels"])
    return dict(rname=[lv.index(s) for s in a["seqnames"]], strand=a["strand"], start=a["start"], width=a["width"])


def write_expected(path):
    reads = load_reads()
    out = {}
    for tag, genes in (("rand7", test_regions()), ("annot", annot_regions())):
        for case in sweep_pileup():
            kw = {k: v for k, v in case.items() if k != "ss"}
            out[f"{tag}|" + case_key("count", case)] = countR(reads, genes, ss=case["ss"], **kw).ravel(order="F")
            prof = profileR(reads, genes, ss=case["ss"], **kw)
            out[f"{tag}|" + case_key("profile", case)] = np.concatenate([p.ravel(order="F") for p in prof])
        for case in sweep_coverage():
            cov = coverageR(reads, genes, **case)
            out[f"{tag}|" + case_key("coverage", case)] = np.concatenate(cov)
    np.savez_compressed(path, **out)
    print("wrote", path, len(out), "arrays")
