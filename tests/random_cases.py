"""Seeded random scenarios (BAM + regions + call) and an independent numpy statement of SURVEY App. A.

Test infrastructure.  `spec_counts` computes the expected result of a call straight from the Python read dicts that
were written into the BAM: no BGZF, no BAM parsing, no index, no sweep.  It is the third voice beside the oracle
(oracle/bsg_oracle.cpp) and the CUDA path, and it covers what the reference's own fixture never exercises (SURVEY
section 4, last bullet): arbitrary CIGARs, flag-0x4 reads with a position, arbitrary flag masks, `*` regions,
zero-width / off-chromosome / negative-start regions, binsize > 1, out-of-range mapqual.
"""
import numpy as np

import bamwriter as W
from bamsignals_b200 import GRanges, core_args

OPS = "MIDNSHP=X"


def random_cigar(rng):
    k = int(rng.integers(0, 7))
    if k == 0:
        return "*"
    ops = []
    for _ in range(k):
        op = OPS[int(rng.choice(9, p=[.4, .08, .1, .08, .1, .04, .02, .1, .08]))]
        ln = int(rng.integers(0, 4)) if rng.random() < .05 else int(rng.integers(1, 3000 if op == "N" else 120))
        ops.append(f"{ln}{op}")
    return "".join(ops)


def scenario(seed):
    """-> dict(refs, reads, gr, fn, kw, writer_kw)"""
    rng = np.random.default_rng(1000 + seed)
    n_ref = int(rng.integers(1, 5))
    refs = [(f"ctg{i}", int(rng.choice([600, 20000, 300000]))) for i in range(n_ref)]
    n_reads = int(rng.choice([0, 1, 60, 2500, 6000]))
    reads = []
    per = np.sort(rng.integers(0, n_ref, n_reads))
    flag_pool = [0, 16, 99, 147, 83, 163, 1123, 1187, 4, 20, 73, 133, 2064, 256, 1024, 1040, 67, 115]
    for tid in range(n_ref):
        m = int((per == tid).sum())
        L = refs[tid][1]
        if rng.random() < .5:                       # clustered: many reads share a position
            centres = rng.integers(0, L, max(1, m // 40))
            pos = np.clip(centres[rng.integers(0, len(centres), m)] + rng.integers(-3, 4, m), 0, L - 1)
        else:
            pos = rng.integers(0, L, m)
        for p in np.sort(pos):
            flag = int(rng.integers(0, 4096)) if rng.random() < .2 else int(rng.choice(flag_pool))
            tl = int(rng.choice([0, 1, -1, 35, -35, 149, -149, 150, -150, 401, -401, 1000, -1000, 1001, 250000, -250000]))
            if rng.random() < .3:
                tl = int(rng.integers(-700, 700))
            mq = int(rng.choice([0, 1, 29, 30, 60, 254, 255])) if rng.random() < .5 else int(rng.integers(0, 61))
            reads.append(dict(tid=tid, pos=int(p), flag=flag, mapq=mq, cigar=random_cigar(rng), tlen=tl,
                              l_seq=int(rng.integers(0, 90)) if rng.random() < .3 else 0))
    for _ in range(int(rng.integers(0, 4))):
        reads.append(dict(tid=-1, pos=-1, flag=4, mapq=0, cigar="*", tlen=0))
    # regions
    R = int(rng.choice([0, 1, 3, 40, 400]))
    names, start, width = [], [], []
    for _ in range(R):
        t = int(rng.integers(0, n_ref))
        L = refs[t][1]
        names.append(refs[t][0])
        u = rng.random()
        if u < .1:
            start.append(int(rng.integers(-300, 2))); width.append(int(rng.integers(0, L + 600)))
        elif u < .2:
            start.append(int(rng.integers(max(1, L - 200), L + 400))); width.append(int(rng.integers(0, 900)))
        elif u < .3:
            start.append(1); width.append(L)
        else:
            start.append(int(rng.integers(1, L + 1))); width.append(int(rng.choice([0, 1, 2, 7, 200, 1000, 5000, 40000])))
    gr = GRanges(names, start, width, rng.choice(["+", "-", "*"], R).tolist(), seqlevels=[n for n, _ in refs][::-1])
    fn = str(rng.choice(["bamCount", "bamProfile", "bamCoverage"]))
    kw = dict(mapqual=int(rng.choice([0, 0, 1, 30, 60, 61, 255, 256, -5])),
              filteredFlag=int(rng.choice([-1, -1, 0, 4, 16, 1024, 1040, 20, 2048 + 256, int(rng.integers(0, 4096))])))
    if fn == "bamCoverage":
        kw["paired_end"] = str(rng.choice(["ignore", "extend"]))
    else:
        kw["paired_end"] = str(rng.choice(["ignore", "filter", "midpoint"]))
        kw["shift"] = int(rng.choice([0, 0, 1, -1, 75, -75, 5000, -5000, 100000]))
        kw["ss"] = bool(rng.integers(0, 2))
        if fn == "bamProfile":
            kw["binsize"] = int(rng.choice([1, 1, 2, 3, 4, 5, 7, 50, 200, 1000, 65536, 1000000]))
    if kw["paired_end"] != "ignore" and rng.random() < .6:
        lo = int(rng.choice([0, 1, 35, 150, 401]))
        kw["tlenFilter"] = (lo, lo + int(rng.choice([0, 1, 115, 600, 300000])))
    writer_kw = dict(block_payload=int(rng.choice([120, 997, 4000, 0xFF00])), cut_mid_record=bool(rng.integers(0, 2)),
                     level=int(rng.choice([0, 1, 6, 9])))
    return dict(refs=refs, reads=reads, gr=gr, fn=fn, kw=kw, writer_kw=writer_kw)


def write(sc, path):
    W.write_bam(path, sc["refs"], sc["reads"], **sc["writer_kw"])
    return path


def _columns(sc):
    rd = [r for r in sc["reads"] if r["tid"] >= 0]
    col = lambda k: np.array([r[k] for r in rd], dtype=np.int64)       # noqa: E731
    rlen = np.array([W.ref_len(W.cigar_ops(r["cigar"]), r["flag"]) or 1 for r in rd], dtype=np.int64)   # App. A.2
    return col("tid"), col("pos"), col("flag"), col("mapq"), col("tlen"), rlen


def spec_counts(sc):
    """Expected flat int32 result of the scenario's call, from SURVEY App. A.2-A.8 alone."""
    fn, gr = sc["fn"], sc["gr"]
    ca = core_args(fn, **sc["kw"])
    tid, pos, flag, mapq, tlen, rlen = _columns(sc)
    name2tid = {n: i for i, (n, _) in enumerate(sc["refs"])}
    # A.3 filter, in 32-bit two's complement like the reference's `int` flags
    req, flt = ca["requiredF"] & 0xFFFFFFFF, ca["filteredF"] & 0xFFFFFFFF
    nflag = ~flag & 0xFFFFFFFF
    keep = (mapq >= ca["mapqual"]) & ((req & nflag) == 0) & ((flt & nflag) != 0)
    if ca["tlen_filter"] is not None:
        keep &= (np.abs(tlen) >= ca["tlen_filter"][0]) & (np.abs(tlen) <= ca["tlen_filter"][1])
    neg = (flag & 16) != 0
    end = pos + rlen - 1
    out = []
    if fn == "bamCoverage":
        s, e = pos.copy(), end.copy()                                  # A.6
        if ca["tspan"]:
            a = neg & (tlen < 0)
            s[a] = e[a] + tlen[a] + 1
            b = ~neg & (tlen > 0)
            e[b] = s[b] + tlen[b] - 1
        for i in range(len(gr)):
            rid, loc, ln = name2tid[gr.seqnames[i]], int(gr.start[i]) - 1, int(gr.width[i])
            sel = keep & (tid == rid) & (s < loc + ln) & (e >= loc)    # A.7
            d = np.zeros(ln + 1, dtype=np.int64)
            np.add.at(d, np.maximum(s[sel], loc) - loc, 1)
            np.add.at(d, np.minimum(e[sel], loc + ln - 1) - loc + 1, -1)
            c = np.cumsum(d[:ln])
            out.append(c[::-1] if gr.strand[i] < 0 else c)
    else:
        off = ca["shift"] + (np.abs(tlen) // 2 if ca["pe_mid"] else 0)  # A.4
        pos5 = np.where(neg, end - off, pos + off)
        ss, bs = ca["ss"], ca["binsize"]
        mult = 2 if ss else 1
        for i in range(len(gr)):
            rid, loc, ln = name2tid[gr.seqnames[i]], int(gr.start[i]) - 1, int(gr.width[i])
            nb = 1 if bs <= 0 else -(-ln // bs)
            sel = keep & (tid == rid) & (pos5 >= loc) & (pos5 < loc + ln)      # A.5
            rel, anti = pos5[sel] - loc, neg[sel].astype(np.int64)
            if gr.strand[i] < 0:
                rel, anti = ln - 1 - rel, 1 - anti
            b = np.zeros_like(rel) if bs <= 0 else rel // bs
            v = np.zeros(nb * mult, dtype=np.int64)
            np.add.at(v, b * mult + anti if ss else b, 1)
            out.append(v)
    return np.concatenate(out).astype(np.int32) if out else np.zeros(0, dtype=np.int32)


def flat(res, fn, ss=False):
    """bamCount array / CountSignals -> the flat layout of bsg_output_layout (column-major (2, w) when ss)."""
    if fn == "bamCount":
        return np.asarray(res).reshape(-1, order="F").astype(np.int32)
    parts = [np.asarray(x).reshape(-1, order="F") for x in res.as_list()]
    return np.concatenate(parts).astype(np.int32) if parts else np.zeros(0, dtype=np.int32)
