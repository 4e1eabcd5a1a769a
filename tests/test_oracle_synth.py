"""CPU checks of the oracle beyond the reference fixture: its three access modes agree on synthetic and hand-crafted
BAMs (so the index query / maxgap chunking / sweep really are pure pruning), and bam_endpos follows
the SAM rule for every CIGAR operator."""
import os
import sys

import numpy as np
import pytest

import bamwriter as W
import edge_cases as E
import oracle_api as O

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import workloads as WL  # noqa: E402


def same(a, b):
    assert len(a) == len(b)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


@pytest.fixture(scope="module")
def variety(tmp_path_factory):
    p = str(tmp_path_factory.mktemp("edge") / "variety.bam")
    reads = E.write_variety(p, block_payload=4000)
    return p, reads


def test_endpos_rule(variety):
    p, reads = variety
    d = O.dump_reads(p)
    assert len(d["pos"]) == len(reads)
    for i, r in enumerate(reads):
        rl = W.ref_len(W.cigar_ops(r["cigar"]), r["flag"]) or 1
        assert d["endpos"][i] == r["pos"] + rl, (i, r)
        assert d["flag"][i] == r["flag"] and d["mapq"][i] == r["mapq"] and d["tlen"][i] == r["tlen"]


@pytest.mark.parametrize("straddle", [False, True])
def test_modes_agree_on_variety(tmp_path, straddle):
    p = str(tmp_path / "v.bam")
    E.write_variety(p, block_payload=997, cut_mid_record=straddle)
    gr = E.variety_regions()
    for kw in (dict(), dict(ss=True, shift=33), dict(paired_end="midpoint", tlenFilter=(0, 500), ss=True),
               dict(filteredFlag=1024, mapqual=20), dict(filteredFlag=4)):
        a = O.bamCount(p, gr, mode=O.INDEXED, **kw)
        assert np.array_equal(a, O.bamCount(p, gr, mode=O.SCAN, **kw))
        assert np.array_equal(a, O.bamCount(p, gr, mode=O.BRUTE, **kw))
        assert np.array_equal(a, O.bamCount(p, gr, mode=O.INDEXED, maxgap=0, nthreads=3, **kw))
    for kw in (dict(binsize=1, ss=True), dict(binsize=50, shift=-20)):
        same(O.bamProfile(p, gr, mode=O.INDEXED, **kw).as_list(), O.bamProfile(p, gr, mode=O.BRUTE, **kw).as_list())
    for kw in (dict(), dict(paired_end="extend", tlenFilter=(0, 600))):
        same(O.bamCoverage(p, gr, mode=O.INDEXED, **kw).as_list(), O.bamCoverage(p, gr, mode=O.BRUTE, **kw).as_list())


def test_coverage_closed_form(variety):
    """out[j] = #{kept reads with s <= x_j <= e} — checked by brute force in numpy."""
    p, reads = variety
    gr = E.variety_regions(seed=5, n=30)
    got = O.bamCoverage(p, gr).as_list()
    names = [n for n, _ in E.REFS]
    for i in range(len(gr)):
        tid = names.index(gr.seqnames[i])
        loc, w = int(gr.start[i]) - 1, int(gr.width[i])
        cov = np.zeros(w + 1, dtype=np.int64)
        for r in reads:
            if r["tid"] != tid:
                continue
            s = r["pos"]
            e = s + (W.ref_len(W.cigar_ops(r["cigar"]), r["flag"]) or 1) - 1
            a, b = max(s, loc), min(e, loc + w - 1)
            if a <= b:
                cov[a - loc] += 1
                cov[b - loc + 1] -= 1
        want = np.cumsum(cov[:w])
        if gr.strand[i] < 0:
            want = want[::-1]
        assert np.array_equal(got[i], want)


def test_long_cigar_placeholder(tmp_path):
    p = str(tmp_path / "cg.bam")
    W.write_bam(p, [("chrA", 200000)], [E.long_cigar_read()])
    d = O.dump_reads(p)
    assert d["endpos"][0] == 1000 + 70000
    from bamsignals_b200 import GRanges
    cov = O.bamCoverage(p, GRanges(["chrA"], [1], [100000])).as_list()[0]
    assert cov[999] == 0 and cov[1000] == 1 and cov[70999] == 1 and cov[71000] == 0


@pytest.mark.parametrize("preset", ["c2", "c3", "c4", "c5"])
def test_generator_and_index(tmp_path_factory, preset):
    """bamgen output: sorted, index consistent (indexed == scan), each config's call runs."""
    d = str(tmp_path_factory.mktemp("gen"))
    gs = 0.004 if preset == "c3" else 0.001
    bam, info = WL.make_bam(preset, gs, d, unplaced=5 if preset == "c4" else 0)
    rd = O.dump_reads(bam)
    assert len(rd["pos"]) == info["records"]
    key = rd["tid"].astype(np.uint32).astype(np.uint64) << np.uint64(32) | rd["pos"].astype(np.uint32).astype(np.uint64)
    assert (key[1:] >= key[:-1]).all()
    gr, kw, fn = WL.regions(preset, gs)
    a = getattr(O, fn)(bam, gr, mode=O.INDEXED, **kw)
    b = getattr(O, fn)(bam, gr, mode=O.SCAN, **kw)
    assert np.array_equal(WL.as_flat(a), WL.as_flat(b))
    assert WL.as_flat(a).sum() > 0
    if preset == "c4":
        assert (rd["tid"] == -1).sum() == 5 and (rd["flag"] & 4).sum() > 5
