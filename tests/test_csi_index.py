"""CSI indexes: the reference gets them for free from htslib's bam_index_load
(src/bamsignals.cpp:207), which looks for <bam>.csi before <bam>.bai.  The oracle restates htslib's CSI query
(generalised reg2bins + per-bin loffset); the library rebuilds its linear index from the loffsets.  Contigs longer
than 2^29 bp can only be indexed this way."""
import os
import shutil

import numpy as np
import pytest

import bamsignals_b200 as B
import bamwriter as W
import edge_cases as E
import oracle_api as O

BIG_REFS = [("chrBig", 1_500_000_000), ("chrSmall", 100_000)]


def big_reads(seed=3, n=4000):
    """reads on both sides of the 2^29 bp limit of the BAI binning scheme"""
    rng = np.random.default_rng(seed)
    centres = [1_000, 300_000_000, (1 << 29) - 500, (1 << 29) + 70_000, 900_000_000, 1_499_990_000]
    pos = np.sort(np.concatenate([c + rng.integers(0, 9_000, n // len(centres)) for c in centres]))
    reads = [dict(tid=0, pos=int(p), flag=int(rng.choice([0, 16, 99, 147])), mapq=int(rng.integers(0, 61)),
                  cigar=str(rng.choice(["50M", "20M5D25M", "10M300N40M", "5S45M"])), tlen=int(rng.integers(-400, 400))) for p in pos]
    reads += [dict(tid=1, pos=int(p), flag=0, mapq=30, cigar="40M", tlen=0) for p in np.sort(rng.integers(0, 90_000, 500))]
    return reads


def big_regions(seed=4, n=60):
    rng = np.random.default_rng(seed)
    centres = np.array([1_000, 300_000_000, (1 << 29) - 500, (1 << 29) + 70_000, 900_000_000, 1_499_990_000])
    start = centres[rng.integers(0, len(centres), n)] + rng.integers(-3_000, 9_000, n)
    start = np.clip(start, 1, None)
    names = ["chrBig"] * n
    names[::7] = ["chrSmall"] * len(names[::7])
    start = np.where(np.array(names) == "chrSmall", rng.integers(1, 80_000, n), start)
    return B.GRanges(names, start, rng.integers(1, 12_000, n), rng.choice(["+", "-", "*"], n).tolist())


def flat(x):
    return np.concatenate([a.ravel(order="F") for a in x.as_list()]) if hasattr(x, "as_list") else np.asarray(x).ravel(order="F")


# ---------------------------------------------------------------------------------------------------------------------
# oracle (CPU)
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("min_shift,depth", [(14, 5), (12, 6), (16, 4), (14, 6)])
def test_oracle_csi_matches_bai_and_scan(tmp_path, min_shift, depth):
    """The same BAM indexed both ways: indexed access through the CSI (htslib's loffset rule) returns what indexed
    access through the BAI and an index-free scan return."""
    a, b = str(tmp_path / "a.bam"), str(tmp_path / "b.bam")
    E.write_variety(a, block_payload=3000, index="bai")
    E.write_variety(b, block_payload=3000, index="csi", min_shift=min_shift, depth=depth)
    assert os.path.exists(b + ".csi") and not os.path.exists(b + ".bai")
    gr = E.variety_regions()
    for kw in (dict(ss=True, shift=33), dict(paired_end="midpoint", tlenFilter=(0, 500)), dict(filteredFlag=1024, mapqual=20)):
        want = O.bamCount(a, gr, **kw)
        assert np.array_equal(O.bamCount(b, gr, **kw), want)
        assert np.array_equal(O.bamCount(b, gr, mode=O.SCAN, **kw), want)
    same = O.bamCoverage(a, gr, paired_end="extend")
    assert np.array_equal(flat(O.bamCoverage(b, gr, paired_end="extend")), flat(same))


def test_oracle_prefers_csi_like_htslib(tmp_path):
    """<bam>.csi wins over <bam>.bai (hts_idx_load's search order): a deliberately useless .bai next to a good .csi
    must not be looked at."""
    p = str(tmp_path / "v.bam")
    E.write_variety(p, index="both")
    open(p + ".bai", "wb").write(b"garbage")
    gr = E.variety_regions()
    assert np.array_equal(O.bamCount(p, gr), O.bamCount(p, gr, mode=O.SCAN))


def test_oracle_long_contig(tmp_path):
    """positions beyond 2^29: indexed (CSI, depth 6) == scan == a brute-force count in numpy"""
    p = str(tmp_path / "big.bam")
    reads = big_reads()
    W.write_bam(p, BIG_REFS, reads, index="csi", min_shift=14, depth=6)
    gr = big_regions()
    got = O.bamCount(p, gr, mapqual=10)
    assert np.array_equal(got, O.bamCount(p, gr, mapqual=10, mode=O.SCAN))
    tid = np.array([r["tid"] for r in reads]); pos = np.array([r["pos"] for r in reads])
    rl = np.array([W.ref_len(W.cigar_ops(r["cigar"]), r["flag"]) or 1 for r in reads])
    neg = np.array([(r["flag"] & 16) != 0 for r in reads]); mq = np.array([r["mapq"] for r in reads])
    p5 = np.where(neg, pos + rl - 1, pos)
    want = []
    for i in range(len(gr)):
        t = 0 if gr.seqnames[i] == "chrBig" else 1
        s0 = int(gr.start[i]) - 1
        want.append(int(((tid == t) & (mq >= 10) & (p5 >= s0) & (p5 < s0 + int(gr.width[i]))).sum()))
    assert np.array_equal(got, np.array(want, dtype=np.int32))
    assert int(got.sum()) > 100


# ---------------------------------------------------------------------------------------------------------------------
# C ABI (CPU part: the index is parsed before any device work)
# ---------------------------------------------------------------------------------------------------------------------
def test_library_parses_csi_before_gpu_work(tmp_path):
    if B.lib().bsg_device_count() > 0:
        pytest.skip("CPU-only check")
    p = str(tmp_path / "v.bam")
    E.write_variety(p, index="csi", min_shift=12, depth=6)
    gr = E.variety_regions()
    with pytest.raises(B.BamsignalsError) as e:          # a good index gets as far as the device check
        B.bamCount(p, gr)
    assert e.value.code == -6
    raw = open(p + ".csi", "rb").read()
    open(p + ".csi", "wb").write(raw[:len(raw) // 2])    # a truncated one is a format error, reported first
    B.lib().bsg_shutdown()
    with pytest.raises(B.BamsignalsError) as e:
        B.bamCount(p, gr)
    assert e.value.code == -4
    os.remove(p + ".csi")
    B.lib().bsg_shutdown()
    with pytest.raises(B.BamsignalsError) as e:
        B.bamCount(p, gr)
    assert e.value.code == -2 and "BAM indexing file is not available" in str(e.value)


# ---------------------------------------------------------------------------------------------------------------------
# GPU parity
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("min_shift,depth", [(14, 5), (12, 6), (16, 4)])
@pytest.mark.parametrize("gpu_inflate", [1, -1])
def test_gpu_csi_parity(tmp_path, min_shift, depth, gpu_inflate):
    p = str(tmp_path / "v.bam")
    E.write_variety(p, block_payload=2500, cut_mid_record=True, index="csi", min_shift=min_shift, depth=depth)
    gr = E.variety_regions()
    opts = B.default_opts(gpu_inflate=gpu_inflate)
    for kw in (dict(ss=True, shift=33), dict(paired_end="midpoint", tlenFilter=(0, 500), ss=True), dict(filteredFlag=1024, mapqual=20)):
        assert np.array_equal(B.bamCount(p, gr, opts=opts, **kw), O.bamCount(p, gr, mode=O.SCAN, **kw)), kw
        assert np.array_equal(flat(B.bamProfile(p, gr, binsize=7, opts=opts, **kw)), flat(O.bamProfile(p, gr, binsize=7, **kw))), kw
    assert np.array_equal(flat(B.bamCoverage(p, gr, paired_end="extend", opts=opts)), flat(O.bamCoverage(p, gr, paired_end="extend")))


@pytest.mark.gpu
def test_gpu_long_contig(tmp_path):
    """coordinates beyond 2^29 bp (CSI depth 6): count, profile and coverage against the oracle's index-free scan"""
    p = str(tmp_path / "big.bam")
    W.write_bam(p, BIG_REFS, big_reads(), index="csi", min_shift=14, depth=6, block_payload=4000)
    gr = big_regions()
    assert np.array_equal(B.bamCount(p, gr, ss=True, shift=40), O.bamCount(p, gr, ss=True, shift=40, mode=O.SCAN))
    assert np.array_equal(flat(B.bamProfile(p, gr, binsize=50, ss=True)), flat(O.bamProfile(p, gr, binsize=50, ss=True, mode=O.SCAN)))
    assert np.array_equal(flat(B.bamCoverage(p, gr)), flat(O.bamCoverage(p, gr, mode=O.SCAN)))
    assert int(B.bamCount(p, gr).sum()) > 100


def test_cram_is_rejected_by_name(tmp_path):
    p = str(tmp_path / "x.cram")
    open(p, "wb").write(b"CRAM\x03\x00" + bytes(64))
    with pytest.raises(B.BamsignalsError) as e:
        B.bamCount(p, E.variety_regions())
    assert e.value.code == -4 and "CRAM input is not supported" in str(e.value)
