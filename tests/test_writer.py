"""bsg_write_sam_as_bam_and_index = the reference's writeSamAsBamAndIndex (src/bamsignals.cpp:496-534), used by its
tests to turn SAM text into the fixture BAM (tests/testthat/utils.R:48-120).  Host-only, so everything here runs on
CPU: the reference's own reads, written as SAM the way utils.R does and pushed through our writer, must give a BAM that
decodes to the same records - and the same counts - as the fixture BAM the reference itself produced."""
import os
import struct
import zlib

import numpy as np
import pytest

import bamsignals_b200 as B
import oracle_api as O
import spec_r

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def reads_to_sam(path, reads):
    """tests/testthat/utils.R:63-111: references sorted by name, reads sorted by (rname, pos), CIGAR '<qwidth>M',
    RNEXT '=', SEQ and QUAL '*'"""
    lv = [str(s) for s in reads["rname_levels"]]
    rn = np.array([lv[i] for i in reads["rname"]])
    reflen = {s: int((reads["pos"][rn == s] + reads["qwidth"][rn == s]).max()) + 1 for s in sorted(set(rn))}
    pnext = reads["pnext"] if "pnext" in reads else np.zeros(len(rn), dtype=np.int64)
    order = np.lexsort((reads["pos"], rn))
    with open(path, "w") as f:
        f.write("@HD\tVN:1.0\tSO:coordinate\n")
        for s in sorted(reflen):
            f.write(f"@SQ\tSN:{s}\tLN:{reflen[s]}\n")
        for k, i in enumerate(order):
            f.write(f"{k}\t{int(reads['flag'][i])}\t{rn[i]}\t{int(reads['pos'][i])}\t{int(reads['mapq'][i])}\t"
                    f"{int(reads['qwidth'][i])}M\t=\t{int(pnext[i])}\t{int(reads['isize'][i])}\t*\t*\n")


def bam_records(path):
    """minimal BAM reader: yields the raw bytes of every record"""
    raw, data, p = b"", open(path, "rb").read(), 0
    while p < len(data):
        bsize = struct.unpack_from("<H", data, p + 16)[0] + 1
        raw += zlib.decompress(data[p + 18:p + bsize - 8], -15)
        p += bsize
    assert raw[:4] == b"BAM\1"
    l_text = struct.unpack_from("<i", raw, 4)[0]
    q = 8 + l_text
    n_ref = struct.unpack_from("<i", raw, q)[0]
    q += 4
    for _ in range(n_ref):
        q += 8 + struct.unpack_from("<i", raw, q)[0]
    recs = []
    while q < len(raw):
        bs = struct.unpack_from("<i", raw, q)[0]
        recs.append(raw[q + 4:q + 4 + bs])
        q += 4 + bs
    return recs


def test_fixture_reads_roundtrip(tmp_path, fixture_bam):
    reads = dict(np.load(os.path.join(ROOT, "tests", "golden", "randomReads.npz"), allow_pickle=True))
    sam, bam = str(tmp_path / "r.sam"), str(tmp_path / "r.bam")
    reads_to_sam(sam, reads)
    assert B.writeSamAsBamAndIndex(sam, bam) is True
    assert os.path.exists(bam + ".bai")
    ours, ref = O.dump_reads(bam), O.dump_reads(fixture_bam)
    for k in ("tid", "pos", "endpos", "flag", "mapq", "tlen"):
        a, b = ours[k], ref[k]
        # reads at the same (tid, pos) may come in a different order: compare as sorted multisets per key
        key_a = np.lexsort((ours["tlen"], ours["mapq"], ours["flag"], ours["endpos"], ours["pos"], ours["tid"]))
        key_b = np.lexsort((ref["tlen"], ref["mapq"], ref["flag"], ref["endpos"], ref["pos"], ref["tid"]))
        assert np.array_equal(a[key_a], b[key_b]), k
    # and the counting results through the index are those of the reference's own fixture
    g = spec_r.test_regions(seed=5, n=60)
    gr = B.GRanges([["chr1", "chr2", "chr3"][i] for i in g["rname"]], g["start"], g["width"], g["strand"])
    for kw in (dict(ss=True, shift=50), dict(paired_end="midpoint", tlenFilter=(0, 300)), dict(mapqual=20, filteredFlag=16)):
        assert np.array_equal(O.bamCount(bam, gr, **kw), O.bamCount(fixture_bam, gr, **kw)), kw
        assert np.array_equal(O.bamCount(bam, gr, **kw), O.bamCount(bam, gr, mode=O.SCAN, **kw)), kw
    cov_a, cov_b = O.bamCoverage(bam, gr, paired_end="extend"), O.bamCoverage(fixture_bam, gr, paired_end="extend")
    assert all(np.array_equal(x, y) for x, y in zip(cov_a.as_list(), cov_b.as_list()))


def test_record_encoding(tmp_path):
    """SEQ / QUAL / CIGAR / optional fields as the SAM spec (4.2) lays them out"""
    sam, bam = str(tmp_path / "e.sam"), str(tmp_path / "e.bam")
    open(sam, "w").write(
        "@HD\tVN:1.6\tSO:coordinate\n@SQ\tSN:c1\tLN:5000\n@SQ\tSN:c2\tLN:70000\n@PG\tID:x\n"
        "a\t99\tc1\t101\t30\t5S10M2D3I10M\t=\t301\t250\tACGTNacgtRYKMACGTACGTACGTACG\tIIIIIIIIIIIIIIIIIIIIIIIIIIII\tNM:i:3\tXS:Z:hello\tXB:B:c,1,-2,3\tXF:f:1.5\tXA:A:q\tXI:i:-70000\tXU:i:70000\tXH:H:1AE3\n"
        "b\t16\tc2\t65000\t0\t*\t*\t0\t0\t*\t*\n"
        "u\t4\t*\t0\t0\t*\t*\t0\t0\tACG\t*\n")
    B.writeSamAsBamAndIndex(sam, bam)
    r = bam_records(bam)
    assert len(r) == 3
    tid, pos, l_name, mapq, bin_, n_cig, flag, l_seq, ntid, npos, tlen = struct.unpack_from("<iiBBHHHiiii", r[0], 0)
    assert (tid, pos, l_name, mapq, n_cig, flag, l_seq, ntid, npos, tlen) == (0, 100, 2, 30, 5, 99, 28, 0, 300, 250)
    assert bin_ == 4681                                   # reg2bin(100, 122)
    q = 32
    assert r[0][q:q + 2] == b"a\0"
    q += 2
    assert list(struct.unpack_from("<5I", r[0], q)) == [5 << 4 | 4, 10 << 4 | 0, 2 << 4 | 2, 3 << 4 | 1, 10 << 4 | 0]
    q += 20
    seq = "".join("=ACMGRSVTWYHKDBN"[b >> 4] + "=ACMGRSVTWYHKDBN"[b & 15] for b in r[0][q:q + 14])
    assert seq == "ACGTNACGTRYKMACGTACGTACGTACG"
    q += 14
    assert r[0][q:q + 28] == bytes([ord("I") - 33] * 28)
    q += 28
    assert r[0][q:] == (b"NMC\x03" + b"XSZhello\0" + b"XBBc" + struct.pack("<i3b", 3, 1, -2, 3) + b"XFf" + struct.pack("<f", 1.5) +
                        b"XAAq" + b"XIi" + struct.pack("<i", -70000) + b"XUI" + struct.pack("<I", 70000) + b"XHH1AE3\0")
    tid, pos, _, _, bin_, n_cig, flag, l_seq, ntid, npos, tlen = struct.unpack_from("<iiBBHHHiiii", r[1], 0)
    assert (tid, pos, n_cig, flag, l_seq, ntid, npos) == (1, 64999, 0, 16, 0, -1, -1)
    tid, pos, _, _, bin_, _, flag, l_seq, _, _, _ = struct.unpack_from("<iiBBHHHiiii", r[2], 0)
    assert (tid, pos, bin_, flag, l_seq) == (-1, -1, 4680, 4, 3)
    assert r[2][32 + 2 + 2:] == b"\xff\xff\xff"            # QUAL '*'
    assert O.header(bam) == (["c1", "c2"], [5000, 70000])


def test_writer_errors(tmp_path):
    with pytest.raises(B.BamsignalsError) as e:
        B.writeSamAsBamAndIndex(str(tmp_path / "nope.sam"), str(tmp_path / "x.bam"))
    assert e.value.code == -1 and "Fail to open SAM file" in str(e.value)              # src/bamsignals.cpp:507
    sam = str(tmp_path / "u.sam")
    open(sam, "w").write("@SQ\tSN:c\tLN:1000\nr1\t0\tc\t500\t9\t10M\t*\t0\t0\t*\t*\nr2\t0\tc\t400\t9\t10M\t*\t0\t0\t*\t*\n")
    with pytest.raises(B.BamsignalsError) as e:
        B.writeSamAsBamAndIndex(sam, str(tmp_path / "u.bam"))
    assert e.value.code == -5
    for bad in ("r1\t0\tc\t500\t9\t10Q\t*\t0\t0\t*\t*\n", "r1\t0\tzzz\t500\t9\t10M\t*\t0\t0\t*\t*\n", "r1\t0\tc\t500\n",
                "r1\tx\tc\t500\t9\t10M\t*\t0\t0\t*\t*\n", "r1\t0\tc\t500\t9\t4M\t*\t0\t0\tACGT\tII\n"):
        open(sam, "w").write("@SQ\tSN:c\tLN:1000\n" + bad)
        with pytest.raises(B.BamsignalsError) as e:
            B.writeSamAsBamAndIndex(sam, str(tmp_path / "b.bam"))
        assert e.value.code == -4, bad
    with pytest.raises(B.BamsignalsError) as e:
        B.writeSamAsBamAndIndex(sam, str(tmp_path / "no_such_dir" / "b.bam"))
    assert e.value.code == -1


def test_many_blocks_and_long_records(tmp_path):
    """records that fill many BGZF blocks (and one longer than a block) keep index offsets consistent:
    indexed access == scan on every region"""
    rng = np.random.default_rng(3)
    sam, bam = str(tmp_path / "m.sam"), str(tmp_path / "m.bam")
    pos = np.sort(rng.integers(1, 400000, 6000))
    with open(sam, "w") as f:
        f.write("@SQ\tSN:c\tLN:500000\n")
        for i, p in enumerate(pos):
            l = 70000 if i == 3000 else int(rng.integers(30, 200))
            f.write(f"q{i}\t{int(rng.choice([0, 16]))}\tc\t{p}\t{int(rng.integers(0, 60))}\t{l}M\t*\t0\t0\t{'A' * l}\t{'I' * l}\n")
    B.writeSamAsBamAndIndex(sam, bam)
    gr = B.GRanges(["c"] * 80, rng.integers(1, 390000, 80), rng.integers(1, 9000, 80), rng.choice(["+", "-", "*"], 80).tolist())
    assert np.array_equal(O.bamCount(bam, gr, ss=True), O.bamCount(bam, gr, ss=True, mode=O.SCAN))
    assert int(O.bamCount(bam, gr).sum()) > 1000


@pytest.mark.gpu
def test_gpu_counts_on_written_bam(tmp_path):
    reads = dict(np.load(os.path.join(ROOT, "tests", "golden", "randomReads.npz"), allow_pickle=True))
    sam, bam = str(tmp_path / "r.sam"), str(tmp_path / "r.bam")
    reads_to_sam(sam, reads)
    B.writeSamAsBamAndIndex(sam, bam)
    g = spec_r.test_regions(seed=8, n=50)
    gr = B.GRanges([["chr1", "chr2", "chr3"][i] for i in g["rname"]], g["start"], g["width"], g["strand"])
    assert np.array_equal(B.bamCount(bam, gr, ss=True, shift=20), O.bamCount(bam, gr, ss=True, shift=20))
    a, b = B.bamCoverage(bam, gr, paired_end="extend"), O.bamCoverage(bam, gr, paired_end="extend")
    assert all(np.array_equal(x, y) for x, y in zip(a.as_list(), b.as_list()))
