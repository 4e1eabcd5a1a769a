"""Pins the CPU oracle (oracle/bsg_oracle.cpp, a restatement of src/bamsignals.cpp) against the reference's own test
oracle (tests/testthat/utils.R restated in tests/spec_r.py) on the reference's own fixtures, over the full parameter
sweep of tests/testthat/test_methods.R:33-104 — and against the committed golden vectors.  CPU only."""
import os

import numpy as np
import pytest

import oracle_api as O
import spec_r
from bamsignals_b200.api import GRanges

LEVELS = ["chr1", "chr2", "chr3"]


def to_gr(genes):
    return GRanges([LEVELS[i] for i in genes["rname"]], genes["start"], genes["width"], genes["strand"])


@pytest.fixture(scope="module")
def reads():
    return spec_r.load_reads()


@pytest.fixture(scope="module")
def expected():
    z = np.load(os.path.join(spec_r.GOLDEN, "expected_fixture.npz"))
    return {k: z[k] for k in z.files}


REGION_SETS = {"rand7": spec_r.test_regions(), "annot": spec_r.annot_regions(),
               "rand99": spec_r.test_regions(seed=99, n=40)}


def test_fixture_facts(fixture_bam):
    """99,000 records, three contigs; the RData table is the same reads."""
    names, lens = O.header(fixture_bam)
    assert names == LEVELS and lens == [10237, 10279, 10238]
    d = O.dump_reads(fixture_bam)
    assert len(d["pos"]) == 99000
    r = spec_r.load_reads()
    order = np.lexsort((r["flag"], r["pos"], r["rname"]))
    o2 = np.lexsort((d["flag"], d["pos"], d["tid"]))
    assert (r["rname"][order] == d["tid"][o2]).all()
    assert (r["pos"][order] - 1 == d["pos"][o2]).all()
    assert (r["flag"][order] == d["flag"][o2]).all()


@pytest.mark.parametrize("tag", list(REGION_SETS))
@pytest.mark.parametrize("mode", [O.INDEXED, O.SCAN, O.BRUTE])
def test_pileup_sweep(fixture_bam, reads, expected, tag, mode):
    genes = REGION_SETS[tag]
    gr = to_gr(genes)
    for case in spec_r.sweep_pileup():
        kw = dict(mapqual=case["mapqual"], shift=case["shift"], ss=case["ss"], paired_end=case["paired_end"],
                  tlenFilter=case["tlenFilter"])
        got = O.bamCount(fixture_bam, gr, mode=mode, **kw)
        skw = {k: v for k, v in case.items() if k != "ss"}
        want = spec_r.countR(reads, genes, ss=case["ss"], **skw)
        assert np.array_equal(got, want), ("count", case)
        gotp = O.bamProfile(fixture_bam, gr, mode=mode, **kw).as_list()
        wantp = spec_r.profileR(reads, genes, ss=case["ss"], **skw)
        for a, b in zip(gotp, wantp):
            assert a.shape == b.shape and np.array_equal(a, b), ("profile", case)
        if tag in ("rand7", "annot"):
            assert np.array_equal(got.ravel(order="F"), expected[f"{tag}|" + spec_r.case_key("count", case)])
            assert np.array_equal(np.concatenate([p.ravel(order="F") for p in gotp]),
                                  expected[f"{tag}|" + spec_r.case_key("profile", case)])


@pytest.mark.parametrize("tag", list(REGION_SETS))
@pytest.mark.parametrize("mode", [O.INDEXED, O.SCAN, O.BRUTE])
def test_coverage_sweep(fixture_bam, reads, expected, tag, mode):
    genes = REGION_SETS[tag]
    gr = to_gr(genes)
    for case in spec_r.sweep_coverage():
        got = O.bamCoverage(fixture_bam, gr, mode=mode, mapqual=case["mapqual"], paired_end=case["paired_end"],
                            tlenFilter=case["tlenFilter"]).as_list()
        want = spec_r.coverageR(reads, genes, **case)
        for a, b in zip(got, want):
            assert np.array_equal(a, b), case
        if tag in ("rand7", "annot"):
            assert np.array_equal(np.concatenate(got), expected[f"{tag}|" + spec_r.case_key("coverage", case)])


def test_filtered_flag_16(fixture_bam, reads):
    """test_methods.R:85-104: filteredFlag=16 on all-'+' regions equals the sense row of the ss count."""
    genes = dict(REGION_SETS["rand7"])
    genes["strand"] = ["+"] * len(genes["start"])
    gr = to_gr(genes)
    for case in spec_r.sweep_pileup():
        if case["ss"]:
            continue
        skw = {k: v for k, v in case.items() if k != "ss"}
        want = spec_r.countR(reads, genes, ss=True, **skw)[0]
        got = O.bamCount(fixture_bam, gr, mapqual=case["mapqual"], shift=case["shift"], ss=False,
                         paired_end=case["paired_end"], tlenFilter=case["tlenFilter"], filteredFlag=16)
        assert np.array_equal(got, want), case


def test_survey_known_answers(fixture_bam):
    """Known answers recorded in SURVEY.md App. C (derived there by an independent throw-away restatement)."""
    a = spec_r.load_annot()
    gr = GRanges(a["seqnames"], a["start"], a["width"], a["strand"], seqlevels=a["seqlevels"])
    c = O.bamCount(fixture_bam, gr)
    assert c.tolist() == [2570, 2449, 2343, 2001, 2129, 1937, 2389, 2129, 2117, 2418, 2462, 2518, 2453, 2135, 2299,
                          2533, 2183, 2402, 2403, 2140]
    css = O.bamCount(fixture_bam, gr, ss=True)
    assert css[:, :3].T.tolist() == [[1269, 1301], [1211, 1238], [1194, 1149]]
    c2 = O.bamCount(fixture_bam, gr, shift=75, mapqual=20, paired_end="midpoint", filteredFlag=1024)
    assert c2.tolist() == [595, 594, 569, 474, 518, 479, 560, 487, 491, 573, 598, 597, 610, 496, 561, 616, 516, 569,
                           575, 523]
    cov = O.bamCoverage(fixture_bam, gr, paired_end="extend")
    assert cov[0][:10].tolist() == [234, 236, 237, 242, 243, 244, 246, 247, 250, 253] and int(cov[0].sum()) == 197141
    p = O.bamProfile(fixture_bam, gr, binsize=20, ss=True)
    assert p[0].shape == (2, 31) and p[0][:, :6].T.tolist() == [[68, 11], [77, 23], [73, 17], [82, 33], [62, 37], [47, 42]]
    tot = lambda ff: int(O.bamCount(fixture_bam, gr, filteredFlag=ff).sum())
    assert (tot(-1), tot(1024), tot(1040), tot(0)) == (46010, 43383, 44777, 0)


def test_errors(fixture_bam, tmp_path):
    gr = GRanges(["chrZ"], [1], [10])
    with pytest.raises(O.OracleError, match="chromosome chrZ not present in the bam file"):
        O.bamCount(fixture_bam, gr)
    with pytest.raises(O.OracleError, match="Fail to open BAM file"):
        O.bamCount(str(tmp_path / "nope.bam"), GRanges(["chr1"], [1], [10]))
    import shutil
    shutil.copyfile(fixture_bam, tmp_path / "noidx.bam")
    with pytest.raises(O.OracleError, match="BAM indexing file is not available"):
        O.bamCount(str(tmp_path / "noidx.bam"), GRanges(["chr1"], [1], [10]))


def test_threads_same(fixture_bam):
    genes = spec_r.test_regions(seed=5, n=200)
    gr = to_gr(genes)
    a = O.bamProfile(fixture_bam, gr, ss=True, shift=10, nthreads=1, maxgap=0).as_list()
    b = O.bamProfile(fixture_bam, gr, ss=True, shift=10, nthreads=4, maxgap=0).as_list()
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
