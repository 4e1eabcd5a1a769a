"""The parity pin that EXECUTES reference code: oracle/_ref/libbamsignals_ref.so is the reference's own
src/bamsignals.cpp (pileup_core, coverage_core, parseRegions, allocateList, overlapAndPileup, Pileupper, Coverager,
cumsum: src/bamsignals.cpp:92-494) compiled UNCHANGED against stand-ins for <Rcpp.h> and <htslib/*.h>
(oracle/Makefile `ref`, oracle/ref_compat/).  Here the restated oracle (impl="port") must equal it (impl="ref") bit for
bit on the reference's fixture over the whole sweep of tests/testthat/test_methods.R:33-104, on the committed golden
vectors, on the reference's R test oracle restated in numpy (tests/spec_r.py), and on the 60 randomised scenarios
(all CIGAR operators, arbitrary flags, odd regions).  CPU only.  What _ref cannot pin is htslib itself (an un-vendored
dependency): bam_endpos and the iterator come from the stand-in, written from the SAM specification."""
import os
import warnings

import numpy as np
import pytest

import oracle_api as O
import random_cases as RC
import spec_r
from bamsignals_b200.api import GRanges

pytestmark = pytest.mark.skipif(not O.ref_available(), reason="oracle/_ref not built (needs /root/reference at build time)")

LEVELS = ["chr1", "chr2", "chr3"]
REGION_SETS = {"rand7": spec_r.test_regions(), "annot": spec_r.annot_regions(), "rand99": spec_r.test_regions(seed=99, n=40)}


def to_gr(genes):
    return GRanges([LEVELS[i] for i in genes["rname"]], genes["start"], genes["width"], genes["strand"])


@pytest.fixture(scope="module")
def reads():
    return spec_r.load_reads()


@pytest.fixture(scope="module")
def expected():
    z = np.load(os.path.join(spec_r.GOLDEN, "expected_fixture.npz"))
    return {k: z[k] for k in z.files}


def test_ref_is_the_reference_source():
    """The recipe compiles the file where it lies under /root/reference: no copy of it exists in this repository."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    mk = open(os.path.join(root, "oracle", "Makefile")).read()
    assert "$(REFERENCE)/src/bamsignals.cpp" in mk
    for dirpath, _, files in os.walk(root):
        if any(part in dirpath for part in (".git", "gpurun_out")):
            continue
        for f in files:
            if f.endswith((".cpp", ".cu", ".h", ".hpp", ".cuh")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "this function overlaps each read with the regions it might fall into" not in txt, f


@pytest.mark.parametrize("tag", list(REGION_SETS))
def test_pileup_sweep_ref(fixture_bam, reads, expected, tag):
    """48 bamCount + 48 bamProfile cases (test_methods.R:33-67): reference == port == R test oracle == golden."""
    genes = REGION_SETS[tag]
    gr = to_gr(genes)
    for case in spec_r.sweep_pileup():
        kw = dict(mapqual=case["mapqual"], shift=case["shift"], ss=case["ss"], paired_end=case["paired_end"],
                  tlenFilter=case["tlenFilter"])
        ref = O.bamCount(fixture_bam, gr, impl="ref", **kw)
        assert np.array_equal(ref, O.bamCount(fixture_bam, gr, **kw)), ("count", case)
        skw = {k: v for k, v in case.items() if k != "ss"}
        assert np.array_equal(ref, spec_r.countR(reads, genes, ss=case["ss"], **skw)), ("count vs R oracle", case)
        refp = O.bamProfile(fixture_bam, gr, impl="ref", **kw).as_list()
        portp = O.bamProfile(fixture_bam, gr, **kw).as_list()
        wantp = spec_r.profileR(reads, genes, ss=case["ss"], **skw)
        assert len(refp) == len(portp) == len(wantp)
        for a, b, c in zip(refp, portp, wantp):
            assert a.shape == b.shape == c.shape and np.array_equal(a, b) and np.array_equal(a, c), ("profile", case)
        if tag in ("rand7", "annot"):
            assert np.array_equal(ref.ravel(order="F"), expected[f"{tag}|" + spec_r.case_key("count", case)])
            assert np.array_equal(np.concatenate([p.ravel(order="F") for p in refp]),
                                  expected[f"{tag}|" + spec_r.case_key("profile", case)])


@pytest.mark.parametrize("tag", list(REGION_SETS))
def test_coverage_sweep_ref(fixture_bam, reads, expected, tag):
    """8 bamCoverage cases (test_methods.R:69-83)."""
    genes = REGION_SETS[tag]
    gr = to_gr(genes)
    for case in spec_r.sweep_coverage():
        kw = dict(mapqual=case["mapqual"], paired_end=case["paired_end"], tlenFilter=case["tlenFilter"])
        ref = O.bamCoverage(fixture_bam, gr, impl="ref", **kw).as_list()
        port = O.bamCoverage(fixture_bam, gr, **kw).as_list()
        want = spec_r.coverageR(reads, genes, **case)
        for a, b, c in zip(ref, port, want):
            assert np.array_equal(a, b) and np.array_equal(a, c), case
        if tag in ("rand7", "annot"):
            assert np.array_equal(np.concatenate(ref), expected[f"{tag}|" + spec_r.case_key("coverage", case)])


def test_filtered_flag_ref(fixture_bam, reads):
    """24 flag cases (test_methods.R:85-104) + the binsize variants of bamProfile."""
    genes = dict(REGION_SETS["rand7"])
    genes["strand"] = ["+"] * len(genes["start"])
    gr = to_gr(genes)
    for case in spec_r.sweep_pileup():
        if case["ss"]:
            continue
        kw = dict(mapqual=case["mapqual"], shift=case["shift"], ss=False, paired_end=case["paired_end"],
                  tlenFilter=case["tlenFilter"], filteredFlag=16)
        ref = O.bamCount(fixture_bam, gr, impl="ref", **kw)
        assert np.array_equal(ref, O.bamCount(fixture_bam, gr, **kw)), case
        skw = {k: v for k, v in case.items() if k != "ss"}
        assert np.array_equal(ref, spec_r.countR(reads, genes, ss=True, **skw)[0]), case
    gr = to_gr(REGION_SETS["annot"])
    for bs in (2, 3, 20, 200, 5000):
        for ss in (False, True):
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                a = O.bamProfile(fixture_bam, gr, binsize=bs, ss=ss, shift=-33, filteredFlag=1024, impl="ref").as_list()
                b = O.bamProfile(fixture_bam, gr, binsize=bs, ss=ss, shift=-33, filteredFlag=1024).as_list()
            assert all(x.shape == y.shape and np.array_equal(x, y) for x, y in zip(a, b)), (bs, ss)


def test_known_answers_ref(fixture_bam):
    """SURVEY App. C's known answers, now produced by reference code."""
    a = spec_r.load_annot()
    gr = GRanges(a["seqnames"], a["start"], a["width"], a["strand"], seqlevels=a["seqlevels"])
    c = O.bamCount(fixture_bam, gr, impl="ref")
    assert c.tolist() == [2570, 2449, 2343, 2001, 2129, 1937, 2389, 2129, 2117, 2418, 2462, 2518, 2453, 2135, 2299,
                          2533, 2183, 2402, 2403, 2140]
    cov = O.bamCoverage(fixture_bam, gr, paired_end="extend", impl="ref")
    assert cov[0][:10].tolist() == [234, 236, 237, 242, 243, 244, 246, 247, 250, 253] and int(cov[0].sum()) == 197141
    tot = lambda ff: int(O.bamCount(fixture_bam, gr, filteredFlag=ff, impl="ref").sum())
    assert (tot(-1), tot(1024), tot(1040), tot(0)) == (46010, 43383, 44777, 0)


def test_errors_ref(fixture_bam, tmp_path):
    """The reference's own error strings (src/bamsignals.cpp:119,204,209,243) out of the reference's own code."""
    with pytest.raises(O.OracleError, match="chromosome chrZ not present in the bam file"):
        O.bamCount(fixture_bam, GRanges(["chrZ"], [1], [10]), impl="ref")
    with pytest.raises(O.OracleError, match="Fail to open BAM file"):
        O.bamCount(str(tmp_path / "nope.bam"), GRanges(["chr1"], [1], [10]), impl="ref")
    import shutil
    shutil.copyfile(fixture_bam, tmp_path / "noidx.bam")
    with pytest.raises(O.OracleError, match="BAM indexing file is not available"):
        O.bamCount(str(tmp_path / "noidx.bam"), GRanges(["chr1"], [1], [10]), impl="ref")
    with pytest.raises(O.OracleError, match="negative 'ext' values don't make sense"):
        O.pileup_core(fixture_bam, GRanges(["chr1"], [1], [10]), (0, -5), pe_mid=True, impl="ref")


def test_threads_and_maxgap_ref(fixture_bam):
    gr = to_gr(spec_r.test_regions(seed=5, n=200))
    a = O.bamProfile(fixture_bam, gr, ss=True, shift=10, impl="ref").as_list()
    for kw in (dict(nthreads=4), dict(maxgap=0), dict(nthreads=3, maxgap=100000)):
        b = O.bamProfile(fixture_bam, gr, ss=True, shift=10, impl="ref", **kw).as_list()
        assert all(np.array_equal(x, y) for x, y in zip(a, b)), kw


@pytest.mark.parametrize("seed", range(60))
def test_random_scenarios_ref(tmp_path, seed):
    """Everything the reference's fixture leaves open (CIGAR operators, flag-0x4 reads, arbitrary flag masks, odd regions,
    stored / fixed / dynamic DEFLATE blocks): reference code == port == the numpy statement of SURVEY App. A."""
    sc = RC.scenario(seed)
    bam = RC.write(sc, str(tmp_path / "r.bam"))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = RC.flat(getattr(O, sc["fn"])(bam, sc["gr"], impl="ref", **sc["kw"]), sc["fn"])
        port = RC.flat(getattr(O, sc["fn"])(bam, sc["gr"], **sc["kw"]), sc["fn"])
    assert ref.shape == port.shape and np.array_equal(ref, port), (seed, sc["fn"], sc["kw"])
    assert np.array_equal(ref, RC.spec_counts(sc)), (seed, "numpy spec")
