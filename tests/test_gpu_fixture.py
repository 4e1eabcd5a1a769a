"""GPU parity on the reference's own toy BAM: the CUDA path (through the C ABI) against the CPU oracle and the
committed golden vectors, over the parameter sweep of the reference's test-suite (tests/testthat/test_methods.R)
plus what that suite leaves out (binsize > 1, '*' strand, flag quirks, tiles, empty and odd regions)."""
import os

import numpy as np
import pytest

import bamsignals_b200 as B
import oracle_api as O
import spec_r

pytestmark = pytest.mark.gpu
LEVELS = ["chr1", "chr2", "chr3"]


def to_gr(genes):
    return B.GRanges([LEVELS[i] for i in genes["rname"]], genes["start"], genes["width"], genes["strand"])


def same_list(a, b):
    assert len(a) == len(b)
    for x, y in zip(a, b):
        assert x.shape == y.shape and np.array_equal(x, y)


@pytest.fixture(scope="module")
def expected():
    z = np.load(os.path.join(spec_r.GOLDEN, "expected_fixture.npz"))
    return {k: z[k] for k in z.files}


@pytest.mark.parametrize("tag", ["rand7", "annot"])
def test_reference_sweep_vs_golden_and_oracle(fixture_bam, expected, tag):
    genes = spec_r.test_regions() if tag == "rand7" else spec_r.annot_regions()
    gr = to_gr(genes)
    for case in spec_r.sweep_pileup():
        kw = dict(mapqual=case["mapqual"], shift=case["shift"], ss=case["ss"], paired_end=case["paired_end"],
                  tlenFilter=case["tlenFilter"])
        got = B.bamCount(fixture_bam, gr, **kw)
        assert np.array_equal(got.ravel(order="F"), expected[f"{tag}|" + spec_r.case_key("count", case)]), case
        assert np.array_equal(got, O.bamCount(fixture_bam, gr, **kw))
        gp = B.bamProfile(fixture_bam, gr, **kw).as_list()
        assert np.array_equal(np.concatenate([p.ravel(order="F") for p in gp]),
                              expected[f"{tag}|" + spec_r.case_key("profile", case)]), case
    for case in spec_r.sweep_coverage():
        gc = B.bamCoverage(fixture_bam, gr, mapqual=case["mapqual"], paired_end=case["paired_end"],
                           tlenFilter=case["tlenFilter"]).as_list()
        assert np.array_equal(np.concatenate(gc), expected[f"{tag}|" + spec_r.case_key("coverage", case)]), case


def test_config1_bamcount_100x1kb(fixture_bam):
    """BASELINE config C1: bamCount over 100 random 1 kb regions, mapqual=0, ss=FALSE."""
    rng = np.random.default_rng(1)
    gr = B.GRanges(rng.choice(LEVELS, 100).tolist(), rng.integers(1, 9001, 100), [1000] * 100,
                   rng.choice(["+", "-", "*"], 100).tolist())
    got = B.bamCount(fixture_bam, gr, mapqual=0, ss=False)
    assert got.shape == (100,) and np.array_equal(got, O.bamCount(fixture_bam, gr, mapqual=0, ss=False))
    t = B.timings()
    assert t["records"] > 0 and t["n_launches"] >= 3                      # decode + filter, join, count (a tiny job inflates on the host)
    got = B.bamCount(fixture_bam, gr, mapqual=0, ss=False, opts=B.default_opts(gpu_inflate=1))
    assert np.array_equal(got, O.bamCount(fixture_bam, gr, mapqual=0, ss=False)) and B.timings()["n_launches"] >= 7


@pytest.mark.parametrize("binsize", [2, 3, 20, 200, 5000])
@pytest.mark.parametrize("ss", [False, True])
def test_binsize_gt1_and_star_strand(fixture_bam, binsize, ss):
    rng = np.random.default_rng(binsize)
    n = 60
    gr = B.GRanges(rng.choice(LEVELS, n).tolist(), rng.integers(1, 9000, n), rng.integers(1, 3000, n),
                   rng.choice(["+", "-", "*"], n).tolist())
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        got = B.bamProfile(fixture_bam, gr, binsize=binsize, ss=ss, shift=37, mapqual=10).as_list()
    want = O.bamProfile(fixture_bam, gr, binsize=binsize, ss=ss, shift=37, mapqual=10).as_list()
    same_list(got, want)


@pytest.mark.parametrize("ff", [-1, 0, 16, 1024, 1040, 83, 2048, 65535])
def test_filtered_flag_quirks(fixture_bam, ff):
    """only reads carrying ALL filteredFlag bits are dropped; 0 drops everything; -1 nothing."""
    gr = to_gr(spec_r.annot_regions())
    for ss in (False, True):
        assert np.array_equal(B.bamCount(fixture_bam, gr, ss=ss, filteredFlag=ff), O.bamCount(fixture_bam, gr, ss=ss, filteredFlag=ff))
    same_list(B.bamCoverage(fixture_bam, gr, filteredFlag=ff).as_list(), O.bamCoverage(fixture_bam, gr, filteredFlag=ff).as_list())
    if ff == 0:
        assert int(B.bamCount(fixture_bam, gr, filteredFlag=0).sum()) == 0


def test_whole_chromosome_tiles(fixture_bam):
    """Regions wider than one tile (8192 ints): profile ss (2 x 10237), coverage, both strands."""
    gr = B.GRanges(LEVELS + LEVELS, [1] * 6, [10237, 10279, 10238] * 2, ["+", "+", "*", "-", "-", "-"])
    for kw in (dict(ss=True), dict(ss=False, shift=-50), dict(ss=True, binsize=3, paired_end="midpoint")):
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            same_list(B.bamProfile(fixture_bam, gr, **kw).as_list(), O.bamProfile(fixture_bam, gr, **kw).as_list())
    for kw in (dict(), dict(paired_end="extend", tlenFilter=(0, 400)), dict(mapqual=30)):
        same_list(B.bamCoverage(fixture_bam, gr, **kw).as_list(), O.bamCoverage(fixture_bam, gr, **kw).as_list())
    # every output element of the wide profile is accounted for: total equals the bamCount of the same regions
    prof = B.bamProfile(fixture_bam, gr, ss=False).as_list()
    assert [int(p.sum()) for p in prof] == B.bamCount(fixture_bam, gr).tolist()


def test_odd_regions(fixture_bam):
    """Zero-width, width-1, off-the-end, duplicated and heavily overlapping regions; input order is preserved."""
    gr = B.GRanges(["chr2", "chr1", "chr1", "chr3", "chr1", "chr1", "chr2", "chr2"],
                   [500, 1, 1, 10000, 5000, 5000, 20000, 500],
                   [0, 1, 30000, 5000, 700, 700, 100, 0],
                   ["+", "-", "*", "-", "+", "-", "+", "-"])
    for ss in (False, True):
        assert np.array_equal(B.bamCount(fixture_bam, gr, ss=ss), O.bamCount(fixture_bam, gr, ss=ss))
        same_list(B.bamProfile(fixture_bam, gr, ss=ss).as_list(), O.bamProfile(fixture_bam, gr, ss=ss).as_list())
    same_list(B.bamCoverage(fixture_bam, gr).as_list(), O.bamCoverage(fixture_bam, gr).as_list())
    empty = B.GRanges([], [], [], [])
    assert B.bamCount(fixture_bam, empty).shape == (0,)
    assert len(B.bamProfile(fixture_bam, empty)) == 0 and len(B.bamCoverage(fixture_bam, empty)) == 0


def test_large_shift_and_negative_shift(fixture_bam):
    gr = to_gr(spec_r.test_regions(seed=3, n=50))
    for shift in (-400, -1, 1, 999, 5000):
        same_list(B.bamProfile(fixture_bam, gr, shift=shift, ss=True).as_list(),
                  O.bamProfile(fixture_bam, gr, shift=shift, ss=True).as_list())


def test_out_ptrs_variant_and_overwrite(fixture_bam):
    """The R-pointer output variant fills every per-region vector, and pre-existing garbage is overwritten."""
    import ctypes as C
    from bamsignals_b200 import api
    gr = to_gr(spec_r.annot_regions())
    m = api.marshal_regions(gr)
    off = api.output_layout(m.width, 1, True)
    bufs = [np.full(int(off[i + 1] - off[i]), 12345, dtype=np.int32) for i in range(len(gr))]
    ptrs = (C.POINTER(C.c_int32) * len(gr))(*[api._p(b, C.c_int32) for b in bufs])
    rc = B.lib().bsg_pileup(fixture_bam.encode(), m.R, m.levels, m.n_levels, api._p(m.seq_idx, C.c_int32),
                            api._p(m.loc, C.c_int32), api._p(m.width, C.c_int32), api._p(m.strand, C.c_int8), None,
                            0, 1, 0, 1, 0, -1, 0, 16385, None, api._p(off, C.c_int64), ptrs, None)
    assert rc == 0, B.lib().bsg_last_error()
    want = O.bamProfile(fixture_bam, gr, ss=True).as_list()
    for b, w in zip(bufs, want):
        assert np.array_equal(b, w.ravel(order="F"))


def test_staged_session_matches(fixture_bam):
    gr = to_gr(spec_r.test_regions(seed=11, n=64))
    with B.Stage(fixture_bam, gr, ext_hint=1100) as st:
        for shift, ss, binsize in ((0, 0, 1), (75, 1, 1), (100, 1, 10), (0, 1, -1)):
            flat = st.pileup((0, 1000), 5, binsize, shift, ss, 66, 1024, True)
            want = O.pileup_core(fixture_bam, gr, (0, 1000), 5, binsize, shift, bool(ss), 66, 1024, True)
            assert np.array_equal(flat, np.concatenate([w.ravel(order="F") for w in want]))
            assert B.timings()["ms_device"] > 0 and B.timings()["n_launches"] >= 3   # decode+filter, join, count
        flat = st.coverage((0, 1000), 0, 66, -1, True)
        assert np.array_equal(flat, np.concatenate(O.coverage_core(fixture_bam, gr, (0, 1000), 0, 66, -1, True)))
        assert st.pileup(None, want_output=False) is None


@pytest.mark.skipif(not O.ref_available(), reason="oracle/_ref not built")
def test_sweep_vs_reference_code(fixture_bam):
    """The CUDA path against the reference's OWN engine (oracle/_ref: src/bamsignals.cpp compiled unchanged) over the
    whole sweep of tests/testthat/test_methods.R:33-104 on a stranded region set."""
    gr = to_gr(spec_r.test_regions(seed=21, n=30))
    for case in spec_r.sweep_pileup():
        kw = dict(mapqual=case["mapqual"], shift=case["shift"], ss=case["ss"], paired_end=case["paired_end"],
                  tlenFilter=case["tlenFilter"])
        assert np.array_equal(B.bamCount(fixture_bam, gr, **kw), O.bamCount(fixture_bam, gr, impl="ref", **kw)), case
        same_list(B.bamProfile(fixture_bam, gr, **kw).as_list(), O.bamProfile(fixture_bam, gr, impl="ref", **kw).as_list())
    for case in spec_r.sweep_coverage():
        kw = dict(mapqual=case["mapqual"], paired_end=case["paired_end"], tlenFilter=case["tlenFilter"])
        same_list(B.bamCoverage(fixture_bam, gr, **kw).as_list(), O.bamCoverage(fixture_bam, gr, impl="ref", **kw).as_list())
    for ff in (16, 1024, 1040, 0):
        assert np.array_equal(B.bamCount(fixture_bam, gr, filteredFlag=ff), O.bamCount(fixture_bam, gr, filteredFlag=ff, impl="ref"))


def test_interleaved_sessions_and_calls(fixture_bam):
    """Two staged sessions on one device, interleaved with each other and with a plain call: each owns its resident
    bytes; cached tiles are rebuilt when another call has replaced them on the device."""
    gr_a = to_gr(spec_r.test_regions(seed=31, n=40))
    gr_b = to_gr(spec_r.test_regions(seed=32, n=77))
    gr_c = to_gr(spec_r.test_regions(seed=33, n=13))
    want_a = np.concatenate([w.ravel(order="F") for w in O.pileup_core(fixture_bam, gr_a, None, 0, 1, 10, True)])
    want_b = np.concatenate(O.coverage_core(fixture_bam, gr_b, None))
    with B.Stage(fixture_bam, gr_a, ext_hint=10) as sa:
        assert np.array_equal(sa.pileup(None, 0, 1, 10, True), want_a)
        with B.Stage(fixture_bam, gr_b, ext_hint=0) as sb:
            assert np.array_equal(sb.coverage(None), want_b)
            assert np.array_equal(sa.pileup(None, 0, 1, 10, True), want_a)          # same parameters: cached tiles must be re-uploaded
            same_list(B.bamProfile(fixture_bam, gr_c, binsize=7, ss=True).as_list(),
                      O.bamProfile(fixture_bam, gr_c, binsize=7, ss=True).as_list())
            assert np.array_equal(sb.coverage(None), want_b)
            assert np.array_equal(sa.pileup(None, 0, 1, 10, True), want_a)
        assert np.array_equal(sa.pileup(None, 0, 1, 10, True), want_a)              # after the other session is closed


def test_shutdown_ends_open_sessions(fixture_bam):
    gr = to_gr(spec_r.test_regions(seed=34, n=10))
    st = B.Stage(fixture_bam, gr)
    want = O.bamCount(fixture_bam, gr)
    assert np.array_equal(st.pileup(None, binsize=-1), want)
    B.lib().bsg_shutdown()
    with pytest.raises(B.BamsignalsError) as e:
        st.pileup(None, binsize=-1)
    assert e.value.code == -8 and "bsg_shutdown" in str(e.value)
    st.close()
    assert np.array_equal(B.bamCount(fixture_bam, gr), want)                        # the library re-initialises itself


def test_inconsistent_layout_is_refused(fixture_bam):
    """out_offsets computed for another binsize / ss would make the result scatter overrun the caller's buffer."""
    import ctypes as C
    from bamsignals_b200 import api
    gr = to_gr(spec_r.annot_regions())
    m = api.marshal_regions(gr)
    for off in (api.output_layout(m.width, 2, True), api.output_layout(m.width, 1, False)):
        flat = np.zeros(int(api.output_layout(m.width, 1, True)[-1]), dtype=np.int32)
        rc = B.lib().bsg_pileup(fixture_bam.encode(), m.R, m.levels, m.n_levels, api._p(m.seq_idx, C.c_int32),
                                api._p(m.loc, C.c_int32), api._p(m.width, C.c_int32), api._p(m.strand, C.c_int8), None,
                                0, 1, 0, 1, 0, -1, 0, 16385, api._p(flat, C.c_int32), api._p(off, C.c_int64), None, None)
        assert rc == -8 and b"bsg_output_layout" in B.lib().bsg_last_error()
        assert not flat.any()


def test_opts_struct_versions(fixture_bam):
    """A zero-initialised bsg_opts (struct_size set) means defaults - CRC check on; a shorter struct is a prefix."""
    import ctypes as C
    gr = to_gr(spec_r.annot_regions())
    want = O.bamCount(fixture_bam, gr)
    o = B.BsgOpts()
    o.struct_size = C.sizeof(B.BsgOpts)
    assert np.array_equal(B.bamCount(fixture_bam, gr, opts=o), want)
    o.struct_size = 8                                   # struct_size + n_devices only
    assert np.array_equal(B.bamCount(fixture_bam, gr, opts=o), want)
    o.struct_size = 0
    with pytest.raises(B.BamsignalsError) as e:
        B.bamCount(fixture_bam, gr, opts=o)
    assert e.value.code == -8
