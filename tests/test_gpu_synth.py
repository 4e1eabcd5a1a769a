"""GPU parity beyond the reference fixture: scaled-down versions of BASELINE.json's configs C2..C5 from the synthetic
generator, hand-crafted edge-case BAMs (CIGAR variety, flag-0x4 reads, straddling records, unsorted / corrupt
input), batching and threading options that must not change results."""
import os
import sys

import numpy as np
import pytest

import bamsignals_b200 as B
import bamwriter as W
import edge_cases as E
import oracle_api as O

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import workloads as WL  # noqa: E402

pytestmark = pytest.mark.gpu


def same(a, b):
    assert len(a) == len(b)
    for x, y in zip(a, b):
        assert x.shape == y.shape and np.array_equal(x, y)


@pytest.fixture(scope="module")
def gen_dir(tmp_path_factory):
    return str(tmp_path_factory.mktemp("gen"))


@pytest.mark.parametrize("preset,gs,record", [("c2", 0.002, "compact"), ("c2", 0.001, "realistic"), ("c3", 0.004, "compact"),
                                              ("c3", 0.002, "realistic"), ("c4", 0.002, "compact"), ("c5", 0.002, "compact")])
def test_scaled_configs(gen_dir, preset, gs, record):
    """host-inflate path (zlib worker pool -> pinned staging -> H2D); the device-inflate twin is further down"""
    bam, info = WL.make_bam(preset, gs, gen_dir, record=record, unplaced=7)
    gr, kw, fn = WL.regions(preset, gs)
    got = getattr(B, fn)(bam, gr, opts=B.default_opts(gpu_inflate=-1), **kw)
    t = B.timings()
    want = getattr(O, fn)(bam, gr, nthreads=8, **kw)
    assert np.array_equal(WL.as_flat(got), WL.as_flat(want))
    assert t["records"] > 0 and t["records"] <= info["records"]
    assert int(WL.as_flat(got).sum()) > 0


@pytest.mark.parametrize("batch_bytes,threads", [(1 << 16, 1), (1 << 20, 3), (1 << 22, 0)])
def test_batching_does_not_change_results(gen_dir, batch_bytes, threads):
    bam, _ = WL.make_bam("c4", 0.002, gen_dir, unplaced=7)
    gr, kw, fn = WL.regions("c4", 0.002)
    ref = WL.as_flat(getattr(O, fn)(bam, gr, nthreads=8, **kw))
    got = getattr(B, fn)(bam, gr, opts=B.default_opts(batch_bytes=batch_bytes, inflate_threads=threads, gpu_inflate=-1), **kw)
    assert np.array_equal(WL.as_flat(got), ref)
    assert B.timings()["n_batches"] >= (2 if batch_bytes < (1 << 22) else 1) and B.timings()["ms_inflate_gpu"] == 0


def test_all_apis_on_paired_synthetic(gen_dir):
    """Every paired.end mode on a paired-end synthetic BAM with soft clips, indels, splices and flag-0x4 mates."""
    bam, _ = WL.make_bam("c3", 0.004, gen_dir)
    rng = np.random.default_rng(9)
    n = 300
    L = WL.contig_lens("c3", 0.004)[0]
    gr = B.GRanges(["chr1"] * n, rng.integers(1, L - 20000, n), rng.integers(1, 20000, n), rng.choice(["+", "-", "*"], n).tolist())
    for pe in ("ignore", "filter", "midpoint"):
        for tf in (None, (100, 400)):
            kw = dict(paired_end=pe, tlenFilter=tf, shift=13, mapqual=5)
            assert np.array_equal(B.bamCount(bam, gr, ss=True, **kw), O.bamCount(bam, gr, ss=True, **kw))
            same(B.bamProfile(bam, gr, binsize=1, ss=True, **kw).as_list(), O.bamProfile(bam, gr, binsize=1, ss=True, **kw).as_list())
    for pe in ("ignore", "extend"):
        for ff in (-1, 1024, 4):
            same(B.bamCoverage(bam, gr, paired_end=pe, filteredFlag=ff).as_list(), O.bamCoverage(bam, gr, paired_end=pe, filteredFlag=ff).as_list())


@pytest.mark.parametrize("straddle", [False, True])
def test_cigar_and_flag_variety(tmp_path, straddle):
    p = str(tmp_path / "v.bam")
    E.write_variety(p, block_payload=997, cut_mid_record=straddle)
    gr = E.variety_regions()
    for kw in (dict(), dict(ss=True, shift=33), dict(paired_end="midpoint", tlenFilter=(0, 500), ss=True),
               dict(filteredFlag=1024, mapqual=20), dict(filteredFlag=4), dict(filteredFlag=20)):
        assert np.array_equal(B.bamCount(p, gr, **kw), O.bamCount(p, gr, **kw)), kw
    for kw in (dict(binsize=1, ss=True), dict(binsize=50, shift=-20), dict(binsize=7, ss=True, paired_end="filter")):
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            same(B.bamProfile(p, gr, **kw).as_list(), O.bamProfile(p, gr, **kw).as_list())
    for kw in (dict(), dict(paired_end="extend", tlenFilter=(0, 600)), dict(mapqual=59)):
        same(B.bamCoverage(p, gr, **kw).as_list(), O.bamCoverage(p, gr, **kw).as_list())


def test_long_cigar_placeholder(tmp_path):
    p = str(tmp_path / "cg.bam")
    W.write_bam(p, [("chrA", 200000)], [E.long_cigar_read()])
    gr = B.GRanges(["chrA"] * 2, [1, 60000], [100000, 20000], ["+", "-"])
    same(B.bamCoverage(p, gr).as_list(), O.bamCoverage(p, gr).as_list())
    assert B.bamCount(p, gr).tolist() == O.bamCount(p, gr).tolist() == [1, 0]
    assert B.bamCount(p, gr, ss=True, shift=0).tolist() == O.bamCount(p, gr, ss=True).tolist()


def test_empty_bam_and_empty_chromosome(tmp_path):
    p = str(tmp_path / "empty.bam")
    W.write_bam(p, [("chrA", 1000), ("chrB", 1000)], [])
    gr = B.GRanges(["chrA", "chrB"], [1, 10], [100, 50], ["+", "-"])
    assert B.bamCount(p, gr, ss=True).tolist() == [[0, 0], [0, 0]]
    assert all(int(x.sum()) == 0 for x in B.bamProfile(p, gr).as_list())
    assert [len(x) for x in B.bamCoverage(p, gr).as_list()] == [100, 50]
    p2 = str(tmp_path / "one.bam")
    W.write_bam(p2, [("chrA", 1000), ("chrB", 1000)], [dict(tid=1, pos=20, flag=16, mapq=9, cigar="10M", tlen=0)])
    assert B.bamCount(p2, gr, ss=True).tolist() == O.bamCount(p2, gr, ss=True).tolist()
    same(B.bamCoverage(p2, gr).as_list(), O.bamCoverage(p2, gr).as_list())


def test_unsorted_input_is_an_error(tmp_path):
    p = str(tmp_path / "unsorted.bam")
    reads = [dict(tid=0, pos=x, flag=0, mapq=30, cigar="20M", tlen=0) for x in (100, 500, 300, 900)]
    W.write_bam(p, [("chrA", 5000)], reads)
    with pytest.raises(B.BamsignalsError) as e:
        B.bamCount(p, B.GRanges(["chrA"], [1], [5000]))
    assert e.value.code == -5


def test_corrupt_bgzf_is_an_error(tmp_path, fixture_bam):
    raw = bytearray(open(fixture_bam, "rb").read())
    raw[200000] ^= 0xFF                                   # inside some block's deflate stream
    p = str(tmp_path / "corrupt.bam")
    open(p, "wb").write(raw)
    open(p + ".bai", "wb").write(open(fixture_bam + ".bai", "rb").read())
    gr = B.GRanges(["chr1", "chr2", "chr3"], [1, 1, 1], [10000] * 3)
    for mode in (-1, 1):
        with pytest.raises(B.BamsignalsError) as e:
            B.bamCount(p, gr, opts=B.default_opts(gpu_inflate=mode))
        assert e.value.code == -4
    open(p, "wb").write(raw[:700000])                     # truncated mid-block
    for mode in (-1, 1):
        with pytest.raises(B.BamsignalsError) as e:
            B.bamCount(p, gr, opts=B.default_opts(gpu_inflate=mode))
        assert e.value.code == -4
    # the library is still usable after an error
    assert np.array_equal(B.bamCount(fixture_bam, gr), O.bamCount(fixture_bam, gr))


def test_properties_at_scale(gen_dir):
    """Size-independent properties on a larger input than the oracle is asked to check bin by bin:
    profile sums == counts, ss rows sum to the unstranded profile, coverage of + and - regions mirror each other,
    filteredFlag=0 gives zeros, binned profile == reshaped sum of the bp profile."""
    bam, info = WL.make_bam("c2", 0.02, gen_dir)
    gr, kw, _ = WL.regions("c2", 0.02)
    prof = B.bamProfile(bam, gr, **kw)
    cnt = B.bamCount(bam, gr, ss=True, shift=75)
    assert np.array_equal(np.stack([p.sum(axis=1) for p in prof.as_list()], axis=1), cnt)
    uns = B.bamProfile(bam, gr, binsize=1, ss=False, shift=75).alignSignals()
    assert np.array_equal(prof.alignSignals().sum(axis=0), uns)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        b40 = B.bamProfile(bam, gr, binsize=40, ss=False, shift=75).alignSignals()
    assert np.array_equal(uns.reshape(50, 40, -1).sum(axis=1), b40)
    assert int(B.bamCount(bam, gr, filteredFlag=0).sum()) == 0
    flipped = B.GRanges.from_codes(gr.seqlevels, gr.seq_idx, gr.start, gr.width, -gr.strand)
    c1 = B.bamCoverage(bam, gr).alignSignals()
    c2 = B.bamCoverage(bam, flipped).alignSignals()
    assert np.array_equal(c1, c2[::-1])
    assert B.timings()["records"] > 0.3 * info["records"]


# ---- GPU-side inflate + record walk (opts.gpu_inflate = 1): same results, bit for bit ---------------------------------
GPUI = dict(gpu_inflate=1)


@pytest.mark.parametrize("preset,gs,record", [("c2", 0.002, "compact"), ("c2", 0.001, "realistic"), ("c3", 0.004, "compact"),
                                              ("c4", 0.002, "compact"), ("c5", 0.002, "compact")])
def test_gpu_inflate_scaled_configs(gen_dir, preset, gs, record):
    bam, info = WL.make_bam(preset, gs, gen_dir, record=record, unplaced=7)
    gr, kw, fn = WL.regions(preset, gs)
    got = getattr(B, fn)(bam, gr, opts=B.default_opts(**GPUI), **kw)
    t = B.timings()
    want = getattr(O, fn)(bam, gr, nthreads=8, **kw)
    assert np.array_equal(WL.as_flat(got), WL.as_flat(want))
    assert t["ms_inflate_gpu"] > 0 and t["records"] > 0


@pytest.mark.parametrize("batch_bytes", [1 << 16, 1 << 20, 0])
def test_gpu_inflate_batching(gen_dir, batch_bytes):
    bam, _ = WL.make_bam("c4", 0.002, gen_dir, unplaced=7)
    gr, kw, fn = WL.regions("c4", 0.002)
    ref = WL.as_flat(getattr(O, fn)(bam, gr, nthreads=8, **kw))
    got = getattr(B, fn)(bam, gr, opts=B.default_opts(batch_bytes=batch_bytes, **GPUI), **kw)
    assert np.array_equal(WL.as_flat(got), ref)


def test_gpu_inflate_reference_fixture(fixture_bam):
    """The reference's own BAM (written by htslib, deflate level 6) through the GPU inflate path."""
    import spec_r
    genes = spec_r.test_regions()
    gr = B.GRanges([["chr1", "chr2", "chr3"][i] for i in genes["rname"]], genes["start"], genes["width"], genes["strand"])
    o = B.default_opts(**GPUI)
    for case in list(spec_r.sweep_pileup())[::5]:
        kw = dict(mapqual=case["mapqual"], shift=case["shift"], ss=case["ss"], paired_end=case["paired_end"], tlenFilter=case["tlenFilter"])
        assert np.array_equal(B.bamCount(fixture_bam, gr, opts=o, **kw), O.bamCount(fixture_bam, gr, **kw))
        same(B.bamProfile(fixture_bam, gr, opts=o, **kw).as_list(), O.bamProfile(fixture_bam, gr, **kw).as_list())
    same(B.bamCoverage(fixture_bam, gr, opts=o, paired_end="extend").as_list(), O.bamCoverage(fixture_bam, gr, paired_end="extend").as_list())
    wide = B.GRanges(["chr1", "chr2", "chr3"], [1, 1, 1], [10237, 10279, 10238], ["+", "-", "*"])
    same(B.bamCoverage(fixture_bam, wide, opts=o).as_list(), O.bamCoverage(fixture_bam, wide).as_list())


@pytest.mark.parametrize("level,straddle,payload", [(0, False, 3000), (1, True, 997), (6, False, 0xFF00), (9, True, 20000), (6, False, 40)])
def test_gpu_inflate_deflate_variants(tmp_path, level, straddle, payload):
    """Stored blocks (level 0), fixed-Huffman blocks (tiny payloads), dynamic blocks at several levels, straddling records."""
    p = str(tmp_path / "v.bam")
    reads = E.variety_reads(n=1500)
    W.write_bam(p, E.REFS, reads, block_payload=payload, cut_mid_record=straddle, level=level)
    gr = E.variety_regions()
    o = B.default_opts(**GPUI)
    assert np.array_equal(B.bamCount(p, gr, ss=True, shift=33, opts=o), O.bamCount(p, gr, ss=True, shift=33))
    same(B.bamCoverage(p, gr, opts=o).as_list(), O.bamCoverage(p, gr).as_list())


def test_gpu_inflate_corrupt_stream(tmp_path, fixture_bam):
    """A flipped byte inside a DEFLATE stream is caught on the device: either the stream no longer decodes to ISIZE
    bytes, or the CRC32 kernel sees the damage (htslib checks the CRC as well)."""
    raw = bytearray(open(fixture_bam, "rb").read())
    gr = B.GRanges(["chr1", "chr2", "chr3"], [1, 1, 1], [10000] * 3)
    for where in (200000, 333333, 901234):
        bad = bytearray(raw)
        bad[where] ^= 0x5A
        p = str(tmp_path / f"corrupt{where}.bam")
        open(p, "wb").write(bad)
        open(p + ".bai", "wb").write(open(fixture_bam + ".bai", "rb").read())
        with pytest.raises(B.BamsignalsError) as e:
            B.bamCount(p, gr, opts=B.default_opts(**GPUI))
        assert e.value.code == -4
    assert np.array_equal(B.bamCount(fixture_bam, gr, opts=B.default_opts(**GPUI)), O.bamCount(fixture_bam, gr))


# ---- streaming finalisation: tiles are counted + shipped batch by batch ------------------------------------------------
@pytest.mark.parametrize("preset,gs", [("c2", 0.004), ("c3", 0.004), ("c4", 0.002), ("c5", 0.002)])
@pytest.mark.parametrize("stream", [1, -1])
def test_streaming_finalisation_small_batches(gen_dir, preset, gs, stream):
    """Tiny batches + a 1-int streaming threshold force the mid-pipeline path (filter new rows, join + count the tiles
    the frontier has passed, ship them) on every batch; -1 is the single pass at the end.  Both must match the oracle."""
    bam, _ = WL.make_bam(preset, gs, gen_dir, unplaced=7)
    gr, kw, fn = WL.regions(preset, gs)
    want = WL.as_flat(getattr(O, fn)(bam, gr, nthreads=8, **kw))
    for batch in (1 << 16, 1 << 19):
        got = getattr(B, fn)(bam, gr, opts=B.default_opts(batch_bytes=batch, stream_min_ints=stream, gpu_inflate=1), **kw)
        t = B.timings()
        assert np.array_equal(WL.as_flat(got), want)
        if stream > 0 and batch == (1 << 16):
            assert t["n_launches"] > 3 * t["n_batches"]          # decode + inflate/walks per batch, plus the count launches


def test_streaming_with_shuffled_and_overlapping_regions(gen_dir):
    """Caller order != genomic order, duplicated and nested regions, '-' strand tiles wider than one tile."""
    bam, _ = WL.make_bam("c3", 0.004, gen_dir)
    L = WL.contig_lens("c3", 0.004)[0]
    rng = np.random.default_rng(12)
    n = 400
    start = rng.integers(1, L - 40000, n)
    width = rng.integers(1, 40000, n)
    start[::7] = start[0]
    width[::7] = width[0]
    gr = B.GRanges(["chr1"] * n, start, width, rng.choice(["+", "-", "*"], n).tolist())
    o = B.default_opts(batch_bytes=1 << 17, stream_min_ints=1, gpu_inflate=1)
    same(B.bamCoverage(bam, gr, paired_end="extend", opts=o).as_list(), O.bamCoverage(bam, gr, paired_end="extend", nthreads=8).as_list())
    same(B.bamProfile(bam, gr, binsize=3, ss=True, shift=40, opts=o).as_list(), O.bamProfile(bam, gr, binsize=3, ss=True, shift=40, nthreads=8).as_list())
    assert np.array_equal(B.bamCount(bam, gr, ss=True, paired_end="midpoint", opts=o), O.bamCount(bam, gr, ss=True, paired_end="midpoint", nthreads=8))


# ---- the default pipeline's own boundaries, crossed and compared with the oracle --------------------------------------
def test_default_streaming_threshold_across_batch_boundaries(gen_dir):
    """Everything at its DEFAULT except the batch size, which is the largest power of two below a third of the input, so
    that the call crosses several batch boundaries (records straddling them included), the short first batch, and the
    default 4 Mi-int streaming threshold (C2 at 1/50: 8 M result ints): what a full-size call does every 1 GiB."""
    bam, info = WL.make_bam("c2", 0.02, gen_dir, unplaced=7)
    gr, kw, fn = WL.regions("c2", 0.02)
    want = WL.as_flat(getattr(O, fn)(bam, gr, nthreads=8, **kw))
    plan = B.debug_plan(bam, gr, ext=75, cap=1)
    batch = 1 << (int(plan["bytes_inflated"] // 3).bit_length() - 1)
    for inflate in (1, -1):
        got = getattr(B, fn)(bam, gr, opts=B.default_opts(batch_bytes=batch, gpu_inflate=inflate), **kw)
        t = B.timings()
        assert np.array_equal(WL.as_flat(got), want), inflate
        assert t["n_batches"] >= 3 and t["out_elems"] > (4 << 20)
    got = getattr(B, fn)(bam, gr, **kw)                        # and the true default: one short first batch + the rest
    assert np.array_equal(WL.as_flat(got), want)
    if O.ref_available():                                      # the same job through the reference's own engine
        assert np.array_equal(WL.as_flat(getattr(O, fn)(bam, gr, nthreads=8, impl="ref", **kw)), want)


@pytest.mark.parametrize("preset,gs", [("c3", 0.01), ("c4", 0.004), ("c5", 0.004)])
def test_default_options_mid_scale(gen_dir, preset, gs):
    """C3 / C4 / C5 at a scale where a call has several default-sized streaming portions; all defaults."""
    bam, _ = WL.make_bam(preset, gs, gen_dir, unplaced=7)
    gr, kw, fn = WL.regions(preset, gs)
    want = WL.as_flat(getattr(O, fn)(bam, gr, nthreads=8, **kw))
    plan = B.debug_plan(bam, gr, ext=1000, cap=1)
    batch = 1 << (int(plan["bytes_inflated"] // 3).bit_length() - 1)
    assert np.array_equal(WL.as_flat(getattr(B, fn)(bam, gr, opts=B.default_opts(batch_bytes=batch), **kw)), want)
    assert B.timings()["n_batches"] >= 3
    assert np.array_equal(WL.as_flat(getattr(B, fn)(bam, gr, **kw)), want)


@pytest.mark.parametrize("density,result_pack,what", [(1.0, 0, "list"), (6.0, 0, "list+fallback"), (6.0, 2, "list"), (6.0, 1 << 20, "fallback"),
                                                      (6.0, -1, "int32")])
def test_result_narrowing_is_exact(gen_dir, density, result_pack, what):
    """opts.result_pack: a large result crosses PCIe as bytes + (index, value) pairs for the elements above 254; portions
    with more such elements than the list holds travel as int32.  Deep coverage (density 6: hundreds of reads per base)
    puts most elements above 254; every variant must return the oracle's integers."""
    bam, _ = WL.make_bam("c3", 0.004, gen_dir, density=density)
    gr, kw, fn = WL.regions("c3", 0.004)
    want = WL.as_flat(getattr(O, fn)(bam, gr, nthreads=8, **kw))
    if density > 1:
        assert int((want > 254).sum()) > 70000 and int((want <= 254).sum()) > 1000      # > 1/64 of the 4 M elements
    got = getattr(B, fn)(bam, gr, opts=B.default_opts(result_pack=result_pack, stream_min_ints=1 << 16), **kw)
    assert np.array_equal(WL.as_flat(got), want), what
    # per-region destinations (the R list layout) take the same path
    got2 = B.bamProfile(bam, gr, binsize=1, ss=True, opts=B.default_opts(result_pack=result_pack))
    want2 = O.bamProfile(bam, gr, binsize=1, ss=True)
    same(got2.as_list(), want2.as_list())


@pytest.mark.parametrize("scheme", [1, 2])
@pytest.mark.parametrize("preset,gs,straddle", [("c4", 0.002, 0.05), ("c5", 0.002, 0.6), ("c2", 0.002, 0.0)])
def test_walk_schemes_agree(gen_dir, preset, gs, straddle, scheme):
    """opts.walk_scheme: the per-span record walk and the block-parallel one (speculative chain per BGZF block, spans
    linked through the blocks) must both find every record - also when most blocks begin inside a record (straddle 0.6:
    the speculation fails, the true chain is re-walked) and across many small batches."""
    bam, _ = WL.make_bam(preset, gs, gen_dir, straddle=straddle, unplaced=3)
    gr, kw, fn = WL.regions(preset, gs)
    want = WL.as_flat(getattr(O, fn)(bam, gr, nthreads=8, **kw))
    for batch in (0, 1 << 18):
        got = getattr(B, fn)(bam, gr, opts=B.default_opts(gpu_inflate=1, walk_scheme=scheme, batch_bytes=batch), **kw)
        assert np.array_equal(WL.as_flat(got), want), (scheme, batch)
        assert B.timings()["records"] > 0
