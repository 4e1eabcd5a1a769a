"""In-library multi-GPU (opts.n_devices > 1): regions are sharded across devices inside ONE C-ABI call, every device
writes its own slice of the caller's result; must be bit-identical to the oracle.  Needs >= 2 GPUs
(run with `gpurun --gpus 2`); skipped on a single-GPU box."""
import os
import sys

import numpy as np
import pytest

import bamsignals_b200 as B
import oracle_api as O

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import workloads as WL  # noqa: E402

pytestmark = pytest.mark.gpu


def ndev():
    return B.lib().bsg_device_count()


@pytest.fixture(scope="module")
def gen_dir(tmp_path_factory):
    return str(tmp_path_factory.mktemp("gen"))


@pytest.mark.parametrize("preset,gs", [("c2", 0.004), ("c3", 0.004), ("c4", 0.002), ("c5", 0.002)])
@pytest.mark.parametrize("inflate", [1, -1])
def test_multi_device_matches_oracle(gen_dir, preset, gs, inflate):
    if ndev() < 2:
        pytest.skip("needs two GPUs")
    devs = list(range(min(ndev(), 8)))
    bam, info = WL.make_bam(preset, gs, gen_dir, unplaced=3)
    gr, kw, fn = WL.regions(preset, gs)
    want = WL.as_flat(getattr(O, fn)(bam, gr, nthreads=8, **kw))
    got = getattr(B, fn)(bam, gr, opts=B.default_opts(devices=devs, gpu_inflate=inflate), **kw)
    t = B.timings()
    assert np.array_equal(WL.as_flat(got), want)
    assert t["n_devices"] == len(devs) and t["records"] > 0
    one = getattr(B, fn)(bam, gr, opts=B.default_opts(devices=[devs[-1]], gpu_inflate=inflate), **kw)
    assert np.array_equal(WL.as_flat(one), want)


def test_multi_device_fixture_and_errors(fixture_bam):
    if ndev() < 2:
        pytest.skip("needs two GPUs")
    import spec_r
    g = spec_r.test_regions(seed=21, n=101)
    gr = B.GRanges([["chr1", "chr2", "chr3"][i] for i in g["rname"]], g["start"], g["width"], g["strand"])
    o = B.default_opts(devices=[0, 1])
    for kw in (dict(ss=True, shift=40), dict(paired_end="midpoint", tlenFilter=(50, 300))):
        assert np.array_equal(B.bamCount(fixture_bam, gr, opts=o, **kw), O.bamCount(fixture_bam, gr, **kw))
        for a, b in zip(B.bamProfile(fixture_bam, gr, opts=o, **kw).as_list(), O.bamProfile(fixture_bam, gr, **kw).as_list()):
            assert np.array_equal(a, b)
    for a, b in zip(B.bamCoverage(fixture_bam, gr, opts=o, paired_end="extend").as_list(), O.bamCoverage(fixture_bam, gr, paired_end="extend").as_list()):
        assert np.array_equal(a, b)
    with pytest.raises(B.BamsignalsError, match="listed twice"):
        B.bamCount(fixture_bam, gr, opts=B.default_opts(devices=[0, 0]))
    with pytest.raises(B.BamsignalsError, match="chromosome chrQ not present"):
        B.bamCount(fixture_bam, B.GRanges(["chrQ"], [1], [5]), opts=o)
    # fewer regions than devices, and empty region sets
    assert B.bamCount(fixture_bam, gr[:1], opts=o).tolist() == O.bamCount(fixture_bam, gr[:1]).tolist()
    assert B.bamCount(fixture_bam, B.GRanges([], [], [], []), opts=o).shape == (0,)


def test_multi_device_splits_wide_regions(gen_dir):
    """A few regions much wider than a shard (C4's shape): the library cuts them into bin-aligned pieces and spreads
    the pieces over the devices; '+', '-' and '*' regions, strand-specific and binned profiles, and coverage must all
    come back exactly as one device (and the oracle) computes them."""
    if ndev() < 2:
        pytest.skip("needs two GPUs")
    devs = list(range(min(ndev(), 8)))
    bam, info = WL.make_bam("c4", 0.002, gen_dir, unplaced=3)
    lens = WL.contig_lens("c4", 0.002)
    names = WL.NAMES[:6]
    gr = B.GRanges(names, [1, 101, 7, 1, 50, 1], [lens[0], lens[1] - 200, lens[2] - 10, lens[3], lens[4] - 100, lens[5]],
                   ["+", "-", "*", "-", "+", "-"])
    o = B.default_opts(devices=devs)
    for kw in (dict(binsize=1, ss=True, shift=30), dict(binsize=7, ss=False), dict(binsize=200, ss=True, paired_end="midpoint", tlenFilter=(70, 200))):
        got = WL.as_flat(B.bamProfile(bam, gr, opts=o, **kw))
        assert np.array_equal(got, WL.as_flat(O.bamProfile(bam, gr, nthreads=8, **kw))), kw
        assert B.timings()["n_devices"] == len(devs)
    got = WL.as_flat(B.bamCoverage(bam, gr, opts=o, paired_end="extend"))
    assert np.array_equal(got, WL.as_flat(O.bamCoverage(bam, gr, nthreads=8, paired_end="extend")))
    assert np.array_equal(B.bamCount(bam, gr, opts=o, ss=True), O.bamCount(bam, gr, ss=True))
