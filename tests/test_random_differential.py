"""Randomised differential tests: seeded scenarios from tests/random_cases.py.

CPU (`-m "not gpu"`): the oracle, in all three access modes, against the independent numpy statement of SURVEY
App. A — this pins the oracle on everything the reference's fixture leaves unpinned (CIGAR operators, flag-0x4 reads,
arbitrary flag masks, odd regions).  GPU (`-m gpu`): the CUDA path through the C ABI against the oracle on the same
scenarios, with the device-inflate and the host-inflate pipeline, forced streaming finalisation and tiny batches.
"""
import warnings

import numpy as np
import pytest

import bamsignals_b200 as B
import oracle_api as O
import random_cases as RC

N_CPU = 60
N_GPU = 48


def _call(mod, sc, bam, **extra):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return RC.flat(getattr(mod, sc["fn"])(bam, sc["gr"], **sc["kw"], **extra), sc["fn"])


@pytest.mark.parametrize("seed", range(N_CPU))
def test_oracle_matches_numpy_spec(tmp_path, seed):
    sc = RC.scenario(seed)
    bam = RC.write(sc, str(tmp_path / "r.bam"))
    want = RC.spec_counts(sc)
    for mode in (O.INDEXED, O.SCAN, O.BRUTE):
        got = _call(O, sc, bam, mode=mode)
        assert got.shape == want.shape and np.array_equal(got, want), (seed, sc["fn"], sc["kw"], mode)
    got = _call(O, sc, bam, mode=O.INDEXED, maxgap=0, nthreads=3)
    assert np.array_equal(got, want), (seed, sc["fn"], sc["kw"], "maxgap=0")


def test_scenarios_cover_the_space():
    """The generator is only useful if the scenarios are not degenerate: most must count something, every API, every
    paired.end mode and both strandedness settings must occur."""
    fns, pes, nonzero, ss = set(), set(), 0, set()
    for seed in range(N_CPU):
        sc = RC.scenario(seed)
        fns.add(sc["fn"]); pes.add(sc["kw"]["paired_end"]); ss.add(sc["kw"].get("ss"))
        nonzero += int(RC.spec_counts(sc).sum() > 0)
    assert fns == {"bamCount", "bamProfile", "bamCoverage"}
    assert pes == {"ignore", "filter", "midpoint", "extend"} and ss >= {True, False}
    assert nonzero >= N_CPU // 4


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(N_GPU))
def test_gpu_matches_oracle_and_spec(tmp_path, seed):
    sc = RC.scenario(seed)
    bam = RC.write(sc, str(tmp_path / "r.bam"))
    want = _call(O, sc, bam)
    assert np.array_equal(want, RC.spec_counts(sc))
    variants = [B.default_opts(gpu_inflate=1), B.default_opts(gpu_inflate=-1),
                B.default_opts(gpu_inflate=1, stream_min_ints=1, batch_bytes=1 << 16),
                B.default_opts(gpu_inflate=-1, stream_min_ints=1, batch_bytes=1 << 14, inflate_threads=2)]
    for k, o in enumerate(variants):
        got = _call(B, sc, bam, opts=o)
        assert got.shape == want.shape and np.array_equal(got, want), (seed, sc["fn"], sc["kw"], k)
    # resident session: same call twice, then with the other strandedness / a stricter filter (tiles are rebuilt)
    ca = B.core_args(sc["fn"], **sc["kw"])
    cov = sc["fn"] == "bamCoverage"
    ext = (ca["tlen_filter"][1] if ca["tspan"] else 0) if cov else abs(ca["shift"]) + (ca["tlen_filter"][1] if ca["pe_mid"] else 0)
    with B.Stage(bam, sc["gr"], ext_hint=ext, opts=variants[seed % 2]) as st:
        run = st.coverage if cov else st.pileup
        for _ in range(2):
            assert np.array_equal(run(**ca), want), (seed, "staged")
        ca2 = dict(ca, mapqual=30, filteredF=1024) if cov else dict(ca, ss=not ca["ss"], mapqual=30)
        core = O.coverage_core if cov else O.pileup_core
        want2 = [np.asarray(x).reshape(-1, order="F") for x in core(bam, sc["gr"], **ca2)]
        want2 = np.concatenate(want2).astype(np.int32) if want2 else np.zeros(0, np.int32)
        assert np.array_equal(run(**ca2), want2), (seed, "staged, second parameter set")
        assert np.array_equal(run(**ca), want), (seed, "staged, back to the first parameter set")
    if ext > 0:
        with B.Stage(bam, sc["gr"], ext_hint=ext - 1) as st:
            with pytest.raises(B.BamsignalsError) as e:
                (st.coverage if cov else st.pileup)(**ca)
            assert e.value.code == -8 and "halo" in str(e.value)
