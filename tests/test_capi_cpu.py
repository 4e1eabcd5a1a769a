"""CPU-only checks of the C-ABI library: it loads, exports every symbol include/bamsignals_cuda.h declares, its
pure-host helpers work, and the compute entry points fail loudly (never fall back) when no GPU is present."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import bamsignals_b200 as B
from bamsignals_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "bamsignals_cuda.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bsg_[a-z_]+)\s*\(", src)))


def test_header_symbols_exported():
    L = B.lib()
    syms = declared_symbols()
    assert {"bsg_pileup", "bsg_coverage", "bsg_last_error", "bsg_get_timings", "bsg_shutdown", "bsg_version",
            "bsg_output_layout", "bsg_stage_open", "bsg_pileup_staged", "bsg_coverage_staged", "bsg_stage_close",
            "bsg_device_count"} <= set(syms)
    for s in syms:
        assert hasattr(L, s), f"{s} declared in the header but not exported"
    assert b"sm_100a" in L.bsg_version()


def test_struct_sizes_match_header():
    # bsg_opts: 4+4+64+4 (+4 pad) +8 +4*3 +32 ; the library rejects a struct_size smaller than its own
    assert C.sizeof(api.BsgOpts) == 136
    assert C.sizeof(api.BsgTimings) == 9 * 8 + 8 + 11 * 8 + 8 * 8


def test_output_layout_matches_python():
    w = np.array([0, 1, 5, 200, 201, 1000], dtype=np.int32)
    for binsize in (-1, 1, 7, 200):
        for ss in (False, True):
            off = np.zeros(len(w) + 1, dtype=np.int64)
            tot = B.lib().bsg_output_layout(len(w), api._p(w, C.c_int32), binsize, int(ss), api._p(off, C.c_int64))
            want = api.output_layout(w, binsize, ss)
            assert tot == want[-1] and np.array_equal(off, want)


def test_errors_before_any_gpu_work(fixture_bam, tmp_path):
    """File / index / chromosome errors carry the reference's messages (src/bamsignals.cpp:204,209,119)."""
    with pytest.raises(B.BamsignalsError, match="Fail to open BAM file") as e:
        B.bamCount(str(tmp_path / "nope.bam"), B.GRanges(["chr1"], [1], [10]))
    assert e.value.code == -1
    import shutil
    shutil.copyfile(fixture_bam, tmp_path / "noidx.bam")
    with pytest.raises(B.BamsignalsError, match="BAM indexing file is not available for file") as e:
        B.bamCount(str(tmp_path / "noidx.bam"), B.GRanges(["chr1"], [1], [10]))
    assert e.value.code == -2
    with pytest.raises(B.BamsignalsError, match="chromosome chrZ not present in the bam file") as e:
        B.bamCount(fixture_bam, B.GRanges(["chrZ"], [1], [10]))
    assert e.value.code == -3
    (tmp_path / "junk.bam").write_bytes(b"this is not a bam file at all, not even gzip" * 10)
    with pytest.raises(B.BamsignalsError) as e:
        B.bamCount(str(tmp_path / "junk.bam"), B.GRanges(["chr1"], [1], [10]))
    assert e.value.code == -4


def test_no_cpu_fallback(fixture_bam):
    if B.lib().bsg_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(B.BamsignalsError, match="no CUDA device") as e:
        B.bamCount(fixture_bam, B.GRanges(["chr1"], [1], [10]))
    assert e.value.code == -6


def test_r_level_argument_checks(fixture_bam):
    gr = B.GRanges(["chr1"], [1], [10])
    with pytest.raises(ValueError, match="binsize greater or equal to 1"):
        B.bamProfile(fixture_bam, gr, binsize=0)
    with pytest.raises(ValueError, match="tlenFilter must be NULL or vector of 2 positive integers"):
        B.bamCount(fixture_bam, gr, paired_end="filter", tlenFilter=(1, 2, 3))
    with pytest.raises(ValueError, match="smaller or equal"):
        B.bamCount(fixture_bam, gr, paired_end="filter", tlenFilter=(5, 2))
    with pytest.raises(ValueError, match="should be one of"):
        B.bamCoverage(fixture_bam, gr, paired_end="midpoint")
    with pytest.raises(TypeError, match="must provide a GRanges object"):
        B.pileup_core(fixture_bam, "chr1:1-10", None)


def test_countsignals_container():
    """tests/testthat/test_CountSignals.R:20-64 restated."""
    sig = [np.arange(6, dtype=np.int32).reshape((2, 3), order="F"), np.arange(6, 12, dtype=np.int32).reshape((2, 3), order="F")]
    cs = B.CountSignals(sig, True)
    assert len(cs) == 2 and cs.width().tolist() == [3, 3]
    assert cs[0] is sig[0] and len(cs[[0, 1]]) == 2 and len(cs[np.array([True, False])]) == 1
    with pytest.raises(IndexError):
        cs[-1]
    with pytest.raises(IndexError):
        cs[5]
    assert cs.alignSignals().shape == (2, 3, 2)
    assert "strand-specific" in repr(cs)
    cu = B.CountSignals([np.zeros(4, np.int32), np.zeros(5, np.int32)], False)
    with pytest.raises(ValueError):
        cu.alignSignals()
    with pytest.raises(ValueError):
        B.CountSignals([np.zeros(4, np.int32)], True)
