"""CPU unit test of the GPU inflate kernel's decode core: bamsignals_b200/csrc/inflate_core.cuh (bit reader, decode
tables, symbols -> token queue) is compiled as plain C++ into tests/host_inflate_harness.cpp, which emulates the
warp-cooperative materialisation and must reproduce the CRC32 + ISIZE of every BGZF block of a file."""
import os
import subprocess
import sys

import pytest

import bamwriter as W
import edge_cases as E

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import workloads as WL  # noqa: E402


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("harness") / "harness")
    subprocess.check_call(["g++", "-O1", "-g", "-fsanitize=address,undefined", "-o", exe,
                           os.path.join(ROOT, "tests", "host_inflate_harness.cpp"), "-lz"])
    return exe


def run(exe, path):
    r = subprocess.run([exe, path], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert " bad 0" in r.stdout
    return int(r.stdout.split()[1])


def test_reference_fixture(harness, fixture_bam):
    assert run(harness, fixture_bam) == 72          # 72 BGZF blocks


@pytest.mark.parametrize("level,straddle,payload", [(0, False, 3000), (1, True, 997), (6, False, 0xFF00), (9, True, 20000), (6, False, 40)])
def test_deflate_variants(harness, tmp_path, level, straddle, payload):
    """stored blocks, fixed-Huffman blocks (tiny payloads), dynamic blocks at several levels"""
    p = str(tmp_path / "v.bam")
    W.write_bam(p, E.REFS, E.variety_reads(n=1500), block_payload=payload, cut_mid_record=straddle, level=level)
    assert run(harness, p) > 3


def test_generator_output(harness, tmp_path):
    bam, _ = WL.make_bam("c3", 0.002, str(tmp_path), record="realistic")
    assert run(harness, bam) > 10


def test_fuzzed_streams_stay_in_bounds(harness, fixture_bam, tmp_path):
    """Damaged DEFLATE payloads (bit flips, random byte runs, truncation, header bytes): the decode core the GPU kernel
    shares must reject them, or leave the damage to the CRC32 check, without ever leaving its buffers (ASan / UBSan
    abort the harness otherwise) and without hanging."""
    p = str(tmp_path / "v.bam")
    W.write_bam(p, E.REFS, E.variety_reads(n=1500), block_payload=20000, level=6)
    for path, n in ((fixture_bam, 1500), (p, 1500)):
        r = subprocess.run([harness, "--fuzz", str(n), path], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout + r.stderr
        assert "runtime error" not in r.stderr and "AddressSanitizer" not in r.stderr, r.stderr
        out = r.stdout.split()
        assert out[0] == "fuzz" and int(out[1]) == n
        assert int(out[3]) > n // 2            # most damage is caught by the decoder itself
