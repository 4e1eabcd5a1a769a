"""Malformed input must be refused, never crash: the library's BAM-header / BGZF / .bai / .csi readers and the fetch
planner (bamsignals_b200/csrc/bamio.cpp, plan.cpp) are compiled into tests/host_plan_harness.cpp with
-fsanitize=address,undefined and fed mutated files.  The harness must exit 0 every time - either "ok" or a clean
"error <code>" - within a few seconds and a bounded amount of memory ("truncated/corrupt BGZF must
error, not crash").  CPU only."""
import os
import shutil
import struct
import subprocess

import numpy as np
import pytest

import bamwriter as W
import edge_cases as E

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "bamsignals_b200", "csrc")
EXTREMES = [0, 1, 0x7F, 0x80, 0xFF, 0x7FFF, 0x8000, 0xFFFF, 0x7FFFFFFF, 0x80000000, 0xFFFFFFFF, 0xFFFFFFFE, 37450, 4681]


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("planharness") / "plan_harness")
    subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined",
                           "-o", exe, os.path.join(ROOT, "tests", "host_plan_harness.cpp"), os.path.join(CSRC, "bamio.cpp"),
                           os.path.join(CSRC, "plan.cpp"), "-lz", "-lpthread"])
    return exe


@pytest.fixture(scope="module")
def seeds(tmp_path_factory):
    """Small valid inputs: (bam bytes, index suffix, index bytes)."""
    d = tmp_path_factory.mktemp("seeds")
    out = []
    for tag, kw in (("bai", dict(index="bai")), ("csi", dict(index="csi", min_shift=12, depth=4))):
        p = str(d / f"{tag}.bam")
        W.write_bam(p, E.REFS, E.variety_reads(seed=3, n=400), block_payload=2000, cut_mid_record=True, **kw)
        out.append((open(p, "rb").read(), "." + tag, open(p + "." + tag, "rb").read()))
    return out


def run(exe, bam, ext=50):
    env = dict(os.environ, ASAN_OPTIONS="max_allocation_size_mb=3072:allocator_may_return_null=0:detect_leaks=1")
    r = subprocess.run([exe, bam, str(ext)], capture_output=True, text=True, errors="replace", timeout=60, env=env)
    assert r.returncode == 0, (r.returncode, r.stdout[-500:], r.stderr[-3000:])
    assert r.stdout.startswith("ok ") or r.stdout.startswith("error -"), r.stdout
    return r.stdout


def mutate(rng, data: bytes, start=0) -> bytes:
    b = bytearray(data)
    kind = rng.integers(0, 5)
    if kind == 0:                                   # flip a few bytes
        for _ in range(int(rng.integers(1, 4))):
            b[int(rng.integers(start, len(b)))] ^= 1 << int(rng.integers(0, 8))
    elif kind == 1:                                 # overwrite an aligned 32-bit field with an extreme value
        o = start + 4 * int(rng.integers(0, max(1, (len(b) - start) // 4 - 1)))
        struct.pack_into("<I", b, o, EXTREMES[int(rng.integers(0, len(EXTREMES)))])
    elif kind == 2:                                 # overwrite an aligned 64-bit field
        o = start + 4 * int(rng.integers(0, max(1, (len(b) - start) // 4 - 2)))
        vals = [0, 1, 1 << 16, (1 << 48) - 1, (1 << 64) - 1, 0x7FFFFFFFFFFFFFFF, len(data) << 16]
        struct.pack_into("<Q", b, o, vals[int(rng.integers(0, len(vals)))])
    elif kind == 3:                                 # truncate
        del b[int(rng.integers(start, len(b))):]
    else:                                           # random garbage run
        o = int(rng.integers(start, len(b)))
        n = int(rng.integers(1, 64))
        b[o:o + n] = rng.integers(0, 256, len(b[o:o + n]), dtype=np.uint8).tobytes()
    return bytes(b)


def test_valid_seeds_plan_everything(harness, seeds, tmp_path):
    for bam, suf, idx in seeds:
        p = str(tmp_path / ("v" + suf + ".bam"))
        open(p, "wb").write(bam); open(p + suf, "wb").write(idx)
        out = run(harness, p, ext=100000)
        assert out.startswith("ok refs 3 ") and " records 800" in out, out      # 2 x 400 placed reads; the 17 unplaced ones are never fetched


def test_large_index_takes_the_radix_sort(harness, tmp_path):
    """An index with more than 16 Ki entry points (the radix-sort path of load_index): whole-contig plans must still
    reach every placed record, and the harness checks that the entry points are strictly ascending."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import workloads as WL
    bam, info = WL.make_bam("c2", 0.03, str(tmp_path))
    out = run(harness, bam, ext=0)
    f = out.split()
    assert f[0] == "ok" and int(f[f.index("records") + 1]) == info["records"], out
    assert int(f[f.index("entries") + 1]) > (1 << 14), out


@pytest.mark.parametrize("which", ["bai", "csi"])
def test_mutated_index(harness, seeds, tmp_path, which):
    bam, suf, idx = seeds[0 if which == "bai" else 1]
    if which == "csi":                              # mutate the uncompressed CSI, then BGZF-wrap it again
        import zlib
        raw, o = b"", 0
        while o < len(idx):
            bs = struct.unpack_from("<H", idx, o + 16)[0] + 1
            raw += zlib.decompress(idx[o + 18:o + bs - 8], -15)
            o += bs
    rng = np.random.default_rng(11)
    p = str(tmp_path / "m.bam")
    open(p, "wb").write(bam)
    refused = 0
    for k in range(150):
        if which == "bai":
            m = mutate(rng, idx, start=4 if k % 3 else 0)
        else:
            r = mutate(rng, raw, start=4 if k % 3 else 0)
            m = b"".join(W.bgzf_block(r[i:i + 0xFF00]) for i in range(0, max(1, len(r)), 0xFF00)) + W.EOF_BLOCK
            if k % 10 == 9:
                m = mutate(rng, idx)               # damage the BGZF container itself
        open(p + suf, "wb").write(m)
        refused += run(harness, p).startswith("error")
    assert refused > 10                             # the fuzzer does hit the checks


def test_mutated_bam_header_and_blocks(harness, seeds, tmp_path):
    bam, suf, idx = seeds[0]
    rng = np.random.default_rng(5)
    p = str(tmp_path / "h.bam")
    open(p + suf, "wb").write(idx)
    refused = 0
    for k in range(120):
        open(p, "wb").write(mutate(rng, bam, start=0 if k % 2 else 18))
        refused += run(harness, p).startswith("error")
    assert refused > 10


def test_crafted_headers(harness, tmp_path):
    """Length fields of the BAM header and the index set to values that must not be trusted."""
    def header(l_text, n_ref, l_name, text=b"@HD\tVN:1.6\tSO:coordinate\n"):
        h = b"BAM\1" + struct.pack("<i", l_text) + text + struct.pack("<i", n_ref)
        return h + struct.pack("<i", l_name) + b"chrA\0" + struct.pack("<i", 1000)
    t = b"@HD\tVN:1.6\tSO:coordinate\n"
    cases = [header(len(t), 1, 5), header(-1, 1, 5), header(0x7FFFFFFF, 1, 5), header(len(t), 0x7FFFFFFF, 5),
             header(len(t), -5, 5), header(len(t), 1, 0x7FFFFFFF), header(len(t), 1, -1), header(len(t), 1, 0),
             header(len(t), 2, 5), b"BAM\1", b"", b"CRAM" + bytes(40)]
    bai_ok = b"BAI\1" + struct.pack("<i", 1) + struct.pack("<i", 0) + struct.pack("<i", 0)
    idx_cases = [bai_ok, b"BAI\1" + struct.pack("<i", 0x7FFFFFFF), b"BAI\1" + struct.pack("<ii", 1, 0x7FFFFFFF),
                 b"BAI\1" + struct.pack("<iiIi", 1, 1, 4681, 0x7FFFFFFF), b"BAI\1" + struct.pack("<iii", 1, 0, 0x7FFFFFFF),
                 b"CSI\1" + struct.pack("<iiii", 14, 9, 0, 0x7FFFFFFF), b"CSI\1" + struct.pack("<iii", 30, 9, 0),
                 b"CSI\1" + struct.pack("<iiii", 1, 9, 0, 1) + struct.pack("<iIQi", 1, 19173960, 1 << 16, 0),
                 b"CSI\1" + struct.pack("<iiii", 14, 5, 0x7FFFFFFF, 1), b"BAI", b"XXXX" + bytes(16)]
    p = str(tmp_path / "c.bam")
    outs = []
    for hb in cases:
        open(p, "wb").write((W.bgzf_block(hb) if hb[:4] != b"CRAM" else hb) + W.EOF_BLOCK)
        open(p + ".bai", "wb").write(bai_ok)
        outs.append(run(harness, p))
    assert outs[0].startswith("ok refs 1 ") and all(o.startswith("error -4") for o in outs[1:7] + outs[8:]), outs
    open(p, "wb").write(W.bgzf_block(cases[0]) + W.EOF_BLOCK)
    for ib in idx_cases:
        open(p + ".bai", "wb").write(ib)
        outs.append(run(harness, p))
    io = outs[len(cases):]
    assert io[0].startswith("ok refs 1 ") and io[7].startswith("ok refs 1 ")      # io[7]: a far-away leaf bin is ignored
    assert all(o.startswith("error -4") for o in io[1:7] + io[8:]), io


# ---- the SAM -> BAM + BAI writer (bamwrite.cpp) --------------------------------------------------------------------------
@pytest.fixture(scope="module")
def writer_harness(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("writerharness") / "writer_harness")
    subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined",
                           "-o", exe, os.path.join(ROOT, "tests", "host_writer_harness.cpp"), os.path.join(CSRC, "bamwrite.cpp"), "-lz"])
    return exe


SAM = ("@HD\tVN:1.6\tSO:coordinate\n@SQ\tSN:chrA\tLN:100000\n@SQ\tSN:chrB\tLN:5000\n"
       "r1\t99\tchrA\t100\t30\t5S20M3D2I10M\t=\t300\t250\tACGTACGTACGTACGTACGTACGTACGTACGTACGTA\tIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIII\tNM:i:3\tXS:Z:hello\tXA:A:c\tXF:f:1.5\tXB:B:c,1,-2,3\n"
       "r2\t147\tchrA\t300\t30\t30M\t=\t100\t-250\t*\t*\tXH:H:1AE3\tXI:B:I,70000,1\n"
       "r3\t16\tchrA\t70000\t0\t10M5000N10M\t*\t0\t0\t*\t*\n"
       "r4\t0\tchrB\t1\t255\t*\t*\t0\t0\t*\t*\n"
       "r5\t4\t*\t0\t0\t*\t*\t0\t0\tACGT\t!!!!\n")


def run_writer(exe, sam, bam):
    env = dict(os.environ, ASAN_OPTIONS="max_allocation_size_mb=3072:detect_leaks=1")
    r = subprocess.run([exe, sam, bam], capture_output=True, text=True, errors="replace", timeout=60, env=env)
    assert r.returncode == 0, (r.returncode, r.stdout[-300:], r.stderr[-3000:])
    assert r.stdout.startswith("ok") or r.stdout.startswith("error -"), r.stdout
    return r.stdout


def test_writer_on_mutated_sam(writer_harness, harness, tmp_path):
    sam, bam = str(tmp_path / "m.sam"), str(tmp_path / "m.bam")
    open(sam, "w").write(SAM)
    assert run_writer(writer_harness, sam, bam).startswith("ok")
    assert run(harness, bam).startswith("ok refs 2 ") and " records 4" in run(harness, bam)    # our reader accepts our writer's output
    rng = np.random.default_rng(23)
    raw = SAM.encode()
    tokens = [b"\t", b"\n", b"*", b"-", b"99999999999999999999", b"2147483648", b"-2147483649", b"0M", b"65536M", b"=", b"@SQ\tSN:chrA\tLN:-1\n",
              b"XX:B:i,", b"XX:Z:", b"XX:i:", b"", b"\x00", b"\xff", b"chrB", b"chrC", b"536870913", b"1S" * 40000]
    refused = 0
    for k in range(250):
        b = bytearray(raw)
        for _ in range(int(rng.integers(1, 4))):
            kind = int(rng.integers(0, 4))
            o = int(rng.integers(0, len(b)))
            if kind == 0:
                b[o] = int(rng.integers(0, 256))
            elif kind == 1:
                del b[o:o + int(rng.integers(1, 12))]
            elif kind == 2:
                b[o:o] = tokens[int(rng.integers(0, len(tokens)))]
            else:                                   # swap two lines (unsorted input must be refused, not indexed)
                lines = bytes(b).split(b"\n")
                i, j = int(rng.integers(0, len(lines))), int(rng.integers(0, len(lines)))
                lines[i], lines[j] = lines[j], lines[i]
                b = bytearray(b"\n".join(lines))
        open(sam, "wb").write(bytes(b))
        out = run_writer(writer_harness, sam, bam)
        refused += out.startswith("error")
        if out.startswith("ok"):                    # whatever the writer accepts, the reader must be able to plan
            assert run(harness, bam).startswith("ok"), bytes(b)
    assert refused > 25
