"""The R shim (rshim/bamsignals_shim.cpp) compiled and RUN without R: tests/mock_rcpp/Rcpp.h provides the handful of
Rcpp types the shim uses, tests/rshim_mock_driver.cpp builds a GRanges-shaped S4 object (factor Rles for seqnames and
strand, levels deliberately in another order than the BAM header) and prints what the shim returns.

CPU: against a recording stub of the C ABI (tests/mock_rcpp/stub_bsg.cpp) the shim must pass exactly the arguments the
reference's pileup_core / coverage_core would have used (src/bamsignals.cpp:92-135, 444-494) and return allocateList's
layout (:139-192); against the real library it must re-raise the reference's error strings.  GPU: the shim over the real
library returns the oracle's counts.  (The file sorts last on purpose: it needs a C++ compiler on the test box.)"""
import json
import os
import subprocess

import numpy as np
import pytest

import bamsignals_b200 as B
import oracle_api as O
import spec_r

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMMON = ["g++", "-std=c++17", "-O1", "-g", "-Wall", "-I", os.path.join(ROOT, "tests", "mock_rcpp"), "-I", os.path.join(ROOT, "include"),
          os.path.join(ROOT, "tests", "rshim_mock_driver.cpp"), os.path.join(ROOT, "rshim", "bamsignals_shim.cpp")]


@pytest.fixture(scope="module")
def stub_driver(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("rshim") / "driver_stub")
    subprocess.check_call(COMMON + ["-fsanitize=address,undefined", os.path.join(ROOT, "tests", "mock_rcpp", "stub_bsg.cpp"), "-o", exe])
    return exe


@pytest.fixture(scope="module")
def real_driver():
    """Built in-tree (tests/_build/, git-ignored) under a name that hashes its sources, so that a binary built on the
    CPU box travels to the GPU box with the snapshot and is reused there; $ORIGIN keeps the rpath relocatable."""
    import hashlib
    srcs = COMMON[-2:] + [os.path.join(ROOT, "tests", "mock_rcpp", "Rcpp.h"), os.path.join(ROOT, "include", "bamsignals_cuda.h")]
    h = hashlib.sha1(b"".join(open(f, "rb").read() for f in srcs)).hexdigest()[:12]
    out = os.path.join(ROOT, "tests", "_build")
    os.makedirs(out, exist_ok=True)
    exe = os.path.join(out, f"rshim_driver_real_{h}")
    if not os.path.exists(exe):
        libdir = os.path.join(ROOT, "bamsignals_b200")
        r = subprocess.run(COMMON + ["-L", libdir, "-lbamsignals_cuda", "-Wl,-rpath,$ORIGIN/../../bamsignals_b200", "-o", exe + ".tmp"],
                           capture_output=True, text=True)
        if r.returncode != 0:
            pytest.skip("cannot link the shim against libbamsignals_cuda.so here: " + r.stderr[-300:])
        os.replace(exe + ".tmp", exe)
    return exe


def call(exe, tmp_path, fn, bam, regions, mapqual=0, binsize=1, shift=0, ss=False, requiredF=0, filteredF=-1, flag=0, tlen=None,
         log=None):
    spec = tmp_path / "spec.txt"
    t = "" if tlen is None else f" {tlen[0]} {tlen[1]}"
    lines = [f"{fn} {bam} {mapqual} {binsize} {shift} {int(ss)} {requiredF} {filteredF} {int(flag)} {0 if tlen is None else 2}{t} {len(regions)}"]
    lines += [f"{n} {s} {w} {st}" for n, s, w, st in regions]
    spec.write_text("\n".join(lines) + "\n")
    env = dict(os.environ)
    if log:
        env["BSG_STUB_LOG"] = str(log)
    r = subprocess.run([exe, str(spec)], capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0, r.stderr[-2000:]
    return json.loads(r.stdout)


REGIONS = [("chr2", 101, 25, "-"), ("chr2", 1, 7, "+"), ("chr1", 50, 30, "*"), ("chr3", 10, 0, "+"), ("chr3", 700, 41, "-")]


def test_shim_passes_the_reference_arguments(stub_driver, tmp_path):
    log = tmp_path / "calls.jsonl"
    out = call(stub_driver, tmp_path, "pileup", "x.bam", REGIONS, mapqual=5, binsize=10, shift=75, ss=True, requiredF=66,
               filteredF=1024, flag=1, tlen=(0, 1000), log=log)
    out_c = call(stub_driver, tmp_path, "coverage", "x.bam", REGIONS, mapqual=7, requiredF=66, filteredF=-1, flag=1, tlen=(70, 200), log=log)
    out_n = call(stub_driver, tmp_path, "pileup", "x.bam", REGIONS, binsize=-1, ss=False, log=log)
    recs = [json.loads(x) for x in log.read_text().splitlines()]
    gr = B.GRanges([r[0] for r in REGIONS], [r[1] for r in REGIONS], [r[2] for r in REGIONS], [r[3] for r in REGIONS])
    for rec in recs:
        # names, not level codes, identify the chromosome (src/bamsignals.cpp:117); loc = start - 1 (:131); strand +1/-1/0 (:126-128)
        got = [(rec["levels"][si], loc + 1, w, {1: "+", -1: "-", 0: "*"}[st]) for si, loc, w, st in rec["regions"]]
        assert got == REGIONS
        assert rec["levels"] != gr.seqlevels                       # the driver really permutes the factor levels
        assert rec["out_null"] == 1 and rec["opts_null"] == 1 and rec["maxgap"] == 16385
    p, c, n = recs
    assert (p["fn"], p["tlen"], p["mapqual"], p["binsize"], p["shift"], p["ss"], p["requiredF"], p["filteredF"], p["pe_mid"]) == \
        ("bsg_pileup", [0, 1000], 5, 10, 75, 1, 66, 1024, 1)
    assert (c["fn"], c["tlen"], c["mapqual"], c["requiredF"], c["filteredF"], c["tspan"]) == ("bsg_coverage", [70, 200], 7, 66, -1, 1)
    assert n["tlen"] is None and n["binsize"] == -1 and n["ss"] == 0          # integer() => NULL filter (:456)
    # allocateList's layout (:139-192): per-region (2, ceil(w/binsize)) matrices with dimnames when ss ...
    widths = np.array([r[2] for r in REGIONS])
    assert [e["dim"] for e in out["result"]] == [[2, int(-(-w // 10))] for w in widths]
    assert all(e["dimnames0"] == ["sense", "antisense"] for e in out["result"])
    assert p["total"] == int(B.api.output_layout(widths.astype(np.int32), 10, True)[-1])
    for i, e in enumerate(out["result"]):                          # ... each region got ITS slice
        assert e["v"] == [1000 * i + k for k in range(2 * int(-(-widths[i] // 10)))]
    # ... plain vectors of width(gr) for coverage ...
    assert [e["dim"] for e in out_c["result"]] == [[int(w)] for w in widths] and all(e["dimnames0"] == [] for e in out_c["result"])
    # ... and ONE vector with an element per region for bamCount (:148-169)
    assert len(out_n["result"]) == 1 and out_n["result"][0]["dim"] == [len(REGIONS)]
    assert out_n["result"][0]["v"] == [1000 * i for i in range(len(REGIONS))]


def test_shim_bamcount_matrix_and_errors(stub_driver, tmp_path):
    log = tmp_path / "calls.jsonl"
    out = call(stub_driver, tmp_path, "pileup", "x.bam", REGIONS[:3], binsize=-1, ss=True, log=log)
    e = out["result"][0]
    assert e["dim"] == [2, 3] and e["dimnames0"] == ["sense", "antisense"] and e["v"] == [0, 1, 1000, 1001, 2000, 2001]
    assert call(stub_driver, tmp_path, "pileup", "fail.bam", REGIONS[:3], log=log) == {"error": "Fail to open BAM file fail.bam"}
    assert call(stub_driver, tmp_path, "pileup", "x.bam", REGIONS[:3], flag=99, log=log) == {"error": "must provide a GRanges object"}   # :94
    assert call(stub_driver, tmp_path, "coverage", "x.bam", REGIONS[:3], flag=99, log=log) == {"error": "must provide a GRanges object"}
    out = call(stub_driver, tmp_path, "writesam", "in.sam", [("out.bam", 1, 1, "+")], log=log)
    assert out["ok"] is True
    assert json.loads(log.read_text().splitlines()[-1]) == {"fn": "bsg_write_sam_as_bam_and_index", "sam": "in.sam", "bam": "out.bam"}


def test_shim_reraises_the_library_errors(real_driver, tmp_path, fixture_bam):
    missing = str(tmp_path / "nope.bam")
    assert call(real_driver, tmp_path, "pileup", missing, REGIONS) == {"error": "Fail to open BAM file " + missing}            # :204
    noidx = str(tmp_path / "noindex.bam")
    open(noidx, "wb").write(open(fixture_bam, "rb").read())
    assert call(real_driver, tmp_path, "coverage", noidx, REGIONS) == {"error": "BAM indexing file is not available for file " + noidx}   # :209
    bad = [("chrZ", 1, 10, "+")]
    assert call(real_driver, tmp_path, "pileup", fixture_bam, bad) == {"error": "chromosome chrZ not present in the bam file"}  # :119
    if B.lib().bsg_device_count() < 1:
        assert "no CUDA device" in call(real_driver, tmp_path, "pileup", fixture_bam, REGIONS)["error"]


@pytest.mark.gpu
def test_shim_over_the_real_library_matches_the_oracle(real_driver, tmp_path, fixture_bam):
    g = spec_r.test_regions(seed=21, n=30)
    names = [["chr1", "chr2", "chr3"][i] for i in g["rname"]]
    regions = list(zip(names, g["start"], g["width"], g["strand"])) + [("chr2", 1, 10279, "-"), ("chr1", 300, 0, "*")]
    gr = B.GRanges([r[0] for r in regions], [r[1] for r in regions], [r[2] for r in regions], [r[3] for r in regions])
    # bamCount(ss=TRUE, shift=75, paired.end="midpoint", filteredFlag=1024)
    out = call(real_driver, tmp_path, "pileup", fixture_bam, regions, mapqual=10, binsize=-1, shift=75, ss=True, requiredF=66,
               filteredF=1024, flag=1, tlen=(0, 1000))
    want = O.bamCount(fixture_bam, gr, mapqual=10, shift=75, ss=True, paired_end="midpoint", filteredFlag=1024)
    assert out["result"][0]["dim"] == [2, len(regions)] and out["result"][0]["v"] == want.ravel(order="F").tolist()
    # bamProfile(binsize=20, ss=TRUE)
    out = call(real_driver, tmp_path, "pileup", fixture_bam, regions, binsize=20, ss=True)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want = O.bamProfile(fixture_bam, gr, binsize=20, ss=True).as_list()
    assert [e["v"] for e in out["result"]] == [w.ravel(order="F").tolist() for w in want]
    assert [e["dim"] for e in out["result"]] == [list(w.shape) for w in want]
    # bamCoverage(paired.end="extend")
    out = call(real_driver, tmp_path, "coverage", fixture_bam, regions, requiredF=66, flag=1, tlen=(0, 1000))
    want = O.bamCoverage(fixture_bam, gr, paired_end="extend").as_list()
    assert [e["v"] for e in out["result"]] == [w.tolist() for w in want]
