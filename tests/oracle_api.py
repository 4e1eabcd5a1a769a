"""ctypes binding of the CPU oracle with the same call shapes as bamsignals_b200.api.  Two implementations:
  impl="port"  oracle/liboracle.so: the restated reference algorithm (oracle/bsg_oracle.cpp)
  impl="ref"   oracle/_ref/libbamsignals_ref.so: the reference's OWN src/bamsignals.cpp compiled unchanged against stand-ins
               for <Rcpp.h> and <htslib/*.h> (oracle/Makefile, oracle/ref_compat/); built where /root/reference exists
TEST INFRASTRUCTURE: imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm."""
import ctypes as C
import os

import numpy as np

from bamsignals_b200.api import (CountSignals, GRanges, _match_arg, _p, _tlen, _trunc, flagMask, marshal_regions,
                                 output_layout, split_signals)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_PATH = os.path.join(ROOT, "oracle", "liboracle.so")
_REF_PATH = os.path.join(ROOT, "oracle", "_ref", "libbamsignals_ref.so")
_lib = None
_ref = None
SCAN, INDEXED, BRUTE = 0, 1, 2
_last_impl = "port"


def _sigs(L, prefix):
    i32p, i64p, i8p = C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_int8)
    reg = [C.c_char_p, C.c_int64, C.POINTER(C.c_char_p), C.c_int32, i32p, i32p, i32p, i8p]
    getattr(L, prefix + "_pileup").argtypes = reg + [i32p] + [C.c_int32] * 8 + [i32p, i64p, C.c_int32, C.c_int32]
    getattr(L, prefix + "_coverage").argtypes = reg + [i32p] + [C.c_int32] * 5 + [i32p, i64p, C.c_int32, C.c_int32]
    getattr(L, prefix + "_last_error").restype = C.c_char_p
    getattr(L, prefix + "_stats").argtypes = [C.POINTER(C.c_uint64)] * 3


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(_PATH)
        i32p = C.POINTER(C.c_int32)
        _sigs(L, "oracle")
        L.oracle_dump_reads.argtypes = [C.c_char_p, C.c_int64] + [i32p] * 6
        L.oracle_dump_reads.restype = C.c_int64
        L.oracle_header.argtypes = [C.c_char_p, C.c_int32, C.c_char_p, i32p]
        _lib = L
    return _lib


def ref_available():
    return os.path.exists(_REF_PATH)


def ref_lib():
    """The reference's own engine (oracle/_ref); raises if it was not built (it needs /root/reference at build time)."""
    global _ref
    if _ref is None:
        if not ref_available():
            raise OracleError("oracle/_ref/libbamsignals_ref.so is missing: run `make -C oracle ref` where /root/reference exists")
        L = C.CDLL(_REF_PATH)
        _sigs(L, "ref")
        _ref = L
    return _ref


class OracleError(RuntimeError):
    pass


def _entry(impl, what):
    global _last_impl
    if impl not in ("port", "ref"):
        raise ValueError("impl must be 'port' or 'ref'")
    _last_impl = impl
    L, prefix = (ref_lib(), "ref") if impl == "ref" else (lib(), "oracle")

    def call(*a):
        if getattr(L, prefix + "_" + what)(*a) != 0:
            raise OracleError(getattr(L, prefix + "_last_error")().decode())
    return call


def pileup_core(bampath, gr, tlen_filter, mapqual=0, binsize=1, shift=0, ss=False, requiredF=0, filteredF=-1,
                pe_mid=False, maxgap=16385, mode=INDEXED, nthreads=1, impl="port"):
    m = marshal_regions(gr)
    off = output_layout(m.width, int(binsize), bool(ss))
    flat = np.full(int(off[-1]), -77, dtype=np.int32)
    tl = None if tlen_filter is None else np.asarray(tlen_filter, dtype=np.int32)
    _entry(impl, "pileup")(os.fsencode(bampath), m.R, m.levels, m.n_levels, _p(m.seq_idx, C.c_int32),
                           _p(m.loc, C.c_int32), _p(m.width, C.c_int32), _p(m.strand, C.c_int8),
                           None if tl is None else _p(tl, C.c_int32), int(mapqual), int(binsize), int(shift),
                           int(bool(ss)), int(requiredF), int(filteredF), int(bool(pe_mid)), int(maxgap),
                           _p(flat, C.c_int32), _p(off, C.c_int64), mode, nthreads)
    if binsize <= 0:
        return [flat.reshape((2, -1), order="F") if ss else flat]
    return split_signals(flat, off, bool(ss))


def coverage_core(bampath, gr, tlen_filter, mapqual=0, requiredF=0, filteredF=-1, tspan=False, maxgap=16385,
                  mode=INDEXED, nthreads=1, impl="port"):
    m = marshal_regions(gr)
    off = output_layout(m.width, 1, False)
    flat = np.full(int(off[-1]), -77, dtype=np.int32)
    tl = None if tlen_filter is None else np.asarray(tlen_filter, dtype=np.int32)
    _entry(impl, "coverage")(os.fsencode(bampath), m.R, m.levels, m.n_levels, _p(m.seq_idx, C.c_int32),
                             _p(m.loc, C.c_int32), _p(m.width, C.c_int32), _p(m.strand, C.c_int8),
                             None if tl is None else _p(tl, C.c_int32), int(mapqual), int(requiredF),
                             int(filteredF), int(bool(tspan)), int(maxgap), _p(flat, C.c_int32),
                             _p(off, C.c_int64), mode, nthreads)
    return split_signals(flat, off, False)


def bamCount(bampath, gr, mapqual=0, shift=0, ss=False, paired_end="ignore", tlenFilter=None, filteredFlag=-1, **kw):
    pe = _match_arg(paired_end, ("ignore", "filter", "midpoint"))
    return pileup_core(bampath, gr, _tlen(tlenFilter, pe), _trunc(mapqual), -1, _trunc(shift), ss, flagMask(pe),
                       _trunc(filteredFlag), pe == "midpoint", **kw)[0]


def bamProfile(bampath, gr, binsize=1, mapqual=0, shift=0, ss=False, paired_end="ignore", tlenFilter=None,
               filteredFlag=-1, **kw):
    pe = _match_arg(paired_end, ("ignore", "filter", "midpoint"))
    return CountSignals(pileup_core(bampath, gr, _tlen(tlenFilter, pe), _trunc(mapqual), _trunc(binsize),
                                    _trunc(shift), ss, flagMask(pe), _trunc(filteredFlag), pe == "midpoint", **kw), ss)


def bamCoverage(bampath, gr, mapqual=0, paired_end="ignore", tlenFilter=None, filteredFlag=-1, **kw):
    pe = _match_arg(paired_end, ("ignore", "extend"))
    return CountSignals(coverage_core(bampath, gr, _tlen(tlenFilter, pe), _trunc(mapqual), flagMask(pe),
                                      _trunc(filteredFlag), pe == "extend", **kw), False)


def dump_reads(bampath, cap=1 << 24):
    cols = [np.empty(cap, dtype=np.int32) for _ in range(6)]
    n = lib().oracle_dump_reads(os.fsencode(bampath), cap, *[_p(c, C.c_int32) for c in cols])
    if n < 0:
        raise OracleError(lib().oracle_last_error().decode())
    if n > cap:
        return dump_reads(bampath, int(n))
    return dict(zip(("tid", "pos", "endpos", "tlen", "flag", "mapq"), [c[:n] for c in cols]))


def stats():
    """Counters of the last call (records the readers handed out, bytes inflated, index queries)."""
    a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
    if _last_impl == "ref":
        ref_lib().ref_stats(C.byref(a), C.byref(b), C.byref(c))
    else:
        lib().oracle_stats(C.byref(a), C.byref(b), C.byref(c))
    return dict(records=a.value, bytes_inflated=b.value, queries=c.value)


def header(bampath, cap=4096):
    names = C.create_string_buffer(256 * cap)
    lens = np.zeros(cap, dtype=np.int32)
    n = lib().oracle_header(os.fsencode(bampath), cap, names, _p(lens, C.c_int32))
    if n < 0:
        raise OracleError(lib().oracle_last_error().decode())
    return [names.raw[256 * i:256 * (i + 1)].split(b"\0")[0].decode() for i in range(n)], lens[:n].tolist()
