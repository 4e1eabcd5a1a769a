"""Host-side mirror of the bamsignals R interface for the counting path, on top of libbamsignals_cuda.so.

R is not installable offline, so the reference's R layer (R/wrappers.R, R/zzzCountSignals.R) is mirrored here in
Python with the same names, argument meaning, defaults and error behaviour; the compute goes through the C ABI of
include/bamsignals_cuda.h exactly as the R shim (rshim/) would.  There is no CPU fallback: if the
CUDA library is missing or no GPU is present the calls raise.

  bamCount / bamProfile / bamCoverage    R/wrappers.R:106-173
  pileup_core / coverage_core            R/RcppExports.R:12-18 (.Call boundary, src/bamsignals.cpp:444,474)
  flagMask / tlenFilter helpers          R/wrappers.R:76-98
  CountSignals                           R/zzzCountSignals.R:27-113
  GRanges                                the slots parseRegions reads (src/bamsignals.cpp:92-135)
"""
from __future__ import annotations

import ctypes as C
import os
import warnings
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

_LIB_PATH = os.environ.get("BSG_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libbamsignals_cuda.so")
_lib = None


class BamsignalsError(RuntimeError):
    """Raised for any non-zero return of the C ABI; .code holds the BSG_E* value."""

    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


# ------------------------------------------------------------------------------------------------------------------
# GRanges: just the slots the native code reads (ranges@start/width, seqnames, strand)
# ------------------------------------------------------------------------------------------------------------------
class GRanges:
    """Minimal stand-in for GenomicRanges::GRanges: seqnames, 1-based start, width, strand in {'+','-','*'}."""

    def __init__(self, seqnames: Sequence[str], start: Sequence[int], width: Sequence[int],
                 strand: Optional[Sequence[str]] = None, seqlevels: Optional[Sequence[str]] = None):
        n = len(start)
        if len(seqnames) != n or len(width) != n:
            raise ValueError("seqnames, start and width must have the same length")
        seqnames = [str(s) for s in seqnames]
        if seqlevels is None:
            seqlevels = list(dict.fromkeys(seqnames))
        self.seqlevels = [str(s) for s in seqlevels]
        lut = {s: i for i, s in enumerate(self.seqlevels)}
        try:
            self.seq_idx = np.fromiter((lut[s] for s in seqnames), dtype=np.int32, count=n)
        except KeyError as e:
            raise ValueError(f"seqname {e} not in seqlevels") from None
        self.start = np.ascontiguousarray(start, dtype=np.int32)
        self.width = np.ascontiguousarray(width, dtype=np.int32)
        if (self.width < 0).any():
            raise ValueError("negative widths are not allowed")
        if strand is None:
            self.strand = np.zeros(n, dtype=np.int8)
        elif isinstance(strand, np.ndarray) and strand.dtype == np.int8:
            self.strand = np.ascontiguousarray(strand)
        else:
            self.strand = np.fromiter(((1 if s == "+" else -1 if s == "-" else 0) for s in strand), dtype=np.int8, count=n)

    @classmethod
    def from_codes(cls, seqlevels, seq_idx, start, width, strand_i8):
        """Vectorised constructor for large region sets (no per-element Python work)."""
        self = cls.__new__(cls)
        self.seqlevels = [str(s) for s in seqlevels]
        self.seq_idx = np.ascontiguousarray(seq_idx, dtype=np.int32)
        self.start = np.ascontiguousarray(start, dtype=np.int32)
        self.width = np.ascontiguousarray(width, dtype=np.int32)
        self.strand = np.ascontiguousarray(strand_i8, dtype=np.int8)
        return self

    def __len__(self):
        return len(self.start)

    def __getitem__(self, idx):
        idx = np.arange(len(self))[idx]
        return GRanges.from_codes(self.seqlevels, self.seq_idx[idx], self.start[idx], self.width[idx], self.strand[idx])

    @property
    def seqnames(self):
        return [self.seqlevels[i] for i in self.seq_idx]


# ------------------------------------------------------------------------------------------------------------------
# CountSignals container (R/zzzCountSignals.R)
# ------------------------------------------------------------------------------------------------------------------
class CountSignals:
    """List-like container: signal i is a (w,) int32 vector, or a (2, w) matrix [sense; antisense] when ss.

    The (2, w) matrices are Fortran-ordered views into the flat result buffer, i.e. the memory layout of R's
    IntegerMatrix(2, w) (src/bamsignals.cpp:178-181): [sense_0, antisense_0, sense_1, ...].  When built from the
    library's flat result (`flat` + `offsets`) the per-region views are created on first use."""

    rownames = ("sense", "antisense")

    def __init__(self, signals: Optional[List[np.ndarray]], ss: bool, flat: Optional[np.ndarray] = None,
                 offsets: Optional[np.ndarray] = None):
        self.ss = bool(ss)
        self._signals = signals
        self._flat, self._offsets = flat, offsets
        if signals is not None:
            for s in signals:                                # checkList, src/CountSignals.cpp:4-16
                if self.ss and not (s.ndim == 2 and s.shape[0] == 2):
                    raise ValueError("strand-specific signals must be matrices with two rows")
                if not self.ss and s.ndim != 1:
                    raise ValueError("signals must be vectors")

    def _view(self, i):
        v = self._flat[self._offsets[i]:self._offsets[i + 1]]
        return v.reshape((2, -1), order="F") if self.ss else v

    @property
    def signals(self):
        if self._signals is None:
            self._signals = [self._view(i) for i in range(len(self._offsets) - 1)]
        return self._signals

    def __len__(self):                                       # R/zzzCountSignals.R:46
        return len(self._signals) if self._signals is not None else len(self._offsets) - 1

    def width(self):                                         # R/zzzCountSignals.R:54-56 (fastWidth)
        if self._signals is None:
            return np.diff(self._offsets) // (2 if self.ss else 1)
        return np.array([s.shape[-1] for s in self._signals], dtype=np.int64)

    def __getitem__(self, i):                                # R/zzzCountSignals.R:68-77
        if isinstance(i, (int, np.integer)):
            if i < 0 or i >= len(self):
                raise IndexError("subscript out of bounds")
            return self._signals[i] if self._signals is not None else self._view(int(i))
        idx = np.arange(len(self))[i]
        return CountSignals([self[int(k)] for k in idx], self.ss)

    def as_list(self):                                       # R/zzzCountSignals.R:83-96
        return list(self.signals)

    def alignSignals(self):                                  # R/zzzCountSignals.R:107-113
        w = self.width()
        if len(w) and (w != w[0]).any():
            raise ValueError("all signals must have the same width")
        if len(self) == 0:
            return np.zeros((0,), dtype=np.int32)
        if self._signals is None:                            # contiguous result: one reshape, no copies
            a = self._flat.reshape((len(self), -1))
            return a.T if not self.ss else a.reshape((len(self), -1, 2)).transpose(2, 1, 0)
        return np.stack(self._signals, axis=-1)              # [w, R] or [2, w, R] like simplify2array

    def __repr__(self):
        kind = "strand-specific" if self.ss else "strand-unspecific"
        return f"CountSignals object with {len(self)} {kind} signals"


# ------------------------------------------------------------------------------------------------------------------
# argument helpers (R/wrappers.R:76-98)
# ------------------------------------------------------------------------------------------------------------------
def flagMask(paired_end: str) -> int:
    return 66 if paired_end != "ignore" else 0


def _match_arg(value, choices):
    if isinstance(value, (tuple, list)):
        value = value[0]
    if value not in choices:
        raise ValueError(f"'arg' should be one of {', '.join(repr(c) for c in choices)}")
    return value


@dataclass
class Marshalled:
    """ctypes views of a GRanges for the C ABI (kept alive by this object)."""
    R: int
    levels: C.Array
    n_levels: int
    seq_idx: np.ndarray
    loc: np.ndarray
    width: np.ndarray
    strand: np.ndarray


def marshal_regions(gr: GRanges) -> Marshalled:
    if not isinstance(gr, GRanges):
        raise TypeError("must provide a GRanges object")             # src/bamsignals.cpp:94
    levels = (C.c_char_p * max(1, len(gr.seqlevels)))(*[s.encode() for s in gr.seqlevels])
    loc = np.ascontiguousarray(gr.start - 1, dtype=np.int32)         # 1-based -> 0-based, src/bamsignals.cpp:131
    return Marshalled(len(gr), levels, len(gr.seqlevels), gr.seq_idx, loc, gr.width, gr.strand)


def output_layout(width: np.ndarray, binsize: int, ss: bool):
    """Offsets (int64, R+1) of every region's slice in the flat result buffer (allocateList, :139-192)."""
    mult = 2 if ss else 1
    if binsize <= 0:
        per = np.full(len(width), mult, dtype=np.int64)
    else:
        per = mult * ((width.astype(np.int64) + binsize - 1) // binsize)
    off = np.zeros(len(width) + 1, dtype=np.int64)
    np.cumsum(per, out=off[1:])
    return off


def split_signals(flat: np.ndarray, offsets: np.ndarray, ss: bool):
    """Zero-copy per-region views shaped like the R list elements."""
    if ss:
        return [flat[offsets[i]:offsets[i + 1]].reshape((2, -1), order="F") for i in range(len(offsets) - 1)]
    return [flat[offsets[i]:offsets[i + 1]] for i in range(len(offsets) - 1)]


def _p(a, ty):
    return a.ctypes.data_as(C.POINTER(ty))


# ------------------------------------------------------------------------------------------------------------------
# the native library
# ------------------------------------------------------------------------------------------------------------------
class BsgOpts(C.Structure):
    _fields_ = [("struct_size", C.c_int32), ("n_devices", C.c_int32), ("devices", C.c_int32 * 16),
                ("inflate_threads", C.c_int32), ("batch_bytes", C.c_int64), ("verify_crc", C.c_int32),
                ("use_cache", C.c_int32), ("gpu_inflate", C.c_int32), ("stream_min_ints", C.c_int32),
                ("result_pack", C.c_int32), ("walk_scheme", C.c_int32), ("reserved", C.c_int32 * 5)]


class BsgTimings(C.Structure):
    _fields_ = [("records", C.c_int64), ("records_kept", C.c_int64), ("bytes_compressed", C.c_int64),
                ("bytes_inflated", C.c_int64), ("candidates", C.c_int64), ("out_elems", C.c_int64),
                ("n_tiles", C.c_int64), ("n_batches", C.c_int64), ("n_launches", C.c_int64),
                ("n_devices", C.c_int32), ("upload_mode", C.c_int32),
                ("ms_total", C.c_double), ("ms_plan", C.c_double), ("ms_fetch", C.c_double),
                ("ms_h2d", C.c_double), ("ms_d2h", C.c_double),
                ("ms_decode", C.c_double), ("ms_filter", C.c_double), ("ms_join", C.c_double),
                ("ms_count", C.c_double), ("ms_inflate_gpu", C.c_double), ("ms_kernels", C.c_double),
                ("ms_device", C.c_double), ("bytes_d2h", C.c_double), ("reserved", C.c_double * 6)]


def lib():
    """Load libbamsignals_cuda.so (built in-tree by __graft_entry__.build()); fail loudly if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise ImportError(f"{_LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`. "
                          "There is no CPU fallback for the counting path.")
    L = C.CDLL(_LIB_PATH)
    i32p, i64p, i8p = C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.POINTER(C.c_int8)
    region_args = [C.c_char_p, C.c_int64, C.POINTER(C.c_char_p), C.c_int32, i32p, i32p, i32p, i8p]
    L.bsg_pileup.argtypes = region_args + [i32p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                           C.c_int32, C.c_int32, i32p, i64p, C.POINTER(C.POINTER(C.c_int32)),
                                           C.POINTER(BsgOpts)]
    L.bsg_pileup.restype = C.c_int
    L.bsg_coverage.argtypes = region_args + [i32p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                             i32p, i64p, C.POINTER(C.POINTER(C.c_int32)), C.POINTER(BsgOpts)]
    L.bsg_coverage.restype = C.c_int
    L.bsg_last_error.restype = C.c_char_p
    L.bsg_version.restype = C.c_char_p
    L.bsg_get_timings.argtypes = [C.POINTER(BsgTimings)]
    L.bsg_get_timings.restype = C.c_int
    L.bsg_shutdown.restype = None
    L.bsg_output_layout.argtypes = [C.c_int64, i32p, C.c_int32, C.c_int32, i64p]
    L.bsg_output_layout.restype = C.c_int64
    L.bsg_device_count.restype = C.c_int
    L.bsg_stage_open.argtypes = [C.POINTER(C.c_void_p)] + region_args + [C.c_int32, C.POINTER(BsgOpts)]
    L.bsg_stage_open.restype = C.c_int
    L.bsg_pileup_staged.argtypes = [C.c_void_p, i32p] + [C.c_int32] * 7 + [i32p, i64p]
    L.bsg_pileup_staged.restype = C.c_int
    L.bsg_coverage_staged.argtypes = [C.c_void_p, i32p] + [C.c_int32] * 4 + [i32p, i64p]
    L.bsg_coverage_staged.restype = C.c_int
    L.bsg_stage_close.argtypes = [C.c_void_p]
    L.bsg_stage_close.restype = None
    L.bsg_debug_plan.argtypes = region_args + [C.c_int64, i64p, i64p, i64p, i32p, i32p, C.c_int64]
    L.bsg_debug_plan.restype = C.c_int64
    L.bsg_write_sam_as_bam_and_index.argtypes = [C.c_char_p, C.c_char_p]
    L.bsg_write_sam_as_bam_and_index.restype = C.c_int
    _lib = L
    return L


def _check(rc):
    if rc != 0:
        raise BamsignalsError(rc, lib().bsg_last_error().decode(errors="replace"))


def default_opts(**kw) -> BsgOpts:
    o = BsgOpts()
    o.struct_size = C.sizeof(BsgOpts)
    o.verify_crc = 1
    for k, v in kw.items():
        if k == "devices":
            o.n_devices = len(v)
            for i, d in enumerate(v):
                o.devices[i] = d
        else:
            setattr(o, k, v)
    return o


def debug_plan(bampath, gr, ext=0, cap=1 << 22) -> dict:
    """bsg_debug_plan: what the fetch planner decides for `gr` with halo `ext`, and the (tid, pos) of every record inside
    the planned ranges.  Host-only diagnostic (used by the CPU tests of the index logic); it counts nothing."""
    m = marshal_regions(gr)
    nseg, cb, ub = C.c_int64(), C.c_int64(), C.c_int64()
    tid, pos = np.empty(cap, dtype=np.int32), np.empty(cap, dtype=np.int32)
    n = lib().bsg_debug_plan(os.fsencode(bampath), m.R, m.levels, m.n_levels, _p(m.seq_idx, C.c_int32), _p(m.loc, C.c_int32),
                             _p(m.width, C.c_int32), _p(m.strand, C.c_int8), int(ext), C.byref(nseg), C.byref(cb), C.byref(ub),
                             _p(tid, C.c_int32), _p(pos, C.c_int32), cap)
    if n < 0:
        _check(int(n))
    if n > cap:
        return debug_plan(bampath, gr, ext, int(n))
    return dict(records=int(n), segments=nseg.value, bytes_compressed=cb.value, bytes_inflated=ub.value, tid=tid[:n], pos=pos[:n])


def writeSamAsBamAndIndex(sampath, bampath) -> bool:
    """bamsignals:::writeSamAsBamAndIndex (R/RcppExports.R:20-22, src/bamsignals.cpp:496-534): SAM text ->
    coordinate-sorted BAM + .bai.  Host-only."""
    _check(lib().bsg_write_sam_as_bam_and_index(os.fsencode(os.path.expanduser(sampath)), os.fsencode(os.path.expanduser(bampath))))
    return True


def timings() -> dict:
    t = BsgTimings()
    lib().bsg_get_timings(C.byref(t))
    return {f: getattr(t, f) for f, _ in BsgTimings._fields_ if f not in ("reserved", "pad")}


def pileup_core(bampath, gr, tlen_filter, mapqual=0, binsize=1, shift=0, ss=False, requiredF=0, filteredF=-1,
                pe_mid=False, maxgap=16385, opts: Optional[BsgOpts] = None, _lazy=False):
    """The .Call('bamsignals_pileup_core') entry point (src/bamsignals.cpp:444-461); binsize <= 0 means bamCount.
    Returns the R `List`: one flat vector / (2,R) matrix for bamCount, else one array per region."""
    m = marshal_regions(gr)
    off = output_layout(m.width, int(binsize), bool(ss))
    flat = np.empty(int(off[-1]), dtype=np.int32)
    tl = None if tlen_filter is None else np.asarray(tlen_filter, dtype=np.int32)
    rc = lib().bsg_pileup(os.fsencode(bampath), m.R, m.levels, m.n_levels, _p(m.seq_idx, C.c_int32),
                          _p(m.loc, C.c_int32), _p(m.width, C.c_int32), _p(m.strand, C.c_int8),
                          None if tl is None else _p(tl, C.c_int32), int(mapqual), int(binsize), int(shift),
                          int(bool(ss)), int(requiredF), int(filteredF), int(bool(pe_mid)), int(maxgap),
                          _p(flat, C.c_int32), _p(off, C.c_int64), None,
                          None if opts is None else C.byref(opts))
    _check(rc)
    if binsize <= 0:
        return [flat.reshape((2, -1), order="F") if ss else flat]
    if _lazy:
        return flat, off
    return split_signals(flat, off, bool(ss))


def coverage_core(bampath, gr, tlen_filter, mapqual=0, requiredF=0, filteredF=-1, tspan=False, maxgap=16385,
                  opts: Optional[BsgOpts] = None, _lazy=False):
    """The .Call('bamsignals_coverage_core') entry point (src/bamsignals.cpp:474-494)."""
    m = marshal_regions(gr)
    off = output_layout(m.width, 1, False)
    flat = np.empty(int(off[-1]), dtype=np.int32)
    tl = None if tlen_filter is None else np.asarray(tlen_filter, dtype=np.int32)
    rc = lib().bsg_coverage(os.fsencode(bampath), m.R, m.levels, m.n_levels, _p(m.seq_idx, C.c_int32),
                            _p(m.loc, C.c_int32), _p(m.width, C.c_int32), _p(m.strand, C.c_int8),
                            None if tl is None else _p(tl, C.c_int32), int(mapqual), int(requiredF), int(filteredF),
                            int(bool(tspan)), int(maxgap), _p(flat, C.c_int32), _p(off, C.c_int64), None,
                            None if opts is None else C.byref(opts))
    _check(rc)
    if _lazy:
        return flat, off
    return split_signals(flat, off, False)


class Stage:
    """Resident session (bsg_stage_*): the byte ranges `gr` needs (with halo `ext_hint`) are fetched, inflated and
    uploaded once; every pileup()/coverage() call then runs the device path only.  want_output=False leaves the
    result in HBM (kernel-only timing)."""

    def __init__(self, bampath, gr, ext_hint=0, opts: Optional[BsgOpts] = None):
        self._m = marshal_regions(gr)
        self._h = C.c_void_p()
        m = self._m
        _check(lib().bsg_stage_open(C.byref(self._h), os.fsencode(bampath), m.R, m.levels, m.n_levels,
                                    _p(m.seq_idx, C.c_int32), _p(m.loc, C.c_int32), _p(m.width, C.c_int32),
                                    _p(m.strand, C.c_int8), int(ext_hint), None if opts is None else C.byref(opts)))

    def pileup(self, tlen_filter, mapqual=0, binsize=1, shift=0, ss=False, requiredF=0, filteredF=-1, pe_mid=False,
               want_output=True):
        off = output_layout(self._m.width, int(binsize), bool(ss))
        flat = np.empty(int(off[-1]), dtype=np.int32) if want_output else None
        tl = None if tlen_filter is None else np.asarray(tlen_filter, dtype=np.int32)
        _check(lib().bsg_pileup_staged(self._h, None if tl is None else _p(tl, C.c_int32), int(mapqual), int(binsize),
                                       int(shift), int(bool(ss)), int(requiredF), int(filteredF), int(bool(pe_mid)),
                                       None if flat is None else _p(flat, C.c_int32), _p(off, C.c_int64)))
        return flat

    def coverage(self, tlen_filter, mapqual=0, requiredF=0, filteredF=-1, tspan=False, want_output=True):
        off = output_layout(self._m.width, 1, False)
        flat = np.empty(int(off[-1]), dtype=np.int32) if want_output else None
        tl = None if tlen_filter is None else np.asarray(tlen_filter, dtype=np.int32)
        _check(lib().bsg_coverage_staged(self._h, None if tl is None else _p(tl, C.c_int32), int(mapqual),
                                         int(requiredF), int(filteredF), int(bool(tspan)),
                                         None if flat is None else _p(flat, C.c_int32), _p(off, C.c_int64)))
        return flat

    def close(self):
        if self._h:
            lib().bsg_stage_close(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def core_args(fn, **kw):
    """R-level keyword arguments of bamCount / bamProfile / bamCoverage -> the arguments of the native entry point
    (R/wrappers.R:112-116, :143-147, :165-169)."""
    if fn == "bamCoverage":
        pe = _match_arg(kw.get("paired_end", "ignore"), ("ignore", "extend"))
        return dict(tlen_filter=_tlen(kw.get("tlenFilter"), pe), mapqual=_trunc(kw.get("mapqual", 0)),
                    requiredF=flagMask(pe), filteredF=_trunc(kw.get("filteredFlag", -1)), tspan=pe == "extend")
    pe = _match_arg(kw.get("paired_end", "ignore"), ("ignore", "filter", "midpoint"))
    return dict(tlen_filter=_tlen(kw.get("tlenFilter"), pe), mapqual=_trunc(kw.get("mapqual", 0)),
                binsize=-1 if fn == "bamCount" else _trunc(kw.get("binsize", 1)), shift=_trunc(kw.get("shift", 0)),
                ss=bool(kw.get("ss", False)), requiredF=flagMask(pe), filteredF=_trunc(kw.get("filteredFlag", -1)),
                pe_mid=pe == "midpoint")


# ------------------------------------------------------------------------------------------------------------------
# the user-facing functions (R/wrappers.R:106-173)
# ------------------------------------------------------------------------------------------------------------------
def _trunc(x):
    return int(x)          # Rcpp's input_parameter<int> truncates R doubles


def bamCount(bampath, gr, mapqual=0, shift=0, ss=False, paired_end=("ignore", "filter", "midpoint"),
             tlenFilter=None, filteredFlag=-1, verbose=False, opts=None):
    pe = _match_arg(paired_end, ("ignore", "filter", "midpoint"))
    bampath = os.path.expanduser(bampath)
    pu = pileup_core(bampath, gr, _tlen(tlenFilter, pe), _trunc(mapqual), -1, _trunc(shift), ss, flagMask(pe),
                     _trunc(filteredFlag), pe == "midpoint", opts=opts)
    return pu[0]                                                                      # R/wrappers.R:118


def _tlen(tf, pe):
    """The tlenFilter() helper of R/wrappers.R:84-98 (named differently: `tlenFilter` is also an argument name)."""
    if pe == "ignore":
        return None
    if tf is None:
        return (0, 1000)
    if len(tf) != 2 or tf[0] < 0 or tf[1] < 0:
        raise ValueError("tlenFilter must be NULL or vector of 2 positive integers")
    if tf[0] > tf[1]:
        raise ValueError("tlenFilter[1] must be smaller or equal to tlenFilter[2]")
    return (int(tf[0]), int(tf[1]))


def bamProfile(bampath, gr, binsize=1, mapqual=0, shift=0, ss=False, paired_end=("ignore", "filter", "midpoint"),
               tlenFilter=None, filteredFlag=-1, verbose=False, opts=None):
    if binsize < 1:
        raise ValueError("provide a binsize greater or equal to 1")                  # R/wrappers.R:136-137
    if binsize > 1 and isinstance(gr, GRanges) and (gr.width % int(binsize) != 0).any():
        warnings.warn("some ranges' widths are not a multiple of the selected binsize, "
                      "some bins will correspond to less than binsize basepairs")     # R/wrappers.R:138-141
    pe = _match_arg(paired_end, ("ignore", "filter", "midpoint"))
    bampath = os.path.expanduser(bampath)
    flat, off = pileup_core(bampath, gr, _tlen(tlenFilter, pe), _trunc(mapqual), _trunc(binsize), _trunc(shift), ss,
                            flagMask(pe), _trunc(filteredFlag), pe == "midpoint", opts=opts, _lazy=True)
    return CountSignals(None, ss, flat=flat, offsets=off)


def bamCoverage(bampath, gr, mapqual=0, paired_end=("ignore", "extend"), tlenFilter=None, filteredFlag=-1,
                verbose=False, opts=None):
    pe = _match_arg(paired_end, ("ignore", "extend"))
    bampath = os.path.expanduser(bampath)
    flat, off = coverage_core(bampath, gr, _tlen(tlenFilter, pe), _trunc(mapqual), flagMask(pe), _trunc(filteredFlag),
                              pe == "extend", opts=opts, _lazy=True)
    return CountSignals(None, False, flat=flat, offsets=off)
