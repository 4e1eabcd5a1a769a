// R-side shim: the two .Call entry points of the bamsignals package re-implemented on top of libbamsignals_cuda.so.
//
// Drop this file in place of the package's src/bamsignals.cpp (keep src/RcppExports.cpp, src/bamsignals_init.c,
// src/CountSignals.cpp and every R file unchanged) and link with rshim/Makevars.  The exported C++ functions keep the
// names, argument order and defaults that src/RcppExports.cpp:34-70 and R/RcppExports.R:12-18 expect:
//   pileup_core  (replaces /root/reference src/bamsignals.cpp:444-461)
//   coverage_core(replaces /root/reference src/bamsignals.cpp:474-494)
// What stays on the R side of the boundary is exactly what touches R objects: reading the GRanges slots
// (the job of parseRegions, :92-135) and allocating the result vectors (the job of allocateList, :139-192).
// The per-region htslib loops (overlapAndPileup, :240-291) and cumsum (:464-470) are replaced by ONE call into the
// C ABI of include/bamsignals_cuda.h.
//
// STATUS: R, Rcpp and Rhtslib are not installable offline, so the R package has not been built with this file.  It IS
// compiled and run by tests/test_zz_rshim_mock.py against a minimal stand-in for <Rcpp.h> (tests/mock_rcpp/): the
// arguments it passes through the C ABI, the layout it returns and the errors it re-raises are checked there; the same
// C ABI is exercised from Python (bamsignals_b200/api.py) by the parity tests.
#include <Rcpp.h>

#include <string>
#include <vector>

#include "bamsignals_cuda.h"

using namespace Rcpp;

namespace {

struct RegionArrays {
    std::vector<std::string> levels;
    std::vector<const char*> level_ptrs;
    std::vector<int32_t> seq_idx, loc, width;
    std::vector<int8_t> strand;
};

// Expand an S4 Rle of a factor into per-range codes (0-based level index).
std::vector<int> expand_factor_rle(const RObject& rle, CharacterVector* levels_out) {
    IntegerVector run_len = as<IntegerVector>(rle.slot("lengths"));
    IntegerVector run_val = as<IntegerVector>(rle.slot("values"));
    *levels_out = as<CharacterVector>(run_val.attr("levels"));
    std::vector<int> codes;
    for (R_xlen_t r = 0; r < run_len.size(); ++r) codes.insert(codes.end(), size_t(run_len[r]), run_val[r] - 1);
    return codes;
}

RegionArrays read_granges(RObject& gr) {
    if (!gr.inherits("GRanges")) stop("must provide a GRanges object");
    RObject ranges = as<RObject>(gr.slot("ranges"));
    IntegerVector start = as<IntegerVector>(ranges.slot("start"));
    IntegerVector width = as<IntegerVector>(ranges.slot("width"));
    CharacterVector chr_levels, strand_levels;
    std::vector<int> chr = expand_factor_rle(as<RObject>(gr.slot("seqnames")), &chr_levels);
    std::vector<int> str = expand_factor_rle(as<RObject>(gr.slot("strand")), &strand_levels);
    RegionArrays a;
    for (R_xlen_t i = 0; i < chr_levels.size(); ++i) a.levels.push_back(as<std::string>(chr_levels[i]));
    for (auto& s : a.levels) a.level_ptrs.push_back(s.c_str());
    const R_xlen_t n = start.size();
    a.seq_idx.resize(n); a.loc.resize(n); a.width.resize(n); a.strand.resize(n);
    for (R_xlen_t i = 0; i < n; ++i) {
        a.seq_idx[i] = chr[i];
        a.loc[i] = start[i] - 1;                       // 1-based closed -> 0-based
        a.width[i] = width[i];
        const std::string s = as<std::string>(strand_levels[str[i]]);
        a.strand[i] = s == "+" ? 1 : (s == "-" ? -1 : 0);
    }
    return a;
}

// Result list with the reference's layout and attributes; fills the per-region pointers and flat offsets.
List make_result(const RegionArrays& a, int binsize, bool ss, std::vector<int32_t*>* ptrs, std::vector<int64_t>* offsets) {
    const R_xlen_t n = R_xlen_t(a.width.size());
    const int mult = ss ? 2 : 1;
    List dn(2);
    if (ss) dn[0] = CharacterVector::create("sense", "antisense");
    offsets->assign(size_t(n) + 1, 0);
    bsg_output_layout(n, a.width.data(), binsize, ss ? 1 : 0, offsets->data());
    ptrs->resize(size_t(n));
    if (binsize <= 0) {                                // bamCount: one vector / 2 x n matrix for all regions
        List res(1);
        int* base;
        if (ss) { IntegerMatrix m(2, n); m.attr("dimnames") = dn; res[0] = m; base = m.begin(); }
        else { IntegerVector v(n); res[0] = v; base = v.begin(); }
        for (R_xlen_t i = 0; i < n; ++i) (*ptrs)[i] = base + mult * i;
        return res;
    }
    List res(n);
    for (R_xlen_t i = 0; i < n; ++i) {
        const int w = int(((*offsets)[i + 1] - (*offsets)[i]) / mult);
        if (ss) { IntegerMatrix m(2, w); m.attr("dimnames") = dn; res[i] = m; (*ptrs)[i] = m.begin(); }
        else { IntegerVector v(w); res[i] = v; (*ptrs)[i] = v.begin(); }
    }
    return res;
}

void check(int rc) {
    if (rc != BSG_OK) stop(std::string(bsg_last_error()));   // same message strings as the reference's Rcpp::stop calls
}

}  // namespace

// [[Rcpp::export]]
List pileup_core(std::string bampath, RObject& gr, IntegerVector& tlen_filter, int mapqual = 0, int binsize = 1,
                 int shift = 0, bool ss = false, int requiredF = 0, int filteredF = -1, bool pe_mid = false,
                 int maxgap = 16385) {
    RegionArrays a = read_granges(gr);
    std::vector<int32_t*> ptrs;
    std::vector<int64_t> offsets;
    List res = make_result(a, binsize, ss, &ptrs, &offsets);
    check(bsg_pileup(bampath.c_str(), int64_t(a.loc.size()), a.level_ptrs.data(), int32_t(a.level_ptrs.size()),
                     a.seq_idx.data(), a.loc.data(), a.width.data(), a.strand.data(),
                     tlen_filter.size() == 0 ? nullptr : tlen_filter.begin(), mapqual, binsize, shift, ss, requiredF,
                     filteredF, pe_mid, maxgap, nullptr, offsets.data(), ptrs.data(), nullptr));
    return res;
}

// [[Rcpp::export]]
List coverage_core(std::string bampath, RObject& gr, IntegerVector& tlen_filter, int mapqual = 0, int requiredF = 0,
                   int filteredF = -1, bool tspan = false, int maxgap = 16385) {
    RegionArrays a = read_granges(gr);
    std::vector<int32_t*> ptrs;
    std::vector<int64_t> offsets;
    List res = make_result(a, 1, false, &ptrs, &offsets);
    check(bsg_coverage(bampath.c_str(), int64_t(a.loc.size()), a.level_ptrs.data(), int32_t(a.level_ptrs.size()),
                       a.seq_idx.data(), a.loc.data(), a.width.data(), a.strand.data(),
                       tlen_filter.size() == 0 ? nullptr : tlen_filter.begin(), mapqual, requiredF, filteredF, tspan,
                       maxgap, nullptr, offsets.data(), ptrs.data(), nullptr));
    return res;
}

// this is used only for the reference's tests (tests/testthat/utils.R:116); same name and signature as
// src/bamsignals.cpp:498, so src/RcppExports.cpp:71-82 and R/RcppExports.R:20-22 bind it unchanged
// [[Rcpp::export]]
bool writeSamAsBamAndIndex(const std::string& sampath, const std::string& bampath) {
    check(bsg_write_sam_as_bam_and_index(sampath.c_str(), bampath.c_str()));
    return true;
}
