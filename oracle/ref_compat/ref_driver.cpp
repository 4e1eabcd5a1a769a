// ref_driver — C entry points around the reference's OWN pileup_core / coverage_core (src/bamsignals.cpp:444-461,
// :474-494), compiled unchanged from /root/reference into oracle/_ref/libbamsignals_ref.so (see oracle/Makefile).
// TEST INFRASTRUCTURE ONLY: this is the checker the restated oracle (../bsg_oracle.cpp) and the CUDA path are compared
// with, and (bench.py --impl reference) the timed reference CPU arm.  Never linked into or loaded by the product.
//
// What happens here is what R does around the two .Call symbols (R/wrappers.R:115,146,168, src/RcppExports.cpp:34-70):
// build the GRanges S4 object (ranges@start/width, seqnames and strand as factor-Rle, as GenomicRanges lays them out),
// call the reference function, and read the returned list - one IntegerVector / 2-row IntegerMatrix for bamCount, one
// per region otherwise (allocateList, :139-192) - into the flat bsg_output_layout() buffer the tests compare.
// nthreads > 1: the regions are cut into contiguous shards in (seqlevel, start) order and every shard is one independent
// call of the reference function on its own thread (own file handle, own index load - what N R processes would do).
#include <Rcpp.h>

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

// the reference's exported functions (src/bamsignals.cpp:444, :474; declared the way src/RcppExports.cpp:10-11,53 does)
Rcpp::List pileup_core(std::string bampath, Rcpp::RObject& gr, Rcpp::IntegerVector& tlen_filter, int mapqual, int binsize,
                       int shift, bool ss, int requiredF, int filteredF, bool pe_mid, int maxgap);
Rcpp::List coverage_core(std::string bampath, Rcpp::RObject& gr, Rcpp::IntegerVector& tlen_filter, int mapqual,
                         int requiredF, int filteredF, bool tspan, int maxgap);
extern "C" void hts_compat_stats(uint64_t* records, uint64_t* queries, uint64_t* bytes, int reset);

namespace {

thread_local std::string g_err;

Rcpp::SEXP int_node(std::vector<int> v) {
    // two ints of slack capacity, as Rcpp::IntegerVector(n) of the stand-in has: for an EMPTY GRanges the reference's
    // RleIter constructor reads rlens[0] of a zero-length vector (src/bamsignals.cpp:72-79); in R that read lands in
    // the allocation behind the object header, here in the slack
    auto n = std::make_shared<Rcpp::Node>();
    n->ints.reserve(v.size() + 2);
    n->ints.assign(v.begin(), v.end());
    if (v.empty()) { n->ints.push_back(0); n->ints.pop_back(); }
    return n;
}

// factor-Rle: lengths, values (1-based codes) with attribute "levels"
Rcpp::SEXP factor_rle(const std::vector<int>& codes, const std::vector<std::string>& levels) {
    std::vector<int> lens, vals;
    for (size_t i = 0; i < codes.size(); ++i) {
        if (!vals.empty() && vals.back() == codes[i] + 1) ++lens.back();
        else { vals.push_back(codes[i] + 1); lens.push_back(1); }
    }
    auto lv = std::make_shared<Rcpp::Node>();
    lv->strs = levels;
    Rcpp::SEXP values = int_node(std::move(vals));
    values->attrs["levels"] = lv;
    auto rle = std::make_shared<Rcpp::Node>();
    rle->klass = {"Rle"};
    rle->attrs["lengths"] = int_node(std::move(lens));
    rle->attrs["values"] = values;
    return rle;
}

Rcpp::RObject make_granges(const std::vector<int64_t>& idx, const char* const* seq_levels, int n_levels, const int32_t* seq_idx,
                           const int32_t* loc, const int32_t* width, const int8_t* strand) {
    std::vector<int> start, w, chr, str;
    for (int64_t i : idx) {
        start.push_back(int(uint32_t(loc[i]) + 1u));         // GRanges starts are 1-based (the reference subtracts 1, :131)
        w.push_back(width[i]);
        chr.push_back(seq_idx[i]);
        str.push_back(strand[i] > 0 ? 0 : (strand[i] < 0 ? 1 : 2));
    }
    std::vector<std::string> levels;
    for (int k = 0; k < n_levels; ++k) levels.push_back(seq_levels[k]);
    auto ranges = std::make_shared<Rcpp::Node>();
    ranges->klass = {"IRanges"};
    ranges->attrs["start"] = int_node(std::move(start));
    ranges->attrs["width"] = int_node(std::move(w));
    auto gr = std::make_shared<Rcpp::Node>();
    gr->klass = {"GRanges"};
    gr->attrs["ranges"] = ranges;
    gr->attrs["seqnames"] = factor_rle(chr, levels);
    gr->attrs["strand"] = factor_rle(str, {"+", "-", "*"});
    return Rcpp::RObject(gr);
}

struct Call {
    bool coverage;
    int mapqual, binsize, shift, ss, requiredF, filteredF, flag, maxgap;     // flag = pe_mid | tspan
    const int32_t* tlen_filter;
};

void run_shard(const char* bam, const std::vector<int64_t>& idx, const char* const* seq_levels, int n_levels,
               const int32_t* seq_idx, const int32_t* loc, const int32_t* width, const int8_t* strand, const Call& c,
               int32_t* out, const int64_t* out_offsets) {
    Rcpp::RObject gr = make_granges(idx, seq_levels, n_levels, seq_idx, loc, width, strand);
    Rcpp::IntegerVector tl(c.tlen_filter ? 2 : 0);           // integer() = no filter (R/wrappers.R:84-98)
    if (c.tlen_filter) { tl[0] = c.tlen_filter[0]; tl[1] = c.tlen_filter[1]; }
    Rcpp::List res = c.coverage ? coverage_core(bam, gr, tl, c.mapqual, c.requiredF, c.filteredF, c.flag != 0, c.maxgap)
                                : pileup_core(bam, gr, tl, c.mapqual, c.binsize, c.shift, c.ss != 0, c.requiredF, c.filteredF,
                                              c.flag != 0, c.maxgap);
    const std::vector<Rcpp::SEXP>& items = res.sexp()->items;
    if (!c.coverage && c.binsize <= 0) {                     // bamCount: one vector / 2-row matrix for all regions (:148-169)
        const int mult = c.ss ? 2 : 1;
        if (items.size() != 1 || items[0]->ints.size() != idx.size() * size_t(mult)) throw std::runtime_error("unexpected bamCount result shape");
        for (size_t k = 0; k < idx.size(); ++k) {
            if (out_offsets[idx[k] + 1] - out_offsets[idx[k]] != mult) throw std::runtime_error("out_offsets do not match the result layout");
            memcpy(out + out_offsets[idx[k]], items[0]->ints.data() + k * size_t(mult), sizeof(int32_t) * size_t(mult));
        }
        return;
    }
    if (items.size() != idx.size()) throw std::runtime_error("unexpected result length");
    for (size_t k = 0; k < idx.size(); ++k) {
        const std::vector<int>& v = items[k]->ints;
        if (int64_t(v.size()) != out_offsets[idx[k] + 1] - out_offsets[idx[k]]) throw std::runtime_error("out_offsets do not match the result layout");
        if (!v.empty()) memcpy(out + out_offsets[idx[k]], v.data(), sizeof(int32_t) * v.size());
    }
}

int run(const char* bam, int64_t R, const char* const* seq_levels, int n_levels, const int32_t* seq_idx, const int32_t* loc,
        const int32_t* width, const int8_t* strand, const Call& c, int32_t* out, const int64_t* out_offsets, int nthreads) {
    try {
        for (int64_t i = 0; i < R; ++i)
            if (seq_idx[i] < 0 || seq_idx[i] >= n_levels) throw std::runtime_error("region refers to an unknown seqlevel index");
        std::vector<int64_t> all(static_cast<size_t>(R));
        std::iota(all.begin(), all.end(), int64_t(0));
        if (nthreads <= 1 || R < 2) {
            run_shard(bam, all, seq_levels, n_levels, seq_idx, loc, width, strand, c, out, out_offsets);
            return 0;
        }
        std::sort(all.begin(), all.end(), [&](int64_t a, int64_t b) {
            return seq_idx[a] != seq_idx[b] ? seq_idx[a] < seq_idx[b] : (loc[a] != loc[b] ? loc[a] < loc[b] : a < b); });
        const int T = int(std::min<int64_t>(nthreads, R));
        std::vector<std::string> errs{size_t(T), std::string()};
        std::vector<std::thread> th;
        for (int t = 0; t < T; ++t)
            th.emplace_back([&, t] {
                try {
                    std::vector<int64_t> idx(all.begin() + R * t / T, all.begin() + R * (t + 1) / T);
                    std::sort(idx.begin(), idx.end());       // the caller's order within the shard
                    run_shard(bam, idx, seq_levels, n_levels, seq_idx, loc, width, strand, c, out, out_offsets);
                } catch (std::exception& e) { errs[size_t(t)] = e.what(); }
            });
        for (auto& t : th) t.join();
        for (auto& e : errs) if (!e.empty()) throw std::runtime_error(e);
        return 0;
    } catch (std::exception& e) {
        g_err = e.what();
        return -1;
    }
}

}  // namespace

extern "C" {

const char* ref_last_error() { return g_err.c_str(); }

int ref_pileup(const char* bam, int64_t R, const char* const* seq_levels, int n_levels, const int32_t* seq_idx,
               const int32_t* loc, const int32_t* width, const int8_t* strand, const int32_t* tlen_filter, int mapqual,
               int binsize, int shift, int ss, int requiredF, int filteredF, int pe_mid, int maxgap, int32_t* out,
               const int64_t* out_offsets, int /*mode*/, int nthreads) {
    const Call c{false, mapqual, binsize, shift, ss, requiredF, filteredF, pe_mid, maxgap, tlen_filter};
    return run(bam, R, seq_levels, n_levels, seq_idx, loc, width, strand, c, out, out_offsets, nthreads);
}

int ref_coverage(const char* bam, int64_t R, const char* const* seq_levels, int n_levels, const int32_t* seq_idx,
                 const int32_t* loc, const int32_t* width, const int8_t* strand, const int32_t* tlen_filter, int mapqual,
                 int requiredF, int filteredF, int tspan, int maxgap, int32_t* out, const int64_t* out_offsets, int /*mode*/,
                 int nthreads) {
    const Call c{true, mapqual, 1, 0, 0, requiredF, filteredF, tspan, maxgap, tlen_filter};
    return run(bam, R, seq_levels, n_levels, seq_idx, loc, width, strand, c, out, out_offsets, nthreads);
}

void ref_stats(uint64_t* records, uint64_t* bytes_inflated, uint64_t* queries) {
    hts_compat_stats(records, queries, bytes_inflated, 1);
}

}  // extern "C"
