// Stand-in for libbamsignals_cuda.so that only RECORDS what the R shim passes through the C ABI (one JSON object per
// call appended to $BSG_STUB_LOG) and fills every output element with a value that encodes (region, element).
// TEST INFRASTRUCTURE for tests/test_zz_rshim_mock.py; implements exactly the entry points rshim/ uses.
#include <cstdio>
#include <cstdlib>
#include <string>

#include "bamsignals_cuda.h"

static std::string g_err;

static void log_regions(FILE* f, int64_t R, const char* const* lv, int32_t nl, const int32_t* si, const int32_t* loc,
                        const int32_t* w, const int8_t* st) {
    fprintf(f, "\"levels\":[");
    for (int i = 0; i < nl; ++i) fprintf(f, "%s\"%s\"", i ? "," : "", lv[i]);
    fprintf(f, "],\"regions\":[");
    for (int64_t i = 0; i < R; ++i) fprintf(f, "%s[%d,%d,%d,%d]", i ? "," : "", si[i], loc[i], w[i], int(st[i]));
    fprintf(f, "]");
}

static int fill(int64_t R, int32_t* out, const int64_t* off, int32_t* const* ptrs) {
    if (!off || (!out && !ptrs)) { g_err = "stub: no output"; return BSG_EARG; }
    for (int64_t i = 0; i < R; ++i)
        for (int64_t k = 0; k < off[i + 1] - off[i]; ++k) (out ? out + off[i] : ptrs[i])[k] = int32_t(1000 * i + k);
    return BSG_OK;
}

extern "C" {

int bsg_pileup(const char* bampath, int64_t R, const char* const* seq_levels, int32_t n_levels, const int32_t* seq_idx,
               const int32_t* loc, const int32_t* width, const int8_t* strand, const int32_t* tlen_filter, int32_t mapqual,
               int32_t binsize, int32_t shift, int32_t ss, int32_t requiredF, int32_t filteredF, int32_t pe_mid, int32_t maxgap,
               int32_t* out, const int64_t* out_offsets, int32_t* const* out_ptrs, const bsg_opts* opts) {
    FILE* f = fopen(getenv("BSG_STUB_LOG"), "a");
    fprintf(f, "{\"fn\":\"bsg_pileup\",\"bam\":\"%s\",", bampath);
    log_regions(f, R, seq_levels, n_levels, seq_idx, loc, width, strand);
    fprintf(f, ",\"tlen\":%s", tlen_filter ? "[" : "null");
    if (tlen_filter) fprintf(f, "%d,%d]", tlen_filter[0], tlen_filter[1]);
    fprintf(f, ",\"mapqual\":%d,\"binsize\":%d,\"shift\":%d,\"ss\":%d,\"requiredF\":%d,\"filteredF\":%d,\"pe_mid\":%d,\"maxgap\":%d,"
               "\"out_null\":%d,\"opts_null\":%d,\"total\":%lld}\n", mapqual, binsize, shift, ss, requiredF, filteredF, pe_mid, maxgap,
            out == nullptr, opts == nullptr, (long long)out_offsets[R]);
    fclose(f);
    if (std::string(bampath) == "fail.bam") { g_err = "Fail to open BAM file fail.bam"; return BSG_EOPEN; }
    return fill(R, out, out_offsets, out_ptrs);
}

int bsg_coverage(const char* bampath, int64_t R, const char* const* seq_levels, int32_t n_levels, const int32_t* seq_idx,
                 const int32_t* loc, const int32_t* width, const int8_t* strand, const int32_t* tlen_filter, int32_t mapqual,
                 int32_t requiredF, int32_t filteredF, int32_t tspan, int32_t maxgap, int32_t* out, const int64_t* out_offsets,
                 int32_t* const* out_ptrs, const bsg_opts* opts) {
    FILE* f = fopen(getenv("BSG_STUB_LOG"), "a");
    fprintf(f, "{\"fn\":\"bsg_coverage\",\"bam\":\"%s\",", bampath);
    log_regions(f, R, seq_levels, n_levels, seq_idx, loc, width, strand);
    fprintf(f, ",\"tlen\":%s", tlen_filter ? "[" : "null");
    if (tlen_filter) fprintf(f, "%d,%d]", tlen_filter[0], tlen_filter[1]);
    fprintf(f, ",\"mapqual\":%d,\"requiredF\":%d,\"filteredF\":%d,\"tspan\":%d,\"maxgap\":%d,\"out_null\":%d,\"opts_null\":%d,\"total\":%lld}\n",
            mapqual, requiredF, filteredF, tspan, maxgap, out == nullptr, opts == nullptr, (long long)out_offsets[R]);
    fclose(f);
    return fill(R, out, out_offsets, out_ptrs);
}

int64_t bsg_output_layout(int64_t R, const int32_t* width, int32_t binsize, int32_t ss, int64_t* offsets) {
    const int64_t mult = ss ? 2 : 1;
    int64_t acc = 0;
    for (int64_t i = 0; i < R; ++i) { offsets[i] = acc; acc += binsize <= 0 ? mult : mult * ((int64_t(width[i]) + binsize - 1) / binsize); }
    offsets[R] = acc;
    return acc;
}

int bsg_write_sam_as_bam_and_index(const char* sampath, const char* bampath) {
    FILE* f = fopen(getenv("BSG_STUB_LOG"), "a");
    fprintf(f, "{\"fn\":\"bsg_write_sam_as_bam_and_index\",\"sam\":\"%s\",\"bam\":\"%s\"}\n", sampath, bampath);
    fclose(f);
    return BSG_OK;
}

const char* bsg_last_error(void) { return g_err.c_str(); }

}  // extern "C"
