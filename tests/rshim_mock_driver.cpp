// Runs rshim/bamsignals_shim.cpp on plain C++ objects through tests/mock_rcpp/Rcpp.h.  TEST INFRASTRUCTURE.
//
//   rshim_mock_driver <spec-file>
// spec (whitespace separated):  fn bampath mapqual binsize shift ss requiredF filteredF flag n_tlen [tmin tmax] R
//                               then R lines: seqname start width strand        (fn: pileup | coverage | writesam)
// Builds a GRanges-shaped S4 object (ranges@start/width, seqnames and strand as factor Rles whose levels are listed in
// REVERSE order of first appearance, so that a shim confusing level codes with BAM reference ids cannot pass), calls the
// shim's exported function and prints the returned R list as JSON: {"result":[{"dim":[..],"dimnames0":[..],"v":[..]}..]}
// or {"error":"<message of Rcpp::stop>"}.
#include <Rcpp.h>

#include <algorithm>
#include <cstdio>
#include <fstream>
#include <iostream>

using namespace Rcpp;

List pileup_core(std::string bampath, RObject& gr, IntegerVector& tlen_filter, int mapqual, int binsize, int shift, bool ss,
                 int requiredF, int filteredF, bool pe_mid, int maxgap);
List coverage_core(std::string bampath, RObject& gr, IntegerVector& tlen_filter, int mapqual, int requiredF, int filteredF,
                   bool tspan, int maxgap);
bool writeSamAsBamAndIndex(const std::string& sampath, const std::string& bampath);

static SEXP factor_rle(const std::vector<std::string>& values) {
    std::vector<std::string> levels;
    for (auto& v : values) if (std::find(levels.begin(), levels.end(), v) == levels.end()) levels.push_back(v);
    std::reverse(levels.begin(), levels.end());
    auto rle = std::make_shared<Node>(), vals = std::make_shared<Node>(), lens = std::make_shared<Node>(), lv = std::make_shared<Node>();
    lv->strs = levels;
    for (size_t i = 0; i < values.size(); ++i) {
        if (i && values[i] == values[i - 1]) { ++lens->ints.back(); continue; }
        vals->ints.push_back(int(std::find(levels.begin(), levels.end(), values[i]) - levels.begin()) + 1);   // 1-based codes
        lens->ints.push_back(1);
    }
    vals->attrs["levels"] = lv;
    rle->attrs["values"] = vals;
    rle->attrs["lengths"] = lens;
    return rle;
}

static std::string esc(const std::string& s) {
    std::string o;
    for (char c : s) { if (c == '"' || c == '\\') o += '\\'; o += c; }
    return o;
}

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    std::ifstream in(argv[1]);
    std::string fn, bam;
    int mapqual, binsize, shift, ss, requiredF, filteredF, flag, n_tlen;
    in >> fn >> bam >> mapqual >> binsize >> shift >> ss >> requiredF >> filteredF >> flag >> n_tlen;
    IntegerVector tlen(n_tlen);
    for (int i = 0; i < n_tlen; ++i) in >> tlen[i];
    long R;
    in >> R;
    std::vector<std::string> names(R), strands(R);
    auto start = std::make_shared<Node>(), width = std::make_shared<Node>();
    start->ints.resize(R); width->ints.resize(R);
    for (long i = 0; i < R; ++i) in >> names[i] >> start->ints[i] >> width->ints[i] >> strands[i];
    auto ranges = std::make_shared<Node>();
    ranges->attrs["start"] = start;
    ranges->attrs["width"] = width;
    auto grn = std::make_shared<Node>();
    grn->klass = {"GRanges", "GenomicRanges"};
    grn->attrs["ranges"] = ranges;
    grn->attrs["seqnames"] = factor_rle(names);
    grn->attrs["strand"] = factor_rle(strands);
    RObject gr(grn);
    try {
        if (fn == "writesam") {
            std::cout << "{\"result\":[],\"ok\":" << (writeSamAsBamAndIndex(bam, names.empty() ? bam + ".bam" : names[0]) ? "true" : "false") << "}\n";
            return 0;
        }
        RObject notgr;
        List res = fn == "pileup"   ? pileup_core(bam, flag == 99 ? notgr : gr, tlen, mapqual, binsize, shift, ss != 0, requiredF, filteredF, flag == 1, 16385)
                                    : coverage_core(bam, flag == 99 ? notgr : gr, tlen, mapqual, requiredF, filteredF, flag == 1, 16385);
        std::cout << "{\"result\":[";
        for (R_xlen_t i = 0; i < res.size(); ++i) {
            SEXP e = res[i].get();
            std::cout << (i ? "," : "") << "{\"dim\":[";
            if (e->nrow) std::cout << e->nrow << "," << (e->ints.size() / size_t(e->nrow)); else std::cout << e->ints.size();
            std::cout << "],\"dimnames0\":[";
            auto dn = e->attrs.find("dimnames");
            if (dn != e->attrs.end() && !dn->second->items.empty() && dn->second->items[0])
                for (size_t k = 0; k < dn->second->items[0]->strs.size(); ++k) std::cout << (k ? "," : "") << '"' << esc(dn->second->items[0]->strs[k]) << '"';
            std::cout << "],\"v\":[";
            for (size_t k = 0; k < e->ints.size(); ++k) std::cout << (k ? "," : "") << e->ints[k];
            std::cout << "]}";
        }
        std::cout << "]}\n";
    } catch (Rcpp::exception& e) {
        std::cout << "{\"error\":\"" << esc(e.what()) << "\"}\n";
    }
    return 0;
}
