// bsg_oracle — CPU restatement of the bamsignals counting path.  TEST INFRASTRUCTURE ONLY.
//
// This file is the parity oracle and the timed CPU baseline for the B200 build.  It is never linked
// into, loaded by, or called from the product library (libbamsignals_cuda.so); only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
//
// What it restates (all citations relative to /root/reference):
//   * region model GArray                      src/bamsignals.cpp:32-50
//   * output layout (allocateList)             src/bamsignals.cpp:139-192
//   * sort + maxgap chunking + sweep           src/bamsignals.cpp:222-291 (overlapAndPileup)
//   * read filter + 5'/midpoint coordinate     src/bamsignals.cpp:326-346 (Pileupper::setRead)
//   * bin increment                            src/bamsignals.cpp:349-363 (Pileupper::pileup)
//   * coverage interval + difference update    src/bamsignals.cpp:392-438 (Coverager)
//   * prefix sum                               src/bamsignals.cpp:464-470 (cumsum)
//   * drivers and `ext`                        src/bamsignals.cpp:444-461, 474-494
// The reference delegates BGZF/BAM/BAI work to htslib (Rhtslib >= 1.13.1, DESCRIPTION:30; version not
// pinned by the reference, htslib sources not in the tree).  The parts of htslib whose behaviour reaches
// the results are restated from the published SAM/BAM specification and htslib >= 1.10 behaviour:
//   * bam_endpos: pos + rlen, rlen = 0 for flag 0x4 reads else sum of M/D/N/=/X op lengths, 0 -> 1
//   * bam_itr_queryi / bam_itr_next: BAI bins + linear index -> chunk list; a record is returned iff
//     tid == query tid, pos < end and endpos > beg (beg clamped at 0)
// Parity pin: the oracle is checked in tests/ against (a) an independent numpy restatement of the
// reference's own R test oracle (tests/testthat/utils.R:178-311) fed with the reference's
// randomReads.RData, over the full parameter sweep of tests/testthat/test_methods.R:33-104 on the
// reference's own randomBam.bam, and (b) committed golden vectors derived from those.  CIGAR ops other
// than M and flag-0x4 reads are NOT covered by any reference fixture: for those parity is pinned only to
// the SAM specification rule above ("parity unpinned by the reference" for those record kinds).
//
// Three access modes, identical results:
//   mode 0 "scan"    : stream the whole file once, one chunk per chromosome (no index needed)
//   mode 1 "indexed" : the reference's access pattern (maxgap chunking, one BAI query per chunk, 10-block
//                      cache) — this is the timed CPU baseline
//   mode 2 "brute"   : every read tested against every region of its chromosome, no sweep (small inputs)
#include "bam_reader.hpp"

namespace {

using namespace bsgo;


// ---------------------------------------------------------------------------------------------
// regions, parameters, per-read state
// ---------------------------------------------------------------------------------------------
struct Region {            // GArray, src/bamsignals.cpp:32-50
    int32_t rid, loc, len, strand;
    int32_t* out;
    int64_t end() const { return int64_t(loc) + len; }
};

struct Params {
    bool coverage = false;
    int mapqual = 0, binsize = 1, shift = 0;
    bool ss = false, midpoint = false, tspan = false;
    uint32_t required = 0, filtered = 0xffffffffu;
    bool have_tlen = false; int tmin = 0, tmax = 0;
    int ext = 0, maxgap = 16385;
};

struct ReadState { int64_t pos5 = 0; bool neg = false; int64_t s = 0, e = 0; };

// filter shared by both pileuppers: src/bamsignals.cpp:328-333 and :394-399
inline bool passes(const Params& P, const Rec& r) {
    if (int(r.mapq) < P.mapqual) return false;
    if (P.required & ~uint32_t(r.flag)) return false;          // some required bit missing
    if (!(P.filtered & ~uint32_t(r.flag))) return false;       // every filtered bit present (0 drops all)
    if (P.have_tlen) {
        int64_t a = r.tlen < 0 ? -int64_t(r.tlen) : r.tlen;
        if (a < P.tmin || a > P.tmax) return false;
    }
    return true;
}

// returns read_end (0-based inclusive) or -1; src/bamsignals.cpp:326-346 / :392-415
inline int64_t set_read(const Params& P, const Rec& r, ReadState& st) {
    if (!passes(P, r)) return -1;
    int64_t read_end = int64_t(r.endpos) - 1;
    st.neg = (r.flag & 0x10) != 0;
    if (!P.coverage) {
        int64_t a = r.tlen < 0 ? -int64_t(r.tlen) : r.tlen;
        int64_t offset = P.midpoint ? a / 2 + P.shift : P.shift;
        st.pos5 = st.neg ? read_end - offset : int64_t(r.pos) + offset;
    } else {
        st.s = r.pos; st.e = read_end;
        if (P.tspan) {
            if (st.neg && r.tlen < 0) st.s = st.e + r.tlen + 1;
            else if (!st.neg && r.tlen > 0) st.e = st.s + r.tlen - 1;
        }
    }
    return read_end;
}

inline void apply(const Params& P, const ReadState& st, Region& g) {
    if (!P.coverage) {                                           // src/bamsignals.cpp:349-363
        int64_t rel = st.pos5 - g.loc;
        if (rel < 0 || rel >= g.len) return;
        int anti = st.neg ? 1 : 0;
        if (g.strand < 0) { rel = g.len - rel - 1; anti = 1 - anti; }
        if (P.ss) ++g.out[2 * (rel / P.binsize) + anti]; else ++g.out[rel / P.binsize];
    } else {                                                     // src/bamsignals.cpp:418-438
        if (st.s >= g.end() || st.e < g.loc) return;
        // Deliberate difference: for a ZERO-WIDTH region the reference's test above (:420) still passes for a read that
        // spans `loc`, and :424/:432 then increment array[0] of an empty vector - a write past the end of the R
        // object.  The defined result is the empty vector; nothing is written here.
        if (g.len <= 0) return;
        if (g.strand >= 0) {
            int64_t p = st.s - g.loc; ++g.out[p > 0 ? p : 0];
            p = st.e + 1 - g.loc; if (p < g.len) --g.out[p];
        } else {
            int64_t p = g.end() - 1 - st.e; ++g.out[p > 0 ? p : 0];
            p = g.end() - st.s; if (p < g.len) --g.out[p];
        }
    }
}

struct Counters { std::atomic<uint64_t> records{0}, bytes{0}, queries{0}; };

// The sweep of src/bamsignals.cpp:252-289 over regs[first,last) (already sorted), fed by `fetch`.
// `whole_chrom` = scan mode: chunks are never split by maxgap and records come from a sequential pass.
void sweep_indexed(const std::string& path, const Bai& bai, std::vector<Region>& regs, size_t first, size_t last,
                   const Params& P, Counters& ctr) {
    BgzfIn in(path);
    std::vector<uint8_t> buf; Rec r; ReadState st;
    uint64_t nrec = 0;
    size_t processed = first;
    while (processed < last) {
        size_t cs = processed;
        int rid = regs[cs].rid;
        int64_t start = int64_t(regs[cs].loc) - P.ext, end = regs[cs].end() + P.ext;
        size_t ce = cs + 1;
        for (; ce < last; ++ce) {
            int64_t ns = int64_t(regs[ce].loc) - P.ext;
            if (regs[ce].rid != rid || ns - end > P.maxgap) break;
            end = std::max(end, regs[ce].end() + P.ext);
        }
        int64_t qbeg = start < 0 ? 0 : start;
        std::vector<Chunk> chunks = query_chunks(bai, rid, qbeg, end);
        ctr.queries++;
        size_t cur = cs; bool done = false;
        for (size_t c = 0; c < chunks.size() && !done; ++c) {
            in.seek(chunks[c].beg);
            while (in.tell() < chunks[c].end) {
                if (!next_record(in, buf, r)) { done = true; break; }
                ++nrec;
                if (r.tid != rid || r.pos >= end) { done = true; break; }      // iterator end rule
                if (!(int64_t(r.endpos) > qbeg)) continue;                      // iterator overlap rule
                int64_t read_end = set_read(P, r, st);
                if (read_end < 0) continue;
                int64_t ov_start = int64_t(r.pos) - P.ext, ov_end = read_end + P.ext;
                while (cur < ce && ov_start >= regs[cur].end()) ++cur;
                if (cur == ce) { done = true; break; }
                for (size_t g = cur; g < ce && regs[g].loc <= ov_end; ++g) apply(P, st, regs[g]);
            }
        }
        processed = ce;
    }
    ctr.records += nrec; ctr.bytes += in.bytes_inflated;
}

void sweep_scan(const std::string& path, std::vector<Region>& regs, const Params& P, bool brute, Counters& ctr) {
    BgzfIn in(path, 2);
    Header h = read_header(in);
    in.seek(h.first_record_voff);
    std::vector<uint8_t> buf; Rec r; ReadState st;
    // index of first region per rid
    size_t n = regs.size(), cur = 0, ce = 0; int cur_rid = -2; uint64_t nrec = 0;
    while (next_record(in, buf, r)) {
        ++nrec;
        if (r.tid < 0) continue;
        if (r.tid != cur_rid) {
            cur_rid = r.tid;
            cur = std::lower_bound(regs.begin(), regs.end(), cur_rid, [](const Region& g, int t) { return g.rid < t; }) - regs.begin();
            ce = cur; while (ce < n && regs[ce].rid == cur_rid) ++ce;
        }
        if (cur == ce) continue;
        int64_t read_end = set_read(P, r, st);
        if (read_end < 0) continue;
        if (brute) {
            size_t lo = std::lower_bound(regs.begin(), regs.end(), cur_rid, [](const Region& g, int t) { return g.rid < t; }) - regs.begin();
            for (size_t g = lo; g < ce; ++g) apply(P, st, regs[g]);
            continue;
        }
        int64_t ov_start = int64_t(r.pos) - P.ext, ov_end = read_end + P.ext;
        while (cur < ce && ov_start >= regs[cur].end()) ++cur;
        for (size_t g = cur; g < ce && regs[g].loc <= ov_end; ++g) apply(P, st, regs[g]);
    }
    ctr.records += nrec; ctr.bytes += in.bytes_inflated;
}

int64_t layout(int64_t R, const int32_t* width, int binsize, int ss, int64_t* offsets) {
    int64_t mult = ss ? 2 : 1, acc = 0;
    for (int64_t i = 0; i < R; ++i) {
        offsets[i] = acc;
        if (binsize <= 0) acc += mult;                                         // src/bamsignals.cpp:148-169
        else acc += mult * ((int64_t(width[i]) + binsize - 1) / binsize);      // src/bamsignals.cpp:175
    }
    offsets[R] = acc;
    return acc;
}

struct Stats { uint64_t records, bytes_inflated, queries; double seconds; };
thread_local Stats g_stats;

int run(const char* bam, int64_t R, const char* const* seq_levels, int n_levels, const int32_t* seq_idx,
        const int32_t* loc, const int32_t* width, const int8_t* strand, const int32_t* tlen_filter,
        Params P, int32_t* out, const int64_t* out_offsets, int mode, int nthreads) {
    std::string path(bam);
    Header hdr;
    { BgzfIn in(path, 2); hdr = read_header(in); }
    Bai bai;
    bai = load_bai(path);                               // the reference always requires the index (:207-210)
    std::unordered_map<std::string, int> name2id;
    for (size_t i = 0; i < hdr.names.size(); ++i) name2id.emplace(hdr.names[i], int(i));
    std::vector<int> level_rid(n_levels, -2);
    std::vector<Region> regs(R);
    int maxw = -1;
    for (int64_t i = 0; i < R; ++i) {
        int lv = seq_idx[i];
        if (lv < 0 || lv >= n_levels) fail("region refers to an unknown seqlevel index");
        if (level_rid[lv] == -2) {
            auto it = name2id.find(seq_levels[lv]);
            level_rid[lv] = it == name2id.end() ? -1 : it->second;
        }
        if (level_rid[lv] < 0) fail(std::string("chromosome ") + seq_levels[lv] + " not present in the bam file"); // :119
        regs[i] = Region{level_rid[lv], loc[i], width[i], strand[i], out + out_offsets[i]};
        maxw = std::max(maxw, width[i]);
    }
    int64_t total = out_offsets[R];
    std::fill(out, out + total, 0);                    // allocateList zero-initialises (:152,157,178,184)
    if (!P.coverage && P.binsize <= 0) P.binsize = maxw;   // :162-167
    if (P.have_tlen) { P.tmin = tlen_filter[0]; P.tmax = tlen_filter[1]; }
    if (!P.coverage) P.ext = std::abs(P.shift) + (P.midpoint ? tlen_filter[1] : 0);    // :457
    else P.ext = P.tspan ? tlen_filter[1] : 0;                                         // :487
    if (P.ext < 0) fail("negative 'ext' values don't make sense");                     // :243
    std::sort(regs.begin(), regs.end(), [](const Region& a, const Region& b) {          // :222-226, :246
        return a.rid != b.rid ? a.rid < b.rid : a.loc < b.loc; });
    Counters ctr;
    if (R > 0) {
        if (mode == 0 || mode == 2) sweep_scan(path, regs, P, mode == 2, ctr);
        else {
            // Split the sorted list into contiguous shards at chunk boundaries (disjoint regions => disjoint
            // output memory); one shard per thread.  nthreads == 1 is the reference's behaviour.
            std::vector<size_t> cuts{0};
            if (nthreads > 1) {
                std::vector<size_t> starts;                     // chunk starts, as the reference forms them
                size_t i = 0;
                while (i < regs.size()) {
                    starts.push_back(i);
                    int rid = regs[i].rid; int64_t end = regs[i].end() + P.ext; size_t j = i + 1;
                    for (; j < regs.size(); ++j) {
                        if (regs[j].rid != rid || int64_t(regs[j].loc) - P.ext - end > P.maxgap) break;
                        end = std::max(end, regs[j].end() + P.ext);
                    }
                    i = j;
                }
                for (int t = 1; t < nthreads; ++t) {
                    size_t target = regs.size() * t / nthreads;
                    auto it = std::lower_bound(starts.begin(), starts.end(), target);
                    size_t c = it == starts.end() ? regs.size() : *it;
                    if (c > cuts.back() && c < regs.size()) cuts.push_back(c);
                }
            }
            cuts.push_back(regs.size());
            std::vector<std::thread> th; std::vector<std::string> errs(cuts.size());
            for (size_t s = 0; s + 1 < cuts.size(); ++s)
                th.emplace_back([&, s] {
                    try { sweep_indexed(path, bai, regs, cuts[s], cuts[s + 1], P, ctr); }
                    catch (OracleError& e) { errs[s] = e.msg; }
                });
            for (auto& t : th) t.join();
            for (auto& e : errs) if (!e.empty()) fail(e);
        }
    }
    if (P.coverage)                                     // cumsum, src/bamsignals.cpp:464-470, :490-492
        for (auto& g : regs) {
            if (g.len < 2) continue;
            int32_t acc = g.out[0];
            for (int32_t i = 1; i < g.len; ++i) g.out[i] = (acc += g.out[i]);
        }
    g_stats.records = ctr.records; g_stats.bytes_inflated = ctr.bytes; g_stats.queries = ctr.queries;
    return 0;
}

}  // namespace

extern "C" {

const char* oracle_last_error() { return g_err.c_str(); }

int64_t oracle_layout(int64_t R, const int32_t* width, int binsize, int ss, int64_t* offsets) {
    return layout(R, width, binsize, ss, offsets);
}

// pileup_core restated (src/bamsignals.cpp:444-461); binsize <= 0 means bamCount.
int oracle_pileup(const char* bam, int64_t R, const char* const* seq_levels, int n_levels, const int32_t* seq_idx,
                  const int32_t* loc, const int32_t* width, const int8_t* strand, const int32_t* tlen_filter,
                  int mapqual, int binsize, int shift, int ss, int requiredF, int filteredF, int pe_mid, int maxgap,
                  int32_t* out, const int64_t* out_offsets, int mode, int nthreads) {
    try {
        Params P; P.coverage = false; P.mapqual = mapqual; P.binsize = binsize; P.shift = shift; P.ss = ss != 0;
        P.required = uint32_t(requiredF); P.filtered = uint32_t(filteredF); P.midpoint = pe_mid != 0;
        P.have_tlen = tlen_filter != nullptr; P.maxgap = maxgap;
        return run(bam, R, seq_levels, n_levels, seq_idx, loc, width, strand, tlen_filter, P, out, out_offsets, mode, nthreads);
    } catch (OracleError& e) { g_err = e.msg; return -1; }
}

// coverage_core restated (src/bamsignals.cpp:474-494)
int oracle_coverage(const char* bam, int64_t R, const char* const* seq_levels, int n_levels, const int32_t* seq_idx,
                    const int32_t* loc, const int32_t* width, const int8_t* strand, const int32_t* tlen_filter,
                    int mapqual, int requiredF, int filteredF, int tspan, int maxgap,
                    int32_t* out, const int64_t* out_offsets, int mode, int nthreads) {
    try {
        Params P; P.coverage = true; P.mapqual = mapqual; P.binsize = 1; P.ss = false;
        P.required = uint32_t(requiredF); P.filtered = uint32_t(filteredF); P.tspan = tspan != 0;
        P.have_tlen = tlen_filter != nullptr; P.maxgap = maxgap;
        return run(bam, R, seq_levels, n_levels, seq_idx, loc, width, strand, tlen_filter, P, out, out_offsets, mode, nthreads);
    } catch (OracleError& e) { g_err = e.msg; return -1; }
}

void oracle_stats(uint64_t* records, uint64_t* bytes_inflated, uint64_t* queries) {
    *records = g_stats.records; *bytes_inflated = g_stats.bytes_inflated; *queries = g_stats.queries;
}

// Dump the fields the counting path reads for every record (test helper for the decode kernel):
// cols[6] = tid,pos,endpos,tlen,flag,mapq as int32 arrays of capacity cap; returns record count.
int64_t oracle_dump_reads(const char* bam, int64_t cap, int32_t* tid, int32_t* pos, int32_t* endpos, int32_t* tlen,
                          int32_t* flag, int32_t* mapq) {
    try {
        BgzfIn in(bam, 2); Header h = read_header(in); in.seek(h.first_record_voff);
        std::vector<uint8_t> buf; Rec r; int64_t n = 0;
        while (next_record(in, buf, r)) {
            if (n < cap) { tid[n] = r.tid; pos[n] = r.pos; endpos[n] = r.endpos; tlen[n] = r.tlen; flag[n] = r.flag; mapq[n] = r.mapq; }
            ++n;
        }
        return n;
    } catch (OracleError& e) { g_err = e.msg; return -1; }
}

int oracle_header(const char* bam, int cap, char* names /*cap x 256*/, int32_t* lens) {
    try {
        BgzfIn in(bam, 2); Header h = read_header(in);
        for (int i = 0; i < int(h.names.size()) && i < cap; ++i) { snprintf(names + 256 * i, 256, "%s", h.names[i].c_str()); lens[i] = h.lens[i]; }
        return int(h.names.size());
    } catch (OracleError& e) { g_err = e.msg; return -1; }
}

}  // extern "C"
