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
#include <algorithm>
#include <atomic>
#include <cerrno>
#include <climits>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>
#include <zlib.h>

namespace {

thread_local std::string g_err;
struct OracleError { std::string msg; };
[[noreturn]] void fail(const std::string& m) { throw OracleError{m}; }

// ---------------------------------------------------------------------------------------------
// BGZF reader with a small LRU-less block cache (the reference asks htslib for a 10-block cache,
// src/bamsignals.cpp:200,212).
// ---------------------------------------------------------------------------------------------
struct Block { int64_t coff = -1; uint32_t csize = 0; std::vector<uint8_t> data; };

class BgzfIn {
public:
    explicit BgzfIn(const std::string& path, int cache_blocks = 10) : cache_(cache_blocks) {
        fp_ = fopen(path.c_str(), "rb");
        if (!fp_) fail("Fail to open BAM file " + path);                 // src/bamsignals.cpp:204
        memset(&zs_, 0, sizeof zs_);
        if (inflateInit2(&zs_, -15) != Z_OK) fail("zlib init failed");
    }
    ~BgzfIn() { if (fp_) fclose(fp_); inflateEnd(&zs_); }
    BgzfIn(const BgzfIn&) = delete;

    void seek(uint64_t voff) { load(int64_t(voff >> 16)); upos_ = uint32_t(voff & 0xffff); }
    uint64_t tell() {
        if (cur_ && upos_ == cur_->data.size() && cur_->csize) { // normalise to the next block start
            return uint64_t(cur_->coff + cur_->csize) << 16;
        }
        return cur_ ? (uint64_t(cur_->coff) << 16 | upos_) : 0;
    }
    // read exactly n bytes; returns false on clean EOF at a record boundary (n bytes, zero available)
    bool read(void* dst, size_t n) {
        uint8_t* out = static_cast<uint8_t*>(dst);
        size_t got = 0;
        while (got < n) {
            if (!cur_) { if (!load(0)) return false; }
            size_t avail = cur_->data.size() - upos_;
            if (avail == 0) {
                if (cur_->csize == 0 || !load(cur_->coff + cur_->csize)) {
                    if (got == 0) return false;
                    fail("truncated BAM record");
                }
                continue;
            }
            size_t take = std::min(avail, n - got);
            memcpy(out + got, cur_->data.data() + upos_, take);
            got += take; upos_ += uint32_t(take);
        }
        return true;
    }
    uint64_t bytes_inflated = 0;

private:
    bool load(int64_t coff) {
        for (auto& b : cache_) if (b.coff == coff && b.csize) { cur_ = &b; upos_ = 0; return true; }
        Block& b = cache_[next_slot_]; next_slot_ = (next_slot_ + 1) % cache_.size();
        b.coff = -1; b.csize = 0; b.data.clear();
        if (fseeko(fp_, coff, SEEK_SET) != 0) { cur_ = nullptr; return false; }
        uint8_t h[18];
        size_t k = fread(h, 1, 18, fp_);
        if (k == 0) { cur_ = nullptr; return false; }
        if (k != 18 || h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4)) fail("bad BGZF header");
        uint32_t xlen = h[10] | h[11] << 8;
        // locate the BC subfield; htslib-written files have it first with XLEN 6
        std::vector<uint8_t> extra(xlen);
        memcpy(extra.data(), h + 12, std::min<size_t>(6, xlen));
        if (xlen > 6 && fread(extra.data() + 6, 1, xlen - 6, fp_) != xlen - 6) fail("bad BGZF header");
        int bsize = -1;
        for (uint32_t p = 0; p + 4 <= xlen;) {
            uint32_t slen = extra[p + 2] | extra[p + 3] << 8;
            if (extra[p] == 'B' && extra[p + 1] == 'C' && slen == 2 && p + 6 <= xlen) bsize = extra[p + 4] | extra[p + 5] << 8;
            p += 4 + slen;
        }
        if (bsize < 0) fail("BGZF block without BC field");
        uint32_t total = uint32_t(bsize) + 1, hdr = 12 + xlen;
        if (total < hdr + 8) fail("bad BGZF block size");
        cbuf_.resize(total - hdr);
        if (fread(cbuf_.data(), 1, cbuf_.size(), fp_) != cbuf_.size()) fail("truncated BGZF block");
        const uint8_t* tail = cbuf_.data() + cbuf_.size() - 8;
        uint32_t crc = tail[0] | tail[1] << 8 | tail[2] << 16 | uint32_t(tail[3]) << 24;
        uint32_t isize = tail[4] | tail[5] << 8 | tail[6] << 16 | uint32_t(tail[7]) << 24;
        if (isize > 65536) fail("bad BGZF ISIZE");
        b.data.resize(isize);
        inflateReset(&zs_);
        zs_.next_in = cbuf_.data(); zs_.avail_in = uInt(cbuf_.size() - 8);
        uint8_t none = 0;                                   // zlib rejects a null next_out, even for the empty EOF block
        zs_.next_out = isize ? b.data.data() : &none; zs_.avail_out = isize;
        int rc = inflate(&zs_, Z_FINISH);
        if (!(rc == Z_STREAM_END && zs_.avail_out == 0)) fail("BGZF inflate failed");
        if (uint32_t(crc32(crc32(0, nullptr, 0), b.data.data(), isize)) != crc) fail("BGZF CRC mismatch");
        bytes_inflated += isize;
        b.coff = coff; b.csize = total; cur_ = &b; upos_ = 0;
        return true;
    }
    FILE* fp_ = nullptr;
    z_stream zs_;
    std::vector<Block> cache_;
    size_t next_slot_ = 0;
    Block* cur_ = nullptr;
    uint32_t upos_ = 0;
    std::vector<uint8_t> cbuf_;
};

inline int32_t rd_i32(const uint8_t* p) { int32_t v; memcpy(&v, p, 4); return v; }
inline uint32_t rd_u32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
inline uint16_t rd_u16(const uint8_t* p) { uint16_t v; memcpy(&v, p, 2); return v; }

struct Header { std::vector<std::string> names; std::vector<int32_t> lens; uint64_t first_record_voff = 0; };

Header read_header(BgzfIn& in) {
    Header h; uint8_t b[8];
    in.seek(0);
    if (!in.read(b, 8) || memcmp(b, "BAM\1", 4) != 0) fail("not a BAM file");
    int32_t l_text = rd_i32(b + 4);
    std::vector<uint8_t> skip(l_text > 0 ? l_text : 0);
    if (l_text > 0 && !in.read(skip.data(), l_text)) fail("truncated BAM header");
    if (!in.read(b, 4)) fail("truncated BAM header");
    int32_t n_ref = rd_i32(b);
    for (int i = 0; i < n_ref; ++i) {
        if (!in.read(b, 4)) fail("truncated BAM header");
        int32_t l_name = rd_i32(b);
        std::string nm(l_name, '\0');
        if (!in.read(&nm[0], l_name) || !in.read(b, 4)) fail("truncated BAM header");
        nm.resize(strlen(nm.c_str()));
        h.names.push_back(nm); h.lens.push_back(rd_i32(b));
    }
    h.first_record_voff = in.tell();
    return h;
}

// One alignment record, only the fields the counting path reads.
struct Rec { int32_t tid, pos, tlen; uint16_t flag; uint8_t mapq; int32_t endpos; };

// htslib bam_endpos (external; call site src/bamsignals.cpp:16-18)
inline int32_t end_position(int32_t pos, uint16_t flag, const uint8_t* cigar, uint32_t n_cigar) {
    int64_t rlen = 0;
    if (!(flag & 0x4))
        for (uint32_t k = 0; k < n_cigar; ++k) {
            uint32_t c = rd_u32(cigar + 4 * k), op = c & 0xf;
            if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) rlen += c >> 4;   // M D N = X
        }
    if (rlen == 0) rlen = 1;
    return int32_t(pos + rlen);
}

bool next_record(BgzfIn& in, std::vector<uint8_t>& buf, Rec& r) {
    uint8_t b4[4];
    if (!in.read(b4, 4)) return false;
    int32_t bs = rd_i32(b4);
    if (bs < 32) fail("corrupt BAM record (block_size < 32)");
    buf.resize(bs);
    if (!in.read(buf.data(), bs)) fail("truncated BAM record");
    const uint8_t* p = buf.data();
    r.tid = rd_i32(p); r.pos = rd_i32(p + 4);
    uint32_t l_name = p[8]; r.mapq = p[9];
    uint32_t n_cigar = rd_u16(p + 12); r.flag = rd_u16(p + 14);
    r.tlen = rd_i32(p + 28);
    if (32 + l_name + 4ull * n_cigar > uint64_t(bs)) fail("corrupt BAM record (cigar beyond record)");
    r.endpos = end_position(r.pos, r.flag, p + 32 + l_name, n_cigar);
    return true;
}

// ---------------------------------------------------------------------------------------------
// BAI / CSI index and the htslib-style region query
// ---------------------------------------------------------------------------------------------
struct Chunk { uint64_t beg, end; };
struct RefIndex {
    std::unordered_map<uint32_t, std::vector<Chunk>> bins;
    std::unordered_map<uint32_t, uint64_t> loff;    // CSI only: per-bin "loffset"
    std::vector<uint64_t> linear;                    // BAI only
};
struct Bai {                                         // a BAI (min_shift 14, depth 5) or a CSI index
    std::vector<RefIndex> refs;
    int min_shift = 14, depth = 5;
    bool csi = false;
};

// A .csi file is BGZF-compressed; a .bai is not.
std::vector<uint8_t> read_maybe_bgzf(FILE* fp) {
    std::vector<uint8_t> d;
    uint8_t tmp[65536]; size_t k;
    while ((k = fread(tmp, 1, sizeof tmp, fp)) > 0) d.insert(d.end(), tmp, tmp + k);
    if (d.size() < 18 || d[0] != 0x1f || d[1] != 0x8b) return d;
    std::vector<uint8_t> out;
    size_t p = 0;
    while (p + 18 <= d.size()) {
        const uint8_t* h = d.data() + p;
        if (h[0] != 0x1f || h[1] != 0x8b) fail("bad BGZF header in index");
        const uint32_t xlen = h[10] | h[11] << 8;
        const uint32_t bsize = (h[16] | h[17] << 8) + 1u;        // htslib writes BC first with XLEN 6
        if (xlen != 6 || h[12] != 'B' || h[13] != 'C' || p + bsize > d.size()) fail("unsupported BGZF layout in index");
        uint32_t isize; memcpy(&isize, h + bsize - 4, 4);
        const size_t old = out.size();
        out.resize(old + isize);
        if (isize) {
            z_stream zs; memset(&zs, 0, sizeof zs);
            if (inflateInit2(&zs, -15) != Z_OK) fail("zlib init failed");
            zs.next_in = const_cast<Bytef*>(h + 18); zs.avail_in = bsize - 18 - 8;
            zs.next_out = out.data() + old; zs.avail_out = isize;
            const int rc = inflate(&zs, Z_FINISH);
            inflateEnd(&zs);
            if (rc != Z_STREAM_END || zs.avail_out) fail("index inflate failed");
        }
        p += bsize;
    }
    return out;
}

// htslib looks for <bam>.csi, then <bam minus extension>.csi, then the same two with .bai (hts_idx_load, HTS_FMT_BAI)
Bai load_bai(const std::string& bampath) {
    std::vector<std::string> cand = {bampath + ".csi"};
    const bool ext = bampath.size() > 4 && bampath.compare(bampath.size() - 4, 4, ".bam") == 0;
    if (ext) cand.push_back(bampath.substr(0, bampath.size() - 4) + ".csi");
    cand.push_back(bampath + ".bai");
    if (ext) cand.push_back(bampath.substr(0, bampath.size() - 4) + ".bai");
    FILE* fp = nullptr;
    for (auto& c : cand) { fp = fopen(c.c_str(), "rb"); if (fp) break; }
    if (!fp) fail("BAM indexing file is not available for file " + bampath);   // src/bamsignals.cpp:209
    std::vector<uint8_t> d = read_maybe_bgzf(fp);
    fclose(fp);
    size_t p = 0;
    auto need = [&](size_t n) { if (p + n > d.size()) fail("truncated BAM index"); };
    need(8);
    Bai bai;
    if (memcmp(d.data(), "CSI\1", 4) == 0) {            // CSIv1: magic, min_shift, depth, l_aux, aux, n_ref
        bai.csi = true;
        need(16);
        bai.min_shift = rd_i32(d.data() + 4); bai.depth = rd_i32(d.data() + 8);
        const int32_t l_aux = rd_i32(d.data() + 12);
        p = 16; need(size_t(l_aux) + 4); p += l_aux;
        if (bai.min_shift < 1 || bai.depth < 1 || bai.min_shift + 3 * bai.depth > 40) fail("unsupported CSI geometry");
    } else {
        if (memcmp(d.data(), "BAI\1", 4) != 0) fail("bad BAI magic");
        p = 4;
    }
    int32_t n_ref = rd_i32(d.data() + p); p += 4;
    bai.refs.resize(n_ref);
    for (int r = 0; r < n_ref; ++r) {
        need(4); int32_t n_bin = rd_i32(d.data() + p); p += 4;
        for (int b = 0; b < n_bin; ++b) {
            need(bai.csi ? 16 : 8);
            uint32_t bin = rd_u32(d.data() + p); p += 4;
            if (bai.csi) { uint64_t lo; memcpy(&lo, d.data() + p, 8); p += 8; bai.refs[r].loff[bin] = lo; }
            int32_t n_chunk = rd_i32(d.data() + p); p += 4;
            need(16ull * n_chunk);
            std::vector<Chunk> cs(n_chunk);
            for (int c = 0; c < n_chunk; ++c) { memcpy(&cs[c].beg, d.data() + p, 8); memcpy(&cs[c].end, d.data() + p + 8, 8); p += 16; }
            bai.refs[r].bins[bin] = std::move(cs);
        }
        if (bai.csi) continue;
        need(4); int32_t n_intv = rd_i32(d.data() + p); p += 4;
        need(8ull * n_intv);
        bai.refs[r].linear.resize(n_intv);
        if (n_intv) memcpy(bai.refs[r].linear.data(), d.data() + p, 8ull * n_intv);
        p += 8ull * n_intv;
    }
    return bai;
}

// Binning scheme (SAM spec section 5.3, generalised in CSIv1): bins that may hold records overlapping [beg,end).
// Level l has 8^l bins of 2^(min_shift + 3 (depth - l)) bp; its first bin number is (8^l - 1) / 7.
void reg2bins(int64_t beg, int64_t end, int min_shift, int depth, std::vector<uint32_t>& out) {
    out.clear();
    if (beg >= end) return;
    const int64_t maxpos = 1LL << (min_shift + 3 * depth);
    if (end > maxpos) end = maxpos;
    if (beg >= end) return;
    --end;
    for (int l = 0, t = 0, s = min_shift + 3 * depth; l <= depth; s -= 3, t += 1 << (3 * l), ++l)
        for (int64_t k = t + (beg >> s); k <= t + (end >> s); ++k) out.push_back(uint32_t(k));
}

std::vector<Chunk> query_chunks(const Bai& bai, int tid, int64_t beg, int64_t end) {
    std::vector<Chunk> res;
    if (tid < 0 || tid >= int(bai.refs.size())) return res;
    const RefIndex& ri = bai.refs[tid];
    if (beg < 0) beg = 0;
    if (end <= beg) return res;
    uint64_t min_off = 0;
    if (!bai.csi) {
        if (!ri.linear.empty()) {
            size_t w = size_t(beg >> 14);
            min_off = w < ri.linear.size() ? ri.linear[w] : ri.linear.back();
        }
    } else {
        // htslib's hts_itr_query: the loffset of the leaf bin holding `beg`; if that bin does not exist, of the nearest
        // existing bin to its left on the same level below the same parent, else of the parent, and so on up to bin 0
        uint32_t bin = uint32_t(((1ull << (3 * bai.depth)) - 1) / 7 + (uint64_t(beg) >> bai.min_shift));
        for (;;) {
            auto it = ri.loff.find(bin);
            if (it != ri.loff.end()) { min_off = it->second; break; }
            if (bin == 0) break;
            const uint32_t parent = (bin - 1) >> 3, first = (parent << 3) + 1;
            bin = bin > first ? bin - 1 : parent;
        }
    }
    std::vector<uint32_t> bins; reg2bins(beg, end, bai.min_shift, bai.depth, bins);
    for (uint32_t b : bins) {
        auto it = ri.bins.find(b);
        if (it == ri.bins.end()) continue;
        for (const Chunk& c : it->second) if (c.end > min_off) res.push_back(c);
    }
    std::sort(res.begin(), res.end(), [](const Chunk& a, const Chunk& b) { return a.beg < b.beg; });
    std::vector<Chunk> merged;
    for (const Chunk& c : res) {
        if (!merged.empty() && c.beg <= merged.back().end) merged.back().end = std::max(merged.back().end, c.end);
        else merged.push_back(c);
    }
    return merged;
}

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
