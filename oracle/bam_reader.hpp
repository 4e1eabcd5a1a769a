// BGZF / BAM / BAI / CSI reader of the CPU oracle, restated from the SAM/BAM specification and htslib >= 1.10
// behaviour (htslib is an un-vendored dependency of the reference: LinkingTo Rhtslib, DESCRIPTION:30).
// TEST INFRASTRUCTURE ONLY: shared by the restated oracle (bsg_oracle.cpp) and by the htslib stand-in that lets the
// reference's own src/bamsignals.cpp compile unchanged (ref_compat/hts_compat.cpp -> oracle/_ref/).
#pragma once
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

namespace bsgo {

inline thread_local std::string g_err;
struct OracleError { std::string msg; };
[[noreturn]] inline void fail(const std::string& m) { throw OracleError{m}; }

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

    void set_cache_blocks(int n) { cache_.assign(size_t(n < 1 ? 1 : n), Block()); next_slot_ = 0; cur_ = nullptr; upos_ = 0; }
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

inline Header read_header(BgzfIn& in) {
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

inline bool next_record(BgzfIn& in, std::vector<uint8_t>& buf, Rec& r) {
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
inline std::vector<uint8_t> read_maybe_bgzf(FILE* fp) {
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
inline Bai load_bai(const std::string& bampath) {
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
inline void reg2bins(int64_t beg, int64_t end, int min_shift, int depth, std::vector<uint32_t>& out) {
    out.clear();
    if (beg >= end) return;
    const int64_t maxpos = 1LL << (min_shift + 3 * depth);
    if (end > maxpos) end = maxpos;
    if (beg >= end) return;
    --end;
    for (int l = 0, t = 0, s = min_shift + 3 * depth; l <= depth; s -= 3, t += 1 << (3 * l), ++l)
        for (int64_t k = t + (beg >> s); k <= t + (end >> s); ++k) out.push_back(uint32_t(k));
}

inline std::vector<Chunk> query_chunks(const Bai& bai, int tid, int64_t beg, int64_t end) {
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

}  // namespace bsgo
