// bsg_write_sam_as_bam_and_index: SAM text -> coordinate-sorted BAM + BAI, the job of the reference's (test-only)
// writeSamAsBamAndIndex (src/bamsignals.cpp:496-534: sam_open/sam_read1/bam_write1 followed by bam_index_build with
// min_shift 0, i.e. a .bai).  htslib is not available offline, so SAM parsing (SAM spec 1.4), BAM encoding (4.2), BGZF
// framing (4.1) and BAI construction (5.2) are written out here over zlib.  Host-only: no device work.
#include <algorithm>
#include <cerrno>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>
#include <zlib.h>

#include "common.h"

namespace bsg {
namespace {

void put_u16(std::vector<uint8_t>& v, uint32_t x) { v.push_back(uint8_t(x)); v.push_back(uint8_t(x >> 8)); }
void put_u32(std::vector<uint8_t>& v, uint32_t x) { for (int i = 0; i < 4; ++i) v.push_back(uint8_t(x >> (8 * i))); }
void put_u64(std::vector<uint8_t>& v, uint64_t x) { for (int i = 0; i < 8; ++i) v.push_back(uint8_t(x >> (8 * i))); }

// UCSC binning scheme, SAM spec 5.3
int reg2bin(int64_t beg, int64_t end) {
    --end;
    if (beg >> 14 == end >> 14) return int(((1 << 15) - 1) / 7 + (beg >> 14));
    if (beg >> 17 == end >> 17) return int(((1 << 12) - 1) / 7 + (beg >> 17));
    if (beg >> 20 == end >> 20) return int(((1 << 9) - 1) / 7 + (beg >> 20));
    if (beg >> 23 == end >> 23) return int(((1 << 6) - 1) / 7 + (beg >> 23));
    if (beg >> 26 == end >> 26) return int(((1 << 3) - 1) / 7 + (beg >> 26));
    return 0;
}

// BGZF writer with htslib's block policy: a record that would not fit into the current block starts a new one, so
// records straddle blocks only when they are longer than a block.
class BgzfOut {
public:
    explicit BgzfOut(const std::string& path) : path_(path) {
        fp_ = fopen(path.c_str(), "wb");
        if (!fp_) fail(BSG_EOPEN, "Fail to open BAM file ." + path);                     // src/bamsignals.cpp:516 (sic)
    }
    ~BgzfOut() { if (fp_) fclose(fp_); }
    uint64_t tell() const { return coff_ << 16 | uint64_t(buf_.size()); }                // virtual offset of the next byte
    void flush_try(size_t need) { if (buf_.size() + need > kPayload) flush(); }
    void write(const uint8_t* p, size_t n) {
        while (n) {
            const size_t take = std::min(n, kPayload - buf_.size());
            buf_.insert(buf_.end(), p, p + take);
            p += take; n -= take;
            if (buf_.size() == kPayload) flush();
        }
    }
    void flush() {
        if (buf_.empty()) return;
        emit(buf_.data(), buf_.size());
        buf_.clear();
    }
    void close() {
        flush();
        emit(nullptr, 0);                                                                // the 28-byte EOF marker block
        if (fclose(fp_) != 0) { fp_ = nullptr; fail(BSG_EOPEN, "error writing " + path_); }
        fp_ = nullptr;
    }

private:
    static constexpr size_t kPayload = 0xff00;
    void emit(const uint8_t* data, size_t n) {
        uLong bound = compressBound(uLong(n)) + 64;
        std::vector<uint8_t> comp(bound);
        z_stream zs;
        memset(&zs, 0, sizeof zs);
        if (deflateInit2(&zs, Z_DEFAULT_COMPRESSION, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) fail(BSG_ENOMEM, "zlib deflateInit2 failed");
        zs.next_in = const_cast<Bytef*>(data ? data : reinterpret_cast<const uint8_t*>(""));
        zs.avail_in = uInt(n);
        zs.next_out = comp.data();
        zs.avail_out = uInt(comp.size());
        const int rc = deflate(&zs, Z_FINISH);
        const size_t clen = zs.total_out;
        deflateEnd(&zs);
        if (rc != Z_STREAM_END || clen + 26 > 65536) fail(BSG_EFORMAT, "BGZF block does not compress into 64 KiB");
        std::vector<uint8_t> blk = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0};
        put_u16(blk, uint32_t(clen + 25));
        blk.insert(blk.end(), comp.begin(), comp.begin() + long(clen));
        put_u32(blk, uint32_t(crc32(crc32(0L, Z_NULL, 0), data, uInt(n))));
        put_u32(blk, uint32_t(n));
        if (fwrite(blk.data(), 1, blk.size(), fp_) != blk.size()) fail(BSG_EOPEN, "error writing " + path_);
        coff_ += blk.size();
    }
    std::string path_;
    FILE* fp_ = nullptr;
    std::vector<uint8_t> buf_;
    uint64_t coff_ = 0;
};

std::vector<std::string> split_tab(const std::string& line) {
    std::vector<std::string> f;
    size_t a = 0;
    for (;;) {
        const size_t b = line.find('\t', a);
        f.push_back(line.substr(a, b == std::string::npos ? b : b - a));
        if (b == std::string::npos) break;
        a = b + 1;
    }
    return f;
}

long long to_int(const std::string& s, const char* what, size_t lineno) {
    errno = 0;
    char* end = nullptr;
    const long long v = strtoll(s.c_str(), &end, 10);
    if (s.empty() || *end || errno) fail(BSG_EFORMAT, "SAM line " + std::to_string(lineno) + ": bad " + what + " '" + s + "'");
    return v;
}

// one optional field TAG:TYPE:VALUE -> BAM aux bytes (integers take the smallest type that holds them, as htslib does)
void encode_aux(const std::string& f, std::vector<uint8_t>& out, size_t lineno) {
    if (f.size() < 5 || f[2] != ':' || f[4] != ':') fail(BSG_EFORMAT, "SAM line " + std::to_string(lineno) + ": malformed optional field '" + f + "'");
    out.push_back(uint8_t(f[0])); out.push_back(uint8_t(f[1]));
    const char ty = f[3];
    const std::string val = f.substr(5);
    auto put_int = [&](long long v) {
        if (v >= 0) {
            if (v <= 0xff) { out.push_back('C'); out.push_back(uint8_t(v)); }
            else if (v <= 0xffff) { out.push_back('S'); put_u16(out, uint32_t(v)); }
            else if (v <= 0xffffffffLL) { out.push_back('I'); put_u32(out, uint32_t(v)); }
            else fail(BSG_EFORMAT, "SAM line " + std::to_string(lineno) + ": integer tag out of range");
        } else {
            if (v >= -128) { out.push_back('c'); out.push_back(uint8_t(int8_t(v))); }
            else if (v >= -32768) { out.push_back('s'); put_u16(out, uint32_t(uint16_t(int16_t(v)))); }
            else if (v >= INT_MIN) { out.push_back('i'); put_u32(out, uint32_t(int32_t(v))); }
            else fail(BSG_EFORMAT, "SAM line " + std::to_string(lineno) + ": integer tag out of range");
        }
    };
    if (ty == 'A') { if (val.size() != 1) fail(BSG_EFORMAT, "SAM line " + std::to_string(lineno) + ": bad A tag"); out.push_back('A'); out.push_back(uint8_t(val[0])); }
    else if (ty == 'i') put_int(to_int(val, "integer tag", lineno));
    else if (ty == 'f') { const float x = strtof(val.c_str(), nullptr); uint32_t u; memcpy(&u, &x, 4); out.push_back('f'); put_u32(out, u); }
    else if (ty == 'Z' || ty == 'H') { out.push_back(uint8_t(ty)); out.insert(out.end(), val.begin(), val.end()); out.push_back(0); }
    else if (ty == 'B') {
        if (val.empty()) fail(BSG_EFORMAT, "SAM line " + std::to_string(lineno) + ": bad B tag");
        const char sub = val[0];
        std::vector<std::string> items;
        size_t a = 1;
        while (a < val.size() && val[a] == ',') {
            const size_t b = val.find(',', a + 1);
            items.push_back(val.substr(a + 1, b == std::string::npos ? b : b - a - 1));
            if (b == std::string::npos) break;
            a = b;
        }
        out.push_back('B'); out.push_back(uint8_t(sub)); put_u32(out, uint32_t(items.size()));
        for (const std::string& it : items) {
            if (sub == 'f') { const float x = strtof(it.c_str(), nullptr); uint32_t u; memcpy(&u, &x, 4); put_u32(out, u); continue; }
            const long long v = to_int(it, "array element", lineno);
            if (sub == 'c' || sub == 'C') out.push_back(uint8_t(v));
            else if (sub == 's' || sub == 'S') put_u16(out, uint32_t(v));
            else if (sub == 'i' || sub == 'I') put_u32(out, uint32_t(v));
            else fail(BSG_EFORMAT, "SAM line " + std::to_string(lineno) + ": bad B subtype");
        }
    } else fail(BSG_EFORMAT, "SAM line " + std::to_string(lineno) + ": unknown tag type '" + std::string(1, ty) + "'");
}

struct RefIdx {
    std::map<uint32_t, std::vector<std::pair<uint64_t, uint64_t>>> bins;
    std::vector<uint64_t> linear;
    uint64_t beg = 0, end = 0, n_mapped = 0, n_unmapped = 0;
    bool any = false;
};

}  // namespace

void write_sam_as_bam_and_index(const char* sampath, const char* bampath) {
    const std::string sp = sampath ? sampath : "", bp = bampath ? bampath : "";
    FILE* in = fopen(sp.c_str(), "r");
    if (!in) fail(BSG_EOPEN, "Fail to open SAM file " + sp);                             // src/bamsignals.cpp:507
    std::vector<std::string> lines;
    {
        std::string cur;
        char buf[1 << 16];
        size_t k;
        while ((k = fread(buf, 1, sizeof buf, in)) > 0)
            for (size_t i = 0; i < k; ++i) {
                if (buf[i] == '\n') { if (!cur.empty() && cur.back() == '\r') cur.pop_back(); lines.push_back(cur); cur.clear(); }
                else cur.push_back(buf[i]);
            }
        if (!cur.empty()) lines.push_back(cur);
        fclose(in);
    }
    // ---- header -------------------------------------------------------------------------------------------------------
    std::string text;
    std::vector<std::pair<std::string, int64_t>> refs;
    std::unordered_map<std::string, int> name2id;
    size_t first_aln = 0;
    for (; first_aln < lines.size() && !lines[first_aln].empty() && lines[first_aln][0] == '@'; ++first_aln) {
        const std::string& l = lines[first_aln];
        text += l; text += '\n';
        if (l.compare(0, 3, "@SQ") == 0) {
            std::string sn; int64_t ln = -1;
            for (const std::string& f : split_tab(l)) {
                if (f.compare(0, 3, "SN:") == 0) sn = f.substr(3);
                else if (f.compare(0, 3, "LN:") == 0) ln = to_int(f.substr(3), "@SQ LN", first_aln + 1);
            }
            if (sn.empty() || ln < 0 || ln > INT32_MAX) fail(BSG_EFORMAT, "SAM line " + std::to_string(first_aln + 1) + ": bad @SQ line");
            name2id.emplace(sn, int(refs.size()));
            refs.emplace_back(sn, ln);
        }
    }
    BgzfOut out(bp);
    {
        std::vector<uint8_t> h = {'B', 'A', 'M', 1};
        put_u32(h, uint32_t(text.size()));
        h.insert(h.end(), text.begin(), text.end());
        put_u32(h, uint32_t(refs.size()));
        for (auto& r : refs) {
            put_u32(h, uint32_t(r.first.size() + 1));
            h.insert(h.end(), r.first.begin(), r.first.end());
            h.push_back(0);
            put_u32(h, uint32_t(r.second));
        }
        out.write(h.data(), h.size());
        out.flush();                                  // htslib starts the records in a fresh block
    }
    // ---- records + index -----------------------------------------------------------------------------------------------
    std::vector<RefIdx> idx(refs.size());
    uint64_t n_no_coor = 0;
    int cur_tid = -2, cur_bin = -1;
    uint64_t run_beg = 0, run_end = 0;
    int last_tid = -1; int64_t last_pos = -1;
    auto close_run = [&]() {
        if (cur_tid >= 0 && cur_bin >= 0) idx[size_t(cur_tid)].bins[uint32_t(cur_bin)].emplace_back(run_beg, run_end);
        cur_tid = -2; cur_bin = -1;
    };
    static const char kSeqCode[] = "=ACMGRSVTWYHKDBN";
    std::vector<uint8_t> rec;
    for (size_t li = first_aln; li < lines.size(); ++li) {
        if (lines[li].empty()) continue;
        const std::vector<std::string> f = split_tab(lines[li]);
        if (f.size() < 11) fail(BSG_EFORMAT, "SAM line " + std::to_string(li + 1) + ": fewer than 11 fields");
        const long long flag = to_int(f[1], "FLAG", li + 1);
        int tid = -1;
        if (f[2] != "*") {
            auto it = name2id.find(f[2]);
            if (it == name2id.end()) fail(BSG_EFORMAT, "SAM line " + std::to_string(li + 1) + ": unknown reference '" + f[2] + "'");
            tid = it->second;
        }
        const int64_t pos = to_int(f[3], "POS", li + 1) - 1;
        const long long mapq = to_int(f[4], "MAPQ", li + 1);
        std::vector<uint32_t> cigar;
        int64_t rlen = 0, qlen = 0;
        if (f[5] != "*") {
            long long num = 0; bool have = false;
            for (char ch : f[5]) {
                if (ch >= '0' && ch <= '9') { num = num * 10 + (ch - '0'); have = true; if (num > 0xfffffff) fail(BSG_EFORMAT, "SAM line " + std::to_string(li + 1) + ": CIGAR length too large"); continue; }
                const char* ops = "MIDNSHP=X";
                const char* q = strchr(ops, ch);
                if (!q || !have) fail(BSG_EFORMAT, "SAM line " + std::to_string(li + 1) + ": bad CIGAR '" + f[5] + "'");
                const int op = int(q - ops);
                cigar.push_back(uint32_t(num) << 4 | uint32_t(op));
                if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) rlen += num;
                if (op == 0 || op == 1 || op == 4 || op == 7 || op == 8) qlen += num;
                num = 0; have = false;
            }
            if (have) fail(BSG_EFORMAT, "SAM line " + std::to_string(li + 1) + ": bad CIGAR '" + f[5] + "'");
        }
        if (cigar.size() > 0xffff) fail(BSG_EFORMAT, "SAM line " + std::to_string(li + 1) + ": more than 65535 CIGAR operations are not supported");
        int next_tid = -1;
        if (f[6] == "=") next_tid = tid;
        else if (f[6] != "*") {
            auto it = name2id.find(f[6]);
            if (it == name2id.end()) fail(BSG_EFORMAT, "SAM line " + std::to_string(li + 1) + ": unknown reference '" + f[6] + "'");
            next_tid = it->second;
        }
        const int64_t next_pos = to_int(f[7], "PNEXT", li + 1) - 1;
        const long long tlen = to_int(f[8], "TLEN", li + 1);
        const std::string& seq = f[9];
        const std::string& qual = f[10];
        const size_t l_seq = seq == "*" ? 0 : seq.size();
        if (qual != "*" && qual.size() != l_seq) fail(BSG_EFORMAT, "SAM line " + std::to_string(li + 1) + ": SEQ and QUAL differ in length");
        if (f[0].empty() || f[0].size() > 254) fail(BSG_EFORMAT, "SAM line " + std::to_string(li + 1) + ": bad QNAME");
        if (pos < -1 || pos > INT32_MAX || next_pos < -1 || next_pos > INT32_MAX || mapq < 0 || mapq > 255 || flag < 0 || flag > 0xffff ||
            tlen < INT_MIN || tlen > INT_MAX)
            fail(BSG_EFORMAT, "SAM line " + std::to_string(li + 1) + ": field out of range");
        const bool unmapped = (flag & 4) != 0;
        const int64_t end = pos + ((unmapped || rlen == 0) ? 1 : rlen);                    // bam_endpos
        // coordinate order is what the index (and the counting path) rely on; samtools index refuses such input too
        if (tid >= 0) {
            if (last_tid != -1 && (tid < last_tid || (tid == last_tid && pos < last_pos)))
                fail(BSG_EUNSORTED, "SAM line " + std::to_string(li + 1) + ": records are not coordinate-sorted");
            last_tid = tid; last_pos = pos;
        } else last_tid = INT_MAX;                       // unplaced reads close the file
        rec.clear();
        put_u32(rec, 0);                                                                   // block_size, patched below
        put_u32(rec, uint32_t(tid)); put_u32(rec, uint32_t(int32_t(pos)));
        rec.push_back(uint8_t(f[0].size() + 1)); rec.push_back(uint8_t(mapq));
        put_u16(rec, uint32_t(pos >= 0 ? reg2bin(pos, end) : 4680) & 0xffffu);
        put_u16(rec, uint32_t(cigar.size())); put_u16(rec, uint32_t(flag));
        put_u32(rec, uint32_t(l_seq));
        put_u32(rec, uint32_t(next_tid)); put_u32(rec, uint32_t(int32_t(next_pos))); put_u32(rec, uint32_t(int32_t(tlen)));
        rec.insert(rec.end(), f[0].begin(), f[0].end()); rec.push_back(0);
        for (uint32_t c : cigar) put_u32(rec, c);
        for (size_t i = 0; i < l_seq; i += 2) {
            auto code = [&](char ch) { const char* q = strchr(kSeqCode, ch >= 'a' && ch <= 'z' ? ch - 32 : ch); return q && ch ? uint8_t(q - kSeqCode) : uint8_t(15); };
            rec.push_back(uint8_t(code(seq[i]) << 4 | (i + 1 < l_seq ? code(seq[i + 1]) : 0)));
        }
        for (size_t i = 0; i < l_seq; ++i) rec.push_back(qual == "*" ? uint8_t(0xff) : uint8_t(qual[i] - 33));
        for (size_t k = 11; k < f.size(); ++k) if (!f[k].empty()) encode_aux(f[k], rec, li + 1);
        const uint32_t bs = uint32_t(rec.size() - 4);
        memcpy(rec.data(), &bs, 4);
        out.flush_try(rec.size());
        const uint64_t vb = out.tell();
        out.write(rec.data(), rec.size());
        const uint64_t ve = out.tell();
        // ---- index bookkeeping (SAM spec 5.2; what bam_index_build records) ----------------------------------------------
        if (tid < 0) { close_run(); ++n_no_coor; continue; }
        if (end > (int64_t(1) << 29)) fail(BSG_EFORMAT, "SAM line " + std::to_string(li + 1) + ": position beyond 2^29 cannot be indexed by a BAI");
        RefIdx& ri = idx[size_t(tid)];
        const int bin = reg2bin(pos, end);
        if (tid != cur_tid || bin != cur_bin) { close_run(); cur_tid = tid; cur_bin = bin; run_beg = vb; }
        run_end = ve;
        const size_t w0 = size_t(pos >> 14), w1 = size_t((end - 1) >> 14);
        if (ri.linear.size() <= w1) ri.linear.resize(w1 + 1, 0);
        for (size_t w = w0; w <= w1; ++w) if (!ri.linear[w]) ri.linear[w] = vb;
        if (!ri.any) { ri.any = true; ri.beg = vb; }
        ri.end = ve;
        if (unmapped) ++ri.n_unmapped; else ++ri.n_mapped;
    }
    close_run();
    out.close();
    // ---- BAI -------------------------------------------------------------------------------------------------------------
    std::vector<uint8_t> bai = {'B', 'A', 'I', 1};
    put_u32(bai, uint32_t(refs.size()));
    for (RefIdx& ri : idx) {
        put_u32(bai, uint32_t(ri.bins.size() + (ri.any ? 1 : 0)));
        for (auto& b : ri.bins) {
            put_u32(bai, b.first); put_u32(bai, uint32_t(b.second.size()));
            for (auto& c : b.second) { put_u64(bai, c.first); put_u64(bai, c.second); }
        }
        if (ri.any) {
            put_u32(bai, 37450); put_u32(bai, 2);
            put_u64(bai, ri.beg); put_u64(bai, ri.end); put_u64(bai, ri.n_mapped); put_u64(bai, ri.n_unmapped);
        }
        // windows no record overlaps take the offset of the next record, as samtools-written indexes do
        for (size_t w = ri.linear.size(); w-- > 1;) if (!ri.linear[w - 1]) ri.linear[w - 1] = ri.linear[w];
        put_u32(bai, uint32_t(ri.linear.size()));
        for (uint64_t v : ri.linear) put_u64(bai, v);
    }
    put_u64(bai, n_no_coor);
    const std::string ip = bp + ".bai";
    FILE* fi = fopen(ip.c_str(), "wb");
    if (!fi || fwrite(bai.data(), 1, bai.size(), fi) != bai.size()) { if (fi) fclose(fi); fail(BSG_EOPEN, "error writing " + ip); }
    fclose(fi);
}

}  // namespace bsg
