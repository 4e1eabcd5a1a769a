#include "bamio.h"

#include <algorithm>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

namespace bsg {

// ---------------------------------------------------------------------------------------------------------------
// Inflater
// ---------------------------------------------------------------------------------------------------------------
Inflater::Inflater() {
    z_stream* zs = new z_stream;
    memset(zs, 0, sizeof *zs);
    if (inflateInit2(zs, -15) != Z_OK) { delete zs; fail(BSG_ENOMEM, "zlib inflateInit2 failed"); }
    zs_ = zs;
}
Inflater::~Inflater() {
    z_stream* zs = static_cast<z_stream*>(zs_);
    inflateEnd(zs);
    delete zs;
}
void Inflater::inflate_block(const uint8_t* file, const BlockInfo& b, uint8_t* dst, bool verify_crc) {
    z_stream* zs = static_cast<z_stream*>(zs_);
    if (b.isize == 0) return;
    inflateReset(zs);
    zs->next_in = const_cast<Bytef*>(file + b.coff + b.hdr);
    zs->avail_in = b.csize - b.hdr - 8;
    zs->next_out = dst;
    zs->avail_out = b.isize;
    int rc = inflate(zs, Z_FINISH);
    if (!(rc == Z_STREAM_END && zs->avail_out == 0))
        fail(BSG_EFORMAT, "BGZF inflate failed at file offset " + std::to_string(b.coff));
    if (verify_crc && uint32_t(crc32(crc32(0L, Z_NULL, 0), dst, b.isize)) != b.crc)
        fail(BSG_EFORMAT, "BGZF CRC32 mismatch at file offset " + std::to_string(b.coff));
}

// ---------------------------------------------------------------------------------------------------------------
// BamFile
// ---------------------------------------------------------------------------------------------------------------
BamFile::BamFile(const std::string& path) : path_(path) {
    fd_ = ::open(path.c_str(), O_RDONLY);
    if (fd_ < 0) fail(BSG_EOPEN, "Fail to open BAM file " + path);                       // src/bamsignals.cpp:204
    struct stat st;
    if (fstat(fd_, &st) != 0 || !S_ISREG(st.st_mode)) { ::close(fd_); fd_ = -1; fail(BSG_EOPEN, "Fail to open BAM file " + path); }
    size_ = uint64_t(st.st_size);
    mtime_ns_ = uint64_t(st.st_mtim.tv_sec) * 1000000000ull + uint64_t(st.st_mtim.tv_nsec);
    if (size_ == 0) { ::close(fd_); fd_ = -1; fail(BSG_EOPEN, "Fail to open BAM file " + path); }
    void* p = mmap(nullptr, size_, PROT_READ, MAP_SHARED, fd_, 0);
    if (p == MAP_FAILED) { ::close(fd_); fd_ = -1; fail(BSG_EOPEN, "Fail to open BAM file " + path); }
    data_ = static_cast<const uint8_t*>(p);
    try {
        parse_header();
        load_index();
    } catch (...) {
        munmap(const_cast<uint8_t*>(data_), size_);
        ::close(fd_);
        throw;
    }
}

BamFile::~BamFile() {
    if (data_) munmap(const_cast<uint8_t*>(data_), size_);
    if (fd_ >= 0) ::close(fd_);
}

int BamFile::name2id(const std::string& name) const {
    auto it = name2id_.find(name);
    return it == name2id_.end() ? -1 : it->second;
}

bool BamFile::block_at(uint64_t coff, BlockInfo* b) const {
    if (coff >= size_) return false;
    if (coff + 18 > size_) fail(BSG_EFORMAT, "truncated BGZF block header in " + path_);
    const uint8_t* h = data_ + coff;
    if (h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4)) fail(BSG_EFORMAT, "bad BGZF block header in " + path_);
    uint32_t xlen = rd_u16(h + 10);
    if (coff + 12 + xlen > size_) fail(BSG_EFORMAT, "truncated BGZF block header in " + path_);
    int bsize = -1;
    for (uint32_t p = 0; p + 4 <= xlen;) {
        const uint8_t* e = h + 12 + p;
        uint32_t slen = rd_u16(e + 2);
        if (e[0] == 'B' && e[1] == 'C' && slen == 2 && p + 6 <= xlen) bsize = rd_u16(e + 4);
        p += 4 + slen;
    }
    if (bsize < 0) fail(BSG_EFORMAT, "BGZF block without BC subfield in " + path_);
    b->coff = coff;
    b->csize = uint32_t(bsize) + 1;
    b->hdr = 12 + xlen;
    if (b->csize < b->hdr + 8 || coff + b->csize > size_) fail(BSG_EFORMAT, "truncated BGZF block in " + path_);
    b->crc = rd_u32(h + b->csize - 8);
    b->isize = rd_u32(h + b->csize - 4);
    if (b->isize > 65536) fail(BSG_EFORMAT, "BGZF ISIZE larger than 64 KiB in " + path_);
    return true;
}

void BamFile::parse_header() {
    // Inflate blocks from the start until the header (magic, text, reference table) is complete.
    Inflater inf;
    std::vector<uint8_t> buf;
    std::vector<uint64_t> block_start_u;   // uncompressed offset where each consumed block starts
    std::vector<uint64_t> block_coff;
    uint64_t coff = 0;
    auto more = [&]() {
        BlockInfo b;
        if (!block_at(coff, &b)) fail(BSG_EFORMAT, "truncated BAM header in " + path_);
        block_start_u.push_back(buf.size());
        block_coff.push_back(coff);
        size_t old = buf.size();
        buf.resize(old + b.isize);
        inf.inflate_block(data_, b, buf.data() + old, true);
        coff += b.csize;
    };
    auto need = [&](size_t n) { while (buf.size() < n) more(); };
    need(12);
    if (memcmp(buf.data(), "BAM\1", 4) != 0) fail(BSG_EFORMAT, "not a BAM file: " + path_);
    int32_t l_text = rd_i32(buf.data() + 4);
    if (l_text < 0) fail(BSG_EFORMAT, "corrupt BAM header in " + path_);
    size_t p = 8 + size_t(l_text);
    need(p + 4);
    int32_t n_ref = rd_i32(buf.data() + p);
    p += 4;
    if (n_ref < 0) fail(BSG_EFORMAT, "corrupt BAM header in " + path_);
    for (int i = 0; i < n_ref; ++i) {
        need(p + 4);
        int32_t l_name = rd_i32(buf.data() + p);
        p += 4;
        if (l_name < 0) fail(BSG_EFORMAT, "corrupt BAM header in " + path_);
        need(p + l_name + 4);
        std::string nm(reinterpret_cast<const char*>(buf.data() + p), size_t(l_name));
        nm.resize(strlen(nm.c_str()));
        p += l_name;
        names_.push_back(nm);
        lens_.push_back(rd_i32(buf.data() + p));
        p += 4;
        name2id_.emplace(nm, i);   // first occurrence wins, as in htslib's name hash
    }
    // virtual offset of the first record
    size_t k = block_start_u.size() - 1;
    while (block_start_u[k] > p) --k;
    if (p == buf.size()) first_rec_ = coff << 16;   // header ends exactly at a block end
    else first_rec_ = block_coff[k] << 16 | uint64_t(p - block_start_u[k]);
}

void BamFile::load_index() {
    std::string cand[2] = {path_ + ".bai", ""};
    if (path_.size() > 4 && path_.compare(path_.size() - 4, 4, ".bam") == 0) cand[1] = path_.substr(0, path_.size() - 4) + ".bai";
    std::vector<uint8_t> d;
    bool found = false;
    for (auto& c : cand) {
        if (c.empty()) continue;
        FILE* fp = fopen(c.c_str(), "rb");
        if (!fp) continue;
        uint8_t tmp[1 << 16];
        size_t k;
        while ((k = fread(tmp, 1, sizeof tmp, fp)) > 0) d.insert(d.end(), tmp, tmp + k);
        struct stat st;
        if (fstat(fileno(fp), &st) == 0) {
            index_size_ = uint64_t(st.st_size);
            index_mtime_ns_ = uint64_t(st.st_mtim.tv_sec) * 1000000000ull + uint64_t(st.st_mtim.tv_nsec);
        }
        index_path_ = c;
        fclose(fp);
        found = true;
        break;
    }
    if (!found) fail(BSG_ENOINDEX, "BAM indexing file is not available for file " + path_);   // src/bamsignals.cpp:209
    size_t p = 0;
    auto need = [&](size_t n) { if (p + n > d.size()) fail(BSG_EFORMAT, "truncated BAI index for " + path_); };
    need(8);
    if (memcmp(d.data(), "BAI\1", 4) != 0) fail(BSG_EFORMAT, "bad BAI magic for " + path_);
    int32_t n_ref = rd_i32(d.data() + 4);
    p = 8;
    if (n_ref < 0) fail(BSG_EFORMAT, "corrupt BAI index for " + path_);
    refs_.resize(n_ref);
    entries_.push_back(first_rec_);
    for (int r = 0; r < n_ref; ++r) {
        need(4);
        int32_t n_bin = rd_i32(d.data() + p);
        p += 4;
        for (int b = 0; b < n_bin; ++b) {
            need(8);
            uint32_t bin = rd_u32(d.data() + p);
            int32_t n_chunk = rd_i32(d.data() + p + 4);
            p += 8;
            if (n_chunk < 0) fail(BSG_EFORMAT, "corrupt BAI index for " + path_);
            need(16ull * n_chunk);
            std::vector<VRange> cs(n_chunk);
            for (int c = 0; c < n_chunk; ++c) {
                cs[c].beg = rd_u64(d.data() + p);
                cs[c].end = rd_u64(d.data() + p + 8);
                p += 16;
            }
            if (bin == 37450) {   // pseudo-bin: chunk 0 = (ref_beg, ref_end) offsets, chunk 1 = read counts
                if (n_chunk >= 1 && cs[0].end > cs[0].beg) { entries_.push_back(cs[0].beg); entries_.push_back(cs[0].end); }
            } else {
                for (auto& c : cs) { entries_.push_back(c.beg); entries_.push_back(c.end); }
            }
            if (bin != 37450)
                for (auto& c : cs) {
                    if (c.end <= c.beg) continue;
                    if (refs_[r].ref_end <= refs_[r].ref_beg) { refs_[r].ref_beg = c.beg; refs_[r].ref_end = c.end; }
                    refs_[r].ref_beg = std::min(refs_[r].ref_beg, c.beg);
                    refs_[r].ref_end = std::max(refs_[r].ref_end, c.end);
                }
            if (bin >= 4681 && bin < 37449 + 1) {
                // Leaf bins go into a flat per-window array: the file is coordinate-sorted, so the chunks of a run of
                // consecutive leaf bins lie between the first bin's first chunk and the last bin's last chunk.
                const size_t w = bin - 4681;
                if (refs_[r].leaf.size() <= w) refs_[r].leaf.resize(w + 1, VRange{0, 0});
                VRange lr{~0ull, 0};
                for (auto& c : cs) { lr.beg = std::min(lr.beg, c.beg); lr.end = std::max(lr.end, c.end); }
                if (lr.end > lr.beg) refs_[r].leaf[w] = lr;
            } else {
                refs_[r].bins[bin] = std::move(cs);
            }
        }
        need(4);
        int32_t n_intv = rd_i32(d.data() + p);
        p += 4;
        if (n_intv < 0) fail(BSG_EFORMAT, "corrupt BAI index for " + path_);
        need(8ull * n_intv);
        refs_[r].linear.resize(n_intv);
        for (int i = 0; i < n_intv; ++i) {
            refs_[r].linear[i] = rd_u64(d.data() + p + 8ull * i);
            if (refs_[r].linear[i]) entries_.push_back(refs_[r].linear[i]);
        }
        p += 8ull * n_intv;
    }
    std::sort(entries_.begin(), entries_.end());
    entries_.erase(std::unique(entries_.begin(), entries_.end()), entries_.end());
    // Normalise "end of block" offsets: (coff, usize) and (coff_next, 0) denote the same position; both forms stay
    // in the list (they differ numerically) which is harmless: the planner resolves positions through block tables.
}

int64_t BamFile::bp_per_block(int tid) const {
    if (tid < 0 || tid >= int(refs_.size()) || tid >= int(lens_.size())) return 16384;
    auto meta = refs_[tid].bins.find(37450);
    uint64_t beg = 0, end = 0;
    if (meta != refs_[tid].bins.end() && !meta->second.empty()) { beg = meta->second[0].beg >> 16; end = meta->second[0].end >> 16; }
    else {
        const auto& lin = refs_[tid].linear;
        for (uint64_t v : lin) if (v) { beg = v >> 16; break; }
        for (size_t k = lin.size(); k-- > 0;) if (lin[k]) { end = lin[k] >> 16; break; }
    }
    if (end <= beg) return 16384;
    const double cbytes_per_bp = double(end - beg) / double(std::max<int32_t>(1, lens_[tid]));
    const double bp = 65536.0 / std::max(1e-9, cbytes_per_bp);
    return int64_t(std::min(16.0e6, std::max(16384.0, bp)));
}

uint64_t BamFile::approx_coffset(int tid, int64_t pos) const {
    // offsets grow with (tid, pos) in a coordinate-sorted file; references without reads inherit the next one's start
    for (int t = std::max(tid, 0); t < int(refs_.size()); ++t) {
        const auto& lin = refs_[t].linear;
        size_t w = t == tid ? size_t(std::max<int64_t>(0, pos) >> 14) : 0;
        for (; w < lin.size(); ++w) if (lin[w]) return lin[w] >> 16;
    }
    return size_;
}

bool BamFile::index_unchanged() const {
    struct stat st;
    if (::stat(index_path_.c_str(), &st) != 0) return false;
    return uint64_t(st.st_size) == index_size_ &&
           uint64_t(st.st_mtim.tv_sec) * 1000000000ull + uint64_t(st.st_mtim.tv_nsec) == index_mtime_ns_;
}

void BamFile::query(int tid, int64_t beg, int64_t end, std::vector<VRange>* out) const {
    if (tid < 0 || tid >= int(refs_.size())) return;
    const RefIndex& ri = refs_[tid];
    if (beg < 0) beg = 0;
    if (end <= beg || ri.ref_end <= ri.ref_beg) return;
    // Lower bound L: the linear index gives the first record overlapping the 16 kb window of `beg`; every record
    // overlapping [beg, ...) lies at or behind it (htslib uses the same bound as `min_off`).
    uint64_t lo = ri.ref_beg;
    if (!ri.linear.empty()) {
        const size_t w = size_t(beg >> 14);
        uint64_t v = w < ri.linear.size() ? ri.linear[w] : ri.linear.back();
        // some indexers leave zero entries for windows no read overlaps; walk back to the last filled one
        if (v == 0 && w < ri.linear.size())
            for (size_t k = w; k-- > 0;) if (ri.linear[k]) { v = ri.linear[k]; break; }
        lo = std::max(lo, v);
    }
    // Upper bound U: the first record that lies entirely inside a 16 kb window behind the query (first chunk of the
    // first non-empty leaf bin > window(end-1)) starts at pos >= end, and the file is coordinate-sorted, so nothing at
    // or behind it can overlap [beg, end).  This plays the role of the iterator's "pos >= end -> stop" rule.
    uint64_t hi = ri.ref_end;
    for (size_t w = size_t((end - 1) >> 14) + 1; w < ri.leaf.size(); ++w)
        if (ri.leaf[w].end) { hi = std::min(hi, ri.leaf[w].beg); break; }
    if (hi > lo) out->push_back(VRange{lo, hi});
}

}  // namespace bsg
