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
    // A PRIVATE mapping with write permission that nothing ever writes to: the CUDA driver refuses to page-lock a
    // PROT_READ file mapping (cudaHostRegister: invalid argument, also with cudaHostRegisterReadOnly; probed on the B200
    // boxes, profiles/r2d_pin_probe_*.txt), and page-locked windows of the mapping are what lets the copy engine read the
    // compressed bytes straight from the page cache (engine.cu: ensure_pinned).
    void* p = mmap(nullptr, size_, PROT_READ | PROT_WRITE, MAP_PRIVATE, fd_, 0);
    if (p == MAP_FAILED) { ::close(fd_); fd_ = -1; fail(BSG_EOPEN, "Fail to open BAM file " + path); }
    data_ = static_cast<const uint8_t*>(p);
    try {
        // the reference would read CRAM through htslib (given a reference genome); this library reads BAM only
        if (size_ >= 4 && memcmp(data_, "CRAM", 4) == 0) fail(BSG_EFORMAT, "CRAM input is not supported (convert it to BAM): " + path);
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

namespace {

// Whole index file into memory; a BGZF-compressed one (.csi files are) is inflated block by block.
std::vector<uint8_t> read_index_file(FILE* fp, const std::string& what) {
    std::vector<uint8_t> d;
    uint8_t tmp[1 << 16];
    size_t k;
    while ((k = fread(tmp, 1, sizeof tmp, fp)) > 0) d.insert(d.end(), tmp, tmp + k);
    if (d.size() < 4 || !(d[0] == 0x1f && d[1] == 0x8b)) return d;
    std::vector<uint8_t> out;
    z_stream zs;
    memset(&zs, 0, sizeof zs);
    if (inflateInit2(&zs, -15) != Z_OK) fail(BSG_ENOMEM, "zlib inflateInit2 failed");
    size_t p = 0;
    bool ok = true;
    while (ok && p + 18 <= d.size()) {
        const uint8_t* h = d.data() + p;
        if (h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4)) { ok = false; break; }
        const uint32_t xlen = rd_u16(h + 10);
        int bsize = -1;
        for (uint32_t q = 0; q + 4 <= xlen && p + 12 + q + 4 <= d.size();) {
            const uint8_t* e = h + 12 + q;
            const uint32_t slen = rd_u16(e + 2);
            if (e[0] == 'B' && e[1] == 'C' && slen == 2 && q + 6 <= xlen) bsize = rd_u16(e + 4);
            q += 4 + slen;
        }
        if (bsize < 0 || uint32_t(bsize) + 1 < 12 + xlen + 8 || p + uint32_t(bsize) + 1 > d.size()) { ok = false; break; }
        const uint32_t csize = uint32_t(bsize) + 1, isize = rd_u32(h + csize - 4);
        if (isize > 65536) { ok = false; break; }
        const size_t old = out.size();
        out.resize(old + isize);
        if (isize) {
            inflateReset(&zs);
            zs.next_in = const_cast<Bytef*>(h + 12 + xlen);
            zs.avail_in = csize - 12 - xlen - 8;
            zs.next_out = out.data() + old;
            zs.avail_out = isize;
            const int rc = inflate(&zs, Z_FINISH);
            if (!(rc == Z_STREAM_END && zs.avail_out == 0) ||
                uint32_t(crc32(crc32(0L, Z_NULL, 0), out.data() + old, isize)) != rd_u32(h + csize - 8)) { ok = false; break; }
        }
        p += csize;
    }
    inflateEnd(&zs);
    if (!ok || p != d.size()) fail(BSG_EFORMAT, "corrupt BGZF-compressed index for " + what);
    return out;
}

// Ascending sort of virtual offsets.  A whole-genome index holds a few million of them (chunk bounds of every bin plus the
// linear index) and std::sort spent 115 ms of a 185 ms first open on them; an LSD radix sort over the 16-bit digits that
// are actually in use (three for files below 4 GiB) takes about a tenth of that.
void sort_u64(std::vector<uint64_t>& v) {
    if (v.size() < (size_t(1) << 14)) { std::sort(v.begin(), v.end()); return; }
    uint64_t all = 0;
    for (uint64_t x : v) all |= x;
    std::vector<uint64_t> tmp(v.size());
    std::vector<size_t> cnt(size_t(1) << 16);
    uint64_t* a = v.data();
    uint64_t* b = tmp.data();
    for (int sh = 0; sh < 64; sh += 16) {
        if ((all >> sh) == 0) break;                       // no key has a bit at or above this digit
        if (((all >> sh) & 0xffffu) == 0) continue;        // every key has a zero digit here: the pass would change nothing
        std::fill(cnt.begin(), cnt.end(), size_t(0));
        for (size_t i = 0; i < v.size(); ++i) ++cnt[(a[i] >> sh) & 0xffffu];
        size_t run = 0;
        for (size_t k = 0; k < cnt.size(); ++k) { const size_t c = cnt[k]; cnt[k] = run; run += c; }
        for (size_t i = 0; i < v.size(); ++i) b[cnt[(a[i] >> sh) & 0xffffu]++] = a[i];
        std::swap(a, b);
    }
    if (a != v.data()) memcpy(v.data(), a, v.size() * sizeof(uint64_t));
}

}  // namespace

// Index discovery follows htslib's order for a BAM file: <path>.csi, <path minus .bam>.csi, <path>.bai,
// <path minus .bam>.bai; the format is decided by the magic, not by the file name.  Both formats end up in the same
// RefIndex: a CSI index (SAM spec "CSIv1": min_shift / depth in the header, a per-bin `loffset` instead of the linear
// index) has its linear index rebuilt from the loffsets - loffset(bin) is the linear-index entry of the bin's first
// window (htslib's update_loff), i.e. the offset of the first record overlapping that window.
void BamFile::load_index() {
    std::vector<std::string> cand = {path_ + ".csi"};
    const bool has_ext = path_.size() > 4 && path_.compare(path_.size() - 4, 4, ".bam") == 0;
    if (has_ext) cand.push_back(path_.substr(0, path_.size() - 4) + ".csi");
    cand.push_back(path_ + ".bai");
    if (has_ext) cand.push_back(path_.substr(0, path_.size() - 4) + ".bai");
    std::vector<uint8_t> d;
    bool found = false;
    for (auto& c : cand) {
        FILE* fp = fopen(c.c_str(), "rb");
        if (!fp) continue;
        struct stat st;
        if (fstat(fileno(fp), &st) == 0) {
            index_size_ = uint64_t(st.st_size);
            index_mtime_ns_ = uint64_t(st.st_mtim.tv_sec) * 1000000000ull + uint64_t(st.st_mtim.tv_nsec);
        }
        index_path_ = c;
        try { d = read_index_file(fp, path_); } catch (...) { fclose(fp); throw; }
        fclose(fp);
        found = true;
        break;
    }
    if (!found) fail(BSG_ENOINDEX, "BAM indexing file is not available for file " + path_);   // src/bamsignals.cpp:209
    size_t p = 0;
    auto need = [&](size_t n) { if (p + n > d.size()) fail(BSG_EFORMAT, "truncated BAM index for " + path_); };
    need(8);
    const bool csi = memcmp(d.data(), "CSI\1", 4) == 0;
    if (!csi && memcmp(d.data(), "BAI\1", 4) != 0) fail(BSG_EFORMAT, "bad BAM index magic for " + path_);
    p = 4;
    if (csi) {
        need(12);
        min_shift_ = rd_i32(d.data() + p);
        depth_ = rd_i32(d.data() + p + 4);
        const int32_t l_aux = rd_i32(d.data() + p + 8);
        p += 12;
        if (min_shift_ < 1 || min_shift_ > 30 || depth_ < 1 || depth_ > 9 || min_shift_ + 3 * depth_ > 40 || l_aux < 0)
            fail(BSG_EFORMAT, "unsupported CSI index geometry for " + path_);
        need(size_t(l_aux));
        p += size_t(l_aux);
    }
    need(4);
    const int32_t n_ref = rd_i32(d.data() + p);
    p += 4;
    // every reference takes at least its n_bin field (and n_intv in a BAI): a count the file cannot hold is garbage,
    // and must not size an allocation
    if (n_ref < 0 || uint64_t(n_ref) * (csi ? 4 : 8) > d.size() - p) fail(BSG_EFORMAT, "corrupt BAM index for " + path_);
    const uint32_t first_leaf = uint32_t(((1ull << (3 * depth_)) - 1) / 7), n_leaf = 1u << (3 * depth_);
    meta_bin_ = uint32_t(((1ull << (3 * (depth_ + 1))) - 1) / 7) + 1;                      // 37450 for a BAI
    refs_.resize(n_ref);
    entries_.push_back(first_rec_);
    // Per-window arrays (CSI: rebuilt linear index; both: leaf bins) are sized by bin NUMBERS found in the file.  A
    // window far behind the end of its contig cannot matter to a query (dropping it only loosens the fetched range,
    // which stays a superset), so such entries are ignored instead of sizing an array; the total is bounded too.
    uint64_t windows_total = 0;
    const auto window_cap = [&](int r) {
        const uint64_t len = r < int(lens_.size()) && lens_[r] > 0 ? uint64_t(lens_[r]) : 0;
        return std::min<uint64_t>(uint64_t(1) << 24, (len >> min_shift_) + 4097);
    };
    const auto account = [&](uint64_t grown) {
        windows_total += grown;
        if (windows_total > (uint64_t(1) << 27)) fail(BSG_EFORMAT, "unsupported index geometry (too many windows) for " + path_);
    };
    for (int r = 0; r < n_ref; ++r) {
        RefIndex& ri = refs_[r];
        const uint64_t w_cap = window_cap(r);
        need(4);
        int32_t n_bin = rd_i32(d.data() + p);
        p += 4;
        for (int b = 0; b < n_bin; ++b) {
            need(csi ? 16 : 8);
            const uint32_t bin = rd_u32(d.data() + p);
            uint64_t loff = 0;
            if (csi) { loff = rd_u64(d.data() + p + 4); p += 8; }
            int32_t n_chunk = rd_i32(d.data() + p + 4);
            p += 8;
            if (n_chunk < 0) fail(BSG_EFORMAT, "corrupt BAM index for " + path_);
            need(16ull * n_chunk);
            std::vector<VRange> cs(n_chunk);
            for (int c = 0; c < n_chunk; ++c) {
                cs[c].beg = rd_u64(d.data() + p);
                cs[c].end = rd_u64(d.data() + p + 8);
                p += 16;
            }
            if (bin == meta_bin_) {   // pseudo-bin: chunk 0 = (ref_beg, ref_end) offsets, chunk 1 = read counts
                if (n_chunk >= 1 && cs[0].end > cs[0].beg) { entries_.push_back(cs[0].beg); entries_.push_back(cs[0].end); }
                ri.bins[bin] = std::move(cs);
                continue;
            }
            if (bin >= first_leaf + n_leaf) fail(BSG_EFORMAT, "bin number out of range in the index of " + path_);
            for (auto& c : cs) { entries_.push_back(c.beg); entries_.push_back(c.end); }
            for (auto& c : cs) {
                if (c.end <= c.beg) continue;
                if (ri.ref_end <= ri.ref_beg) { ri.ref_beg = c.beg; ri.ref_end = c.end; }
                ri.ref_beg = std::min(ri.ref_beg, c.beg);
                ri.ref_end = std::max(ri.ref_end, c.end);
            }
            if (csi && loff) {
                // first window of the bin (htslib's hts_bin_bot): level l has 8^l bins of 8^(depth-l) windows each
                int l = 0;
                uint32_t first = 0;
                while (l < depth_ && bin >= first + (1u << (3 * l))) { first += 1u << (3 * l); ++l; }
                const uint64_t w = uint64_t(bin - first) << (3 * (depth_ - l));
                if (w < w_cap) {
                    if (ri.linear.size() <= w) { account(w + 1 - ri.linear.size()); ri.linear.resize(w + 1, 0); }
                    ri.linear[w] = ri.linear[w] ? std::min(ri.linear[w], loff) : loff;
                }
                entries_.push_back(loff);
            }
            if (bin >= first_leaf) {
                // Leaf bins go into a flat per-window array: the file is coordinate-sorted, so the chunks of a run of
                // consecutive leaf bins lie between the first bin's first chunk and the last bin's last chunk.
                const size_t w = bin - first_leaf;
                if (w < w_cap) {
                    if (ri.leaf.size() <= w) { account(w + 1 - ri.leaf.size()); ri.leaf.resize(w + 1, VRange{0, 0}); }
                    VRange lr{~0ull, 0};
                    for (auto& c : cs) { lr.beg = std::min(lr.beg, c.beg); lr.end = std::max(lr.end, c.end); }
                    if (lr.end > lr.beg) ri.leaf[w] = lr;
                }
            } else {
                ri.bins[bin] = std::move(cs);
            }
        }
        if (csi) continue;
        need(4);
        int32_t n_intv = rd_i32(d.data() + p);
        p += 4;
        if (n_intv < 0) fail(BSG_EFORMAT, "corrupt BAM index for " + path_);
        need(8ull * n_intv);
        ri.linear.resize(n_intv);
        for (int i = 0; i < n_intv; ++i) {
            ri.linear[i] = rd_u64(d.data() + p + 8ull * i);
            if (ri.linear[i]) entries_.push_back(ri.linear[i]);
        }
        p += 8ull * n_intv;
    }
    sort_u64(entries_);
    entries_.erase(std::unique(entries_.begin(), entries_.end()), entries_.end());
    // Normalise "end of block" offsets: (coff, usize) and (coff_next, 0) denote the same position; both forms stay
    // in the list (they differ numerically) which is harmless: the planner resolves positions through block tables.
}

int64_t BamFile::bp_per_block(int tid) const {
    if (tid < 0 || tid >= int(refs_.size()) || tid >= int(lens_.size())) return 16384;
    auto meta = refs_[tid].bins.find(meta_bin_);
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
        size_t w = t == tid ? size_t(std::max<int64_t>(0, pos) >> min_shift_) : 0;
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
    // Lower bound L: the linear index gives the first record overlapping the window (16 kb for a BAI) of `beg`; every record
    // overlapping [beg, ...) lies at or behind it (htslib uses the same bound as `min_off`).
    uint64_t lo = ri.ref_beg;
    if (!ri.linear.empty()) {
        const size_t w = size_t(beg >> min_shift_);
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
    for (size_t w = size_t((end - 1) >> min_shift_) + 1; w < ri.leaf.size(); ++w)
        if (ri.leaf[w].end) { hi = std::min(hi, ri.leaf[w].beg); break; }
    if (hi > lo) out->push_back(VRange{lo, hi});
}

}  // namespace bsg
