// Host-side BAM / BGZF / BAI + CSI access for the fetch pipeline (no htslib offline: SURVEY.md App. B).
// Replaces what the reference gets from htslib through `Bamfile` (src/bamsignals.cpp:195-220): open + index load,
// header lookup (bam_name2id, :27) and the region -> virtual-offset query behind bam_itr_queryi (:267).
#pragma once
#include <cstdint>
#include <string>
#include <unordered_map>
#include <vector>

#include "common.h"

namespace bsg {

struct VRange {   // half-open range of BGZF virtual offsets (coffset << 16 | uoffset), record-aligned at both ends
    uint64_t beg, end;
};

struct BlockInfo {   // one BGZF block as found in the file
    uint64_t coff;   // file offset of the block
    uint32_t csize;  // total block bytes (BSIZE + 1)
    uint32_t hdr;    // bytes before the raw DEFLATE stream (12 + XLEN)
    uint32_t isize;  // uncompressed bytes
    uint32_t crc;    // CRC32 of the uncompressed bytes
};

class BamFile {
public:
    explicit BamFile(const std::string& path);
    ~BamFile();
    BamFile(const BamFile&) = delete;
    BamFile& operator=(const BamFile&) = delete;

    const std::string& path() const { return path_; }
    const uint8_t* data() const { return data_; }
    uint64_t size() const { return size_; }
    const std::vector<std::string>& ref_names() const { return names_; }
    const std::vector<int32_t>& ref_lens() const { return lens_; }
    int name2id(const std::string& name) const;   // -1 if absent
    uint64_t first_record_voff() const { return first_rec_; }

    // Parse the block header at `coff`; returns false at EOF (coff == size), throws BSG_EFORMAT on garbage.
    bool block_at(uint64_t coff, BlockInfo* b) const;

    // One record-aligned virtual-offset range that contains every record of `tid` overlapping [beg,end): from the
    // linear-index lower bound to the first record known to start behind `end`.  A superset is all the counting path
    // needs; it is at most a 16 kb window wider than what htslib's iterator reads.
    void query(int tid, int64_t beg, int64_t end, std::vector<VRange>* out) const;

    // Genomic distance (bp) on `tid` that one maximal BGZF block (64 KiB compressed) covers on average, from the
    // index's (ref_beg, ref_end) file offsets: index queries closer than this fetch the same blocks anyway.
    int64_t bp_per_block(int tid) const;

    // Approximate compressed file offset of the first record at or after (tid, pos), from the linear index
    // (used only to balance region shards across devices).
    uint64_t approx_coffset(int tid, int64_t pos) const;

    // Sorted, de-duplicated record-aligned virtual offsets known to the index (linear index entries, chunk
    // begins/ends, first record): independent entry points for parallel inflate + record walking.
    const std::vector<uint64_t>& entry_points() const { return entries_; }

    // Identity of the file contents for the resident-table cache.
    uint64_t mtime_ns() const { return mtime_ns_; }
    bool index_unchanged() const;   // the index on disk still has the size and mtime it had when it was parsed

private:
    struct RefIndex {
        std::unordered_map<uint32_t, std::vector<VRange>> bins;   // levels 0-4 and the pseudo-bin
        std::vector<VRange> leaf;                                   // level 5 (16 kb bins): [first chunk beg, last chunk end) per window
        std::vector<uint64_t> linear;
        uint64_t ref_beg = 0, ref_end = 0;                          // first chunk begin / last chunk end of the reference
    };
    void parse_header();
    void load_index();

    std::string path_;
    int fd_ = -1;
    const uint8_t* data_ = nullptr;
    uint64_t size_ = 0;
    uint64_t mtime_ns_ = 0;
    std::string index_path_;
    uint64_t index_size_ = 0, index_mtime_ns_ = 0;
    std::vector<std::string> names_;
    std::vector<int32_t> lens_;
    std::unordered_map<std::string, int> name2id_;
    uint64_t first_rec_ = 0;
    std::vector<RefIndex> refs_;
    std::vector<uint64_t> entries_;
    int min_shift_ = 14, depth_ = 5;     // index geometry: BAI is fixed at 14 / 5, a CSI header states its own
    uint32_t meta_bin_ = 37450;
};

// Raw-DEFLATE inflate of one BGZF block into dst[0..isize); zlib state is per worker thread.
class Inflater {
public:
    Inflater();
    ~Inflater();
    Inflater(const Inflater&) = delete;
    void inflate_block(const uint8_t* file, const BlockInfo& b, uint8_t* dst, bool verify_crc);

private:
    void* zs_;
};

}  // namespace bsg
