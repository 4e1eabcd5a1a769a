// hts_compat — the htslib entry points the reference's src/bamsignals.cpp calls, implemented over this repo's own
// BGZF/BAM/BAI/CSI reader (../bam_reader.hpp, zlib).  TEST INFRASTRUCTURE ONLY (oracle/_ref): it exists so that the
// reference's engine compiles and runs UNCHANGED here; it is never part of the product library.
//
// htslib (Rhtslib >= 1.13.1, DESCRIPTION:30) is an un-vendored dependency of the reference and absent offline.  What
// is restated, from the SAM/BAM specification and htslib >= 1.10's documented behaviour:
//   * bam_read1: block_size + 32-byte core + variable part -> bam1_t
//   * bam_endpos: pos + (unmapped ? 0 : sum of M/D/N/=/X op lengths), 0 -> 1
//   * hts_idx_load: <bam>.csi, <bam - .bam>.csi, <bam>.bai, <bam - .bam>.bai; BAI or CSIv1 by magic
//   * hts_itr_query: beg clamped at 0; reg2bins; min_off from the linear index (BAI) or the bins' loffset (CSI);
//     chunks with end <= min_off dropped, sorted, overlapping ones merged
//   * hts_itr_next: a chunk is left when the virtual offset reaches its end; iteration stops at the first record with
//     another tid or pos >= end; a record is returned iff endpos > beg and pos < end
//   * bgzf cache: blocks of the last `size / BGZF_MAX_BLOCK_SIZE` distinct addresses are kept inflated
// The fixture writer's entry points (sam_hdr_write, sam_read1, bam_write1, sam_index_build; reference :496-534) are
// declared so that the file links, and report failure: writeSamAsBamAndIndex is not on the counting path
// (its equivalent, bsg_write_sam_as_bam_and_index, is tested against the reference's fixture in tests/test_writer.py).
#include <atomic>

#include "../bam_reader.hpp"
#include "htslib/sam.h"

using bsgo::BgzfIn;

struct BGZF {
    BgzfIn in;
    uint64_t records = 0;                   // handed to the global counters in hts_close (no shared line in the hot loop)
    explicit BGZF(const std::string& path) : in(path, 1) {}
};

namespace {

struct BinEnt { uint32_t bin, n_chunk; uint64_t loff; const uint8_t* chunks; };
struct RefIdx { std::vector<BinEnt> bins; const uint8_t* linear = nullptr; int32_t n_intv = 0; };

std::atomic<uint64_t> g_records{0}, g_queries{0}, g_bytes{0};

inline uint64_t rd_u64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }

}  // namespace

struct hts_idx_t {
    std::vector<uint8_t> raw;               // the index file (BGZF-inflated if it was a .csi); RefIdx points into it
    std::vector<RefIdx> refs;
    int min_shift = 14, depth = 5;
    bool csi = false;
};

struct hts_itr_t {
    int tid = -1;
    int64_t beg = 0, end = 0;
    std::vector<bsgo::Chunk> off;
    int i = -1;
    uint64_t curr_off = 0;
    bool finished = false;
};

extern "C" {

void bgzf_set_cache_size(BGZF* fp, int size) {
    if (fp && size > 0) fp->in.set_cache_blocks(size / BGZF_MAX_BLOCK_SIZE);
}

samFile* hts_open(const char* fn, const char* mode) {
    if (!fn || !mode || strcmp(mode, "rb") != 0) return nullptr;        // only BAM reading is provided
    try {
        BGZF* b = new BGZF(fn);
        htsFile* f = new htsFile();
        memset(f, 0, sizeof *f);
        f->is_bin = 1; f->is_bgzf = 1;
        f->fn = strdup(fn);
        f->fp.bgzf = b;
        return f;
    } catch (bsgo::OracleError&) {
        return nullptr;
    }
}

int hts_close(htsFile* fp) {
    if (!fp) return 0;
    if (fp->fp.bgzf) { g_bytes += fp->fp.bgzf->in.bytes_inflated; g_records += fp->fp.bgzf->records; }
    delete fp->fp.bgzf;
    free(fp->fn);
    delete fp;
    return 0;
}

sam_hdr_t* sam_hdr_read(samFile* fp) {
    if (!fp || !fp->fp.bgzf) return nullptr;
    try {
        bsgo::Header h = bsgo::read_header(fp->fp.bgzf->in);            // leaves the stream at the first record
        sam_hdr_t* s = static_cast<sam_hdr_t*>(calloc(1, sizeof(sam_hdr_t)));
        s->n_targets = int32_t(h.names.size());
        s->target_len = static_cast<uint32_t*>(calloc(h.names.size() + 1, sizeof(uint32_t)));
        s->target_name = static_cast<char**>(calloc(h.names.size() + 1, sizeof(char*)));
        for (size_t i = 0; i < h.names.size(); ++i) { s->target_len[i] = uint32_t(h.lens[i]); s->target_name[i] = strdup(h.names[i].c_str()); }
        return s;
    } catch (bsgo::OracleError&) {
        return nullptr;
    }
}

void sam_hdr_destroy(sam_hdr_t* h) {
    if (!h) return;
    for (int32_t i = 0; i < h->n_targets; ++i) free(h->target_name[i]);
    free(h->target_name); free(h->target_len); free(h->text);
    free(h);
}

int sam_hdr_name2tid(sam_hdr_t* h, const char* ref) {
    if (!h || !ref) return -2;
    for (int32_t i = 0; i < h->n_targets; ++i)
        if (strcmp(h->target_name[i], ref) == 0) return i;
    return -1;
}

bam1_t* bam_init1(void) { return static_cast<bam1_t*>(calloc(1, sizeof(bam1_t))); }
void bam_destroy1(bam1_t* b) { if (b) { free(b->data); free(b); } }

hts_pos_t bam_endpos(const bam1_t* b) {
    hts_pos_t rlen = 0;
    if (!(b->core.flag & BAM_FUNMAP)) {
        const uint8_t* cg = b->data + b->core.l_qname;
        for (uint32_t k = 0; k < b->core.n_cigar; ++k) {
            uint32_t c; memcpy(&c, cg + 4 * k, 4);
            if (bam_cigar_type(bam_cigar_op(c)) & 2) rlen += bam_cigar_oplen(c);
        }
    }
    if (rlen == 0) rlen = 1;
    return b->core.pos + rlen;
}

hts_idx_t* hts_idx_load(const char* fn, int /*fmt*/) {
    if (!fn) return nullptr;
    const std::string bam(fn);
    std::vector<std::string> cand = {bam + ".csi"};
    const bool ext = bam.size() > 4 && bam.compare(bam.size() - 4, 4, ".bam") == 0;
    if (ext) cand.push_back(bam.substr(0, bam.size() - 4) + ".csi");
    cand.push_back(bam + ".bai");
    if (ext) cand.push_back(bam.substr(0, bam.size() - 4) + ".bai");
    FILE* fp = nullptr;
    for (auto& c : cand) { fp = fopen(c.c_str(), "rb"); if (fp) break; }
    if (!fp) return nullptr;
    hts_idx_t* idx = new hts_idx_t();
    try {
        idx->raw = bsgo::read_maybe_bgzf(fp);
        fclose(fp); fp = nullptr;
        const std::vector<uint8_t>& d = idx->raw;
        size_t p = 0;
        auto need = [&](size_t n) { if (p + n > d.size()) bsgo::fail("truncated BAM index"); };
        need(8);
        if (memcmp(d.data(), "CSI\1", 4) == 0) {
            idx->csi = true;
            need(16);
            idx->min_shift = bsgo::rd_i32(d.data() + 4); idx->depth = bsgo::rd_i32(d.data() + 8);
            const int32_t l_aux = bsgo::rd_i32(d.data() + 12);
            if (l_aux < 0) bsgo::fail("bad CSI header");
            p = 16; need(size_t(l_aux) + 4); p += size_t(l_aux);
            if (idx->min_shift < 1 || idx->depth < 1 || idx->min_shift + 3 * idx->depth > 40) bsgo::fail("unsupported CSI geometry");
        } else {
            if (memcmp(d.data(), "BAI\1", 4) != 0) bsgo::fail("bad BAI magic");
            p = 4;
        }
        const int32_t n_ref = bsgo::rd_i32(d.data() + p); p += 4;
        if (n_ref < 0 || size_t(n_ref) > d.size()) bsgo::fail("bad index n_ref");
        idx->refs.resize(size_t(n_ref));
        for (int r = 0; r < n_ref; ++r) {
            RefIdx& ri = idx->refs[size_t(r)];
            need(4); const int32_t n_bin = bsgo::rd_i32(d.data() + p); p += 4;
            if (n_bin < 0 || size_t(n_bin) > d.size()) bsgo::fail("bad index n_bin");
            ri.bins.reserve(size_t(n_bin));
            for (int b = 0; b < n_bin; ++b) {
                need(idx->csi ? 16 : 8);
                BinEnt e; e.bin = bsgo::rd_u32(d.data() + p); p += 4; e.loff = 0;
                if (idx->csi) { e.loff = rd_u64(d.data() + p); p += 8; }
                const int32_t n_chunk = bsgo::rd_i32(d.data() + p); p += 4;
                if (n_chunk < 0) bsgo::fail("bad index n_chunk");
                need(16ull * size_t(n_chunk));
                e.n_chunk = uint32_t(n_chunk); e.chunks = d.data() + p;
                p += 16ull * size_t(n_chunk);
                ri.bins.push_back(e);
            }
            std::sort(ri.bins.begin(), ri.bins.end(), [](const BinEnt& a, const BinEnt& b) { return a.bin < b.bin; });
            if (idx->csi) continue;
            need(4); const int32_t n_intv = bsgo::rd_i32(d.data() + p); p += 4;
            if (n_intv < 0) bsgo::fail("bad index n_intv");
            need(8ull * size_t(n_intv));
            ri.linear = d.data() + p; ri.n_intv = n_intv;
            p += 8ull * size_t(n_intv);
        }
        return idx;
    } catch (bsgo::OracleError&) {
        if (fp) fclose(fp);
        delete idx;
        return nullptr;
    }
}

hts_idx_t* sam_index_load(htsFile* /*fp*/, const char* fn) { return hts_idx_load(fn, HTS_FMT_BAI); }
void hts_idx_destroy(hts_idx_t* idx) { delete idx; }
void hts_itr_destroy(hts_itr_t* iter) { delete iter; }

hts_itr_t* sam_itr_queryi(const hts_idx_t* idx, int tid, hts_pos_t beg, hts_pos_t end) {
    if (!idx) return nullptr;
    if (beg < 0) beg = 0;
    if (end < beg) return nullptr;
    hts_itr_t* it = new hts_itr_t();
    it->tid = tid; it->beg = beg; it->end = end;
    ++g_queries;
    if (tid < 0 || size_t(tid) >= idx->refs.size() || idx->refs[size_t(tid)].bins.empty() || end == beg) { it->finished = true; return it; }
    const RefIdx& ri = idx->refs[size_t(tid)];
    auto find_bin = [&](uint32_t b) -> const BinEnt* {
        auto p = std::lower_bound(ri.bins.begin(), ri.bins.end(), b, [](const BinEnt& e, uint32_t v) { return e.bin < v; });
        return (p != ri.bins.end() && p->bin == b) ? &*p : nullptr;
    };
    uint64_t min_off = 0;
    if (!idx->csi) {
        if (ri.n_intv > 0) {
            const int64_t w = beg >> 14;
            min_off = rd_u64(ri.linear + 8 * size_t(w < ri.n_intv ? w : ri.n_intv - 1));
        }
    } else {
        uint32_t bin = uint32_t(((1ull << (3 * idx->depth)) - 1) / 7 + (uint64_t(beg) >> idx->min_shift));
        for (;;) {
            if (const BinEnt* e = find_bin(bin)) { min_off = e->loff; break; }
            if (bin == 0) break;
            const uint32_t parent = (bin - 1) >> 3, first = (parent << 3) + 1;
            bin = bin > first ? bin - 1 : parent;
        }
    }
    std::vector<uint32_t> bins;
    bsgo::reg2bins(beg, end, idx->min_shift, idx->depth, bins);
    std::vector<bsgo::Chunk> res;
    for (uint32_t b : bins) {
        const BinEnt* e = find_bin(b);
        if (!e) continue;
        for (uint32_t c = 0; c < e->n_chunk; ++c) {
            bsgo::Chunk ck{rd_u64(e->chunks + 16 * size_t(c)), rd_u64(e->chunks + 16 * size_t(c) + 8)};
            if (ck.end > min_off) res.push_back(ck);
        }
    }
    std::sort(res.begin(), res.end(), [](const bsgo::Chunk& a, const bsgo::Chunk& b) { return a.beg < b.beg; });
    for (const bsgo::Chunk& c : res) {
        if (!it->off.empty() && c.beg <= it->off.back().end) it->off.back().end = std::max(it->off.back().end, c.end);
        else it->off.push_back(c);
    }
    if (it->off.empty()) it->finished = true;
    return it;
}

static int read1(BgzfIn& in, bam1_t* b) {
    uint8_t x[36];
    if (!in.read(x, 4)) return -1;
    const int32_t bs = bsgo::rd_i32(x);
    if (bs < 32) bsgo::fail("corrupt BAM record (block_size < 32)");
    if (!in.read(x + 4, 32)) bsgo::fail("truncated BAM record");
    bam1_core_t& c = b->core;
    c.tid = bsgo::rd_i32(x + 4); c.pos = bsgo::rd_i32(x + 8);
    c.l_qname = x[12]; c.qual = x[13]; c.bin = bsgo::rd_u16(x + 14);
    c.n_cigar = bsgo::rd_u16(x + 16); c.flag = bsgo::rd_u16(x + 18);
    c.l_qseq = bsgo::rd_i32(x + 20); c.mtid = bsgo::rd_i32(x + 24); c.mpos = bsgo::rd_i32(x + 28); c.isize = bsgo::rd_i32(x + 32);
    c.l_extranul = 0;
    const uint32_t rest = uint32_t(bs) - 32u;
    if (b->m_data < rest) {
        uint32_t m = rest; m--; m |= m >> 1; m |= m >> 2; m |= m >> 4; m |= m >> 8; m |= m >> 16; m++;     // kroundup32
        b->data = static_cast<uint8_t*>(realloc(b->data, m));
        b->m_data = m;
    }
    b->l_data = int(rest);
    if (rest && !in.read(b->data, rest)) bsgo::fail("truncated BAM record");
    if (uint64_t(c.l_qname) + 4ull * c.n_cigar > rest) bsgo::fail("corrupt BAM record (cigar beyond record)");
    return int(bs) + 4;
}

int sam_itr_next(htsFile* htsfp, hts_itr_t* iter, bam1_t* r) {
    if (!htsfp || !htsfp->fp.bgzf || !iter) return -2;
    if (iter->finished) return -1;
    BgzfIn& in = htsfp->fp.bgzf->in;
    int ret = -1;
    try {
        for (;;) {
            if (iter->curr_off == 0 || iter->curr_off >= iter->off[size_t(iter->i)].end) {      // leave this chunk
                if (iter->i == int(iter->off.size()) - 1) { ret = -1; break; }
                if (iter->i < 0 || iter->off[size_t(iter->i)].end != iter->off[size_t(iter->i) + 1].beg) {
                    in.seek(iter->off[size_t(iter->i) + 1].beg);
                    iter->curr_off = in.tell();
                }
                ++iter->i;
            }
            ret = read1(in, r);
            if (ret < 0) break;                                                                      // end of file
            ++htsfp->fp.bgzf->records;
            iter->curr_off = in.tell();
            const hts_pos_t rb = r->core.pos, re = bam_endpos(r);
            if (r->core.tid != iter->tid || rb >= iter->end) { ret = -1; break; }
            if (re > iter->beg && iter->end > rb) return ret;
        }
    } catch (bsgo::OracleError&) {
        ret = -2;
    }
    iter->finished = true;
    return ret;
}

// ---- fixture-writer entry points: declared for linkage, not provided ----------------------------------------------------
int sam_hdr_write(samFile*, const sam_hdr_t*) { return -1; }
int sam_read1(samFile*, sam_hdr_t*, bam1_t*) { return -1; }
int bam_write1(BGZF*, const bam1_t*) { return -1; }
int sam_index_build(const char*, int) { return -1; }

// counters for the benchmark's bookkeeping (records the iterators handed out, index queries, bytes inflated)
void hts_compat_stats(uint64_t* records, uint64_t* queries, uint64_t* bytes, int reset) {
    if (records) *records = g_records.load();
    if (queries) *queries = g_queries.load();
    if (bytes) *bytes = g_bytes.load();
    if (reset) { g_records = 0; g_queries = 0; g_bytes = 0; }
}

}  // extern "C"
