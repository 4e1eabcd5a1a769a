// bamgen — deterministic synthetic coordinate-sorted BAM + BAI writer (BGZF / BAM / BAI written directly over zlib;
// htslib is not available offline).  Plays the role of the reference's test-only writeSamAsBamAndIndex
// (src/bamsignals.cpp:496-534) and produces the inputs of BASELINE.json's configs C2..C5.
//
// Reads are generated per 16 kb bucket from a counter-based RNG keyed by (seed, contig, bucket), so any 1 Mbp tile
// can be produced independently (and in parallel) and the file is identical for any thread count.
//
//   bamgen --out x.bam --preset c2|c3|c4|c5 [--scale 0.01] [--genome-scale 0.01] [--threads N] [--seed S]
//          [--record compact|realistic] [--level 1] [--straddle 0.05] [--unplaced N]
#include <algorithm>
#include <atomic>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include <zlib.h>

namespace {

const int64_t HG38[24] = {248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636,
                          138394717, 133797422, 135086622, 133275309, 114364328, 107043718, 101991189, 90338345,
                          83257441,  80373285,  58617616,  64444167,  46709983,  50818468,  156040895, 57227415};

struct Rng {
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed) {}
    uint64_t next() {
        uint64_t z = (s += 0x9e3779b97f4a7c15ull);
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
        return z ^ (z >> 31);
    }
    double uni() { return double(next() >> 11) * (1.0 / 9007199254740992.0); }
    int64_t range(int64_t lo, int64_t hi) { return lo + int64_t(next() % uint64_t(hi - lo + 1)); }   // inclusive
    double normal() {
        double u1 = uni(), u2 = uni();
        if (u1 < 1e-300) u1 = 1e-300;
        return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
    }
    double gamma(double shape) {   // Marsaglia-Tsang, shape >= 1
        const double d = shape - 1.0 / 3.0, c = 1.0 / std::sqrt(9.0 * d);
        for (;;) {
            double x = normal(), v = 1.0 + c * x;
            if (v <= 0) continue;
            v = v * v * v;
            const double u = uni();
            if (u < 1 - 0.0331 * x * x * x * x || std::log(u) < 0.5 * x * x + d * (1 - v + std::log(v))) return d * v;
        }
    }
    int64_t poisson(double lam) {
        if (lam <= 0) return 0;
        if (lam < 30) {
            const double L = std::exp(-lam);
            int64_t k = 0;
            double p = 1;
            do { ++k; p *= uni(); } while (p > L);
            return k - 1;
        }
        const int64_t v = int64_t(std::llround(lam + std::sqrt(lam) * normal()));
        return v < 0 ? 0 : v;
    }
};

uint64_t mix(uint64_t a, uint64_t b, uint64_t c) {
    Rng r(a * 0x9e3779b97f4a7c15ull ^ (b + 0x7f4a7c15ull) * 0xbf58476d1ce4e5b9ull ^ (c + 1) * 0x94d049bb133111ebull);
    r.next();
    return r.next();
}

struct Config {
    std::string out;
    std::vector<std::string> names;
    std::vector<int64_t> lens;
    bool paired = false;
    double units = 0;          // reads (SE) or pairs (PE) over the whole genome
    int readlen = 100;
    int tlen_kind = 0;         // 0 none, 1 negbin(149,10)+1, 2 normal(350,50) clipped [150,1000]
    double dup = 0.05, unmapped = 0.0, hotspot_frac = 0.2, straddle = 0.05;
    bool realistic = false;
    int level = 1, threads = 0;
    uint64_t seed = 20260101;
    int64_t unplaced = 0;
};

struct Rec {
    int32_t pos, end;          // end exclusive (bam_endpos)
    uint32_t off, len;         // bytes in the tile's record buffer (including block_size)
};

constexpr int64_t BUCKET = 16384, TILE = 64 * BUCKET;

int reg2bin(int64_t beg, int64_t end) {
    --end;
    if (beg >> 14 == end >> 14) return int(((1 << 15) - 1) / 7 + (beg >> 14));
    if (beg >> 17 == end >> 17) return int(((1 << 12) - 1) / 7 + (beg >> 17));
    if (beg >> 20 == end >> 20) return int(((1 << 9) - 1) / 7 + (beg >> 20));
    if (beg >> 23 == end >> 23) return int(((1 << 6) - 1) / 7 + (beg >> 23));
    if (beg >> 26 == end >> 26) return int(((1 << 3) - 1) / 7 + (beg >> 26));
    return 0;
}

void put32(std::vector<uint8_t>& v, uint32_t x) { for (int i = 0; i < 4; ++i) v.push_back(uint8_t(x >> (8 * i))); }
void put16(std::vector<uint8_t>& v, uint32_t x) { v.push_back(uint8_t(x)); v.push_back(uint8_t(x >> 8)); }

struct Chunk { uint64_t beg, end; };
struct TileOut {
    std::vector<uint8_t> comp;                       // BGZF blocks of this tile
    std::map<uint32_t, std::vector<Chunk>> bins;     // relative virtual offsets (coffset relative to tile start)
    std::vector<std::pair<int64_t, uint64_t>> linear;   // (window, rel voffset)
    uint64_t n_mapped = 0, n_unmapped = 0, n_records = 0, usize = 0;
    bool done = false;
};

void bgzf_block(const uint8_t* src, size_t n, int level, std::vector<uint8_t>& out) {
    uint8_t buf[65536 + 1024];
    z_stream zs;
    memset(&zs, 0, sizeof zs);
    deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY);
    zs.next_in = const_cast<Bytef*>(src); zs.avail_in = uInt(n);
    zs.next_out = buf; zs.avail_out = sizeof buf;
    int rc = deflate(&zs, Z_FINISH);
    size_t clen = zs.total_out;
    deflateEnd(&zs);
    if (rc != Z_STREAM_END || clen + 26 > 65536) {   // incompressible: stored block
        deflateInit2(&zs, 0, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY);
        zs.next_in = const_cast<Bytef*>(src); zs.avail_in = uInt(n);
        zs.next_out = buf; zs.avail_out = sizeof buf;
        deflate(&zs, Z_FINISH);
        clen = zs.total_out;
        deflateEnd(&zs);
    }
    const uint8_t hdr[12] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0};
    out.insert(out.end(), hdr, hdr + 12);
    out.push_back('B'); out.push_back('C'); put16(out, 2); put16(out, uint32_t(clen + 25));
    out.insert(out.end(), buf, buf + clen);
    put32(out, uint32_t(crc32(crc32(0, nullptr, 0), src, uInt(n))));
    put32(out, uint32_t(n));
}

struct Gen {
    const Config& cfg;
    double density;            // units per bp
    explicit Gen(const Config& c) : cfg(c) {
        double total = 0;
        for (auto l : c.lens) total += double(l);
        density = c.units / total;
    }

    int draw_tlen(Rng& r) const {
        if (cfg.tlen_kind == 1) return 1 + int(r.poisson(r.gamma(10.0) * 14.9));
        double v = 350 + 50 * r.normal();
        return int(std::min(1000.0, std::max(150.0, v)));
    }
    int draw_mapq(Rng& r) const {
        const double u = r.uni();
        if (u < 0.10) return 0;
        if (u < 0.25) return int(r.range(1, 29));
        return int(r.range(30, 60));
    }
    // returns reference length
    int draw_cigar(Rng& r, int L, std::vector<uint32_t>& cig) const {
        cig.clear();
        const double u = r.uni();
        auto op = [&](int len, int o) { cig.push_back(uint32_t(len) << 4 | uint32_t(o)); };
        if (u < 0.88 || L < 30) { op(L, 0); return L; }
        if (u < 0.94) { int a = int(r.range(1, 10)), c = int(r.range(1, 10)); op(a, 4); op(L - a - c, 0); op(c, 4); return L - a - c; }
        int x = int(r.range(5, L - 10));
        if (u < 0.97) { int d = int(r.range(1, 5)); op(x, 0); op(d, 2); op(L - x, 0); return L + d; }
        if (u < 0.99) { int i = int(r.range(1, 5)); op(x, 0); op(i, 1); op(L - x - i, 0); return L - i; }
        int n = int(r.range(100, 10000)); op(x, 0); op(n, 3); op(L - x, 0); return L + n;
    }

    void emit(std::vector<uint8_t>& buf, std::vector<Rec>& recs, int tid, int64_t pos, int mapq, int flag,
              const std::vector<uint32_t>& cig, int rlen, int l_seq, int next_tid, int64_t next_pos, int tlen,
              uint64_t name_id, Rng& r) const {
        char name[16];
        const int nl = snprintf(name, sizeof name, "r%09llu", (unsigned long long)(name_id % 1000000000ull)) + 1;
        const int ls = cfg.realistic ? l_seq : 0;
        const uint32_t bs = 32 + nl + 4 * uint32_t(cig.size()) + (ls + 1) / 2 + ls;
        const uint32_t off = uint32_t(buf.size());
        const int64_t end = pos + (rlen > 0 ? rlen : 1);
        put32(buf, bs); put32(buf, uint32_t(tid)); put32(buf, uint32_t(pos));
        buf.push_back(uint8_t(nl)); buf.push_back(uint8_t(mapq)); put16(buf, uint32_t(reg2bin(pos, end)));
        put16(buf, uint32_t(cig.size())); put16(buf, uint32_t(flag)); put32(buf, uint32_t(ls));
        put32(buf, uint32_t(next_tid)); put32(buf, uint32_t(next_pos)); put32(buf, uint32_t(tlen));
        buf.insert(buf.end(), name, name + nl);
        for (uint32_t c : cig) put32(buf, c);
        (void)r;
        Rng q(name_id * 31 + uint64_t(flag));   // sequence bytes must not consume the bucket stream (tiles regenerate buckets)
        for (int i = 0; i < (ls + 1) / 2; ++i) buf.push_back(uint8_t(q.next()));
        for (int i = 0; i < ls; ++i) buf.push_back(uint8_t(20 + q.next() % 20));
        recs.push_back(Rec{int32_t(pos), int32_t(end), off, bs + 4});
    }

    // all records of bucket `b` of contig `tid` whose pos lies in [lo, hi)
    void bucket(int tid, int64_t b, int64_t lo, int64_t hi, std::vector<uint8_t>& buf, std::vector<Rec>& recs) const {
        const int64_t clen = cfg.lens[tid];
        const int64_t b_lo = b * BUCKET, b_hi = std::min(clen, b_lo + BUCKET);
        if (b_lo >= clen) return;
        Rng r(mix(cfg.seed, uint64_t(tid), uint64_t(b)));
        const double lam = density * double(b_hi - b_lo);
        const double lam_bg = lam * (1 - cfg.hotspot_frac);
        int64_t n_bg = int64_t(lam_bg);
        if (r.uni() < lam_bg - double(n_bg)) ++n_bg;
        // one hot-spot per ~50 kb carrying hotspot_frac of the reads of 50 kb
        int64_t n_hot = 0, hot_c = 0;
        if (r.uni() < double(BUCKET) / 50000.0) {
            const double lh = density * 50000.0 * cfg.hotspot_frac;
            n_hot = int64_t(lh);
            if (r.uni() < lh - double(n_hot)) ++n_hot;
            hot_c = r.range(b_lo, b_hi - 1);
        }
        std::vector<uint32_t> cig1, cig2;
        for (int64_t i = 0; i < n_bg + n_hot; ++i) {
            int64_t f;
            if (i < n_bg) f = r.range(b_lo, b_hi - 1);
            else f = std::min(b_hi - 1, std::max(b_lo, hot_c + int64_t(std::llround(150.0 * r.normal()))));
            const uint64_t name_id = mix(cfg.seed ^ 0x5151, uint64_t(tid) << 40 | uint64_t(b), uint64_t(i));
            const bool dup = r.uni() < cfg.dup;
            const int dupf = dup ? 0x400 : 0;
            if (!cfg.paired) {
                const bool neg = r.next() & 1;
                const int rl = draw_cigar(r, cfg.readlen, cig1);
                const int mq = draw_mapq(r);
                if (f + rl > clen || f < lo || f >= hi) continue;
                emit(buf, recs, tid, f, mq, (neg ? 16 : 0) | dupf, cig1, rl, cfg.readlen, -1, -1, 0, name_id, r);
                continue;
            }
            const int T = draw_tlen(r);
            const bool first_left = r.next() & 1;            // is read1 the left (+) mate?
            const int rl1 = draw_cigar(r, cfg.readlen, cig1);   // left mate
            const int rl2 = draw_cigar(r, cfg.readlen, cig2);   // right mate
            const int mq1 = draw_mapq(r), mq2 = draw_mapq(r);
            const bool unm = r.uni() < cfg.unmapped;
            const int64_t p2 = std::max(f, f + T - cfg.readlen);
            if (std::max(f + rl1, p2 + rl2) > clen) continue;
            if (unm) {
                // right mate unmapped, placed at its mate's position (flag 0x4 with a coordinate)
                if (f >= lo && f < hi) {
                    emit(buf, recs, tid, f, mq1, 0x1 | 0x8 | (first_left ? 0x40 : 0x80) | dupf, cig1, rl1, cfg.readlen, tid, f, 0, name_id, r);
                    cig2.clear();
                    emit(buf, recs, tid, f, 0, 0x1 | 0x4 | (first_left ? 0x80 : 0x40), cig2, 0, cfg.readlen, tid, f, 0, name_id, r);
                }
                continue;
            }
            // flags as in the reference fixture (tests/testthat/utils.R:165): 99/147 or 163/83
            const int fl_left = 0x1 | 0x2 | 0x20 | (first_left ? 0x40 : 0x80) | dupf;
            const int fl_right = 0x1 | 0x2 | 0x10 | (first_left ? 0x80 : 0x40) | dupf;
            if (f >= lo && f < hi) emit(buf, recs, tid, f, mq1, fl_left, cig1, rl1, cfg.readlen, tid, p2, T, name_id, r);
            if (p2 >= lo && p2 < hi) emit(buf, recs, tid, p2, mq2, fl_right, cig2, rl2, cfg.readlen, tid, f, -T, name_id, r);
        }
    }

    void tile(int tid, int64_t t, TileOut& out) const {
        const int64_t lo = t * TILE, hi = std::min(cfg.lens[tid], lo + TILE);
        std::vector<uint8_t> buf;
        std::vector<Rec> recs;
        buf.reserve(size_t(density * double(hi - lo) * (cfg.paired ? 2 : 1) * (cfg.realistic ? 60 + 1.5 * cfg.readlen : 56)) + 4096);
        if (cfg.paired && t > 0) bucket(tid, lo / BUCKET - 1, lo, hi, buf, recs);   // right mates crossing in
        for (int64_t b = lo / BUCKET; b * BUCKET < hi; ++b) bucket(tid, b, lo, hi, buf, recs);
        std::stable_sort(recs.begin(), recs.end(), [](const Rec& a, const Rec& b) { return a.pos < b.pos; });
        // lay the records out in sorted order and cut into BGZF blocks
        std::vector<uint8_t> stream;
        stream.reserve(buf.size());
        std::vector<uint64_t> start(recs.size() + 1);
        for (size_t i = 0; i < recs.size(); ++i) {
            start[i] = stream.size();
            stream.insert(stream.end(), buf.begin() + recs[i].off, buf.begin() + recs[i].off + recs[i].len);
        }
        start[recs.size()] = stream.size();
        out.usize = stream.size();
        out.n_records = recs.size();
        // block boundaries (uncompressed): htslib keeps records whole; a fraction of blocks is cut mid-record
        Rng r(mix(cfg.seed ^ 0xb10c, uint64_t(tid), uint64_t(t)));
        std::vector<uint64_t> bstart{0};
        {
            const uint64_t cap = 0xff00;
            size_t i = 0;
            uint64_t cur = 0;
            while (cur < stream.size()) {
                uint64_t end;
                if (r.uni() < cfg.straddle) end = std::min<uint64_t>(stream.size(), cur + cap);
                else {
                    while (i < recs.size() && start[i] < cur) ++i;               // first record starting at/after cur
                    size_t j = i;
                    while (j < recs.size() && start[j + 1] - cur <= cap) ++j;     // whole records that fit
                    end = j > i ? start[j] : std::min<uint64_t>(stream.size(), cur + cap);   // oversize record: cut
                    if (j == i && i < recs.size() && start[i] > cur) end = std::min(start[i], cur + cap);  // straddler's tail first
                }
                if (end <= cur) end = std::min<uint64_t>(stream.size(), cur + cap);
                bstart.push_back(end);
                cur = end;
            }
        }
        std::vector<uint64_t> cstart(bstart.size());
        for (size_t k = 0; k + 1 < bstart.size(); ++k) {
            cstart[k] = out.comp.size();
            bgzf_block(stream.data() + bstart[k], size_t(bstart[k + 1] - bstart[k]), cfg.level, out.comp);
        }
        cstart[bstart.size() - 1] = out.comp.size();
        // index pieces with tile-relative virtual offsets
        auto voff = [&](uint64_t u) {
            size_t k = std::upper_bound(bstart.begin(), bstart.end(), u) - bstart.begin() - 1;
            if (k + 1 == bstart.size()) return cstart[k] << 16;                  // end of tile
            return cstart[k] << 16 | (u - bstart[k]);
        };
        int cur_bin = -1;
        uint64_t run_beg = 0, run_end = 0;
        int64_t last_win = -1;
        for (size_t i = 0; i < recs.size(); ++i) {
            const uint64_t vb = voff(start[i]), ve = voff(start[i + 1]);
            const int bin = reg2bin(recs[i].pos, recs[i].end);
            if (bin != cur_bin) {
                if (cur_bin >= 0) out.bins[uint32_t(cur_bin)].push_back(Chunk{run_beg, run_end});
                cur_bin = bin; run_beg = vb;
            }
            run_end = ve;
            for (int64_t w = recs[i].pos >> 14; w <= (int64_t(recs[i].end) - 1) >> 14; ++w)
                if (w > last_win) { out.linear.emplace_back(w, vb); last_win = w; }
            const uint8_t* p = stream.data() + start[i];
            const uint16_t flag = uint16_t(p[18] | p[19] << 8);
            if (flag & 4) ++out.n_unmapped; else ++out.n_mapped;
        }
        if (cur_bin >= 0) out.bins[uint32_t(cur_bin)].push_back(Chunk{run_beg, run_end});
    }
};

bool preset(Config& c, const std::string& p, double scale, double gscale) {
    auto genome = [&](int n) {
        for (int i = 0; i < n; ++i) {
            c.names.push_back(i < 22 ? "chr" + std::to_string(i + 1) : (i == 22 ? "chrX" : "chrY"));
            c.lens.push_back(std::max<int64_t>(20000, int64_t(double(HG38[i]) * gscale)));
        }
    };
    if (p == "c2") { genome(24); c.paired = false; c.units = 100e6; c.readlen = 100; c.dup = 0.05; c.seed = 20260103; }
    else if (p == "c3") { genome(1); c.paired = true; c.units = 24895642; c.readlen = 150; c.tlen_kind = 2; c.dup = 0.05; c.unmapped = 0.005; c.seed = 20260104; }
    else if (p == "c4") { genome(24); c.paired = true; c.units = 500e6; c.readlen = 50; c.tlen_kind = 1; c.dup = 0.05; c.unmapped = 0.005; c.seed = 20260105; }
    else if (p == "c5") { genome(24); c.paired = false; c.units = 500e6; c.readlen = 100; c.dup = 0.10; c.seed = 20260106; }
    else return false;
    c.units *= scale;
    return true;
}

}  // namespace

int main(int argc, char** argv) {
    Config cfg;
    std::string pre = "c2";
    double scale = 1.0, gscale = 1.0;
    bool seed_set = false;
    uint64_t seed = 0;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto val = [&]() -> std::string { if (i + 1 >= argc) { fprintf(stderr, "missing value for %s\n", a.c_str()); exit(2); } return argv[++i]; };
        if (a == "--out") cfg.out = val();
        else if (a == "--preset") pre = val();
        else if (a == "--scale") scale = atof(val().c_str());
        else if (a == "--genome-scale") gscale = atof(val().c_str());
        else if (a == "--threads") cfg.threads = atoi(val().c_str());
        else if (a == "--seed") { seed = strtoull(val().c_str(), nullptr, 10); seed_set = true; }
        else if (a == "--record") cfg.realistic = val() == "realistic";
        else if (a == "--level") cfg.level = atoi(val().c_str());
        else if (a == "--straddle") cfg.straddle = atof(val().c_str());
        else if (a == "--unplaced") cfg.unplaced = atoll(val().c_str());
        else { fprintf(stderr, "unknown option %s\n", a.c_str()); return 2; }
    }
    if (cfg.out.empty() || !preset(cfg, pre, scale, gscale)) { fprintf(stderr, "usage: bamgen --out x.bam --preset c2|c3|c4|c5 [--scale f] [--genome-scale f] ...\n"); return 2; }
    if (seed_set) cfg.seed = seed;
    if (cfg.threads <= 0) cfg.threads = int(std::thread::hardware_concurrency());
    if (cfg.threads < 1) cfg.threads = 1;

    FILE* fp = fopen(cfg.out.c_str(), "wb");
    if (!fp) { perror("open output"); return 1; }
    // header
    std::vector<uint8_t> hdr;
    std::string text = "@HD\tVN:1.6\tSO:coordinate\n";
    for (size_t i = 0; i < cfg.names.size(); ++i) text += "@SQ\tSN:" + cfg.names[i] + "\tLN:" + std::to_string(cfg.lens[i]) + "\n";
    hdr.insert(hdr.end(), {'B', 'A', 'M', 1});
    put32(hdr, uint32_t(text.size()));
    hdr.insert(hdr.end(), text.begin(), text.end());
    put32(hdr, uint32_t(cfg.names.size()));
    for (size_t i = 0; i < cfg.names.size(); ++i) {
        put32(hdr, uint32_t(cfg.names[i].size() + 1));
        hdr.insert(hdr.end(), cfg.names[i].begin(), cfg.names[i].end());
        hdr.push_back(0);
        put32(hdr, uint32_t(cfg.lens[i]));
    }
    std::vector<uint8_t> comp;
    for (size_t p = 0; p < hdr.size(); p += 0xff00) bgzf_block(hdr.data() + p, std::min<size_t>(0xff00, hdr.size() - p), cfg.level, comp);
    fwrite(comp.data(), 1, comp.size(), fp);
    uint64_t file_off = comp.size();

    // tiles
    struct Job { int tid; int64_t t; };
    std::vector<Job> jobs;
    for (size_t tid = 0; tid < cfg.lens.size(); ++tid)
        for (int64_t t = 0; t * TILE < cfg.lens[tid]; ++t) jobs.push_back(Job{int(tid), t});
    Gen gen(cfg);
    std::vector<TileOut> outs(jobs.size());
    std::atomic<size_t> next{0};
    std::mutex m;
    std::condition_variable cv;
    size_t written = 0;
    const size_t window = size_t(cfg.threads) * 3 + 2;
    std::vector<std::thread> th;
    for (int w = 0; w < cfg.threads; ++w)
        th.emplace_back([&] {
            for (;;) {
                const size_t j = next.fetch_add(1);
                if (j >= jobs.size()) return;
                {
                    std::unique_lock<std::mutex> lk(m);
                    cv.wait(lk, [&] { return j < written + window; });
                }
                gen.tile(jobs[j].tid, jobs[j].t, outs[j]);
                std::lock_guard<std::mutex> g(m);
                outs[j].done = true;
                cv.notify_all();
            }
        });

    // per-reference index, assembled in file order
    struct RefIdx { std::map<uint32_t, std::vector<Chunk>> bins; std::vector<uint64_t> lin; uint64_t beg = 0, end = 0, nm = 0, nu = 0; bool any = false; };
    std::vector<RefIdx> idx(cfg.lens.size());
    uint64_t total_records = 0, total_u = 0;
    for (size_t j = 0; j < jobs.size(); ++j) {
        {
            std::unique_lock<std::mutex> lk(m);
            cv.wait(lk, [&] { return outs[j].done; });
        }
        TileOut& o = outs[j];
        RefIdx& ri = idx[jobs[j].tid];
        const uint64_t base = file_off << 16;
        if (o.n_records) {
            if (!ri.any) { ri.beg = base; ri.any = true; }
            ri.end = (file_off + o.comp.size()) << 16;
            for (auto& kv : o.bins) {
                auto& dst = ri.bins[kv.first];
                for (auto c : kv.second) {
                    c.beg += base; c.end += base;
                    if (!dst.empty() && dst.back().end == c.beg) dst.back().end = c.end; else dst.push_back(c);
                }
            }
            for (auto& lw : o.linear) {
                if (size_t(lw.first) >= ri.lin.size()) ri.lin.resize(size_t(lw.first) + 1, 0);
                if (ri.lin[lw.first] == 0) ri.lin[lw.first] = lw.second + base;
            }
            ri.nm += o.n_mapped; ri.nu += o.n_unmapped;
        }
        fwrite(o.comp.data(), 1, o.comp.size(), fp);
        file_off += o.comp.size();
        total_records += o.n_records; total_u += o.usize;
        {
            std::lock_guard<std::mutex> g(m);
            std::vector<uint8_t>().swap(o.comp);
            o.bins.clear(); std::vector<std::pair<int64_t, uint64_t>>().swap(o.linear);
            written = j + 1;
            cv.notify_all();
        }
    }
    for (auto& t : th) t.join();
    // unplaced unmapped reads (tid -1) at the end
    if (cfg.unplaced > 0) {
        std::vector<uint8_t> buf; std::vector<Rec> recs; std::vector<uint32_t> nocig; Rng r(cfg.seed ^ 0xdead);
        Config c2 = cfg; c2.realistic = false; Gen g2(c2);
        for (int64_t i = 0; i < cfg.unplaced; ++i) g2.emit(buf, recs, -1, -1, 0, 4, nocig, 0, 0, -1, -1, 0, uint64_t(i), r);
        std::vector<uint8_t> cbuf;
        for (size_t p = 0; p < buf.size();) {
            size_t q = p, k = 0;
            while (k < recs.size() && recs[k].off < p) ++k;
            while (k < recs.size() && recs[k].off + recs[k].len - p <= 0xff00) { q = recs[k].off + recs[k].len; ++k; }
            if (q == p) q = std::min(buf.size(), p + 0xff00);
            bgzf_block(buf.data() + p, q - p, cfg.level, cbuf);
            p = q;
        }
        fwrite(cbuf.data(), 1, cbuf.size(), fp);
        file_off += cbuf.size();
        total_records += uint64_t(cfg.unplaced);
    }
    const uint8_t eof[28] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0, 0x1b, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    fwrite(eof, 1, 28, fp);
    fclose(fp);

    // BAI
    std::vector<uint8_t> bai{'B', 'A', 'I', 1};
    put32(bai, uint32_t(idx.size()));
    auto put64 = [&](uint64_t x) { for (int i = 0; i < 8; ++i) bai.push_back(uint8_t(x >> (8 * i))); };
    for (auto& ri : idx) {
        put32(bai, uint32_t(ri.bins.size() + (ri.any ? 1 : 0)));
        for (auto& kv : ri.bins) {
            put32(bai, kv.first); put32(bai, uint32_t(kv.second.size()));
            for (auto& c : kv.second) { put64(c.beg); put64(c.end); }
        }
        if (ri.any) { put32(bai, 37450); put32(bai, 2); put64(ri.beg); put64(ri.end); put64(ri.nm); put64(ri.nu); }
        for (size_t l = ri.lin.size(); l-- > 1;) if (ri.lin[l - 1] == 0) ri.lin[l - 1] = ri.lin[l];   // htslib back-fills
        put32(bai, uint32_t(ri.lin.size()));
        for (auto v : ri.lin) put64(v);
    }
    put64(uint64_t(cfg.unplaced));
    FILE* fi = fopen((cfg.out + ".bai").c_str(), "wb");
    if (!fi) { perror("open index"); return 1; }
    fwrite(bai.data(), 1, bai.size(), fi);
    fclose(fi);
    printf("{\"records\": %llu, \"uncompressed_bytes\": %llu, \"file_bytes\": %llu, \"contigs\": %zu, \"paired\": %s}\n",
           (unsigned long long)total_records, (unsigned long long)total_u, (unsigned long long)(file_off + 28), cfg.lens.size(),
           cfg.paired ? "true" : "false");
    return 0;
}
