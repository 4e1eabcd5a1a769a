// Single-lane DEFLATE decode core shared by the GPU inflate kernel and its host-side unit harness
// (tests/host_inflate_harness.cpp compiles this header with a plain C++ compiler).  Everything here is executed by ONE
// thread per BGZF block: bit reader, decode-table construction, and "phase 1" = Huffman symbols -> token queue.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define BSG_HD __host__ __device__ __forceinline__
#else
#define BSG_HD inline
#endif

namespace bsg {
namespace inflate_core {

#ifndef BSG_LIT_BITS
#define BSG_LIT_BITS 10
#endif
constexpr int kLitBits = BSG_LIT_BITS, kDistBits = 8;
constexpr int kQueue = 256;

// token: literal = byte ; match = 1 << 31 | (dist - 1) << 16 | len ; skip (stored bytes already in place) = 1 << 30 | len
constexpr uint32_t kTokMatch = 0x80000000u, kTokSkip = 0x40000000u;

// decode-table entry: [3:0] code length, [7:4] extra-bit count, [9:8] type, [31:16] literal byte / length base /
// distance base.  Slots no code of <= kLitBits (kDistBits) bits maps to hold kTypeBad with length 0: one type test
// sends literals and lengths down their fast paths and everything rare (end of block, long codes, errors) elsewhere.
enum : uint32_t { kTypeLit = 0u << 8, kTypeLen = 1u << 8, kTypeEob = 2u << 8, kTypeBad = 3u << 8, kTypeMask = 3u << 8 };

BSG_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t sh) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, sh);
#else
    sh &= 31;
    return sh ? (lo >> sh) | (hi << (32 - sh)) : lo;
#endif
}
BSG_HD uint32_t rev_bits(uint32_t code, uint32_t len) {
#if defined(__CUDA_ARCH__)
    return __brev(code) >> (32 - len);
#else
    uint32_t r = 0;
    for (uint32_t i = 0; i < len; ++i) r |= ((code >> i) & 1u) << (len - 1 - i);
    return r;
#endif
}

// Bit reader over aligned 32-bit words with two words of look-ahead: peek() is one funnel shift, consume() is an
// add plus a predicated word rotation whose load was issued a word earlier.  Positions are 32-bit word indices into
// one 4-byte aligned buffer (the batch's compressed bytes), so the state is five 32-bit registers.
struct BitReader {
    const uint32_t* base;  // the buffer (uniform for all blocks of a launch)
    uint32_t wi;           // index of the next word to load
    uint32_t w0, w1, w2, bo;
    BSG_HD void init(const uint32_t* buf, uint32_t byte_off) {
        base = buf;
        wi = byte_off >> 2;
        bo = (byte_off & 3u) * 8u;
        w0 = base[wi]; w1 = base[wi + 1]; w2 = base[wi + 2];
        wi += 3;
    }
    BSG_HD uint32_t peek() const { return funnel_r(w0, w1, bo); }           // next 32 bits
    BSG_HD uint32_t peek_hi() const { return funnel_r(w1, w2, bo); }        // the 32 bits after those
    BSG_HD void consume_short(uint32_t n) {                                   // n <= 32
        bo += n;
        if (bo >= 32) { bo -= 32; w0 = w1; w1 = w2; w2 = base[wi++]; }
    }
    BSG_HD void consume(uint32_t n) {                                         // n <= 64
        bo += n;
        if (bo >= 32) {
            bo -= 32; w0 = w1; w1 = w2; w2 = base[wi++];
            if (bo >= 32) { bo -= 32; w0 = w1; w1 = w2; w2 = base[wi++]; }
        }
    }
    BSG_HD uint64_t bit_pos() const { return uint64_t(wi - 3u) * 32u + bo; }  // bits from the start of the buffer
    BSG_HD uint32_t byte_pos() const { return (wi - 3u) * 4u + (bo >> 3); }   // only valid when bo is a multiple of 8
};

struct Tables {
    uint32_t lit[1 << kLitBits];
    uint32_t dist[1 << kDistBits];
    uint16_t lit_sorted[288];
    uint16_t dist_sorted[32];
    uint16_t lit_count[16], dist_count[16];
    uint8_t lens[320];
};

BSG_HD uint32_t lit_entry(int sym, int len) {
    if (sym < 256) return uint32_t(len) | kTypeLit | (uint32_t(sym) << 16);
    if (sym == 256) return uint32_t(len) | kTypeEob;
    if (sym > 285) return uint32_t(len) | kTypeBad;
    uint32_t eb, base;
    if (sym < 265) { eb = 0; base = uint32_t(sym - 254); }
    else if (sym == 285) { eb = 0; base = 258; }
    else { eb = uint32_t(sym - 261) >> 2; base = ((4u + (uint32_t(sym - 265) & 3u)) << eb) + 3u; }
    return uint32_t(len) | (eb << 4) | kTypeLen | (base << 16);
}
BSG_HD uint32_t dist_entry(int sym, int len) {
    if (sym > 29) return uint32_t(len) | kTypeBad;
    uint32_t eb, base;
    if (sym < 4) { eb = 0; base = uint32_t(sym + 1); }
    else { eb = (uint32_t(sym) >> 1) - 1u; base = ((2u + (uint32_t(sym) & 1u)) << eb) + 1u; }
    return uint32_t(len) | (eb << 4) | (base << 16);
}

// canonical decode of a code longer than the primary table (RFC 1951 3.2.2) from the 32 peeked bits;
// returns the symbol and its code length, or -1.  DIST selects the distance code's count / sorted arrays.
template <bool DIST, class A>
BSG_HD int slow_symbol(uint32_t v, const A& acc, int* len_out) {
    int code = 0, first = 0, index = 0;
    for (int len = 1; len <= 15; ++len) {
        code |= int(v & 1u);
        v >>= 1;
        const int c = int(acc.template count<DIST>(uint32_t(len)));
        if (code - c < first) { *len_out = len; return int(acc.template sorted<DIST>(uint32_t(index + (code - first)))); }
        index += c;
        first += c;
        first <<= 1;
        code <<= 1;
    }
    *len_out = 15;
    return -1;
}

// Serial table construction by the owning thread.  is_dist selects the entry format.  Returns false when the code
// is over-subscribed.
BSG_HD bool build_table(const uint8_t* lens, int n, uint32_t* primary, int bits, uint16_t* count, uint16_t* sorted, bool is_dist) {
    for (int i = 0; i < (1 << bits); ++i) primary[i] = kTypeBad;   // length 0 + kTypeBad = "not in the primary table"
    int cnt[16];
    for (int l = 0; l < 16; ++l) cnt[l] = 0;
    for (int s = 0; s < n; ++s) cnt[lens[s] & 15]++;
    cnt[0] = 0;
    int left = 1, ok = 1;
    for (int l = 1; l <= 15; ++l) { left <<= 1; left -= cnt[l]; if (left < 0) ok = 0; }
    int next[16], offs[16];
    next[0] = 0; offs[0] = 0; next[1] = 0; offs[1] = 0;
    for (int l = 1; l < 15; ++l) { next[l + 1] = (next[l] + cnt[l]) << 1; offs[l + 1] = offs[l] + cnt[l]; }
    for (int l = 0; l < 16; ++l) count[l] = uint16_t(cnt[l]);
    for (int s = 0; s < n; ++s) {
        const int l = lens[s] & 15;
        if (!l) continue;
        const uint32_t code = uint32_t(next[l]++);
        sorted[offs[l]++] = uint16_t(s);
        if (l <= bits) {
            const uint32_t e = is_dist ? dist_entry(s, l) : lit_entry(s, l);
            for (int k = int(rev_bits(code, uint32_t(l))); k < (1 << bits); k += (1 << l)) primary[k] = e;
        }
    }
    return ok != 0;
}

// Block header: reads BFINAL/BTYPE and, for Huffman blocks, the code lengths, then builds the tables.
// Returns 0 = Huffman block ready, 1 = stored block (caller handles LEN/NLEN), 2 = error.
BSG_HD int read_block_header(BitReader& br, Tables& T, int* last) {
    const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    uint32_t v = br.peek();
    *last = int(v & 1u);
    const int btype = int((v >> 1) & 3u);
    br.consume(3);
    if (btype == 0) return 1;
    if (btype == 3) return 2;
    int nlit = 288, ndist = 32;
    if (btype == 1) {
        for (int s = 0; s < 288; ++s) T.lens[s] = uint8_t(s < 144 ? 8 : (s < 256 ? 9 : (s < 280 ? 7 : 8)));
        for (int s = 0; s < 32; ++s) T.lens[288 + s] = 5;
    } else {
        uint8_t* cl = reinterpret_cast<uint8_t*>(T.lit);      // code-length table borrows the (not yet built) lit table
        v = br.peek();
        nlit = int(v & 31u) + 257;
        ndist = int((v >> 5) & 31u) + 1;
        const int ncl = int((v >> 10) & 15u) + 4;
        br.consume(14);
        uint8_t cll[19];
        for (int i = 0; i < 19; ++i) cll[i] = 0;
        for (int i = 0; i < ncl; ++i) { cll[order[i]] = uint8_t(br.peek() & 7u); br.consume(3); }
        int cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0}, next[8];
        for (int i = 0; i < 19; ++i) cnt[cll[i]]++;
        cnt[0] = 0;
        next[1] = 0;
        for (int l = 1; l < 7; ++l) next[l + 1] = (next[l] + cnt[l]) << 1;
        for (int i = 0; i < 128; ++i) cl[i] = 0;
        for (int s = 0; s < 19; ++s) {
            const int l = cll[s];
            if (!l) continue;
            const uint32_t r = rev_bits(uint32_t(next[l]++), uint32_t(l));
            for (uint32_t k = r; k < 128; k += (1u << l)) cl[k] = uint8_t(s | (l << 5));
        }
        if (nlit > 286 || ndist > 30) return 2;
        const int total = nlit + ndist;
        int i = 0;
        while (i < total) {
            v = br.peek();
            const uint32_t e = cl[v & 127u];
            const int l = int(e >> 5), s = int(e & 31u);
            if (!l) return 2;
            if (s < 16) { T.lens[i++] = uint8_t(s); br.consume(uint32_t(l)); continue; }
            int rep, val = 0;
            uint32_t used = uint32_t(l);
            if (s == 16) { if (i == 0) return 2; val = T.lens[i - 1]; rep = 3 + int((v >> l) & 3u); used += 2; }
            else if (s == 17) { rep = 3 + int((v >> l) & 7u); used += 3; }
            else { rep = 11 + int((v >> l) & 127u); used += 7; }
            br.consume(used);
            if (i + rep > total) return 2;
            while (rep--) T.lens[i++] = uint8_t(val);
        }
        if (T.lens[256] == 0) return 2;
        for (int k = ndist - 1; k >= 0; --k) T.lens[288 + k] = T.lens[nlit + k];
        for (int k = nlit; k < 288; ++k) T.lens[k] = 0;
    }
    bool ok = build_table(T.lens, 288, T.lit, kLitBits, T.lit_count, T.lit_sorted, false);
    ok = build_table(T.lens + 288, ndist, T.dist, kDistBits, T.dist_count, T.dist_sorted, true) && ok;
    return ok ? 0 : 2;
}

// Table / queue access of fill_queue.  The host harness (and any generic caller) uses plain arrays; the kernel
// passes shared-memory byte addresses and ld.shared / st.shared so that the addresses stay in registers.
struct ArrayAccess {
    const Tables* T;
    uint32_t* q;
    BSG_HD uint32_t lit(uint32_t byte_off) const { return T->lit[byte_off >> 2]; }
    BSG_HD uint32_t dist(uint32_t byte_off) const { return T->dist[byte_off >> 2]; }
    BSG_HD void put(uint32_t byte_off, uint32_t v) const { q[byte_off >> 2] = v; }
    template <bool DIST> BSG_HD uint32_t count(uint32_t len) const { return DIST ? T->dist_count[len] : T->lit_count[len]; }
    template <bool DIST> BSG_HD uint32_t sorted(uint32_t i) const { return DIST ? T->dist_sorted[i] : T->lit_sorted[i]; }
};

// Phase 1: decode symbols into the queue until it holds kQueue tokens or the block ends.
// *op_dec = bytes decoded so far (updated).  Returns the number of tokens; *eob is set at end-of-block; *bad is
// sticky (errors do not stop the loop: every access stays in bounds, the caller discards the round).
// A match is decoded from ONE 64-bit look-ahead (length code + extra bits <= 20 bits, distance code + extra bits
// <= 28 bits) and consumed once.
template <class A>
BSG_HD int fill_queue(BitReader& br, const A& acc, uint32_t* op_dec, int* eob, int* bad) {
    constexpr uint32_t kLitMask4 = ((1u << kLitBits) - 1u) << 2, kDistMask4 = ((1u << kDistBits) - 1u) << 2;
    uint32_t qo = 0;                  // byte offset of the next queue slot
    uint32_t op = *op_dec;
    int err = 0;
    *eob = 0;
    do {
        const uint32_t v = br.peek();
        uint32_t e = acc.lit((v << 2) & kLitMask4);
        uint32_t type = e & kTypeMask;
        if (type == kTypeLit) {                       // literal with a short code: the common non-match case
            acc.put(qo, e >> 16);
            br.consume_short(e & 15u);
            qo += 4; ++op;
            continue;
        }
        if (type != kTypeLen) {                       // rare: end of block, a code longer than the primary table, garbage
            if (!(e & 15u)) {
                int l;
                const int sym = slow_symbol<false>(v, acc, &l);
                e = sym < 0 ? (uint32_t(l) | kTypeBad) : lit_entry(sym, l);
                type = e & kTypeMask;
            }
            if (type == kTypeLit) {
                acc.put(qo, e >> 16);
                br.consume_short(e & 15u);
                qo += 4; ++op;
                continue;
            }
            if (type != kTypeLen) {
                br.consume_short(e & 15u);
                if (type == kTypeEob) *eob = 1; else err = 1;
                break;
            }
        }
        const uint32_t len = e & 15u, eb = (e >> 4) & 15u, used = len + eb;
        const uint32_t mlen = (e >> 16) + ((v >> len) & ~(~0u << eb));
        const uint32_t v2 = funnel_r(v, br.peek_hi(), used);      // used <= 20
        uint32_t d = acc.dist((v2 << 2) & kDistMask4);
        if (!(d & 15u)) {
            int l;
            const int sym = slow_symbol<true>(v2, acc, &l);
            d = sym < 0 ? (uint32_t(l) | kTypeBad) : dist_entry(sym, l);
        }
        const uint32_t dl = d & 15u, deb = (d >> 4) & 15u;
        uint32_t mdist = (d >> 16) + ((v2 >> dl) & ~(~0u << deb));
        br.consume(used + dl + deb);                               // <= 48
        if ((d & kTypeMask) != 0u || mdist > op) { err = 1; mdist = 1; }
        acc.put(qo, kTokMatch | ((mdist - 1u) << 16) | mlen);
        qo += 4; op += mlen;
    } while (qo < uint32_t(kQueue) * 4u);
    *op_dec = op;
    *bad |= err;
    return int(qo >> 2);
}

}  // namespace inflate_core
}  // namespace bsg

