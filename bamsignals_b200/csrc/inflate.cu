// GPU-side BGZF inflate and record-boundary walk: the host only reads the file and ships
// COMPRESSED bytes; raw-DEFLATE decoding of every BGZF block and the block_size chain walk run on the device.
//
// k_inflate: one warp per BGZF block (RFC 1951 stream, <= 64 KiB out).  Lane 0 owns the bit reader and decodes
// Huffman symbols through per-warp shared-memory tables (10-bit primary table for literal/length codes, 8-bit for
// distances, canonical bit-by-bit fallback for longer codes); literals are stored directly, every match is
// broadcast and copied by all 32 lanes.  Table construction at each dynamic block header uses the whole warp.
// Throughput comes from block-level parallelism: ~64 warps resident per SM, tens of thousands of blocks per batch.
//
// k_walk_count / k_walk_write: one thread per index entry point (BAI linear-index offsets and chunk bounds are
// record-aligned); each walks block_size -> next record until the next entry point; a scan of the counts in between
// gives every walker its slice of the offsets array (two passes, no atomics, deterministic order).
#include "kernels.cuh"

#include <cstdlib>

#include "inflate_core.cuh"

namespace bsg {
namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int kWarpsPerCta = 4;
constexpr int kLitBits = 10, kDistBits = 8;

struct WarpTables {
    uint16_t lit[1 << kLitBits];     // sym | len << 9 ; 0 = not in primary table (code longer than kLitBits)
    uint16_t dist[1 << kDistBits];   // sym | len << 9
    uint16_t lit_sorted[288];        // canonical order, for the slow path
    uint16_t dist_sorted[32];
    uint16_t lit_count[16], dist_count[16];
    uint16_t code[320];              // scratch: bit-reversed canonical code per symbol
    uint8_t lens[320];               // scratch: code lengths (literal/length then distance)
    uint8_t cl[128];                 // code-length alphabet: sym | len << 5
};

struct BitReader {
    const uint32_t* wp;      // next aligned word to load
    uint64_t bb;
    uint32_t nb;
    uint64_t consumed_limit; // total bits available
    uint64_t loaded;         // total bits loaded into bb so far
    __device__ __forceinline__ void init(const uint8_t* in, uint32_t in_len) {
        const uintptr_t a = reinterpret_cast<uintptr_t>(in);
        wp = reinterpret_cast<const uint32_t*>(a & ~uintptr_t(3));
        const uint32_t mis = uint32_t(a & 3);
        bb = uint64_t(*wp++) >> (8 * mis);
        nb = 32 - 8 * mis;
        loaded = nb;
        consumed_limit = uint64_t(in_len) * 8;
    }
    __device__ __forceinline__ void refill() {       // guarantees nb >= 32
        if (nb < 32) {
            bb |= uint64_t(*wp++) << nb;
            nb += 32;
            loaded += 32;
        }
    }
    __device__ __forceinline__ uint32_t peek(uint32_t n) const { return uint32_t(bb) & ((1u << n) - 1u); }
    __device__ __forceinline__ void drop(uint32_t n) { bb >>= n; nb -= n; }
    __device__ __forceinline__ uint32_t take(uint32_t n) { const uint32_t v = peek(n); drop(n); return v; }
    __device__ __forceinline__ bool overrun() const { return loaded - nb > consumed_limit; }
};

__device__ __forceinline__ uint32_t rev_bits(uint32_t code, uint32_t len) { return __brev(code) >> (32 - len); }

// canonical slow path: decode one symbol bit by bit (RFC 1951 3.2.2); returns -1 on an invalid code
__device__ __forceinline__ int slow_decode(BitReader& br, const volatile uint16_t* count, const volatile uint16_t* sorted) {
    int code = 0, first = 0, index = 0;
    for (int len = 1; len <= 15; ++len) {
        code |= int(br.take(1));
        const int c = count[len];
        if (code - c < first) return sorted[index + (code - first)];
        index += c;
        first += c;
        first <<= 1;
        code <<= 1;
    }
    return -1;
}

// Build decode tables for `n` symbols with code lengths lens[0..n) (warp-cooperative).  Returns false if the code is
// over-subscribed.  primary: table of (1 << bits) uint16 entries.
// All table memory is accessed through volatile pointers: the tables are produced by one set of lanes and consumed
// by another with only warp-level barriers in between, and the optimised build mis-ordered plain shared accesses
// here (found with compute-sanitizer; the -G build and the volatile build are bit-exact against zlib).
__device__ bool build_table(const volatile uint8_t* lens, int n, volatile uint16_t* primary, int bits, volatile uint16_t* count,
                            volatile uint16_t* sorted, volatile uint16_t* code_scratch, int lane) {
    __syncwarp();
    for (int i = lane; i < (1 << bits); i += 32) primary[i] = 0;
    int ok = 1;
    if (lane == 0) {
        int cnt[16];
#pragma unroll
        for (int l = 0; l < 16; ++l) cnt[l] = 0;
        for (int s = 0; s < n; ++s) cnt[lens[s] & 15]++;
        cnt[0] = 0;
        int left = 1;
        for (int l = 1; l <= 15; ++l) { left <<= 1; left -= cnt[l]; if (left < 0) ok = 0; }
        int next[16], offs[16];
        next[0] = 0; offs[0] = 0; next[1] = 0; offs[1] = 0;
        for (int l = 1; l < 15; ++l) { next[l + 1] = (next[l] + cnt[l]) << 1; offs[l + 1] = offs[l] + cnt[l]; }
        for (int l = 0; l < 16; ++l) count[l] = uint16_t(cnt[l]);
        for (int s = 0; s < n; ++s) {
            const int l = lens[s] & 15;
            if (l) { code_scratch[s] = uint16_t(rev_bits(uint32_t(next[l]++), uint32_t(l))); sorted[offs[l]++] = uint16_t(s); }
        }
    }
    ok = __shfl_sync(FULL, ok, 0);
    __syncwarp();
    for (int s = lane; s < n; s += 32) {
        const int l = lens[s] & 15;
        if (l && l <= bits) {
            const uint16_t e = uint16_t(s | (l << 9));
            for (int k = code_scratch[s]; k < (1 << bits); k += (1 << l)) primary[k] = e;
        }
    }
    __syncwarp();
    return ok != 0;
}

__constant__ uint8_t kClOrder[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kWarpsPerCta * 32) k_inflate(const InflateBlock* __restrict__ blocks, int n_blocks,
                                                                const uint8_t* __restrict__ comp, uint8_t* __restrict__ raw,
                                                                DeviceScalars* sc) {
    __shared__ WarpTables s_tab[kWarpsPerCta];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * kWarpsPerCta + wid;
    if (b >= n_blocks) return;
    volatile WarpTables& T = s_tab[wid];
    const InflateBlock blk = blocks[b];
    uint8_t* out = raw + blk.out_off;
    const uint32_t out_len = blk.out_len;
    BitReader br;
    if (lane == 0) br.init(comp + blk.in_off, blk.in_len);
    uint32_t op = 0;          // bytes produced (lane 0 authoritative, broadcast at sync points)
    int err = 0, last = 0;
    while (!last && !err) {
        // ---- block header (lane 0), tables (warp) ----------------------------------------------------------------
        int btype = 0;
        if (lane == 0) {
            br.refill();
            last = int(br.take(1));
            btype = int(br.take(2));
        }
        last = __shfl_sync(FULL, last, 0);
        btype = __shfl_sync(FULL, btype, 0);
        if (btype == 0) {                                   // stored block
            uint32_t len = 0;
            const uint8_t* src = nullptr;
            if (lane == 0) {
                br.drop(br.nb & 7);                          // to the byte boundary
                br.refill();
                len = br.take(16);
                br.refill();
                const uint32_t nlen = br.take(16);
                if ((len ^ nlen) != 0xffffu || op + len > out_len) err = 1;
                // bytes still in the bit buffer belong to the stored data: rewind the word pointer onto them
                const uint8_t* p = reinterpret_cast<const uint8_t*>(br.wp) - (br.nb >> 3);
                src = p;
            }
            err = __shfl_sync(FULL, err, 0);
            len = __shfl_sync(FULL, len, 0);
            const uint64_t srcb = __shfl_sync(FULL, (unsigned long long)reinterpret_cast<uintptr_t>(src), 0);
            op = __shfl_sync(FULL, op, 0);
            if (!err) {
                const uint8_t* s = reinterpret_cast<const uint8_t*>(uintptr_t(srcb));
                for (uint32_t j = lane; j < len; j += 32) out[op + j] = s[j];
                op += len;
                if (lane == 0) {                             // restart the bit reader behind the stored bytes
                    const uint64_t used = (br.loaded - br.nb) + uint64_t(len) * 8;
                    const uint8_t* np = s + len;
                    const uint64_t limit = br.consumed_limit;
                    br.init(np, 0);
                    br.consumed_limit = limit;
                    br.loaded = used + br.nb;
                }
            }
            __syncwarp();
            continue;
        }
        if (btype == 3) { err = 1; break; }
        int nlit = 288, ndist = 30;
        if (btype == 1) {                                   // fixed Huffman code
            for (int s = lane; s < 288; s += 32) T.lens[s] = s < 144 ? 8 : (s < 256 ? 9 : (s < 280 ? 7 : 8));
            for (int s = lane; s < 32; s += 32) T.lens[288 + s] = 5;
            ndist = 32;
            __syncwarp();
        } else {                                            // dynamic Huffman code
            if (lane == 0) {
                br.refill();
                nlit = int(br.take(5)) + 257;
                ndist = int(br.take(5)) + 1;
                const int ncl = int(br.take(4)) + 4;
                uint8_t cll[19];
#pragma unroll
                for (int i = 0; i < 19; ++i) cll[i] = 0;
                for (int i = 0; i < ncl; ++i) { br.refill(); cll[kClOrder[i]] = uint8_t(br.take(3)); }
                // code-length code: canonical codes into a 128-entry direct table
                int cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0}, next[8];
                for (int i = 0; i < 19; ++i) cnt[cll[i]]++;
                cnt[0] = 0;
                next[1] = 0;
                for (int l = 1; l < 7; ++l) next[l + 1] = (next[l] + cnt[l]) << 1;
                for (int i = 0; i < 128; ++i) T.cl[i] = 0;
                for (int s = 0; s < 19; ++s) {
                    const int l = cll[s];
                    if (!l) continue;
                    const uint32_t r = rev_bits(uint32_t(next[l]++), uint32_t(l));
                    for (uint32_t k = r; k < 128; k += (1u << l)) T.cl[k] = uint8_t(s | (l << 5));
                }
                // literal/length + distance code lengths
                int i = 0;
                const int total = nlit + ndist;
                if (nlit > 286 || ndist > 30) err = 1;
                while (i < total && !err) {
                    br.refill();
                    const uint32_t e = T.cl[br.peek(7)];
                    const int l = int(e >> 5), s = int(e & 31);
                    if (!l) { err = 1; break; }
                    br.drop(uint32_t(l));
                    if (s < 16) { T.lens[i++] = uint8_t(s); continue; }
                    int rep, val = 0;
                    if (s == 16) { if (i == 0) { err = 1; break; } val = T.lens[i - 1]; rep = 3 + int(br.take(2)); }
                    else if (s == 17) rep = 3 + int(br.take(3));
                    else rep = 11 + int(br.take(7));
                    if (i + rep > total) { err = 1; break; }
                    while (rep--) T.lens[i++] = uint8_t(val);
                }
                if (!err && T.lens[256] == 0) err = 1;
                // move the distance lengths to their fixed place behind 288 literal/length slots
                if (!err) {
                    for (int k = ndist - 1; k >= 0; --k) T.lens[288 + k] = T.lens[nlit + k];
                    for (int k = nlit; k < 288; ++k) T.lens[k] = 0;
                }
            }
            err = __shfl_sync(FULL, err, 0);
            nlit = __shfl_sync(FULL, nlit, 0);
            ndist = __shfl_sync(FULL, ndist, 0);
            if (err) break;
            __syncwarp();
        }
        bool ok = build_table(T.lens, 288, T.lit, kLitBits, T.lit_count, T.lit_sorted, T.code, lane);
        ok = build_table(T.lens + 288, ndist, T.dist, kDistBits, T.dist_count, T.dist_sorted, T.code, lane) && ok;
        if (!ok) { err = 1; break; }

        // ---- symbols: lane 0 decodes, matches are copied by the whole warp ---------------------------------------------
        for (;;) {
            uint32_t mlen = 0, mdist = 0;
            int state = 0;                                  // 0 = match pending, 1 = end of block, 2 = error
            if (lane == 0) {
                for (;;) {
                    br.refill();
                    uint32_t e = T.lit[br.peek(kLitBits)];
                    int sym;
                    if (e) { br.drop(e >> 9); sym = int(e & 511u); }
                    else { sym = slow_decode(br, T.lit_count, T.lit_sorted); if (sym < 0) { state = 2; break; } }
                    if (sym < 256) {
                        if (op >= out_len) { state = 2; break; }
                        out[op++] = uint8_t(sym);
                        continue;
                    }
                    if (sym == 256) { state = 1; break; }
                    if (sym > 285) { state = 2; break; }
                    // length
                    if (sym < 265) mlen = uint32_t(sym - 254);
                    else if (sym == 285) mlen = 258;
                    else {
                        const uint32_t eb = uint32_t(sym - 261) >> 2;
                        mlen = ((4u + (uint32_t(sym - 265) & 3u)) << eb) + 3u + br.take(eb);
                    }
                    // distance
                    br.refill();
                    e = T.dist[br.peek(kDistBits)];
                    int ds;
                    if (e) { br.drop(e >> 9); ds = int(e & 511u); }
                    else { ds = slow_decode(br, T.dist_count, T.dist_sorted); if (ds < 0) { state = 2; break; } }
                    if (ds > 29) { state = 2; break; }
                    if (ds < 4) mdist = uint32_t(ds + 1);
                    else {
                        const uint32_t eb = (uint32_t(ds) >> 1) - 1u;
                        br.refill();
                        mdist = ((2u + (uint32_t(ds) & 1u)) << eb) + 1u + br.take(eb);
                    }
                    if (mdist > op || op + mlen > out_len) { state = 2; break; }
                    break;
                }
                if (br.overrun()) state = 2;
            }
            state = __shfl_sync(FULL, state, 0);
            if (state) { if (state == 2) err = 1; break; }
            mlen = __shfl_sync(FULL, mlen, 0);
            mdist = __shfl_sync(FULL, mdist, 0);
            op = __shfl_sync(FULL, op, 0);
            __syncwarp();                                   // lane 0's literal stores are visible to the copying lanes
            const uint8_t* src = out + op - mdist;
            if (mdist >= mlen) {
                for (uint32_t j = lane; j < mlen; j += 32) out[op + j] = src[j];
            } else {
                for (uint32_t j = lane; j < mlen; j += 32) out[op + j] = src[j % mdist];
            }
            op += mlen;
            __syncwarp();
        }
    }
    op = __shfl_sync(FULL, op, 0);
    if (lane == 0 && (err || op != out_len)) atomicOr(&sc->status, STATUS_BAD_DEFLATE);
}

// ---------------------------------------------------------------------------------------------------------------
// k_inflate_ms: D independent DEFLATE streams (BGZF blocks) per warp, decoded in LOCKSTEP.
//
// The single-stream kernel above is issue-bound: ~120 warp instructions per symbol are spent with one active lane.
// Here the leader lane of each of the D lane-groups (G = 32 / D lanes) runs the same decode step on its own stream at
// the same time, so one issued instruction advances D streams; the group's lanes then copy the match their leader
// decoded.  Every stream's tables are private to its leader lane (built serially by that lane), so no table memory is
// shared between lanes.  All 32 lanes execute every warp-level primitive of the step (full-mask shuffles with a
// per-lane source), which keeps the streams converged.
// ---------------------------------------------------------------------------------------------------------------
template <int LB, int DB>
struct StreamTables {
    uint16_t lit[1 << LB];       // sym | len << 9 ; 0 = code longer than LB bits (slow path)
    uint16_t dist[1 << DB];
    uint16_t lit_sorted[288];
    uint16_t dist_sorted[32];
    uint16_t lit_count[16], dist_count[16];
    uint8_t lens[320];
};

__device__ __forceinline__ int slow_decode_nv(BitReader& br, const uint16_t* count, const uint16_t* sorted) {
    int code = 0, first = 0, index = 0;
    for (int len = 1; len <= 15; ++len) {
        code |= int(br.take(1));
        const int c = count[len];
        if (code - c < first) return sorted[index + (code - first)];
        index += c;
        first += c;
        first <<= 1;
        code <<= 1;
    }
    return -1;
}

// serial build by the owning lane; returns false if over-subscribed
__device__ bool build_serial(const uint8_t* lens, int n, uint16_t* primary, int bits, uint16_t* count, uint16_t* sorted) {
    uint32_t* p32 = reinterpret_cast<uint32_t*>(primary);
    for (int i = 0; i < (1 << bits) / 2; ++i) p32[i] = 0;
    int cnt[16];
#pragma unroll
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
            const uint16_t e = uint16_t(s | (l << 9));
            for (int k = int(rev_bits(code, uint32_t(l))); k < (1 << bits); k += (1 << l)) primary[k] = e;
        }
    }
    return ok != 0;
}

template <int D, int LB, int DB, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k_inflate_ms(const InflateBlock* __restrict__ blocks, int n_blocks,
                                                            const uint8_t* __restrict__ comp, uint8_t* raw, DeviceScalars* sc) {
    constexpr int G = 32 / D;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    using Tab = StreamTables<LB, DB>;
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane / G, glane = lane % G;
    const bool leader = glane == 0;
    const int leader_lane = g * G;
    Tab& T = reinterpret_cast<Tab*>(smem_raw)[wid * D + g];
    const int stream = (blockIdx.x * WARPS + wid) * D + g;
    InflateBlock blk = {0, 0, 0, 0};
    int phase = 2;                                   // 0 = block header next, 1 = symbols, 2 = done
    if (stream < n_blocks) { blk = blocks[stream]; phase = 0; }
    uint8_t* out = raw + blk.out_off;                // same for every lane of the group
    const uint32_t out_len = blk.out_len;
    BitReader br;
    br.wp = nullptr; br.bb = 0; br.nb = 0; br.consumed_limit = 0; br.loaded = 0;
    if (leader && phase == 0) br.init(comp + blk.in_off, blk.in_len);
    uint32_t op = 0;
    int err = 0, last = 0;
    for (;;) {
        // ---- block headers: only the leaders that need one (rare: 1-2 per BGZF block) -------------------------------
        if (leader && phase == 0 && !err) {
            br.refill();
            last = int(br.take(1));
            const int btype = int(br.take(2));
            if (btype == 0) {                                        // stored: copied by the leader alone
                br.drop(br.nb & 7);
                br.refill();
                const uint32_t len = br.take(16);
                br.refill();
                const uint32_t nlen = br.take(16);
                if ((len ^ nlen) != 0xffffu || op + len > out_len) err = 1;
                else {
                    const uint8_t* src = reinterpret_cast<const uint8_t*>(br.wp) - (br.nb >> 3);
                    for (uint32_t j = 0; j < len; ++j) out[op + j] = src[j];
                    op += len;
                    const uint64_t used = (br.loaded - br.nb) + uint64_t(len) * 8, limit = br.consumed_limit;
                    br.init(src + len, 0);
                    br.consumed_limit = limit;
                    br.loaded = used + br.nb;
                    if (last) phase = 2;
                }
            } else if (btype == 3) {
                err = 1;
            } else {
                int nlit = 288, ndist = 32;
                if (btype == 1) {
                    for (int s = 0; s < 288; ++s) T.lens[s] = s < 144 ? 8 : (s < 256 ? 9 : (s < 280 ? 7 : 8));
                    for (int s = 0; s < 32; ++s) T.lens[288 + s] = 5;
                } else {
                    uint8_t* cl = reinterpret_cast<uint8_t*>(T.lit);   // the code-length table borrows the (not yet built) lit table
                    br.refill();
                    nlit = int(br.take(5)) + 257;
                    ndist = int(br.take(5)) + 1;
                    const int ncl = int(br.take(4)) + 4;
                    uint8_t cll[19];
#pragma unroll
                    for (int i = 0; i < 19; ++i) cll[i] = 0;
                    for (int i = 0; i < ncl; ++i) { br.refill(); cll[kClOrder[i]] = uint8_t(br.take(3)); }
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
                    int i = 0;
                    const int total = nlit + ndist;
                    if (nlit > 286 || ndist > 30) err = 1;
                    while (i < total && !err) {
                        br.refill();
                        const uint32_t e = cl[br.peek(7)];
                        const int l = int(e >> 5), s = int(e & 31);
                        if (!l) { err = 1; break; }
                        br.drop(uint32_t(l));
                        if (s < 16) { T.lens[i++] = uint8_t(s); continue; }
                        int rep, val = 0;
                        if (s == 16) { if (i == 0) { err = 1; break; } val = T.lens[i - 1]; rep = 3 + int(br.take(2)); }
                        else if (s == 17) rep = 3 + int(br.take(3));
                        else rep = 11 + int(br.take(7));
                        if (i + rep > total) { err = 1; break; }
                        while (rep--) T.lens[i++] = uint8_t(val);
                    }
                    if (!err && T.lens[256] == 0) err = 1;
                    if (!err) {
                        for (int k = ndist - 1; k >= 0; --k) T.lens[288 + k] = T.lens[nlit + k];
                        for (int k = nlit; k < 288; ++k) T.lens[k] = 0;
                    }
                }
                if (!err) {
                    bool ok = build_serial(T.lens, 288, T.lit, LB, T.lit_count, T.lit_sorted);
                    ok = build_serial(T.lens + 288, ndist, T.dist, DB, T.dist_count, T.dist_sorted) && ok;
                    if (!ok) err = 1; else phase = 1;
                }
            }
        }
        __syncwarp();
        // ---- lockstep symbol steps ----------------------------------------------------------------------------------------------
#pragma unroll 1
        for (int it = 0; it < 128; ++it) {
            uint32_t mlen = 0, mdist = 0;
            if (leader && phase == 1 && !err) {
                br.refill();
                uint32_t e = T.lit[br.peek(LB)];
                int sym;
                if (e) { br.drop(e >> 9); sym = int(e & 511u); }
                else { sym = slow_decode_nv(br, T.lit_count, T.lit_sorted); if (sym < 0) { err = 1; sym = 0; } }
                if (sym < 256) {
                    if (op >= out_len) err = 1; else out[op++] = uint8_t(sym);
                } else if (sym == 256) {
                    phase = last ? 2 : 0;
                } else if (sym > 285) {
                    err = 1;
                } else {
                    if (sym < 265) mlen = uint32_t(sym - 254);
                    else if (sym == 285) mlen = 258;
                    else {
                        const uint32_t eb = uint32_t(sym - 261) >> 2;
                        mlen = ((4u + (uint32_t(sym - 265) & 3u)) << eb) + 3u + br.take(eb);
                    }
                    br.refill();
                    e = T.dist[br.peek(DB)];
                    int ds;
                    if (e) { br.drop(e >> 9); ds = int(e & 511u); }
                    else { ds = slow_decode_nv(br, T.dist_count, T.dist_sorted); if (ds < 0) { err = 1; ds = 0; } }
                    if (ds > 29) { err = 1; ds = 0; }
                    if (ds < 4) mdist = uint32_t(ds + 1);
                    else {
                        const uint32_t eb = (uint32_t(ds) >> 1) - 1u;
                        br.refill();
                        mdist = ((2u + (uint32_t(ds) & 1u)) << eb) + 1u + br.take(eb);
                    }
                    if (mdist > op || op + mlen > out_len || br.overrun()) { err = 1; mlen = 0; }
                }
            }
            // every lane fetches its leader's match (mlen == 0: nothing to copy)
            mlen = __shfl_sync(FULL, mlen, leader_lane);
            mdist = __shfl_sync(FULL, mdist, leader_lane);
            const uint32_t gop = __shfl_sync(FULL, op, leader_lane);
            if (mlen) {
                const uint8_t* src = out + gop - mdist;
                if (mdist >= mlen) {
                    for (uint32_t j = glane; j < mlen; j += G) out[gop + j] = src[j];
                } else {
                    for (uint32_t j = glane; j < mlen; j += G) out[gop + j] = src[j % mdist];
                }
                if (leader) op += mlen;
            }
            __syncwarp();
        }
        if (__all_sync(FULL, !leader || phase == 2 || err != 0)) break;
    }
    if (leader && stream < n_blocks && (err || op != out_len || br.overrun())) atomicOr(&sc->status, STATUS_BAD_DEFLATE);
}

// ---------------------------------------------------------------------------------------------------------------
// k_inflate_q: one warp per BGZF block, two alternating phases per batch of <= 256 symbols.
//   phase 1  lane 0 decodes Huffman symbols into a shared-memory token queue (no output traffic, no warp sync);
//   phase 2  all 32 lanes materialise the queue: token lengths are scanned across the warp to get output positions,
//            every literal of a 32-token chunk is stored by its own lane in ONE store instruction, matches are then
//            replayed in order, each copied by the whole warp.
// Compared with k_inflate this removes the per-symbol warp round trip (shuffles, divergence reconvergence) and the
// single-lane byte stores that made it issue-bound at ~120 warp instructions per symbol.
// token: literal = byte ; match = 1 << 31 | (dist - 1) << 16 | len
// ---------------------------------------------------------------------------------------------------------------
constexpr int kQueue = 256;
constexpr int kQWarps = 4;

struct QTables {
    uint16_t lit[1 << kLitBits];
    uint16_t dist[1 << kDistBits];
    uint16_t lit_sorted[288];
    uint16_t dist_sorted[32];
    uint16_t lit_count[16], dist_count[16];
    uint8_t lens[320];
    uint32_t q[kQueue];
};

__global__ void __launch_bounds__(kQWarps * 32, 8) k_inflate_q(const InflateBlock* __restrict__ blocks, int n_blocks,
                                                                const uint8_t* __restrict__ comp, uint8_t* raw, DeviceScalars* sc) {
    __shared__ QTables s_tab[kQWarps];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * kQWarps + wid;
    if (b >= n_blocks) return;
    QTables& T = s_tab[wid];
    volatile uint32_t* q = T.q;
    const InflateBlock blk = blocks[b];
    uint8_t* out = raw + blk.out_off;
    const uint32_t out_len = blk.out_len;
    BitReader br;
    br.wp = nullptr; br.bb = 0; br.nb = 0; br.consumed_limit = 0; br.loaded = 0;
    if (lane == 0) br.init(comp + blk.in_off, blk.in_len);
    uint32_t op_dec = 0;       // lane 0: bytes decoded so far
    uint32_t pos_base = 0;     // all lanes: bytes materialised so far
    int phase = 0;             // lane 0: 0 = header next, 1 = symbols
    int last = 0;
    for (;;) {
        int nq = 0, state = 0;   // state: 0 = go on, 1 = stream finished, 2 = error
        if (lane == 0) {
            if (phase == 0) {
                br.refill();
                last = int(br.take(1));
                const int btype = int(br.take(2));
                if (btype == 0) {
                    // stored block: queue its bytes as literals (rare: level-0 files), 256 per round
                    br.drop(br.nb & 7);
                    br.refill();
                    const uint32_t len = br.take(16);
                    br.refill();
                    const uint32_t nlen = br.take(16);
                    if ((len ^ nlen) != 0xffffu || op_dec + len > out_len) state = 2;
                    else {
                        const uint8_t* src = reinterpret_cast<const uint8_t*>(br.wp) - (br.nb >> 3);
                        const uint64_t used = (br.loaded - br.nb) + uint64_t(len) * 8, limit = br.consumed_limit;
                        // copy directly (single lane; the other lanes are idle in this round anyway)
                        for (uint32_t j = 0; j < len; ++j) out[op_dec + j] = src[j];
                        op_dec += len;
                        br.init(src + len, 0);
                        br.consumed_limit = limit;
                        br.loaded = used + br.nb;
                        q[0] = 0x40000000u | len;                 // "skip len bytes" token
                        nq = 1;
                        if (last) state = 1;
                    }
                } else if (btype == 3) {
                    state = 2;
                } else {
                    int nlit = 288, ndist = 32, err = 0;
                    if (btype == 1) {
                        for (int s = 0; s < 288; ++s) T.lens[s] = s < 144 ? 8 : (s < 256 ? 9 : (s < 280 ? 7 : 8));
                        for (int s = 0; s < 32; ++s) T.lens[288 + s] = 5;
                    } else {
                        uint8_t* cl = reinterpret_cast<uint8_t*>(T.lit);
                        br.refill();
                        nlit = int(br.take(5)) + 257;
                        ndist = int(br.take(5)) + 1;
                        const int ncl = int(br.take(4)) + 4;
                        uint8_t cll[19];
#pragma unroll
                        for (int i = 0; i < 19; ++i) cll[i] = 0;
                        for (int i = 0; i < ncl; ++i) { br.refill(); cll[kClOrder[i]] = uint8_t(br.take(3)); }
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
                        int i = 0;
                        const int total = nlit + ndist;
                        if (nlit > 286 || ndist > 30) err = 1;
                        while (i < total && !err) {
                            br.refill();
                            const uint32_t e = cl[br.peek(7)];
                            const int l = int(e >> 5), s = int(e & 31);
                            if (!l) { err = 1; break; }
                            br.drop(uint32_t(l));
                            if (s < 16) { T.lens[i++] = uint8_t(s); continue; }
                            int rep, val = 0;
                            if (s == 16) { if (i == 0) { err = 1; break; } val = T.lens[i - 1]; rep = 3 + int(br.take(2)); }
                            else if (s == 17) rep = 3 + int(br.take(3));
                            else rep = 11 + int(br.take(7));
                            if (i + rep > total) { err = 1; break; }
                            while (rep--) T.lens[i++] = uint8_t(val);
                        }
                        if (!err && T.lens[256] == 0) err = 1;
                        if (!err) {
                            for (int k = ndist - 1; k >= 0; --k) T.lens[288 + k] = T.lens[nlit + k];
                            for (int k = nlit; k < 288; ++k) T.lens[k] = 0;
                        }
                    }
                    if (!err) {
                        bool ok = build_serial(T.lens, 288, T.lit, kLitBits, T.lit_count, T.lit_sorted);
                        ok = build_serial(T.lens + 288, ndist, T.dist, kDistBits, T.dist_count, T.dist_sorted) && ok;
                        if (!ok) err = 1;
                    }
                    if (err) state = 2; else phase = 1;
                }
            }
            if (phase == 1 && state == 0) {
                // ---- phase 1: fill the token queue ----------------------------------------------------------------------
                while (nq < kQueue) {
                    br.refill();
                    uint32_t e = T.lit[br.peek(kLitBits)];
                    int sym;
                    if (e) { br.drop(e >> 9); sym = int(e & 511u); }
                    else { sym = slow_decode_nv(br, T.lit_count, T.lit_sorted); if (sym < 0) { state = 2; break; } }
                    if (sym < 256) { q[nq++] = uint32_t(sym); ++op_dec; continue; }
                    if (sym == 256) { phase = 0; if (last) state = 1; break; }
                    if (sym > 285) { state = 2; break; }
                    uint32_t mlen;
                    if (sym < 265) mlen = uint32_t(sym - 254);
                    else if (sym == 285) mlen = 258;
                    else {
                        const uint32_t eb = uint32_t(sym - 261) >> 2;
                        mlen = ((4u + (uint32_t(sym - 265) & 3u)) << eb) + 3u + br.take(eb);
                    }
                    br.refill();
                    e = T.dist[br.peek(kDistBits)];
                    int ds;
                    if (e) { br.drop(e >> 9); ds = int(e & 511u); }
                    else { ds = slow_decode_nv(br, T.dist_count, T.dist_sorted); if (ds < 0) { state = 2; break; } }
                    if (ds > 29) { state = 2; break; }
                    uint32_t mdist;
                    if (ds < 4) mdist = uint32_t(ds + 1);
                    else {
                        const uint32_t eb = (uint32_t(ds) >> 1) - 1u;
                        br.refill();
                        mdist = ((2u + (uint32_t(ds) & 1u)) << eb) + 1u + br.take(eb);
                    }
                    if (mdist > op_dec) { state = 2; break; }
                    q[nq++] = 0x80000000u | ((mdist - 1u) << 16) | mlen;
                    op_dec += mlen;
                }
                if (op_dec > out_len || br.overrun()) state = 2;
            }
        }
        nq = __shfl_sync(FULL, nq, 0);
        state = __shfl_sync(FULL, state, 0);
        if (state == 2) break;
        __syncwarp();
        // ---- phase 2: materialise the queue -----------------------------------------------------------------------------------
        for (int base = 0; base < nq; base += 32) {
            const bool valid = base + lane < nq;
            const uint32_t t = valid ? q[base + lane] : 0u;
            const bool is_match = valid && (t >> 31);
            const bool is_skip = valid && !is_match && (t & 0x40000000u);
            const uint32_t len = !valid ? 0u : (is_match ? (t & 0x1ffu) : (is_skip ? (t & 0xffffffu) : 1u));
            uint32_t incl = len;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t up = __shfl_up_sync(FULL, incl, o);
                if (lane >= o) incl += up;
            }
            const uint32_t pos = pos_base + incl - len;
            if (valid && !is_match && !is_skip) out[pos] = uint8_t(t);
            uint32_t mm = __ballot_sync(FULL, is_match);
            __syncwarp();
            while (mm) {
                const int k = __ffs(mm) - 1;
                mm &= mm - 1;
                const uint32_t mt = __shfl_sync(FULL, t, k);
                const uint32_t mpos = __shfl_sync(FULL, pos, k);
                const uint32_t mlen = mt & 0x1ffu, mdist = ((mt >> 16) & 0x7fffu) + 1u;
                const uint8_t* src = out + mpos - mdist;
                if (mdist >= mlen) {
                    for (uint32_t j = lane; j < mlen; j += 32) out[mpos + j] = src[j];
                } else {
                    for (uint32_t j = lane; j < mlen; j += 32) out[mpos + j] = src[j % mdist];
                }
                __syncwarp();
            }
            pos_base += __shfl_sync(FULL, incl, 31);
        }
        __syncwarp();
        if (state == 1) break;
    }
    const int st = __shfl_sync(FULL, 0, 0);
    (void)st;
    const uint32_t final_dec = __shfl_sync(FULL, op_dec, 0);
    if (lane == 0) {
        // state 2 leaves the loop early; a clean finish must have produced exactly ISIZE bytes
        if (final_dec != out_len || pos_base != out_len) atomicOr(&sc->status, STATUS_BAD_DEFLATE);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// k_inflate_q2: the production inflate kernel.  Same two-phase structure as k_inflate_q, on the lean decode core of
// inflate_core.cuh (32-bit look-ahead bit reader, one table lookup per code with base/extra-bits packed in the
// entry, sticky error flag instead of divergent breaks) and a cheaper materialisation step (one predicated
// load/store per <= 32-byte match; overlapping matches use the period trick instead of a modulo per byte).
// The decode core is unit-tested on the host (tests/host_inflate_harness.cpp) against zlib's CRC32.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kQ2Warps = 4;
struct Q2Smem {
    inflate_core::Tables T;
    uint32_t q[inflate_core::kQueue];
};

__global__ void __launch_bounds__(kQ2Warps * 32, 7) k_inflate_q2(const InflateBlock* __restrict__ blocks, int n_blocks,
                                                                  const uint8_t* __restrict__ comp, uint8_t* raw, DeviceScalars* sc) {
    using namespace inflate_core;
    __shared__ Q2Smem s_mem[kQ2Warps];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * kQ2Warps + wid;
    if (b >= n_blocks) return;
    Tables& T = s_mem[wid].T;
    volatile uint32_t* q = s_mem[wid].q;
    const InflateBlock blk = blocks[b];
    uint8_t* out = raw + blk.out_off;
    const uint32_t out_len = blk.out_len;
    inflate_core::BitReader br;
    br.wp = nullptr; br.p0 = nullptr; br.w0 = br.w1 = br.w2 = br.bo = br.first_bit = 0;
    if (lane == 0) br.init(comp + blk.in_off);
    uint32_t op_dec = 0, pos_base = 0;
    int phase = 0, last = 0;
    uint64_t stored_bits = 0;       // bits consumed before the last reader re-init (stored blocks)
    for (;;) {
        int nq = 0, state = 0;       // state: 0 = go on, 1 = stream finished, 2 = error
        if (lane == 0) {
            if (phase == 0) {
                const int h = read_block_header(br, T, &last);
                if (h == 2) state = 2;
                else if (h == 1) {
                    br.consume((32u - br.bo) & 7u);
                    const uint32_t v = br.peek();
                    br.consume(32);
                    const uint32_t len = v & 0xffffu, nlen = v >> 16;
                    if ((len ^ nlen) != 0xffffu || op_dec + len > out_len) state = 2;
                    else {
                        const uint8_t* src = br.byte_ptr();
                        for (uint32_t j = 0; j < len; ++j) out[op_dec + j] = src[j];
                        op_dec += len;
                        stored_bits += br.bits_used() + uint64_t(len) * 8;
                        br.init(src + len);
                        q[0] = kTokSkip | len;
                        nq = 1;
                        if (last) state = 1;
                    }
                } else phase = 1;
            }
            if (phase == 1 && nq == 0 && state == 0) {
                int eob = 0, bad = 0;
                nq = fill_queue(br, T, q, &op_dec, &eob, &bad);
                if (eob) { phase = 0; if (last) state = 1; }
                if (bad || op_dec > out_len || stored_bits + br.bits_used() > uint64_t(blk.in_len) * 8) state = 2;
            }
        }
        nq = __shfl_sync(FULL, nq, 0);
        state = __shfl_sync(FULL, state, 0);
        if (state == 2) break;
        __syncwarp();
        // ---- phase 2: materialise the queue -----------------------------------------------------------------------------------
        for (int base = 0; base < nq; base += 32) {
            const bool valid = base + lane < nq;
            const uint32_t t = valid ? q[base + lane] : 0u;
            const bool is_match = (t >> 31) != 0;
            const bool is_skip = !is_match && (t & kTokSkip);
            const uint32_t len = is_match ? (t & 0x1ffu) : (is_skip ? (t & 0xffffffu) : (valid ? 1u : 0u));
            uint32_t incl = len;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t up = __shfl_up_sync(FULL, incl, o);
                if (lane >= o) incl += up;
            }
            const uint32_t pos = pos_base + incl - len;
            if (valid && !is_match && !is_skip) out[pos] = uint8_t(t);
            uint32_t mm = __ballot_sync(FULL, is_match);
            __syncwarp();
            while (mm) {
                const int k = __ffs(mm) - 1;
                mm &= mm - 1;
                const uint32_t mt = __shfl_sync(FULL, t, k);
                const uint32_t mpos = __shfl_sync(FULL, pos, k);
                const uint32_t mlen = mt & 0x1ffu, mdist = ((mt >> 16) & 0x7fffu) + 1u;
                uint8_t* dst = out + mpos;
                if (uint32_t(lane) < mlen) {
                    // the first 32 bytes never depend on bytes written in this step
                    const int so = (mdist >= 32u || mdist >= mlen) ? int(lane) - int(mdist) : int(uint32_t(lane) % mdist) - int(mdist);
                    dst[lane] = dst[so];
                }
                if (mlen > 32u) {
                    // later steps read one whole period (>= 32 bytes) back: already written, barrier between steps
                    const uint32_t K = mdist >= 32u ? mdist : mdist * (31u / mdist + 1u);
                    for (uint32_t j = 32u + lane; j - lane < mlen; j += 32u) {
                        __syncwarp();
                        if (j < mlen) dst[j] = dst[int(j) - int(K)];
                    }
                }
                __syncwarp();
            }
            pos_base += __shfl_sync(FULL, incl, 31);
        }
        __syncwarp();
        if (state == 1) break;
    }
    const uint32_t final_dec = __shfl_sync(FULL, op_dec, 0);
    if (lane == 0 && (final_dec != out_len || pos_base != out_len)) atomicOr(&sc->status, STATUS_BAD_DEFLATE);
}

// ---------------------------------------------------------------------------------------------------------------
// k_crc32: BGZF integrity check on the device (htslib verifies the CRC32 of every inflated block; so do we).
// One thread per block, slicing-by-4 over the block's bytes with the four 256-entry tables in shared memory.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_crc32(const InflateBlock* __restrict__ blocks, const uint32_t* __restrict__ want, int n_blocks,
                                               const uint8_t* __restrict__ raw, DeviceScalars* sc) {
    __shared__ uint32_t tab[4][256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        uint32_t c = uint32_t(i);
        for (int k = 0; k < 8; ++k) c = (c & 1u) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
        tab[0][i] = c;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        uint32_t c = tab[0][i];
        for (int t = 1; t < 4; ++t) { c = tab[0][c & 0xffu] ^ (c >> 8); tab[t][i] = c; }
    }
    __syncthreads();
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_blocks) return;
    const InflateBlock blk = blocks[b];
    const uint8_t* p = raw + blk.out_off;
    uint32_t n = blk.out_len, crc = 0xffffffffu;
    while (n && (reinterpret_cast<uintptr_t>(p) & 3)) { crc = tab[0][(crc ^ *p++) & 0xffu] ^ (crc >> 8); --n; }
    const uint32_t* w = reinterpret_cast<const uint32_t*>(p);
    for (; n >= 4; n -= 4) {
        crc ^= *w++;
        crc = tab[3][crc & 0xffu] ^ tab[2][(crc >> 8) & 0xffu] ^ tab[1][(crc >> 16) & 0xffu] ^ tab[0][crc >> 24];
    }
    p = reinterpret_cast<const uint8_t*>(w);
    while (n--) crc = tab[0][(crc ^ *p++) & 0xffu] ^ (crc >> 8);
    if (~crc != want[b]) atomicOr(&sc->status, STATUS_BAD_CRC);
}

// ---------------------------------------------------------------------------------------------------------------
// record-boundary walk
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ld_u32_any(const uint8_t* raw, uint32_t off) {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(raw + (off & ~3u));
    const uint32_t sh = (off & 3u) * 8;
    const uint32_t lo = w[0];
    const uint32_t hi = sh ? w[1] : 0u;
    return __funnelshift_r(lo, hi, sh);
}

template <bool WRITE>
__global__ void __launch_bounds__(128) k_walk(const uint8_t* __restrict__ raw, const uint2* __restrict__ walkers, int n_walkers,
                                              uint32_t* __restrict__ counts, const uint32_t* __restrict__ base,
                                              uint32_t* __restrict__ offs, DeviceScalars* sc) {
    const int w = blockIdx.x * 128 + threadIdx.x;
    if (w >= n_walkers) return;
    const uint2 wk = walkers[w];
    uint32_t p = wk.x, n = 0;
    const uint32_t e = wk.y;
    uint32_t* o = WRITE ? offs + base[w] : nullptr;
    bool bad = false;
    while (p < e) {
        if (p + 4 > e) { bad = true; break; }
        const uint32_t bs = ld_u32_any(raw, p);
        if (bs < 32u || bs > 0x7fffffffu || uint64_t(p) + 4 + bs > e) { bad = true; break; }
        if (WRITE) o[n] = p;
        ++n;
        p += 4 + bs;
    }
    if (!WRITE) counts[w] = n;
    if (bad) atomicOr(&sc->status, STATUS_CORRUPT);
}

// exclusive scan of counts[0..n) into base[0..n), total into *total and (when offs != null) the end sentinel
__global__ void __launch_bounds__(1024) k_scan_counts(const uint32_t* __restrict__ counts, int n, uint32_t* __restrict__ base,
                                                      uint32_t* __restrict__ total, uint32_t* __restrict__ offs, uint32_t end_pos) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int i0 = 0; i0 < n; i0 += 1024) {
        const int i = i0 + threadIdx.x;
        const uint32_t v = i < n ? counts[i] : 0u;
        uint32_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t up = __shfl_up_sync(FULL, incl, o);
            if (lane >= o) incl += up;
        }
        if (lane == 31) s_warp[wid] = incl;
        __syncthreads();
        uint32_t pre = s_carry + incl - v;
        for (int k = 0; k < wid; ++k) pre += s_warp[k];
        if (i < n) base[i] = pre;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = pre + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        *total = s_carry;
        if (offs) offs[s_carry] = end_pos;
    }
}

}  // namespace

template <int D, int LB, int DB, int WARPS>
static void launch_ms(const InflateBlock* d_blocks, int n_blocks, const uint8_t* d_comp, uint8_t* d_raw, DeviceScalars* sc,
                      cudaStream_t s) {
    const size_t smem = sizeof(StreamTables<LB, DB>) * D * WARPS;
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(k_inflate_ms<D, LB, DB, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
        configured = true;
    }
    const int per_cta = D * WARPS;
    const int grid = (n_blocks + per_cta - 1) / per_cta;
    k_inflate_ms<D, LB, DB, WARPS><<<grid, WARPS * 32, smem, s>>>(d_blocks, n_blocks, d_comp, d_raw, sc);
}

void launch_inflate(const InflateBlock* d_blocks, int n_blocks, const uint8_t* d_comp, uint8_t* d_raw, DeviceScalars* sc,
                    cudaStream_t s) {
    if (n_blocks <= 0) return;
    static const int variant = getenv("BSG_INFLATE_KERNEL") ? atoi(getenv("BSG_INFLATE_KERNEL")) : 3;
    if (variant == 3) {
        const int grid = (n_blocks + kQ2Warps - 1) / kQ2Warps;
        k_inflate_q2<<<grid, kQ2Warps * 32, 0, s>>>(d_blocks, n_blocks, d_comp, d_raw, sc);
    } else if (variant == 2) {
        const int grid = (n_blocks + kQWarps - 1) / kQWarps;
        k_inflate_q<<<grid, kQWarps * 32, 0, s>>>(d_blocks, n_blocks, d_comp, d_raw, sc);
    } else if (variant == 1) {
        const int grid = (n_blocks + kWarpsPerCta - 1) / kWarpsPerCta;
        k_inflate<<<grid, kWarpsPerCta * 32, 0, s>>>(d_blocks, n_blocks, d_comp, d_raw, sc);
    } else if (variant == 4 || (variant == 8 && n_blocks < 4096)) {
        launch_ms<4, 10, 8, 2>(d_blocks, n_blocks, d_comp, d_raw, sc, s);
    } else if (variant == 16) {
        launch_ms<16, 9, 7, 1>(d_blocks, n_blocks, d_comp, d_raw, sc, s);
    } else if (variant == 89) {
        launch_ms<8, 9, 7, 2>(d_blocks, n_blocks, d_comp, d_raw, sc, s);
    } else if (variant == 49) {
        launch_ms<4, 9, 7, 4>(d_blocks, n_blocks, d_comp, d_raw, sc, s);
    } else {
        launch_ms<8, 10, 8, 2>(d_blocks, n_blocks, d_comp, d_raw, sc, s);
    }
}

void launch_crc32(const InflateBlock* d_blocks, const uint32_t* d_crc, int n_blocks, const uint8_t* d_raw, DeviceScalars* sc,
                  cudaStream_t s) {
    if (n_blocks <= 0) return;
    k_crc32<<<(n_blocks + 127) / 128, 128, 0, s>>>(d_blocks, d_crc, n_blocks, d_raw, sc);
}

void launch_walk(const uint8_t* d_raw, const uint2* d_walkers, int n_walkers, uint32_t* d_counts, uint32_t* d_base,
                 uint32_t* d_total, uint32_t* d_offs, uint32_t end_pos, DeviceScalars* sc, cudaStream_t s) {
    if (n_walkers <= 0) {
        k_scan_counts<<<1, 1024, 0, s>>>(d_counts, 0, d_base, d_total, d_offs, end_pos);
        return;
    }
    const int grid = (n_walkers + 127) / 128;
    k_walk<false><<<grid, 128, 0, s>>>(d_raw, d_walkers, n_walkers, d_counts, nullptr, nullptr, sc);
    k_scan_counts<<<1, 1024, 0, s>>>(d_counts, n_walkers, d_base, d_total, d_offs, end_pos);
    k_walk<true><<<grid, 128, 0, s>>>(d_raw, d_walkers, n_walkers, nullptr, d_base, d_offs, sc);
}

}  // namespace bsg
