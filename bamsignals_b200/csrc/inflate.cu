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

void launch_inflate(const InflateBlock* d_blocks, int n_blocks, const uint8_t* d_comp, uint8_t* d_raw, DeviceScalars* sc,
                    cudaStream_t s) {
    if (n_blocks <= 0) return;
    const int grid = (n_blocks + kWarpsPerCta - 1) / kWarpsPerCta;
    k_inflate<<<grid, kWarpsPerCta * 32, 0, s>>>(d_blocks, n_blocks, d_comp, d_raw, sc);
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
