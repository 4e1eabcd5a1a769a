// Hand-written sm_100a kernels of the counting path.  Integer / byte work bounded by HBM bandwidth and shared-memory
// atomics: no tensor cores by design (nothing here is a dense contraction).  See kernels.cuh for the data layout and
// DESIGN.md for the per-kernel algorithmic bytes.
#include "kernels.cuh"

#include <climits>

namespace bsg {
namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int kThreads = 256;

__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

// ---------------------------------------------------------------------------------------------------------------
// K1 decode + filter.  One CTA per chunk of 256 consecutive records, one thread per record.
//
// Records start at arbitrary byte offsets, so a per-thread gather from global memory makes every load instruction
// touch a dozen 128-byte lines (measured: 15 % of HBM peak).  Instead the CTA's records are one CONTIGUOUS byte span
// [offs[i0], offs[i0+256]) of the batch: thread 0 issues a single TMA bulk copy (cp.async.bulk, completion on an
// mbarrier) of that span into shared memory, so DRAM is read in full lines exactly once, and the threads then pick
// their fields out of shared memory with aligned 32-bit loads + funnel shifts.  Eight CTAs are resident per SM, so
// the copy of one chunk overlaps the decode and the coalesced SoA stores of the others.  Spans that do not fit the
// shared buffer (records with long SEQ/QUAL: only ~60 of their bytes are needed) take the direct global-load path.
//
// Fixed part of a record (offsets from the block_size field): refID +4, pos +8, l_read_name +12, mapq +13,
// bin +14, n_cigar_op +16, flag +18, l_seq +20, next_refID +24, next_pos +28, tlen +32, read_name +36, cigar after.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kDecThreads = kDecodeChunk;
constexpr int kSpanCap = 96 * kDecThreads;   // bytes of shared staging per CTA (average record <= 96 B)

struct GlobalLd {
    const uint8_t* raw;
    __device__ __forceinline__ uint32_t operator()(uint32_t aligned_off) const {
        return __ldg(reinterpret_cast<const uint32_t*>(raw + aligned_off));
    }
};
struct SharedLd {
    const uint8_t* sm;
    uint32_t lo;
    __device__ __forceinline__ uint32_t operator()(uint32_t aligned_off) const {
        return *reinterpret_cast<const uint32_t*>(sm + (aligned_off - lo));
    }
};

struct Decoded {
    int32_t tid, pos, end, tlen;
    uint32_t flagmq, bad;
};

template <class LD>
__device__ __forceinline__ Decoded decode_one(const LD& ld, uint32_t o, uint32_t rec_end) {
    Decoded d;
    const uint32_t a = o + 4, b = a & ~3u, sh = (a & 3u) * 8;
    const uint32_t w0 = ld(b), w1 = ld(b + 4), w2 = ld(b + 8), w3 = ld(b + 12), w4 = ld(b + 16);
    const uint32_t w7 = ld(b + 28), w8 = ld(b + 32);
    d.tid = int32_t(__funnelshift_r(w0, w1, sh));
    d.pos = int32_t(__funnelshift_r(w1, w2, sh));
    const uint32_t nm = __funnelshift_r(w2, w3, sh);   // l_read_name | mapq << 8 | bin << 16
    const uint32_t cf = __funnelshift_r(w3, w4, sh);   // n_cigar_op | flag << 16
    d.tlen = int32_t(__funnelshift_r(w7, w8, sh));
    const uint32_t l_name = nm & 0xffu, mapq = (nm >> 8) & 0xffu;
    const uint32_t n_cigar = cf & 0xffffu, flag = cf >> 16;
    const uint32_t c = o + 36 + l_name;
    uint32_t rlen = 0;
    d.bad = 0;
    if (c + 4u * n_cigar > rec_end) {
        d.bad = STATUS_CORRUPT;
    } else if (!(flag & 0x4u) && n_cigar) {             // bam_endpos: unmapped reads have no reference length
        // nine reads in ten have a one-operation CIGAR: that case is straight-line code; the rest take a plain loop
        // (unrolled by eight, the loop's remainder ladder cost a one-op read 45 instructions)
        const uint32_t cb = c & ~3u, csh = (c & 3u) * 8;
        uint32_t lo = ld(cb), hi = ld(cb + 4);
        uint32_t op = __funnelshift_r(lo, hi, csh);
        rlen = ((0x18Du >> (op & 0xfu)) & 1u) ? op >> 4 : 0u;    // M, D, N, =, X consume the reference
#pragma unroll 1
        for (uint32_t k = 1; k < n_cigar; ++k) {
            lo = hi;
            hi = ld(cb + 4 * k + 4);
            op = __funnelshift_r(lo, hi, csh);
            if ((0x18Du >> (op & 0xfu)) & 1u) rlen += op >> 4;
        }
    }
    if (rlen == 0) rlen = 1;
    d.end = int32_t(uint32_t(d.pos) + rlen - 1u);
    d.flagmq = flag | (mapq << 16);
    return d;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }

// setRead's filter (src/bamsignals.cpp:328-333, :394-399)
__device__ __forceinline__ bool keep_read(const FilterParams& p, uint32_t fm, int32_t tlen, int32_t* abs_tlen) {
    const uint32_t flag = fm & 0xffffu, mapq = fm >> 16;
    const int32_t a = tlen < 0 ? -tlen : tlen;   // abs() as the reference computes it (INT_MIN stays negative)
    *abs_tlen = a;
    bool keep = int32_t(mapq) >= p.mapqual;
    keep = keep && !(p.required & ~flag);          // every required bit present   (src/bamsignals.cpp:21-23, :328)
    keep = keep && (p.filtered & ~flag);           // not all filtered bits present (:329); filtered == 0 drops all
    keep = keep && (!p.have_tlen || (a >= p.tmin && a <= p.tmax));
    return keep;
}

// K1+K2 fused: the filter and the pileup / coverage coordinate (setRead, :326-346 and :392-415) are applied to the
// decoded fields while they are still in registers, so a read costs its raw bytes in and 16 bytes out
// (tid, pos for the join; c0, c1 for the counting kernels) instead of a 20-byte table row written, read again and
// turned into 8 more bytes by a second kernel.
template <bool COVERAGE>
__global__ void __launch_bounds__(kDecThreads) k_decode(DecodeBatch single, const DecodeBatch* __restrict__ table,
                                                     int n_batches, ReadTable t, FilterParams p, DeviceScalars* sc) {
    extern __shared__ __align__(128) uint8_t sm_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t s_lo, s_bytes;
    __shared__ int s_hlo, s_hhi, s_ghlo, s_ghhi;
    __shared__ int2 s_last[kDecThreads / 32];         // (tid, pos) of every warp's last record
    // which batch does this chunk belong to?
    DecodeBatch B = single;
    int bi = 0;
    if (table) {
        int lo = 0, hi = n_batches - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (int(blockIdx.x) >= table[mid].chunk0) lo = mid; else hi = mid - 1;
        }
        B = table[lo];
        bi = lo;
    }
    const int i0 = (int(blockIdx.x) - B.chunk0) * kDecThreads;
    const int i = i0 + threadIdx.x;
    const bool active = i < B.n;
    if (threadIdx.x == 0) {
        const int i1 = min(i0 + kDecThreads, B.n);
        const uint32_t o_lo = B.offs[i0], o_hi = B.offs[i1];
        const uint32_t lo = o_lo & ~15u;
        const uint32_t bytes = (o_hi - lo + 15u) & ~15u;
        if (bytes <= uint32_t(kSpanCap)) {
            const uint32_t bar_a = smem_u32(&bar);
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(sm_raw)), "l"(B.raw + lo), "r"(bytes), "r"(bar_a) : "memory");
            s_lo = lo; s_bytes = bytes;
        } else {
            s_lo = 0; s_bytes = 0;
        }
        s_hlo = INT_MIN; s_hhi = INT_MIN;
        // the running halo maxima, read ONCE per CTA (L2, in flight during the copy) and handed to the warps through
        // shared memory: a warp whose own maxima do not beat them has nothing to publish, and a stale value only costs
        // a redundant atomicMax.  (Every warp reading them itself put 3.6 M loads on one L2 line: +0.35 ms.)
        const int2 g = __ldcg(reinterpret_cast<const int2*>(&sc->halo_lo));
        s_ghlo = g.x; s_ghhi = g.y;
    }
    uint32_t o = 0, rec_end = 0;
    if (active) { o = __ldg(B.offs + i); rec_end = __ldg(B.offs + i + 1); }   // in flight while the bulk copy lands
    const int lane = threadIdx.x & 31;
    __syncthreads();
    const bool staged = s_bytes != 0;
    if (staged) {
        const uint32_t bar_a = smem_u32(&bar);
        uint32_t done = 0;
        while (!done)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                         : "=r"(done) : "r"(bar_a) : "memory");
    }
    Decoded d;
    d.tid = -1; d.pos = -1; d.end = -1; d.tlen = 0; d.flagmq = 0; d.bad = 0;
    const SharedLd sld{sm_raw, s_lo};
    const GlobalLd gld{B.raw};
    int hlo = INT_MIN, hhi = INT_MIN;
    bool keep = false;
    if (active) {
        d = staged ? decode_one(sld, o, rec_end) : decode_one(gld, o, rec_end);
        int32_t a;
        keep = keep_read(p, d.flagmq, d.tlen, &a);
        const bool neg = (d.flagmq & 0x10u) != 0;
        int32_t o0, o1;
        if (!COVERAGE) {
            const int32_t offset = p.midpoint ? a / 2 + p.shift : p.shift;               // :339
            const int32_t pos5 = neg ? d.end - offset : d.pos + offset;                  // :340-344
            o0 = pos5;
            o1 = keep ? int32_t(neg) : -1;
            if (keep) { hlo = pos5 - d.pos; hhi = d.pos - pos5; }
        } else {
            int32_t s = d.pos, e = d.end;                                                 // :401-403
            if (p.tspan) {
                if (neg && d.tlen < 0) s = e + d.tlen + 1;                                // :408-409
                else if (!neg && d.tlen > 0) e = s + d.tlen - 1;                          // :410-411
            }
            o0 = keep ? s : INT_MAX;
            o1 = keep ? e : INT_MIN;
            if (keep) { hlo = e - d.pos; hhi = d.pos - s; }
        }
        const int64_t row = B.row0 + i;
        t.tid[row] = d.tid;
        t.pos[row] = d.pos;
        t.c0[row] = o0;
        t.c1[row] = o1;
    }
    // One barrier serves the cross-warp sortedness check and the CTA-wide statistics: before it, every warp leaves the
    // key of its last record and its halo maxima in shared memory; the barrier itself counts the kept reads.
    const int wid = threadIdx.x >> 5;
    hlo = __reduce_max_sync(FULL, hlo);
    hhi = __reduce_max_sync(FULL, hhi);
    if (lane == 31) s_last[wid] = make_int2(d.tid, d.pos);
    if (lane == 0 && hlo != INT_MIN) { atomicMax(&s_hlo, hlo); atomicMax(&s_hhi, hhi); }
    uint32_t ptid = __shfl_up_sync(FULL, uint32_t(d.tid), 1);
    int32_t ppos = __shfl_up_sync(FULL, d.pos, 1);
    const int kept_cta = __syncthreads_count(keep);
    // coordinate-sortedness: compare with the previous record (previous lane; lane 0 takes the previous warp's last
    // record; thread 0 re-reads the record before the chunk)
    bool have_prev = active;
    if (lane == 0 && active) {
        if (threadIdx.x > 0) {
            const int2 l = s_last[wid - 1];
            ptid = uint32_t(l.x); ppos = l.y;
        } else if (i > 0) {
            const uint32_t a = __ldg(B.offs + i - 1) + 4, b = a & ~3u, sh = (a & 3u) * 8;
            const uint32_t w0 = gld(b), w1 = gld(b + 4), w2 = gld(b + 8);
            ptid = __funnelshift_r(w0, w1, sh);
            ppos = int32_t(__funnelshift_r(w1, w2, sh));
        } else if (table && bi > 0) {
            // first record of a resident batch: the record before it is the last one of the batch before, which this
            // same launch decodes (its table row may not be written yet): read it from the raw bytes
            const DecodeBatch Bp = table[bi - 1];
            const GlobalLd pld{Bp.raw};
            const uint32_t a = __ldg(Bp.offs + Bp.n - 1) + 4, b = a & ~3u, sh = (a & 3u) * 8;
            const uint32_t w0 = pld(b), w1 = pld(b + 4), w2 = pld(b + 8);
            ptid = __funnelshift_r(w0, w1, sh);
            ppos = int32_t(__funnelshift_r(w1, w2, sh));
        } else if (B.row0 > 0) {          // streaming path: the row was written by an earlier launch on this stream
            ptid = uint32_t(t.tid[B.row0 - 1]);
            ppos = t.pos[B.row0 - 1];
        } else {
            have_prev = false;
        }
    }
    uint32_t bad = d.bad;
    if (have_prev && (uint32_t(d.tid) < ptid || (uint32_t(d.tid) == ptid && d.pos < ppos))) bad |= STATUS_UNSORTED;
    if (bad) atomicOr(&sc->status, bad);
    // halo maxima and the kept-read count: published once per CTA, the maxima only when they beat the value read at the
    // start, the count into one of kKeptSlots counters that sit on different 128-byte lines (atomics on one line
    // serialise in its L2 slice: 1.8 M same-line atomics cost this kernel 3.7 ms)
    if (threadIdx.x == 0 && kept_cta) {
        if (s_hlo > s_ghlo) atomicMax(&sc->halo_lo, s_hlo);
        if (s_hhi > s_ghhi) atomicMax(&sc->halo_hi, s_hhi);
        atomicAdd(&sc->kept[(blockIdx.x & (kKeptSlots - 1)) * kKeptStride], (unsigned long long)kept_cta);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// K3 region join: two binary searches per tile over the (tid,pos)-sorted table.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int64_t lower_bound_tp(const int32_t* __restrict__ tid, const int32_t* __restrict__ pos,
                                                  int64_t n, uint32_t ktid, int64_t kpos) {
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        const uint32_t t = uint32_t(__ldg(tid + mid));
        const int64_t p = __ldg(pos + mid);
        const bool less = t < ktid || (t == ktid && p < kpos);
        if (less) lo = mid + 1; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(kThreads) k_join(ReadTable t, int64_t n, TileTable tiles, int64_t n_tiles,
                                                   const DeviceScalars* sc_in, DeviceScalars* sc) {
    const int64_t i = int64_t(blockIdx.x) * kThreads + threadIdx.x;
    long long cand = 0;
    if (i < n_tiles) {
        const int64_t halo_lo = sc_in->halo_lo, halo_hi = sc_in->halo_hi;
        const uint32_t rid = uint32_t(tiles.rid[i]);
        const int64_t loc = tiles.loc[i], len = tiles.len[i];
        const int64_t lo = lower_bound_tp(t.tid, t.pos, n, rid, loc - halo_lo);
        const int64_t hi = lower_bound_tp(t.tid, t.pos, n, rid, loc + len + halo_hi);
        tiles.cand_lo[i] = lo;
        tiles.cand_hi[i] = hi > lo ? hi : lo;
        cand = hi > lo ? hi - lo : 0;
    }
    // block reduce the candidate count (diagnostic: C in the 8*C + 4*B algorithmic-byte formula)
    __shared__ long long s_c[kThreads / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cand += __shfl_xor_sync(FULL, cand, o);
    if ((threadIdx.x & 31) == 0) s_c[threadIdx.x >> 5] = cand;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long tot = 0;
        for (int k = 0; k < kThreads / 32; ++k) tot += s_c[k];
        if (tot) atomicAdd(&sc->candidates, (unsigned long long)tot);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// K4 bamCount: one warp per region, per-lane counters, no atomics.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_count(TileTable tiles, int64_t n_tiles, const int32_t* __restrict__ c0,
                                                    const int32_t* __restrict__ c1, int ss, int32_t* __restrict__ out) {
    const int64_t w = (int64_t(blockIdx.x) * kThreads + threadIdx.x) >> 5;
    if (w >= n_tiles) return;
    const int lane = threadIdx.x & 31;
    const int64_t lo = tiles.cand_lo[w], hi = tiles.cand_hi[w];
    const int32_t loc = tiles.loc[w];
    const uint32_t len = uint32_t(tiles.len[w]);
    const int flip = tiles.strand[w] < 0 ? 1 : 0;
    int sense = 0, anti = 0;
    for (int64_t i = lo + lane; i < hi; i += 32) {
        const int32_t p5 = __ldg(c0 + i), ng = __ldg(c1 + i);
        const bool ok = ng >= 0 && uint32_t(p5 - loc) < len;                 // src/bamsignals.cpp:351-353
        const int a = (ng ^ flip) & 1;                                        // :355-359
        sense += ok && !a;
        anti += ok && a;
    }
    sense = warp_sum(sense);
    anti = warp_sum(anti);
    if (lane == 0) {
        const int64_t off = tiles.out_off[w];
        if (ss) { out[off] = sense; out[off + 1] = anti; } else { out[off] = sense + anti; }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// shared -> global write-out of one tile.  `sm` was placed so that sm and dst have the same phase modulo four
// ints, which lets the body move as 128-bit shared loads + 128-bit streaming stores.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void copy_out(int32_t* __restrict__ dst, const int32_t* sm, int n) {
    const int head = min(n, int((4 - ((reinterpret_cast<uintptr_t>(dst) >> 2) & 3)) & 3));
    if (int(threadIdx.x) < head) __stcs(dst + threadIdx.x, sm[threadIdx.x]);
    const int body = (n - head) >> 2;
    int4* d4 = reinterpret_cast<int4*>(dst + head);
    const int4* s4 = reinterpret_cast<const int4*>(sm + head);
    for (int k = threadIdx.x; k < body; k += blockDim.x) __stcs(d4 + k, s4[k]);
    for (int k = head + 4 * body + threadIdx.x; k < n; k += blockDim.x) __stcs(dst + k, sm[k]);
}

// exact floor(n / d) for 0 <= n < 2^31 with a precomputed multiplier (host: magic_for())
__device__ __forceinline__ uint32_t fast_div(uint32_t n, uint64_t magic, uint32_t shift) {
    return uint32_t((uint64_t(n) * magic) >> shift);
}

// ---------------------------------------------------------------------------------------------------------------
// K4 bamProfile: one CTA per tile (<= kTileInts output ints), shared-memory histogram.
// k_profile: one read per lane and one shared atomic per counted read (binsize 1..3: neighbouring reads fall into
// different bins anyway).
// ---------------------------------------------------------------------------------------------------------------
template <bool SS, bool BIN1>
__global__ void __launch_bounds__(kThreads) k_profile(TileTable tiles, const int32_t* __restrict__ c0,
                                                      const int32_t* __restrict__ c1, int32_t binsize, uint64_t magic,
                                                      uint32_t mshift, int32_t* __restrict__ out) {
    extern __shared__ __align__(16) int32_t smem[];
    const int64_t tix = blockIdx.x;
    const int32_t loc = tiles.loc[tix], len = tiles.len[tix];
    const int flip = tiles.strand[tix] < 0 ? 1 : 0;
    const int64_t lo = tiles.cand_lo[tix], hi = tiles.cand_hi[tix];
    const int64_t off = tiles.out_off[tix];
    const int nbins = BIN1 ? len : int((uint32_t(len) + uint32_t(binsize) - 1u) / uint32_t(binsize));
    const int nout = SS ? 2 * nbins : nbins;
    int32_t* hist = smem + (off & 3);                      // same 16-byte phase as the destination
    for (int k = threadIdx.x; k < nout; k += kThreads) hist[k] = 0;
    __syncthreads();
    for (int64_t i = lo + threadIdx.x; i < hi; i += kThreads) {
        const int32_t p5 = __ldg(c0 + i), ng = __ldg(c1 + i);
        int32_t rel = p5 - loc;
        if (ng < 0 || uint32_t(rel) >= uint32_t(len)) continue;                    // src/bamsignals.cpp:351-353
        const int a = (ng ^ flip) & 1;
        if (flip) rel = len - 1 - rel;                                             // :356-359
        const int32_t bin = BIN1 ? rel : int32_t(fast_div(uint32_t(rel), magic, mshift));
        atomicAdd(&hist[SS ? 2 * bin + a : bin], 1);                               // :361-362
    }
    __syncthreads();
    copy_out(out + off, hist, nout);
}

// k_profile_agg (binsize >= 4): the table is coordinate-sorted, so the reads a warp looks at fall into very few
// bins (C4: 128 consecutive reads span ~400 bp of 200 bp bins).  Every lane takes FOUR consecutive reads (128-bit
// loads) and folds them into one (key, count) pair; the warp then peels off its smallest key with redux.sync
// (min over keys, add over the matching counts) and lane 0 issues ONE shared atomic per distinct key.  Three rounds
// cover the common case (<= 3 distinct keys per warp); what is left goes out as one atomic per lane.
// Division by binsize: exact for 0 <= n < 2^31 with a 32-bit multiplier m = ceil(2^(31+L) / d), L = ceil(log2 d):
// n / d = umulhi(n, m) >> (L - 1).
template <bool SS>
__global__ void __launch_bounds__(kThreads) k_profile_agg(TileTable tiles, const int32_t* __restrict__ c0,
                                                          const int32_t* __restrict__ c1, int32_t binsize, uint32_t magic,
                                                          uint32_t mshift, int32_t* __restrict__ out) {
    extern __shared__ __align__(16) int32_t smem[];
    const int64_t tix = blockIdx.x;
    const int32_t loc = tiles.loc[tix], len = tiles.len[tix];
    const bool flip = tiles.strand[tix] < 0;
    const int64_t lo = tiles.cand_lo[tix], hi = tiles.cand_hi[tix];
    const int64_t off = tiles.out_off[tix];
    const int nbins = int((uint32_t(len) + uint32_t(binsize) - 1u) / uint32_t(binsize));
    const int nout = SS ? 2 * nbins : nbins;
    int32_t* hist = smem + (off & 3);
    for (int k = threadIdx.x; k < nout; k += kThreads) hist[k] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int4* c0q = reinterpret_cast<const int4*>(c0);
    const int4* c1q = reinterpret_cast<const int4*>(c1);
    const int64_t q_lo = lo >> 2, q_hi = (hi + 3) >> 2;                            // quads of rows
    for (int64_t qb = q_lo + (threadIdx.x & ~31); qb < q_hi; qb += kThreads) {     // whole warps iterate together
        const int64_t q = qb + lane;
        int key = INT_MAX, cnt = 0;
        if (q < q_hi) {
            const int4 P = __ldg(c0q + q);
            int4 N = __ldg(c1q + q);
            if (4 * q < lo || 4 * q + 4 > hi) {                                    // first / last quad of the range
                const int64_t r0 = 4 * q;
                if (r0 < lo || r0 >= hi) N.x = -1;
                if (r0 + 1 < lo || r0 + 1 >= hi) N.y = -1;
                if (r0 + 2 < lo || r0 + 2 >= hi) N.z = -1;
                if (r0 + 3 < lo || r0 + 3 >= hi) N.w = -1;
            }
            const int32_t ps[4] = {P.x, P.y, P.z, P.w}, ns[4] = {N.x, N.y, N.z, N.w};
            int k[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int32_t r = ps[j] - loc;
                const bool ok = ns[j] >= 0 && uint32_t(r) < uint32_t(len);         // src/bamsignals.cpp:351-353
                const int32_t rel = flip ? len - 1 - r : r;                        // :356-359
                const int32_t bin = int32_t(__umulhi(uint32_t(rel), magic) >> mshift);
                const int kk = SS ? 2 * bin + ((ns[j] ^ int(flip)) & 1) : bin;     // :361-362
                k[j] = ok ? kk : INT_MAX;
            }
            key = min(min(k[0], k[1]), min(k[2], k[3]));
            cnt = int(k[0] == key) + int(k[1] == key) + int(k[2] == key) + int(k[3] == key);
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (k[j] != key && k[j] != INT_MAX) atomicAdd(&hist[k[j]], 1);
        }
#pragma unroll 1
        for (int r = 0; r < 3; ++r) {
            const int kmin = __reduce_min_sync(FULL, key);
            if (kmin == INT_MAX) break;                                            // warp-uniform
            const bool mine = key == kmin;
            const int tot = __reduce_add_sync(FULL, mine ? cnt : 0);
            if (lane == 0) atomicAdd(&hist[kmin], tot);
            if (mine) key = INT_MAX;
        }
        if (key != INT_MAX) atomicAdd(&hist[key], cnt);
    }
    __syncthreads();
    copy_out(out + off, hist, nout);
}

// ---------------------------------------------------------------------------------------------------------------
// K5 bamCoverage: one CTA per tile; +1/-1 into a shared difference array, then a block-wide inclusive scan fused with
// the coalesced write-out.  A tile is treated as its own region: the clamp max(start - loc, 0) makes every tile's scan
// self-contained.
// The scan: every WARP owns one contiguous eighth of the tile and runs through it 128 ints at a time (one int4 per
// lane, warp shuffles, its own running carry in a register), leaving its total in shared memory; after ONE barrier every
// warp adds the totals of the warps before it while it streams its own part out.  (Round 1 scanned 1024-int chunks
// across the whole CTA with three barriers per chunk: ncu showed the kernel issue-bound - 66 % of the issue slots, 2.5 TB/s
// of DRAM traffic - on ~6 k warp instructions per tile, half of them the scan.)
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) k_coverage(TileTable tiles, const int32_t* __restrict__ c0,
                                                       const int32_t* __restrict__ c1, int32_t* __restrict__ out) {
    extern __shared__ __align__(16) int32_t smem[];
    __shared__ int32_t s_warp[kThreads / 32];
    const int64_t tix = blockIdx.x;
    const int32_t loc = tiles.loc[tix], len = tiles.len[tix];
    const bool neg_region = tiles.strand[tix] < 0;
    const int64_t lo = tiles.cand_lo[tix], hi = tiles.cand_hi[tix];
    const int64_t off = tiles.out_off[tix];
    const int32_t rend = loc + len;                          // region end, exclusive
    // d[j] <-> out[off + j]; d sits `ph` ints into the 16-byte aligned buffer so that shared and global addresses
    // have the same phase; the ph leading ints are zeros and simply take part in the scan.
    const int ph = int(off & 3);
    int32_t* d = smem + ph;
    const int total = ph + len;
    const int nquads = (total + 3) >> 2;
    int4* s4 = reinterpret_cast<int4*>(smem);
    for (int k = threadIdx.x; k < nquads; k += kThreads) s4[k] = make_int4(0, 0, 0, 0);
    __syncthreads();
    for (int64_t i = lo + threadIdx.x; i < hi; i += kThreads) {
        const int32_t s = __ldg(c0 + i), e = __ldg(c1 + i);
        if (s >= rend || e < loc) continue;                  // src/bamsignals.cpp:420 (dropped reads: s = INT_MAX)
        int32_t up, down;
        if (!neg_region) { up = s - loc; down = e + 1 - loc; }          // :423-428
        else { up = rend - 1 - e; down = rend - s; }                    // :431-436
        atomicAdd(&d[up > 0 ? up : 0], 1);
        if (down < len) atomicAdd(&d[down], -1);
    }
    __syncthreads();
    // pass 1: every warp scans its own contiguous part in place (cumsum, src/bamsignals.cpp:464-470)
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    constexpr int kWarps = kThreads / 32;
    const int qpw = ((nquads + kWarps - 1) / kWarps + 31) & ~31;          // quads per warp, whole warp steps
    const int q_lo = wid * qpw, q_hi = min(nquads, q_lo + qpw);
    int carry = 0;
    for (int q0 = q_lo; q0 < q_hi; q0 += 32) {                            // warp-uniform
        const int q = q0 + lane;
        int4 v = make_int4(0, 0, 0, 0);
        if (q < q_hi) v = s4[q];
        v.y += v.x; v.z += v.y; v.w += v.z;
        int incl = v.w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int up = __shfl_up_sync(FULL, incl, o);
            if (lane >= o) incl += up;
        }
        const int prefix = carry + incl - v.w;
        if (q < q_hi) { v.x += prefix; v.y += prefix; v.z += prefix; v.w += prefix; s4[q] = v; }
        carry += __shfl_sync(FULL, incl, 31);
    }
    if (lane == 0) s_warp[wid] = carry;
    __syncthreads();
    // pass 2: add the totals of the warps in front and stream the warp's own part out (128-bit streaming stores; the
    // first and the last quad of the tile may be partial)
    int before = 0;
    for (int k = 0; k < wid; ++k) before += s_warp[k];
    int32_t* dst = out + off - ph;                           // dst[j] <-> smem[j]; 16-byte aligned by construction
    for (int q = q_lo + lane; q < q_hi; q += 32) {
        int4 v = s4[q];
        v.x += before; v.y += before; v.z += before; v.w += before;
        const int j = 4 * q;
        if (j >= ph && j + 4 <= total) {
            __stcs(reinterpret_cast<int4*>(dst + j), v);
        } else {
            if (j >= ph && j < total) __stcs(dst + j, v.x);
            if (j + 1 >= ph && j + 1 < total) __stcs(dst + j + 1, v.y);
            if (j + 2 >= ph && j + 2 < total) __stcs(dst + j + 2, v.z);
            if (j + 3 >= ph && j + 3 < total) __stcs(dst + j + 3, v.w);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Result narrowing (see kernels.cuh): four elements per thread, one 32-bit store; src is only 4-byte aligned.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_pack_u8(const int32_t* __restrict__ src, int64_t n, uint8_t* __restrict__ dst, uint2* ovf,
                                                 uint32_t cap, uint32_t* cnt) {
    const int64_t i0 = (int64_t(blockIdx.x) * 256 + threadIdx.x) * 4;
    if (i0 >= n) return;
    uint32_t word = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int64_t i = i0 + k;
        if (i < n) {
            const int32_t v = __ldcs(src + i);
            uint32_t b = uint32_t(v);
            if (uint32_t(v) > 254u) {                  // also negative values (none are produced; they would be kept exactly)
                b = 255u;
                const uint32_t slot = atomicAdd(cnt, 1u);
                if (slot < cap) ovf[slot] = make_uint2(uint32_t(i), uint32_t(v));
            }
            word |= b << (8 * k);
        }
    }
    if (i0 + 3 < n) *reinterpret_cast<uint32_t*>(dst + i0) = word;
    else for (int k = 0; k < 4 && i0 + k < n; ++k) dst[i0 + k] = uint8_t(word >> (8 * k));
}
__global__ void k_publish_reset(uint32_t* cnt, uint32_t* host_cnt) {
    *host_cnt = *cnt;
    *cnt = 0u;
    __threadfence_system();
}

void magic_for(int32_t d, uint64_t* magic, uint32_t* shift) {
    uint32_t L = 0;
    while ((1ull << L) < uint64_t(d)) ++L;
    *shift = 32 + L;
    const unsigned __int128 num = (unsigned __int128)1 << (32 + L);
    *magic = uint64_t((num + uint64_t(d) - 1) / uint64_t(d));
}

}  // namespace

void launch_decode(const DecodeBatch& one, ReadTable t, bool coverage, const FilterParams& p, DeviceScalars* sc, cudaStream_t s) {
    if (one.n <= 0) return;
    const int grid = (one.n + kDecThreads - 1) / kDecThreads;
    DecodeBatch b = one;
    b.chunk0 = 0;
    if (coverage) k_decode<true><<<unsigned(grid), kDecThreads, kSpanCap + 32, s>>>(b, nullptr, 1, t, p, sc);
    else k_decode<false><<<unsigned(grid), kDecThreads, kSpanCap + 32, s>>>(b, nullptr, 1, t, p, sc);
}

void launch_decode_table(const DecodeBatch* d_table, int n_batches, int total_chunks, ReadTable t, bool coverage,
                         const FilterParams& p, DeviceScalars* sc, cudaStream_t s) {
    if (n_batches <= 0 || total_chunks <= 0) return;
    if (coverage) k_decode<true><<<unsigned(total_chunks), kDecThreads, kSpanCap + 32, s>>>(DecodeBatch{}, d_table, n_batches, t, p, sc);
    else k_decode<false><<<unsigned(total_chunks), kDecThreads, kSpanCap + 32, s>>>(DecodeBatch{}, d_table, n_batches, t, p, sc);
}

void launch_join(ReadTable t, int64_t n, TileTable tiles, int64_t n_tiles, const DeviceScalars* sc, cudaStream_t s) {
    if (n_tiles <= 0) return;
    const int64_t grid = (n_tiles + kThreads - 1) / kThreads;
    k_join<<<unsigned(grid), kThreads, 0, s>>>(t, n, tiles, n_tiles, sc, const_cast<DeviceScalars*>(sc));
}

void launch_count(TileTable tiles, int64_t n_tiles, const int32_t* c0, const int32_t* c1, int ss, int32_t* out,
                  DeviceScalars*, cudaStream_t s) {
    if (n_tiles <= 0) return;
    const int64_t grid = (n_tiles * 32 + kThreads - 1) / kThreads;
    k_count<<<unsigned(grid), kThreads, 0, s>>>(tiles, n_tiles, c0, c1, ss, out);
}

void launch_profile(TileTable tiles, int64_t n_tiles, const int32_t* c0, const int32_t* c1, int ss, int32_t binsize,
                    int max_tile_ints, int32_t* out, DeviceScalars*, cudaStream_t s) {
    if (n_tiles <= 0) return;
    const size_t smem = size_t(max_tile_ints + 4) * sizeof(int32_t);
    const unsigned grid = unsigned(n_tiles);
    if (binsize >= 4) {
        uint32_t L = 0;
        while ((1ull << L) < uint64_t(binsize)) ++L;
        const uint32_t magic = uint32_t(((1ull << (31 + L)) + uint64_t(binsize) - 1) / uint64_t(binsize));
        if (ss) k_profile_agg<true><<<grid, kThreads, smem, s>>>(tiles, c0, c1, binsize, magic, L - 1, out);
        else k_profile_agg<false><<<grid, kThreads, smem, s>>>(tiles, c0, c1, binsize, magic, L - 1, out);
        return;
    }
    uint64_t magic; uint32_t mshift;
    magic_for(binsize, &magic, &mshift);
#define BSG_LAUNCH_PROFILE(SS, B1) k_profile<SS, B1><<<grid, kThreads, smem, s>>>(tiles, c0, c1, binsize, magic, mshift, out)
    if (binsize == 1) { if (ss) BSG_LAUNCH_PROFILE(true, true); else BSG_LAUNCH_PROFILE(false, true); }
    else              { if (ss) BSG_LAUNCH_PROFILE(true, false); else BSG_LAUNCH_PROFILE(false, false); }
#undef BSG_LAUNCH_PROFILE
}

void launch_coverage(TileTable tiles, int64_t n_tiles, const int32_t* c0, const int32_t* c1, int max_tile_ints,
                     int32_t* out, DeviceScalars*, cudaStream_t s) {
    if (n_tiles <= 0) return;
    const size_t smem = size_t(max_tile_ints + 8) * sizeof(int32_t);
    k_coverage<<<unsigned(n_tiles), kThreads, smem, s>>>(tiles, c0, c1, out);
}

void launch_pack_u8(const int32_t* d_src, int64_t n, uint8_t* d_dst, uint2* ovf, uint32_t cap, uint32_t* d_cnt, cudaStream_t s) {
    if (n <= 0) return;
    k_pack_u8<<<unsigned((n + 1023) / 1024), 256, 0, s>>>(d_src, n, d_dst, ovf, cap, d_cnt);
}
void launch_publish_reset(uint32_t* d_cnt, uint32_t* host_cnt, cudaStream_t s) { k_publish_reset<<<1, 1, 0, s>>>(d_cnt, host_cnt); }

}  // namespace bsg
