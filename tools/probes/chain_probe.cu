// Probe: what does one step of a forward pointer chase through freshly written memory cost on this GPU?
// T threads, each walks its own 64 KiB region: load 4 bytes at p, p += 4 + value (51-byte records).  Variants: no
// prefetch / prefetch.global.L2 of 4 sectors 1 KB ahead / bulk L2 prefetch / 16-byte loads; cold (regions >> L2) vs
// warm (everything fits L2, second run).
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)
constexpr uint32_t kRegion = 65536, kRec = 51;
__global__ void fill(uint8_t* buf, size_t n_regions) {
    const size_t r = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
    if (r >= n_regions) return;
    uint8_t* b = buf + r * kRegion;
    for (uint32_t p = 0; p + kRec <= kRegion; p += kRec) { const uint32_t v = kRec - 4; b[p] = v; b[p + 1] = 0; b[p + 2] = 0; b[p + 3] = 0; }
}
template <int MODE>
__global__ void __launch_bounds__(128) chase(const uint8_t* __restrict__ buf, size_t n_regions, uint32_t* out) {
    const size_t r = blockIdx.x * size_t(blockDim.x) + threadIdx.x;
    if (r >= n_regions) return;
    const uint8_t* b = buf + r * kRegion;
    uint32_t p = 0, n = 0;
    while (p + kRec <= kRegion) {
        const uint32_t* w = reinterpret_cast<const uint32_t*>(b + (p & ~3u));
        const uint32_t sh = (p & 3u) * 8;
        uint32_t lo, hi = 0;
        if (MODE == 3) { lo = __ldcg(w); if (sh) hi = __ldcg(w + 1); }
        else { lo = w[0]; if (sh) hi = w[1]; }
        const uint32_t bs = __funnelshift_r(lo, hi, sh);
        const uint32_t q = p + 4 + bs;
        if (MODE == 1 && (q >> 7) != (p >> 7)) {
            const uint8_t* a = b + (q & ~127u) + 1024;
            if ((q & ~127u) + 1152 <= kRegion)
                asm volatile("prefetch.global.L2 [%0];\n\tprefetch.global.L2 [%0+32];\n\tprefetch.global.L2 [%0+64];\n\tprefetch.global.L2 [%0+96];" ::"l"(a) : "memory");
        }
        if (MODE == 2 && (q >> 10) != (p >> 10)) {
            const uint32_t a = (q & ~1023u) + 2048u;
            if (a + 1024u <= kRegion) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(b + a), "r"(1024u) : "memory");
        }
        if (MODE == 4 && (q >> 7) != (p >> 7)) {           // L1 prefetch of the 4 sectors two lines ahead
            const uint8_t* a = b + (q & ~127u) + 256;
            if ((q & ~127u) + 384 <= kRegion)
                asm volatile("prefetch.global.L1 [%0];\n\tprefetch.global.L1 [%0+32];\n\tprefetch.global.L1 [%0+64];\n\tprefetch.global.L1 [%0+96];" ::"l"(a) : "memory");
        }
        p = q; ++n;
    }
    out[r] = n;
}
template <int MODE> float run(const uint8_t* buf, size_t n, uint32_t* out) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    chase<MODE><<<unsigned((n + 127) / 128), 128>>>(buf, n, out);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}
int main() {
    for (size_t n : {size_t(9472), size_t(512)}) {         // 620 MB (cold) and 33 MB (fits L2)
        uint8_t* buf; uint32_t* out;
        CK(cudaMalloc(&buf, n * kRegion + 4096)); CK(cudaMalloc(&out, n * 4));
        fill<<<unsigned((n + 127) / 128), 128>>>(buf, n); CK(cudaDeviceSynchronize());
        const double steps = double(kRegion / kRec);
        for (int rep = 0; rep < 2; ++rep) {
            const float t0 = run<0>(buf, n, out), t1 = run<1>(buf, n, out), t2 = run<2>(buf, n, out), t3 = run<3>(buf, n, out), t4 = run<4>(buf, n, out);
            printf("regions %zu rep %d: plain %.3f ms (%.0f ns/step), pf.L2 x4 %.3f (%.0f), bulk.L2 %.3f (%.0f), ldcg %.3f (%.0f), pf.L1 x4 %.3f (%.0f)\n", n, rep,
                   t0, t0 * 1e6 / steps, t1, t1 * 1e6 / steps, t2, t2 * 1e6 / steps, t3, t3 * 1e6 / steps, t4, t4 * 1e6 / steps);
        }
        cudaFree(buf); cudaFree(out);
    }
    return 0;
}
