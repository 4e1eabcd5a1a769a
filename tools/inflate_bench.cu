// Micro-benchmark of the device inflate kernel alone (development tool, not part of the library): takes the BGZF blocks of
// a BAM file, uploads them, runs launch_inflate + launch_crc32 and reports ms per launch and GB/s of output.  The CRC32
// kernel is the correctness check (every block against its BGZF trailer).  The kernel source is compiled INTO this
// binary, so that variants (-D switches) can be built side by side and measured in one GPU call:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --use_fast_math -lineinfo [-D...] -o inflate_bench tools/inflate_bench.cu
//   inflate_bench file.bam [max_blocks (0 = two waves)] [reps] [spec: 1 = also follow the record chains]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../bamsignals_b200/csrc/inflate.cu"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 2; } } while (0)

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: %s file.bam [max_blocks] [reps]\n", argv[0]); return 1; }
    int n_sm = 148;
    CK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, 0));
    long max_blocks = argc > 2 ? atol(argv[2]) : 0;
    if (max_blocks <= 0) max_blocks = 2L * bsg::inflate_wave_blocks(n_sm);
    const int reps = argc > 3 ? atoi(argv[3]) : 5;
    const bool want_spec = argc > 4 && atoi(argv[4]) != 0;
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 1; }
    fseek(f, 0, SEEK_END);
    const size_t n = size_t(ftell(f));
    fseek(f, 0, SEEK_SET);
    std::vector<uint8_t> d(n);
    if (fread(d.data(), 1, n, f) != n) { fprintf(stderr, "short read\n"); return 1; }
    fclose(f);
    std::vector<bsg::InflateBlock> blocks;
    std::vector<uint32_t> crcs;
    uint64_t out = 0;
    size_t first = 0, last = 0;
    for (size_t off = 0; off + 28 <= n && long(blocks.size()) < max_blocks;) {
        const uint8_t* h = d.data() + off;
        const uint32_t xlen = h[10] | h[11] << 8, bs = (h[16] | h[17] << 8) + 1u;
        uint32_t isize, crc;
        memcpy(&isize, h + bs - 4, 4);
        memcpy(&crc, h + bs - 8, 4);
        if (isize) {
            if (blocks.empty()) first = off;
            blocks.push_back(bsg::InflateBlock{uint32_t(off - first + 12 + xlen), bs - 12 - xlen - 8, uint32_t(out), isize});
            crcs.push_back(crc);
            out += (isize + 15u) & ~15u;
        }
        off += bs;
        last = off;
    }
    const size_t comp_bytes = last - first;
    uint8_t *d_comp, *d_raw;
    bsg::InflateBlock* d_blocks;
    uint32_t* d_crc;
    bsg::DeviceScalars* d_sc;
    CK(cudaMalloc(&d_comp, comp_bytes + 4096));
    CK(cudaMemset(d_comp, 0, comp_bytes + 4096));
    CK(cudaMalloc(&d_raw, out + 4096));
    CK(cudaMalloc(&d_blocks, blocks.size() * sizeof(bsg::InflateBlock)));
    CK(cudaMalloc(&d_crc, crcs.size() * 4));
    CK(cudaMalloc(&d_sc, sizeof(bsg::DeviceScalars)));
    CK(cudaMemset(d_sc, 0, sizeof(bsg::DeviceScalars)));
    CK(cudaMemcpy(d_comp, d.data() + first, comp_bytes, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_blocks, blocks.data(), blocks.size() * sizeof(bsg::InflateBlock), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_crc, crcs.data(), crcs.size() * 4, cudaMemcpyHostToDevice));
    uint32_t* d_spec = nullptr;
    if (want_spec) CK(cudaMalloc(&d_spec, bsg::inflate_spec_words(int(blocks.size())) * 4));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    float best = 1e30f, sum = 0.f;
    for (int r = 0; r < reps + 1; ++r) {
        CK(cudaMemsetAsync(d_raw, 0xA5, out, 0));                      // also evicts the previous run's output from L2
        CK(cudaEventRecord(e0, 0));
        bsg::launch_inflate(d_blocks, int(blocks.size()), d_comp, d_raw, d_spec, d_sc, 0);
        CK(cudaEventRecord(e1, 0));
        bsg::launch_crc32(d_blocks, d_crc, int(blocks.size()), d_raw, d_sc, 0);
        CK(cudaDeviceSynchronize());
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (r > 0) { best = ms < best ? ms : best; sum += ms; }
    }
    bsg::DeviceScalars sc;
    CK(cudaMemcpy(&sc, d_sc, sizeof(sc), cudaMemcpyDeviceToHost));
    if (d_spec) {       // how many blocks got a chain, how many records
        std::vector<uint32_t> cnt(blocks.size());
        CK(cudaMemcpy(cnt.data(), d_spec, blocks.size() * 4, cudaMemcpyDeviceToHost));
        size_t ok = 0; uint64_t recs = 0;
        for (uint32_t c : cnt) if (c != 0xffffffffu) { ++ok; recs += c; }
        printf("spec: %zu of %zu blocks have a chain, %llu records\n", ok, blocks.size(), (unsigned long long)recs);
    }
    printf("{\"blocks\": %zu, \"comp_mb\": %.1f, \"out_mb\": %.1f, \"ms_min\": %.3f, \"ms_mean\": %.3f, \"out_gbs\": %.1f, \"status\": %u}\n",
           blocks.size(), comp_bytes / 1e6, out / 1e6, best, sum / reps, out / 1e6 / best, sc.status);
    return sc.status ? 3 : 0;
}
