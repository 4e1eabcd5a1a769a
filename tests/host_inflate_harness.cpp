// Host-side unit harness for the GPU inflate kernel's decode core (bamsignals_b200/csrc/inflate_core.cuh): the
// single-lane part (bit reader, tables, symbols -> token queue) is compiled as-is with g++; the warp-cooperative
// materialisation (phase 2 of k_inflate_q2) is emulated lane by lane with SIMT load-then-store semantics.
// Every BGZF block of the given file must reproduce its CRC32 and ISIZE.   usage: harness file.bam
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <algorithm>
#include <cstdlib>
#include <vector>
#include <zlib.h>

#include "../bamsignals_b200/csrc/inflate_core.cuh"

using namespace bsg::inflate_core;

static int inflate_block(const uint8_t* buf, uint32_t in_off, uint32_t in_len, uint8_t* out, uint32_t out_len) {
    static Tables T;
    static uint32_t q[kQueue];
    const uint32_t* base = reinterpret_cast<const uint32_t*>(buf);      // the whole file, 4-byte aligned
    BitReader br;
    br.init(base, in_off);
    uint32_t op_dec = 0, pos_base = 0;
    int phase = 0, last = 0, bad = 0;
    for (;;) {
        int nq = 0, done = 0;
        if (phase == 0) {
            const int h = read_block_header(br, T, reinterpret_cast<uint8_t*>(q), &last);
            if (h == 2) return 2;
            if (h == 1) {
                br.consume((0u - br.bo) & 7u);
                const uint32_t v = br.peek();
                br.consume(32);
                const uint32_t len = v & 0xffffu, nlen = v >> 16;
                if ((len ^ nlen) != 0xffffu || op_dec + len > out_len) return 3;
                const uint32_t src_off = br.byte_pos();
                if ((uint64_t(src_off) + len) * 8 > (uint64_t(in_off) + in_len) * 8) return 3;
                op_dec += len;
                br.init(base, src_off + len);
                q[0] = kTokSkip | len;                 // as the kernel: a stored block is a two-slot queue of its own,
                q[1] = src_off;                        // moved by a copy warp
                nq = 2;
                if (last) done = 1;
            } else phase = 1;
        }
        if (phase == 1 && nq == 0) {
            int eob = 0;
            nq = fill_queue(br, ArrayAccess{&T, q}, &op_dec, &eob, &bad);
            if (eob) { phase = 0; if (last) done = 1; }
            if (bad || op_dec > out_len) return 4;
            if (br.bit_pos() > (uint64_t(in_off) + in_len) * 8) return 6;      // as the kernel: never more than one round past the end
        }
        // phase 2 emulation
        if (nq == 2 && !(q[0] >> 31) && (q[0] & kTokSkip)) {
            memcpy(out + pos_base, buf + q[1], q[0] & 0xffffu);
            pos_base += q[0] & 0xffffu;
            nq = 0;
        }
        if (nq > 0) {
            // the kernel's byte-parallel materialisation (inflate.cu: materialise), lane by lane: token starts in a
            // bitmap (1024-byte windows), owner of a byte = prefix pop-count; the output is built in chunks of 256 bytes:
            // pass 1 resolves every byte of the chunk to a literal, a source in FRONT of the chunk (finished output, loaded
            // before anything of the chunk is written) or a source inside the chunk (sources inside the byte's own 32-byte
            // step are chased through their owner tokens); pass 2 goes through the steps in order via the stage; pass 3
            // copies the stage out
            uint32_t t[32], incl[32]; int32_t s[32], e[32];
            uint32_t run = 0;
            for (int lane = 0; lane < 32; ++lane) {
                const bool valid = lane < nq;
                t[lane] = valid ? q[lane] : 0;
                const uint32_t len = !valid ? 0 : ((t[lane] >> 31) ? (t[lane] & 0x1ffu) : 1u);
                run += len; incl[lane] = run;
            }
            const uint32_t total = run;
            uint8_t* first_byte = out + pos_base;
            const uint32_t a0 = uint32_t(reinterpret_cast<uintptr_t>(first_byte) & 7u);
            uint8_t* al = first_byte - a0;
            for (int lane = 0; lane < 32; ++lane) {
                const uint32_t len = lane < nq ? ((t[lane] >> 31) ? (t[lane] & 0x1ffu) : 1u) : 0u;
                s[lane] = int32_t(incl[lane] - len + a0); e[lane] = s[lane] + int32_t(len);
            }
            const int32_t end_u = int32_t(a0 + total);
            for (int32_t wb = 0; wb < end_u && total; wb += 1024) {
                const int32_t wend = std::min(end_u, wb + 1024);
                uint32_t firstk = 0, bm[32], pre[32];
                for (int k = 0; k < 32; ++k) bm[k] = 0;
                for (int lane = 0; lane < 32; ++lane) {
                    const bool valid = lane < nq;
                    if (valid && e[lane] <= wb) ++firstk;
                    if (valid && e[lane] > wb && s[lane] < wend) {
                        const uint32_t r = uint32_t(std::max(s[lane], wb) - wb);
                        bm[r >> 5] |= 1u << (r & 31u);
                    }
                }
                uint32_t acc = 0;
                for (int k = 0; k < 32; ++k) { pre[k] = acc; acc += uint32_t(__builtin_popcount(bm[k])); }
                for (int32_t cb = wb; cb < wend; cb += 256) {
                    const int32_t clim = std::min(wend, cb + 256), front = std::max(cb, int32_t(a0));
                    uint8_t stage[256];
                    int32_t val[256]; uint8_t kind[256];            // kind: 0 literal byte, 1 source in front of the chunk, 2 on the stage
                    for (int32_t u = front; u < clim; ++u) {        // pass 1 (every load happens before any store of the chunk)
                        const int32_t lim = std::max(u & ~31, int32_t(a0));
                        int32_t xx = u;
                        for (int hops = 0;; ++hops) {
                            if (hops > 40) return 7;
                            const uint32_t r = uint32_t(xx - wb);
                            const int k = int((firstk + pre[r >> 5] + uint32_t(__builtin_popcount(bm[r >> 5] & ((2u << (r & 31u)) - 1u))) - 1u) & 31u);
                            if (!(xx >= s[k] && xx < e[k])) return 8;  // the owner lookup must land on the covering token
                            if (!(t[k] >> 31)) { val[u - cb] = int32_t(t[k] & 0xffu); kind[u - cb] = 0; break; }
                            const int32_t d = int32_t(((t[k] >> 16) & 0x7fffu) + 1u);
                            int32_t y = xx - d;
                            if (y >= s[k]) y = s[k] - d + (xx - s[k]) % d;
                            if (y < lim) {
                                if (y < front) { val[u - cb] = al[y]; kind[u - cb] = 1; }
                                else { val[u - cb] = y - cb; kind[u - cb] = 2; }
                                break;
                            }
                            xx = y;
                        }
                    }
                    for (int32_t u = front; u < clim; ++u)          // pass 2: a stage source was written by an earlier STEP
                        if (kind[u - cb] == 2) {
                            if (val[u - cb] + cb >= (u & ~31)) return 9;
                            stage[u - cb] = stage[val[u - cb]];
                        } else stage[u - cb] = uint8_t(val[u - cb]);
                    for (int32_t u = front; u < clim; ++u) al[u] = stage[u - cb];      // pass 3
                }
            }
            pos_base += total;
        }
        if (done) break;
    }
    if (op_dec != out_len || pos_base != out_len) return 5;
    if (br.bit_pos() > (uint64_t(in_off) + in_len) * 8) return 6;
    return 0;
}

// --fuzz N: damage the compressed payload of every block in turn (bit flips, byte runs, truncation) and decode it from a
// private buffer with the same 4096 bytes of slack the device buffer has.  Any outcome is fine - an error code, a CRC
// mismatch, even a harmless change - as long as the decode core stays inside its buffers (ASan / UBSan watch) and
// terminates.
static int fuzz(const std::vector<uint8_t>& d, size_t n, long iters) {
    struct Blk { size_t off; uint32_t in_off, in_len, isize, crc; };
    std::vector<Blk> blks;
    for (size_t off = 0; off + 28 <= n;) {
        const uint8_t* h = d.data() + off;
        const uint32_t xlen = h[10] | h[11] << 8, bs = (h[16] | h[17] << 8) + 1u;
        Blk b{off, uint32_t(12 + xlen), bs - 12 - xlen - 8, 0, 0};
        memcpy(&b.isize, h + bs - 4, 4);
        memcpy(&b.crc, h + bs - 8, 4);
        if (b.isize) blks.push_back(b);
        off += bs;
    }
    if (blks.empty()) return 0;
    uint64_t rng = 0x9e3779b97f4a7c15ull;
    auto next = [&]() { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; return rng; };
    long rejected = 0, crc_caught = 0, harmless = 0;
    for (long it = 0; it < iters; ++it) {
        const Blk& b = blks[size_t(it) % blks.size()];
        std::vector<uint8_t> pb(size_t(b.in_len) + 4096, 0);
        memcpy(pb.data(), d.data() + b.off + b.in_off, b.in_len);
        const int kind = int(next() % 4);
        if (kind == 0) { for (int k = 0, m = 1 + int(next() % 8); k < m; ++k) pb[next() % b.in_len] ^= uint8_t(1u << (next() % 8)); }
        else if (kind == 1) { size_t a = next() % b.in_len, l = 1 + next() % 64; for (size_t k = a; k < a + l && k < b.in_len; ++k) pb[k] = uint8_t(next()); }
        else if (kind == 2) { size_t a = next() % b.in_len; memset(pb.data() + a, 0, b.in_len - a); }
        else { size_t a = next() % std::min<size_t>(b.in_len, 40); pb[a] = uint8_t(next()); }          // block header area
        std::vector<uint8_t> out(size_t(b.isize) + 64);
        const int rc = inflate_block(pb.data(), 0, b.in_len, out.data(), b.isize);
        if (rc) ++rejected;
        else if (uint32_t(crc32(crc32(0, nullptr, 0), out.data(), b.isize)) != b.crc) ++crc_caught;
        else ++harmless;
    }
    printf("fuzz %ld rejected %ld crc_caught %ld harmless %ld bad 0\n", iters, rejected, crc_caught, harmless);
    return 0;
}

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    long fuzz_iters = 0;
    if (argc >= 4 && strcmp(argv[1], "--fuzz") == 0) { fuzz_iters = atol(argv[2]); argv += 2; }
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 2;
    std::vector<uint8_t> d;
    uint8_t tmp[65536];
    size_t k;
    while ((k = fread(tmp, 1, sizeof tmp, f)) > 0) d.insert(d.end(), tmp, tmp + k);
    fclose(f);
    const size_t n = d.size();
    d.resize(n + 64);
    if (fuzz_iters > 0) return fuzz(d, n, fuzz_iters);
    size_t off = 0;
    int nb = 0, bad = 0;
    while (off + 28 <= n) {
        const uint8_t* h = d.data() + off;
        const uint32_t xlen = h[10] | h[11] << 8, bs = (h[16] | h[17] << 8) + 1u;
        uint32_t isize, crc;
        memcpy(&isize, h + bs - 4, 4);
        memcpy(&crc, h + bs - 8, 4);
        std::vector<uint8_t> out(isize + 64);
        const int rc = inflate_block(d.data(), uint32_t(off + 12 + xlen), bs - 12 - xlen - 8, out.data(), isize);
        if (rc || uint32_t(crc32(crc32(0, nullptr, 0), out.data(), isize)) != crc) {
            if (++bad < 5) fprintf(stderr, "block %d at %zu: rc %d\n", nb, off, rc);
        }
        ++nb;
        off += bs;
    }
    printf("blocks %d bad %d\n", nb, bad);
    return bad != 0;
}
