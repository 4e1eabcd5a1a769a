// Host harness for bamsignals_b200/csrc/hostsimd.cpp (the result scatter's streaming-store copies): every size 0..300 and a
// few large ones, every destination / source misalignment, canaries on both sides.  Built with ASan/UBSan by the test.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>
#include "../bamsignals_b200/csrc/hostsimd.h"

int main() {
    std::vector<int64_t> sizes;
    for (int64_t n = 0; n <= 300; ++n) sizes.push_back(n);
    for (int64_t n : {1023, 1024, 1025, 4096, 8191, 65536 + 7}) sizes.push_back(n);
    long checked = 0;
    for (int64_t n : sizes)
        for (int da = 0; da < 9; ++da)
            for (int sa = 0; sa < 5; ++sa) {
                std::vector<int32_t> dst(size_t(n) + 32, 0x5a5a5a5a), src32(size_t(n) + 16);
                std::vector<uint8_t> src8(size_t(n) + 16);
                for (int64_t i = 0; i < n + 8; ++i) { src8[size_t(sa + i)] = uint8_t(i * 7 + 3); src32[size_t(std::min<int64_t>(sa, 4) + i)] = int32_t(i * 2654435761u); }
                bsg::widen_u8_to_i32(dst.data() + 8 + da, src8.data() + sa, n);
                for (int64_t i = 0; i < n; ++i) if (dst[size_t(8 + da + i)] != int32_t(src8[size_t(sa + i)])) { printf("widen mismatch n=%lld da=%d sa=%d i=%lld\n", (long long)n, da, sa, (long long)i); return 1; }
                if (dst[size_t(7 + da)] != 0x5a5a5a5a || dst[size_t(8 + da + n)] != 0x5a5a5a5a) { printf("widen overrun n=%lld da=%d\n", (long long)n, da); return 1; }
                std::fill(dst.begin(), dst.end(), 0x5a5a5a5a);
                const int so = std::min(sa, 4);
                bsg::copy_i32_stream(dst.data() + 8 + da, src32.data() + so, n);
                for (int64_t i = 0; i < n; ++i) if (dst[size_t(8 + da + i)] != src32[size_t(so + i)]) { printf("copy mismatch n=%lld da=%d i=%lld\n", (long long)n, da, (long long)i); return 1; }
                if (dst[size_t(7 + da)] != 0x5a5a5a5a || dst[size_t(8 + da + n)] != 0x5a5a5a5a) { printf("copy overrun n=%lld da=%d\n", (long long)n, da); return 1; }
                ++checked;
            }
    printf("ok %ld\n", checked);
    return 0;
}
