// Host-side copies of result tiles into the caller's buffers with streaming (non-temporal) stores: the destination is
// written once and not read again by this library, so the stores should not first pull every destination line into the
// cache (regular stores cost a read-for-ownership per line: 3 bytes of memory traffic per byte delivered instead of 2).
#pragma once
#include <cstdint>

namespace bsg {
// dst[i] = src[i] for i < n (uint8 -> int32), dst 4-byte aligned
void widen_u8_to_i32(int32_t* dst, const uint8_t* src, int64_t n);
// dst[i] = src[i] for i < n
void copy_i32_stream(int32_t* dst, const int32_t* src, int64_t n);
}  // namespace bsg
