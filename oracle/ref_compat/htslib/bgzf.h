// Stand-in for <htslib/bgzf.h>: the declarations the reference's src/bamsignals.cpp needs (src/bamsignals.cpp:5,200,212).
// TEST INFRASTRUCTURE (oracle/_ref): htslib is an un-vendored dependency of the reference (LinkingTo: Rhtslib >= 1.13.1,
// DESCRIPTION:30), absent offline; names, types and argument order follow htslib >= 1.10's public header, the
// implementation (../hts_compat.cpp) is this repo's own reader over zlib.
#pragma once
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BGZF_BLOCK_SIZE 0xff00
#define BGZF_MAX_BLOCK_SIZE 0x10000

typedef struct BGZF BGZF;

// htslib: "Set the cache size. Only effective when compiled with -DBGZF_CACHE"; size in bytes, blocks of
// BGZF_MAX_BLOCK_SIZE (the reference asks for 10 blocks, src/bamsignals.cpp:200,212)
void bgzf_set_cache_size(BGZF* fp, int size);

#ifdef __cplusplus
}
#endif
