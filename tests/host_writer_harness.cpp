// Host-only harness around the SAM -> BAM + BAI writer (bamsignals_b200/csrc/bamwrite.cpp), built with
// -fsanitize=address,undefined by tests/test_host_parsers_fuzz.py.  TEST INFRASTRUCTURE.
//   host_writer_harness in.sam out.bam    -> "ok" | "error <code> <message>", exit 0; anything else is a defect.
#include <cstdio>
#include <new>

#include "../bamsignals_b200/csrc/common.h"

namespace bsg { void write_sam_as_bam_and_index(const char* sampath, const char* bampath); }

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    try {
        bsg::write_sam_as_bam_and_index(argv[1], argv[2]);
        printf("ok\n");
    } catch (bsg::Error& e) {
        printf("error %d %s\n", e.code, e.msg.c_str());
    } catch (std::bad_alloc&) {
        printf("error %d out of host memory\n", BSG_ENOMEM);
    }
    return 0;
}
