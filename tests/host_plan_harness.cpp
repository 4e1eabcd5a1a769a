// Host-only harness around the library's BAM / index reader and fetch planner (bamsignals_b200/csrc/bamio.cpp,
// plan.cpp), built with -fsanitize=address,undefined by tests/test_host_parsers_fuzz.py.  TEST INFRASTRUCTURE.
//
//   host_plan_harness <file.bam> [ext]
//
// Opens the file (header + .bai/.csi), plans the fetch for every whole contig plus a few short regions, inflates every
// planned block with CRC check and walks the block_size chain of every segment - the host-side equivalent of what
// bsg_debug_plan does inside the shared library.  Exit 0 with "ok ..." or "error <code> <message>" (a clean refusal);
// anything else (sanitizer report, signal, uncaught exception) is a defect.
#include <cstdio>
#include <cstdlib>
#include <new>
#include <string>
#include <vector>

#include "../bamsignals_b200/csrc/bamio.h"
#include "../bamsignals_b200/csrc/plan.h"
#include "../bamsignals_b200/csrc/pool.h"

using namespace bsg;

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: %s file.bam [ext]\n", argv[0]); return 2; }
    const int64_t ext = argc > 2 ? atoll(argv[2]) : 0;
    try {
        BamFile bam(argv[1]);
        Pool pool(2);
        const auto& names = bam.ref_names();
        const auto& lens = bam.ref_lens();
        std::vector<const char*> levels;
        for (auto& n : names) levels.push_back(n.c_str());
        std::vector<int32_t> seq_idx, loc, width;
        std::vector<int8_t> strand;
        for (size_t t = 0; t < names.size() && t < 64; ++t) {
            const int32_t L = lens[t] > 0 ? lens[t] : 1;
            const int32_t locs[4] = {0, L / 2, L - 1, -100};
            const int32_t wids[4] = {L, 1000, 5000, 300};
            for (int k = 0; k < 4; ++k) { seq_idx.push_back(int32_t(t)); loc.push_back(locs[k]); width.push_back(wids[k]); strand.push_back(int8_t(k % 3 - 1)); }
        }
        {   // the entry points must come out strictly ascending whichever sort produced them
            const auto& ent = bam.entry_points();
            for (size_t i = 1; i < ent.size(); ++i)
                if (ent[i - 1] >= ent[i]) fail(BSG_EARG, "internal error: index entry points are not sorted and unique");
        }
        Regions rg;
        resolve_regions(bam, int64_t(loc.size()), levels.data(), int32_t(levels.size()), seq_idx.data(), loc.data(), width.data(),
                        strand.data(), &rg);
        std::vector<Segment> segs;
        plan_fetch(bam, rg, ext, 1u << 20, pool, &segs);
        Inflater inf;
        std::vector<uint8_t> buf;
        long long n_rec = 0, ub = 0;
        for (const Segment& sg : segs) {
            buf.assign(sg.usize + 8, 0);
            uint64_t u = 0;
            for (const BlockInfo& b : sg.blocks) { inf.inflate_block(bam.data(), b, buf.data() + u, true); u += b.isize; }
            ub += (long long)sg.usize;
            for (uint64_t p = sg.ubeg; p < sg.uend;) {
                if (p + 4 > sg.uend) fail(BSG_EFORMAT, "truncated BAM record");
                const int32_t bs = rd_i32(buf.data() + p);
                if (bs < 32 || p + 4 + uint64_t(bs) > sg.uend) fail(BSG_EFORMAT, "corrupt BAM record chain");
                ++n_rec;
                p += 4 + uint64_t(bs);
            }
        }
        // the helpers the multi-device sharding uses
        for (size_t t = 0; t < names.size() && t < 64; ++t) { (void)bam.bp_per_block(int(t)); (void)bam.approx_coffset(int(t), lens[t] / 2); }
        printf("ok refs %zu segments %zu inflated %lld records %lld entries %zu\n", names.size(), segs.size(), ub, n_rec, bam.entry_points().size());
    } catch (Error& e) {
        printf("error %d %s\n", e.code, e.msg.c_str());
    } catch (std::bad_alloc&) {
        printf("error %d out of host memory\n", BSG_ENOMEM);
    }
    return 0;
}
