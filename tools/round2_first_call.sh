#!/bin/bash
# First GPU call of round 2 (run through gpurun): what round 1 could prepare but not measure any more.
#   1. parity of the device-inflate tests with the four-streams-per-warp inflate instantiation (BSG_INFLATE_STREAMS=4,
#      bamsignals_b200/csrc/inflate.cu: same decode core, lanes 0/8/16/24 decode, 2 warps per CTA)
#   2. the same for the compressed-input prefetch and the loads-first match copies (BSG_INFLATE_VARIANT), then an
#      end-to-end A/B of all of them on one box, variants interleaved, results compared bit for bit
#   3. compute-sanitizer racecheck of every kernel (round 1 ran memcheck only: 0 errors)
set -u
# Budget: about 6-10 GPU-minutes (tests 6 x ~30 s, two A/B passes ~1 min, racecheck capped at 5 min).
O=gpurun_out
mkdir -p $O
BSG_INFLATE_STREAMS=4 timeout 300 python -m pytest tests -m gpu -q -x -k "gpu_inflate or random_differential or fixture" 2>&1 | tail -5 > $O/r2_tests_streams4.log
cat $O/r2_tests_streams4.log
# the other unmeasured inflate experiments (BSG_INFLATE_VARIANT: 1 / 2 = prefetch the next line of compressed input into
# L1 / L2, 4 = loads-first short-match copies, 5 / 6 = both; profiles/r1b_k_inflate_hot_lines.md says why)
for v in 1 2 4 5 6; do
    BSG_INFLATE_VARIANT=$v timeout 300 python -m pytest tests -m gpu -q -x -k "gpu_inflate or random_differential or fixture" 2>&1 | tail -2 > $O/r2_tests_variant$v.log
    echo "variant $v: $(tail -1 $O/r2_tests_variant$v.log)"
done
timeout 900 python tools/e2e_ab.py --preset c2 --reps 5 s2: s4:BSG_INFLATE_STREAMS=4 s1:BSG_INFLATE_STREAMS=1 \
    v1:BSG_INFLATE_VARIANT=1 v2:BSG_INFLATE_VARIANT=2 v4:BSG_INFLATE_VARIANT=4 v5:BSG_INFLATE_VARIANT=5 v6:BSG_INFLATE_VARIANT=6 \
    > $O/r2_ab_inflate_c2.json 2> $O/r2_ab.err
timeout 900 python tools/e2e_ab.py --preset c4 --gscale 0.1 --reps 5 s2: s4:BSG_INFLATE_STREAMS=4 \
    v1:BSG_INFLATE_VARIANT=1 v2:BSG_INFLATE_VARIANT=2 v4:BSG_INFLATE_VARIANT=4 v5:BSG_INFLATE_VARIANT=5 v6:BSG_INFLATE_VARIANT=6 \
    > $O/r2_ab_inflate_c4_g0.1.json 2>> $O/r2_ab.err
cat $O/r2_ab_inflate_c2.json $O/r2_ab_inflate_c4_g0.1.json
# host-side experiment: a short LAST batch (engine.cu, BSG_SHORT_LAST=1) to shrink the count/copy/scatter tail of a call
timeout 600 python tools/e2e_ab.py --preset c2 --reps 5 base: short_last:BSG_SHORT_LAST=1 > $O/r2_ab_short_last_c2.json 2>> $O/r2_ab.err
cat $O/r2_ab_short_last_c2.json
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_smoke.py > $O/r2_racecheck.log 2>&1
echo "racecheck rc=$?" >> $O/r2_racecheck.log
tail -5 $O/r2_racecheck.log
