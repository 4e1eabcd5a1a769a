#!/bin/bash
# First GPU call of round 2 (run through gpurun): what round 1 could prepare but not measure any more.
#   1. parity of the device-inflate tests with the four-streams-per-warp inflate instantiation (BSG_INFLATE_STREAMS=4,
#      bamsignals_b200/csrc/inflate.cu: same decode core, lanes 0/8/16/24 decode, 2 warps per CTA)
#   2. end-to-end A/B on one box, variants interleaved, results compared bit for bit: two vs four (vs one) streams
#   3. compute-sanitizer racecheck of every kernel (round 1 ran memcheck only: 0 errors)
set -u
O=gpurun_out
mkdir -p $O
BSG_INFLATE_STREAMS=4 timeout 300 python -m pytest tests -m gpu -q -x -k "gpu_inflate or random_differential or fixture" 2>&1 | tail -5 > $O/r2_tests_streams4.log
cat $O/r2_tests_streams4.log
timeout 600 python tools/e2e_ab.py --preset c2 --reps 7 s2: s4:BSG_INFLATE_STREAMS=4 s1:BSG_INFLATE_STREAMS=1 > $O/r2_ab_inflate_streams_c2.json 2> $O/r2_ab.err
timeout 600 python tools/e2e_ab.py --preset c4 --gscale 0.1 --reps 7 s2: s4:BSG_INFLATE_STREAMS=4 > $O/r2_ab_inflate_streams_c4_g0.1.json 2>> $O/r2_ab.err
cat $O/r2_ab_inflate_streams_c2.json $O/r2_ab_inflate_streams_c4_g0.1.json
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_smoke.py > $O/r2_racecheck.log 2>&1
echo "racecheck rc=$?" >> $O/r2_racecheck.log
tail -5 $O/r2_racecheck.log
