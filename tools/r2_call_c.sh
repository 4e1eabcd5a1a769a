#!/bin/bash
# Round 2, call C: the warp-specialised inflate kernel + direct DMA upload: parity first, then timings.
set -u
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q -k "gpu_inflate or fixture" 2>&1 | tail -8 > $O/r2c_tests_inflate.log
cat $O/r2c_tests_inflate.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > $O/r2c_tests.log
cat $O/r2c_tests.log
BSG_DEBUG=1 timeout 600 python tools/e2e_ab.py --preset c2 --reps 7 base: 2> $O/r2c_ab_c2.err > $O/r2c_ab_c2.json
cat $O/r2c_ab_c2.json; grep "gpu pipeline" $O/r2c_ab_c2.err | tail -3; grep -i "refused" $O/r2c_ab_c2.err | head -2
timeout 600 python tools/e2e_ab.py --preset c4 --gscale 0.1 --reps 5 base: 2> $O/r2c_ab_c4.err > $O/r2c_ab_c4.json
cat $O/r2c_ab_c4.json
timeout 600 python tools/e2e_ab.py --preset c3 --reps 5 base: 2> $O/r2c_ab_c3.err > $O/r2c_ab_c3.json
cat $O/r2c_ab_c3.json
