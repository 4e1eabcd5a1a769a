#!/bin/bash
# Round 2, call F: byte-parallel materialise + background page-locking (direct DMA upload from the second call on).
set -u
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > $O/r2f_tests.log
cat $O/r2f_tests.log
BSG_DEBUG=1 timeout 600 python tools/e2e_ab.py --preset c2 --reps 9 base: 2> $O/r2f_ab_c2.err > $O/r2f_ab_c2.json
cat $O/r2f_ab_c2.json; grep "gpu pipeline" $O/r2f_ab_c2.err | tail -2; grep -i "refused" $O/r2f_ab_c2.err | head -2
timeout 600 python tools/e2e_ab.py --preset c4 --gscale 0.1 --reps 5 base: 2> $O/r2f_ab_c4.err > $O/r2f_ab_c4.json
cat $O/r2f_ab_c4.json
timeout 600 python tools/e2e_ab.py --preset c3 --reps 5 base: 2> $O/r2f_ab_c3.err > $O/r2f_ab_c3.json
cat $O/r2f_ab_c3.json
ncu --set full --import-source on --clock-control none -k regex:k_inflate_ws -s 1 -c 1 -f -o $O/r2f_k_inflate_ws_c2_g0.5 \
    python tools/e2e_ab.py --gscale 0.5 --reps 1 > $O/r2f_ncu.log 2>&1
tail -2 $O/r2f_ncu.log
