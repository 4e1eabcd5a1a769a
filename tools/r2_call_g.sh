#!/bin/bash
# Round 2: quick inflate iteration: parity of the inflate tests, C2 e2e timing, one ncu capture.
set -u
O=gpurun_out
TAG=${1:-r2g}
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q -k "gpu_inflate or fixture or differential" 2>&1 | tail -3 > $O/${TAG}_tests.log
cat $O/${TAG}_tests.log
BSG_DEBUG=1 timeout 600 python tools/e2e_ab.py --preset c2 --reps 5 base: 2> $O/${TAG}_ab_c2.err > $O/${TAG}_ab_c2.json
cat $O/${TAG}_ab_c2.json; grep "gpu pipeline" $O/${TAG}_ab_c2.err | tail -1
ncu --set full --import-source on --clock-control none -k regex:k_inflate_ws -s 1 -c 1 -f -o $O/${TAG}_k_inflate_ws_c2_g0.5 \
    python tools/e2e_ab.py --gscale 0.5 --reps 1 > $O/${TAG}_ncu.log 2>&1
tail -1 $O/${TAG}_ncu.log
