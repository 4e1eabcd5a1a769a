#!/bin/bash
# Round 2, call T: tests + racecheck of the current build, then same-box A/B current vs libbamsignals_cuda_old.so on C4 / C5 / C2 at full scale.
set -u
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_smoke.py 2>&1 | tail -3
for spec in "c4 1" "c5 1" "c2 1"; do
  set -- $spec
  for rep in 1 2; do
    python tools/ab_lib.py $1 $2 2>/dev/null | tail -1
    BSG_LIB=$PWD/bamsignals_b200/libbamsignals_cuda_old.so python tools/ab_lib.py $1 $2 2>/dev/null | tail -1
  done
done
rm -f /dev/shm/bsg_bench/c4_g1_* /dev/shm/bsg_bench/c5_g1_*
