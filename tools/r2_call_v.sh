#!/bin/bash
# Round 2, call V: GPU tests of the current build, then same-box A/B of k_coverage (registers-resident scan) against
# libbamsignals_cuda_old.so (in-place scan) on C3 at 1/5 and full scale, kernel-only resident steps.
set -u
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for spec in "c3 0.2" "c3 1"; do
  set -- $spec
  for rep in 1 2; do
    echo "new $1 $2: $(python bench.py --profile --preset $1 --gscale $2 --steps 20 --warmup 3 2>/dev/null | tail -1)"
    echo "old $1 $2: $(BSG_LIB=$PWD/bamsignals_b200/libbamsignals_cuda_old.so python bench.py --profile --preset $1 --gscale $2 --steps 20 --warmup 3 2>/dev/null | tail -1)"
  done
done
