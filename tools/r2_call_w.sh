#!/bin/bash
# Round 2, call W: k_coverage with the leaner in-place scan (predicated shuffle rounds, no per-lane bounds, 32 registers)
# at three tile sizes (8188 / 7164 / 6140 ints = 6 / 7 / 8 CTAs per SM) against libbamsignals_cuda_old.so, kernel-only.
set -u
python -m pytest tests -m gpu -x -q -k "cover or random or fixture" 2>&1 | tail -2
for spec in "c3 0.2" "c3 1"; do
  set -- $spec
  for lib in libbamsignals_cuda.so libbamsignals_cuda_t7164.so libbamsignals_cuda_t6140.so libbamsignals_cuda_old.so; do
    echo "$lib $1 $2: $(BSG_LIB=$PWD/bamsignals_b200/$lib python bench.py --profile --preset $1 --gscale $2 --steps 20 --warmup 3 2>/dev/null | tail -1 | cut -c1-230)"
  done
done
