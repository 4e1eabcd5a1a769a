#!/bin/bash
# Round 2, call X: k_coverage with the scanned tile in registers at four geometries (threads x int4 per lane:
# 256x4 = 4092-int tiles, 512x4 = 8188, 512x2 = 4092, 256x2 = 2044) against libbamsignals_cuda_old.so (in-place scan,
# 8192-int tiles), kernel-only resident steps on C3.
set -u
python -m pytest tests -m gpu -x -q -k "cover or random or fixture" 2>&1 | tail -2
for spec in "c3 0.2" "c3 1"; do
  set -- $spec
  for lib in libbamsignals_cuda.so libbamsignals_cuda_c512x4.so libbamsignals_cuda_c512x2.so libbamsignals_cuda_c256x2.so libbamsignals_cuda_old.so; do
    echo "$lib $1 $2: $(BSG_LIB=$PWD/bamsignals_b200/$lib python bench.py --profile --preset $1 --gscale $2 --steps 20 --warmup 3 2>/dev/null | tail -1 | cut -c60-215)"
  done
done
