#!/bin/bash
set -u
for lib in libbamsignals_cuda.so libbamsignals_cuda_old.so; do
  echo "== $lib"
  BSG_LIB=$PWD/bamsignals_b200/$lib BSG_DEBUG=1 python tools/e2e_ab.py --preset c5 --gscale 0.25 --reps 2 base: 2>&1 | tail -60 | grep -E "batch (5|6|7|8):|pipeline host|median"
done
