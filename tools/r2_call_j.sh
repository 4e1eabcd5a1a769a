#!/bin/bash
# Round 2, call J: inflate micro-benchmark variants (tools/inflate_bench.cu built by tools/ib_build.sh into tools/probes/ib_*).
set -u
O=gpurun_out
TAG=${1:-r2j}
mkdir -p $O
BAM=$(python - <<PY
import sys; sys.path.insert(0, "tools"); sys.path.insert(0, ".")
import workloads as W
from bench import data_dir
print(W.make_bam("c2", 0.25, data_dir())[0])
PY
)
echo "bam $BAM" > $O/${TAG}_ib.log
for rep in 1 2; do
for b in tools/probes/ib_*; do
  echo "$(basename $b) $(timeout 120 $b $BAM 0 5 2>&1 | tail -1)" | tee -a $O/${TAG}_ib.log
done
done
