#!/bin/bash
# ncu capture of one inflate micro-benchmark variant:  r2_call_k.sh TAG VARIANT
set -u
O=gpurun_out
TAG=${1:-r2k}; V=${2:-pred}
mkdir -p $O
BAM=$(python - <<PY
import sys; sys.path.insert(0, "tools"); sys.path.insert(0, ".")
import workloads as W
from bench import data_dir
print(W.make_bam("c2", 0.25, data_dir())[0])
PY
)
ncu --set full --import-source on --clock-control none -k regex:k_inflate -c 1 -f -o $O/${TAG}_ib_$V tools/probes/ib_$V $BAM 9472 1 > $O/${TAG}_ncu_$V.log 2>&1
tail -2 $O/${TAG}_ncu_$V.log
