#!/bin/bash
# Round 2, call R: GPU tests, inflate micro-benchmark (with / without record chains), C2 and C4 (quarter) end-to-end medians.
set -u
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
BAM=$(python - <<PY
import sys; sys.path.insert(0, "tools"); sys.path.insert(0, ".")
import workloads as W
from bench import data_dir
print(W.make_bam("c2", 0.25, data_dir())[0])
PY
)
for b in tools/probes/ib_*; do echo "$(basename $b) $($b $BAM 0 5 0 | tail -1)"; done
for b in tools/probes/ib_spec*; do echo "$(basename $b) chains: $($b $BAM 0 5 1 | tail -2 | tr '\n' ' ')"; done
python tools/e2e_ab.py --preset c2 --reps 5 base: 2>/dev/null | tail -1
python tools/e2e_ab.py --preset c4 --gscale 0.25 --reps 3 base: 2>/dev/null | tail -1
