#!/bin/bash
# Round 2, call Y: GPU tests of the final tree, then the C2 end-to-end call (planner changes: ms_plan) five times.
set -u
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python tools/e2e_ab.py --preset c2 --gscale 1 --reps 7 base: 2>/dev/null | tail -1
