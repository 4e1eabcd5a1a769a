#!/bin/bash
# Round 2, call L: GPU tests + C2 end-to-end A/B line + bench (C2) with the current library.
set -u
O=gpurun_out
TAG=${1:-r2l}
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > $O/${TAG}_tests.log
cat $O/${TAG}_tests.log
BSG_DEBUG=1 timeout 600 python tools/e2e_ab.py --preset c2 --reps 5 base: 2> $O/${TAG}_ab_c2.err > $O/${TAG}_ab_c2.json
cat $O/${TAG}_ab_c2.json; grep "gpu pipeline" $O/${TAG}_ab_c2.err | tail -1
timeout 600 python bench.py --no-cold --also none --steps 10 --warmup 3 --cpu-seconds 2 > $O/${TAG}_bench_c2.json 2> $O/${TAG}_bench_c2.err
python - <<PY
import json
d = json.loads(open("$O/${TAG}_bench_c2.json").read().strip().splitlines()[-1])
print("c2 step", round(d["ms_per_step"], 4), {k: (v["ms"], v["frac"]) for k, v in d["roofline"]["kernels"].items()}, "e2e", round(d["e2e"]["ms_per_step"], 1), d["e2e"]["breakdown_ms_rank0"], d["parity"]["equal"])
PY
