#!/bin/bash
# Round 2, call N: bench lines of the other presets (C3 full scale, C4 at quarter scale, C5 at 1/5) with the current library.
set -u
O=gpurun_out
TAG=${1:-r2n}
mkdir -p $O
for spec in "c3 1" "c4 0.25" "c5 0.2"; do
  set -- $spec
  timeout 900 python bench.py --preset $1 --gscale $2 --no-cold --also none --steps 10 --warmup 3 --cpu-seconds 2 > $O/${TAG}_bench_$1.json 2> $O/${TAG}_bench_$1.err
  python - <<PY
import json
try:
    d = json.loads(open("$O/${TAG}_bench_$1.json").read().strip().splitlines()[-1])
    print("$1 g$2 step", round(d["ms_per_step"], 4), {k: (v["ms"], v["frac"]) for k, v in d["roofline"]["kernels"].items()}, "e2e", round(d["e2e"]["ms_per_step"], 1), d["e2e"]["breakdown_ms_rank0"], d["parity"]["equal"])
except Exception as e:
    print("$1 failed", e); print(open("$O/${TAG}_bench_$1.err").read()[-2000:])
PY
done
