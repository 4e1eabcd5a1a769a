#!/bin/bash
# Round 2, call I: unified lock-step decode + L1 discipline in the inflate kernel; new coverage scan; K1 CTA size A/B.
set -u
O=gpurun_out
TAG=${1:-r2i}
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > $O/${TAG}_tests.log
cat $O/${TAG}_tests.log
BSG_DEBUG=1 timeout 600 python tools/e2e_ab.py --preset c2 --reps 5 base: 2> $O/${TAG}_ab_c2.err > $O/${TAG}_ab_c2.json
cat $O/${TAG}_ab_c2.json; grep "gpu pipeline" $O/${TAG}_ab_c2.err | tail -1
for lib in libbamsignals_cuda.so libbamsignals_cuda_dec128.so; do
  BSG_LIB=$PWD/bamsignals_b200/$lib timeout 600 python bench.py --no-cold --also none --steps 10 --warmup 3 --cpu-seconds 2 > $O/${TAG}_bench_c2_$lib.json 2> $O/${TAG}_bench_c2_$lib.err
  python - <<PY
import json
d = json.loads(open("$O/${TAG}_bench_c2_$lib.json").read().strip().splitlines()[-1])
print("$lib", "step", round(d["ms_per_step"], 4), {k: (v["ms"], v["frac"]) for k, v in d["roofline"]["kernels"].items()}, "e2e", round(d["e2e"]["ms_per_step"], 1), d["parity"]["equal"])
PY
done
timeout 600 python bench.py --preset c3 --no-cold --also none --steps 10 --warmup 3 --cpu-seconds 2 > $O/${TAG}_bench_c3.json 2> $O/${TAG}_bench_c3.err
python - <<PY
import json
d = json.loads(open("$O/${TAG}_bench_c3.json").read().strip().splitlines()[-1])
print("c3", "step", round(d["ms_per_step"], 4), {k: (v["ms"], v["frac"]) for k, v in d["roofline"]["kernels"].items()}, "e2e", round(d["e2e"]["ms_per_step"], 1), d["parity"])
PY
ncu --set full --import-source on --clock-control none -k regex:k_inflate_ws -s 1 -c 1 -f -o $O/${TAG}_k_inflate_ws_c2_g0.5 \
    python tools/e2e_ab.py --gscale 0.5 --reps 1 > $O/${TAG}_ncu.log 2>&1
tail -1 $O/${TAG}_ncu.log
ncu --set full --import-source on --clock-control none -k regex:k_coverage -c 1 -f -o $O/${TAG}_k_coverage_c3_g0.2 \
    python bench.py --preset c3 --gscale 0.2 --steps 1 --warmup 1 --profile > $O/${TAG}_ncu_cov.log 2>&1
tail -1 $O/${TAG}_ncu_cov.log
