#!/bin/bash
# Round 2, call B (2 GPUs): whole GPU suite incl. the 2-device tests, the new bench line at N=1 and N=2 (parity checked,
# in-library multi-device call timed and checked under N=2).
set -u
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv,noheader > $O/r2b_gpus.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $O/r2b_tests.log
cat $O/r2b_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 > $O/r2b_bench_n1.json 2> $O/r2b_bench_n1.err
echo "bench n1 rc=$?"; tail -c 600 $O/r2b_bench_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $O/r2b_bench_n2.json 2> $O/r2b_bench_n2.err
echo "bench n2 rc=$?"; tail -c 600 $O/r2b_bench_n2.err
python - <<'PY'
import json
for f in ("gpurun_out/r2b_bench_n1.json", "gpurun_out/r2b_bench_n2.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f, "value %.3g e2e %.1f ms (median %.1f) parity %s inlib %s" % (d["value"], d["e2e"]["ms_per_step"], d["e2e"]["ms_per_call_median_maxrank"], d["parity"], d.get("inlib")))
    print("  breakdown", d["e2e"]["breakdown_ms_rank0"], "cold", d["e2e"].get("cold"))
    for k, v in d.get("configs", {}).items():
        if "e2e" in v:
            print("  ", k, "value %.3g e2e %.1f ms parity %s inlib %s" % (v["value"], v["e2e"]["ms_per_step"], v["parity"], v.get("inlib")), v["e2e"]["breakdown_ms_rank0"])
        else:
            print("  ", k, v)
PY
