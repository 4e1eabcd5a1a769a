#!/bin/bash
# Round 2 evidence pass on ONE B200 (through gpurun): parity tests, the bench lines (C2 default incl. C1 / C4 entries and
# the cold call, the reference arm, C3 / C4 / C5 at full scale), ncu launch lists and one full capture per hot kernel.
# Everything lands in gpurun_out/<tag>_*; what is meant to be judged is summarised under profiles/ afterwards.
set -u
TAG=${1:-r2z}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -q 2>&1 | tail -3 > $O/${TAG}_tests.log; cat $O/${TAG}_tests.log
python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_c2_reference_arm.json 2> $O/${TAG}_bench_c2_ref.err
python bench.py > $O/${TAG}_bench_c2_default.json 2> $O/${TAG}_bench_c2.err
python bench.py --preset c3 --also none --no-cold --cpu-seconds 6 > $O/${TAG}_bench_c3_full.json 2> $O/${TAG}_bench_c3.err
python bench.py --preset c5 --also none --no-cold --cpu-seconds 6 > $O/${TAG}_bench_c5_full.json 2> $O/${TAG}_bench_c5.err
rm -f /dev/shm/bsg_bench/c4_g1_* /dev/shm/bsg_bench/c5_g1_*
for f in c2_default c3_full c5_full; do python - <<PY
import json
try:
    d = json.loads(open("$O/${TAG}_bench_$f.json").read().strip().splitlines()[-1])
    print("$f", "step", round(d["ms_per_step"], 3), {k: (v["ms"], v["frac"]) for k, v in d["roofline"]["kernels"].items()}, "e2e", round(d["e2e"]["ms_per_step"], 1),
          d["e2e"]["breakdown_ms_rank0"], "parity", d["parity"]["equal"], d["parity"].get("full"), "cpu1", round(d["cpu_baseline"]["value"]))
    c4 = d.get("configs", {}).get("c4")
    if c4: print("  configs.c4: step", round(c4["ms_per_step"], 3), "e2e", round(c4["e2e"]["ms_per_step"], 1), c4["e2e"]["breakdown_ms_rank0"], c4["parity"]["equal"], c4["parity"].get("full"))
    if "c1" in d.get("configs", {}): print("  configs.c1:", d["configs"]["c1"]["gpu_ms_per_call_median"], d["configs"]["c1"]["cpu_ms_per_call_median"])
except Exception as e:
    print("$f", "FAILED", e)
PY
done
python - <<PY
import json
try:
    d = json.loads(open("$O/${TAG}_bench_c2_reference_arm.json").read().strip().splitlines()[-1])
    print("reference arm", d["ms_per_step"], d["value"], d["cpu_baseline"]["kind"], d["cpu_baseline"]["cores"])
except Exception as e:
    print("reference arm FAILED", e)
PY
# ncu: launch lists (kernel-only steps, then ONE end-to-end call at full scale), serialised and cold-cache: shares only
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches_c2_g0.2.csv \
    python bench.py --gscale 0.2 --steps 2 --warmup 1 --profile > $O/${TAG}_ncu_a.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${TAG}_launches_e2e_c2_full.csv \
    python tools/e2e_ab.py --preset c2 --reps -1 base: > $O/${TAG}_ncu_b.log 2>&1
# ncu: one full capture per hot kernel
ncu --set full --import-source on --clock-control none -k regex:k_inflate_ws -s 1 -c 1 -f -o $O/${TAG}_k_inflate_ws_c2_g0.5 \
    python tools/e2e_ab.py --gscale 0.5 --reps -1 base: > $O/${TAG}_ncu_c.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_decode -c 1 -f -o $O/${TAG}_k_decode_c2_g0.2 \
    python bench.py --gscale 0.2 --steps 1 --warmup 1 --profile > $O/${TAG}_ncu_d.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_profile -c 1 -f -o $O/${TAG}_k_profile_c2_g0.2 \
    python bench.py --gscale 0.2 --steps 1 --warmup 1 --profile > $O/${TAG}_ncu_e.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_coverage -c 1 -f -o $O/${TAG}_k_coverage_c3_g0.2 \
    python bench.py --preset c3 --gscale 0.2 --steps 1 --warmup 1 --profile > $O/${TAG}_ncu_f.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_walk_link -s 2 -c 1 -f -o $O/${TAG}_k_walk_link_c4_g0.1 \
    python tools/e2e_ab.py --preset c4 --gscale 0.1 --reps -1 base: > $O/${TAG}_ncu_g.log 2>&1
# sanitizers on every kernel of the path (toy BAM)
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_smoke.py > $O/${TAG}_memcheck.log 2>&1; echo "memcheck rc=$?" >> $O/${TAG}_memcheck.log
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_smoke.py > $O/${TAG}_racecheck.log 2>&1; echo "racecheck rc=$?" >> $O/${TAG}_racecheck.log
tail -n 3 $O/${TAG}_memcheck.log; tail -n 3 $O/${TAG}_racecheck.log
ls -la $O | grep ${TAG} | wc -l
