#!/bin/bash
# One measurement pass on a B200 box (run through gpurun): parity tests, the bench lines of every config at full
# scale, the reference arm, the ncu launch lists and one full ncu capture per hot kernel.  Everything lands in
# gpurun_out/<tag>_*; the summaries that are meant to be judged are copied into profiles/ afterwards.
set -u
TAG=${1:-r1b}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -q 2>&1 | tail -3 > $O/${TAG}_tests.log
python bench.py > $O/${TAG}_bench_c2_default.json 2> $O/${TAG}_bench_c2.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_c2_reference_arm.json 2>> $O/${TAG}_bench_c2.err
python bench.py --preset c3 --cpu-seconds 6 > $O/${TAG}_bench_c3_full.json 2> $O/${TAG}_bench_c3.err
python bench.py --preset c5 --cpu-seconds 6 > $O/${TAG}_bench_c5_full.json 2> $O/${TAG}_bench_c5.err
python bench.py --preset c4 --cpu-seconds 6 > $O/${TAG}_bench_c4_full.json 2> $O/${TAG}_bench_c4.err
rm -f /dev/shm/bsg_bench/c4_g1_* /dev/shm/bsg_bench/c5_g1_*
# ncu: launch lists (kernel-only steps, then the end-to-end call), serialised and cold-cache: shares only
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches_c2_g0.2.csv \
    python bench.py --gscale 0.2 --steps 2 --warmup 1 --profile > $O/${TAG}_ncu_a.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches_e2e_c2_g0.2.csv \
    python tools/e2e_ab.py --gscale 0.2 --reps 1 > $O/${TAG}_ncu_b.log 2>&1
# ncu: one full capture per hot kernel
ncu --set full --import-source on --clock-control none -k regex:k_decode -c 1 -f -o $O/${TAG}_k_decode_c2_g0.2 \
    python bench.py --gscale 0.2 --steps 1 --warmup 1 --profile > $O/${TAG}_ncu_c.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_profile -c 1 -f -o $O/${TAG}_k_profile_c2_g0.2 \
    python bench.py --gscale 0.2 --steps 1 --warmup 1 --profile > $O/${TAG}_ncu_d.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_profile_agg -c 1 -f -o $O/${TAG}_k_profile_agg_c4_g0.05 \
    python bench.py --preset c4 --gscale 0.05 --steps 1 --warmup 1 --profile > $O/${TAG}_ncu_e.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_inflate_q2 -c 1 -f -o $O/${TAG}_k_inflate_c2_g0.2 \
    python tools/e2e_ab.py --gscale 0.2 --reps 1 > $O/${TAG}_ncu_f.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_coverage -c 1 -f -o $O/${TAG}_k_coverage_c3_g0.2 \
    python bench.py --preset c3 --gscale 0.2 --steps 1 --warmup 1 --profile > $O/${TAG}_ncu_g.log 2>&1
ls -la $O | tail -30
for f in c2_default c3_full c4_full c5_full; do python - <<PY
import json
try:
    d=json.load(open("$O/${TAG}_bench_$f.json"))
    print("$f", round(d["ms_per_step"],3), {k:(v["ms"],v["gbs"]) for k,v in d["roofline"]["kernels"].items()}, "e2e", round(d["e2e"]["ms_per_step"],1), "cpu1", round(d["cpu_baseline"]["value"]))
except Exception as e:
    print("$f", "FAILED", e)
PY
done
cat $O/${TAG}_tests.log
