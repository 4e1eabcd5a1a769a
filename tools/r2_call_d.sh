#!/bin/bash
# Round 2, call D: (1) which page-locking of the BAM mapping the box accepts + raw H2D / memcpy rates,
# (2) ncu source-level capture of the warp-specialised inflate kernel (second launch of a C2 1/5-scale call).
set -u
O=gpurun_out
mkdir -p $O
python - <<'PY'
import sys
sys.path[:0] = ['.', 'tools']
import workloads as WL
from bench import data_dir
print(WL.make_bam('c2', 0.5, data_dir()))
PY
./tools/probes/pin_probe /dev/shm/bsg_bench/c2_g0.5_d1_compact_s0.05_u0_l1.bam > $O/r2d_pin_probe_shm.txt 2>&1
cat $O/r2d_pin_probe_shm.txt
cp /dev/shm/bsg_bench/c2_g0.5_d1_compact_s0.05_u0_l1.bam /tmp/probe.bam && ./tools/probes/pin_probe /tmp/probe.bam > $O/r2d_pin_probe_disk.txt 2>&1
cat $O/r2d_pin_probe_disk.txt; rm -f /tmp/probe.bam
nproc; free -g | head -2
ncu --set full --import-source on --clock-control none -k regex:k_inflate_ws -s 1 -c 1 -f -o $O/r2d_k_inflate_ws_c2_g0.5 \
    python tools/e2e_ab.py --gscale 0.5 --reps 1 > $O/r2d_ncu.log 2>&1
tail -3 $O/r2d_ncu.log; ls -la $O/*.ncu-rep | tail -2
