#!/bin/bash
# Build one variant of the inflate micro-benchmark (tools/inflate_bench.cu) into tools/probes/ib_<name>.
#   tools/ib_build.sh NAME [GIT_REV|work] [extra nvcc flags...]     (work = the working tree)
set -eu
NAME=$1; REV=${2:-work}; shift; shift || true
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=$ROOT/tools/probes/ib_$NAME
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --use_fast_math -lineinfo"
if [ "$REV" = work ]; then
  nvcc $FLAGS "$@" -o "$OUT" "$ROOT/tools/inflate_bench.cu"
else
  T=$(mktemp -d)
  mkdir -p "$T/tools" "$T/bamsignals_b200" "$T/include"
  git -C "$ROOT" archive "$REV" bamsignals_b200/csrc include | tar -x -C "$T"
  cp "$ROOT/tools/inflate_bench.cu" "$T/tools/"
  nvcc $FLAGS "$@" -o "$OUT" "$T/tools/inflate_bench.cu"
  rm -rf "$T"
fi
echo "built $OUT"
