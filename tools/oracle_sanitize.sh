#!/bin/bash
# The oracle under AddressSanitizer + UndefinedBehaviorSanitizer: builds a sanitised copy of oracle/bsg_oracle.cpp in a
# scratch directory and runs every CPU test that exercises the oracle against it (reference fixture sweep, golden
# vectors, CSI/BAI queries, the randomised scenarios).  A checker with undefined behaviour of its own is no checker:
# the reference's out-of-bounds increment for zero-width coverage regions (src/bamsignals.cpp:420-432) was found this
# way.  libstdc++ is preloaded beside libasan so that C++ exceptions thrown inside the oracle unwind under Python.
set -eu
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=${1:-/tmp/bsg_oracle_asan}
mkdir -p "$OUT"
g++ -O1 -g -std=c++17 -fPIC -shared -fsanitize=address,undefined -fno-sanitize-recover=undefined \
    -o "$OUT/liboracle.so" "$ROOT/oracle/bsg_oracle.cpp" -lz -lpthread
LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libstdc++.so.6)" ASAN_OPTIONS=detect_leaks=0 \
python - "$ROOT" "$OUT" <<'PY'
import sys
root, out = sys.argv[1], sys.argv[2]
sys.path[:0] = [root, root + "/tests", root + "/tools"]
import oracle_api as O
O._PATH = out + "/liboracle.so"
import pytest
sys.exit(pytest.main(["-x", "-q", "-s", "-m", "not gpu", "-p", "no:cacheprovider"] +
                     [f"{root}/tests/{t}" for t in ("test_oracle_pin.py", "test_random_differential.py", "test_oracle_synth.py", "test_csi_index.py")]))
PY
