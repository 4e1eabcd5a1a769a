#!/bin/bash
# One command for a machine that has R (>= 3.5 with Rcpp, GenomicRanges, testthat; no Rhtslib needed) and a B200:
# builds the reference R package with its C++ engine replaced by the shim over libbamsignals_cuda.so and runs the
# reference's own UNCHANGED test suite (tests/testthat.R, tests/testthat/test_methods.R) against it.
#   rshim/build_and_test_with_r.sh /path/to/bamsignals-checkout        (a clone of lamortenera/bamsignals)
# Nothing of the reference is modified except the three files INTEGRATION.md section 1 names: src/bamsignals.cpp (replaced by
# the shim), src/Makevars (replaced), DESCRIPTION (Rhtslib dropped from LinkingTo / Imports).
# Not run in this repository's CI: R is not installable offline here.
set -euo pipefail
REF=${1:?usage: $0 /path/to/bamsignals-checkout}
HERE=$(cd "$(dirname "$0")/.." && pwd)
python -c "import sys; sys.path.insert(0, '$HERE'); import __graft_entry__ as g; g.build()"
WORK=$(mktemp -d)
cp -r "$REF" "$WORK/bamsignals"
cp "$HERE/rshim/bamsignals_shim.cpp" "$WORK/bamsignals/src/bamsignals.cpp"
sed "s|^BSG_HOME ?=.*|BSG_HOME ?= $HERE|" "$HERE/rshim/Makevars" > "$WORK/bamsignals/src/Makevars"
rm -f "$WORK/bamsignals/src/Makevars.win"
sed -i -e 's/,\? *Rhtslib *([^)]*)//' -e 's/,\? *Rhtslib//' -e '/^SystemRequirements:/d' "$WORK/bamsignals/DESCRIPTION"
export LD_LIBRARY_PATH="$HERE/bamsignals_b200:${LD_LIBRARY_PATH:-}"
R CMD INSTALL --no-test-load "$WORK/bamsignals"
Rscript -e 'library(testthat); library(bamsignals); test_dir(file.path("'"$WORK"'", "bamsignals", "tests", "testthat"), reporter = "summary", stop_on_failure = TRUE)'
echo "reference test suite passed against libbamsignals_cuda.so"
