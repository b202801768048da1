#!/bin/sh
# AddressSanitizer + UBSan + LeakSanitizer over the C host rows (ingest.c, radii.c, areas.c, workers.c, host_shim.c):
#   1. the ingest/areas test-suites against an instrumented libfreesasa_b200_host.so, serial and threaded paths forced on;
#   2. a C driver (host_leakcheck.c) that reads, builds trees, writes, splits and frees, with leak detection.
# Usage (dev container, repo root):  sh tests/tools/host_sanitize.sh
set -e
ROOT=$(cd "$(dirname "$0")/../.." && pwd)
CSRC=$ROOT/freesasa_b200/csrc
CC=/usr/bin/gcc
ASAN=$($CC -print-file-name=libasan.so)
TMP=$(mktemp -d)
SRC="$CSRC/host_shim.c $CSRC/radii.c $CSRC/ingest.c $CSRC/areas.c $CSRC/select.c $CSRC/workers.c"
$CC -std=gnu99 -O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer -fPIC -shared -I "$ROOT/include" \
    -o "$TMP/libfreesasa_b200_host.so" $SRC -L "$CSRC" -lfsb200 -Wl,-rpath,"$CSRC" -lm -lpthread
cp "$CSRC/libfreesasa_b200_host.so" "$TMP/plain.so"
trap 'cp "$TMP/plain.so" "$CSRC/libfreesasa_b200_host.so"; touch "$CSRC/libfreesasa_b200_host.so"; rm -rf "$TMP"' EXIT
cp "$TMP/libfreesasa_b200_host.so" "$CSRC/libfreesasa_b200_host.so"; touch "$CSRC/libfreesasa_b200_host.so"
cd "$ROOT"
LD_PRELOAD=$ASAN ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 python -m pytest tests/test_ingest.py tests/test_areas.py tests/test_select.py -x -q -p no:cacheprovider
LD_PRELOAD=$ASAN ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 FREESASA_B200_PARALLEL_MIN_BYTES=1 FREESASA_B200_PARALLEL_MIN_ATOMS=1 \
    FREESASA_B200_THREADS=4 python -m pytest tests/test_ingest.py tests/test_areas.py tests/test_select.py -x -q -p no:cacheprovider
# leak check: the engine entry points are stubbed (no GPU needed; nothing here computes SASA)
printf 'const char *fsb200_last_error(void){return "";}\nint fsb200_lr(){return -1;}\nint fsb200_sr(){return -1;}\nint fsb200_calc_batch(){return -1;}\n' > "$TMP/stub.c"
$CC -std=gnu99 -O1 -g -fsanitize=address,undefined -I "$ROOT/include" "$ROOT/tests/tools/host_leakcheck.c" $SRC "$TMP/stub.c" -o "$TMP/leak" -lm -lpthread
python - "$TMP" <<'PY'
import sys
sys.path.insert(0, ".")
from freesasa_b200 import workloads as w
from tests.test_ingest import EDGE_TEXTS
d = sys.argv[1]
open(d + "/l1.pdb", "w").write(w.pdb_text(3000, seed=3, chains=3, models=3, hydrogens=0.2, hetatm=3, altloc=0.1, unknown=0.1))
open(d + "/l2.pdb", "w", encoding="latin-1").write(EDGE_TEXTS["coords_garbage"] + EDGE_TEXTS["altloc_runs"])
open(d + "/l3.pdb", "w").write(EDGE_TEXTS["short_line_fails"])
open(d + "/l4.pdb", "w").write(EDGE_TEXTS["no_atoms"])
PY
ASAN_OPTIONS=detect_leaks=1 FREESASA_B200_THREADS=1 "$TMP/leak" "$TMP"/l1.pdb "$TMP"/l2.pdb "$TMP"/l3.pdb "$TMP"/l4.pdb
ASAN_OPTIONS=detect_leaks=1 FREESASA_B200_THREADS=4 FREESASA_B200_PARALLEL_MIN_BYTES=1 FREESASA_B200_PARALLEL_MIN_ATOMS=1 \
    "$TMP/leak" "$TMP"/l1.pdb "$TMP"/l2.pdb "$TMP"/l3.pdb "$TMP"/l4.pdb
echo "host sanitize ok"
