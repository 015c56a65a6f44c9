#!/bin/bash
# AddressSanitizer pass over the host-only entry points of libgl_b200.so (drift, kick, domain decomposition pieces,
# Peano-Hilbert order, snapshot writer / reader, reorder): builds an instrumented copy of the library under
# oracle/_ref/asan (git-ignored) and runs the CPU tests that call those entry points against it.  No GPU needed.
#   bash scripts/sanitize_host.sh
# (ThreadSanitizer: same recipe with -fsanitize=thread and a small C++ driver instead of Python - a Python process does
#  not start under libtsan; tests/test_host_logic.py::test_host_worker_pool_native runs the worker pool under TSAN.)
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=$ROOT/oracle/_ref/asan
mkdir -p "$OUT"
cd "$ROOT/gadget_leicester_b200/csrc"
for f in gl_api gl_gadget gl_keys gl_tree gl_gravity gl_sph gl_snapshot gl_timestep; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O1 -g -std=c++17 -fmad=false \
       -Xcompiler -fPIC,-pthread,-fsanitize=address,-fno-omit-frame-pointer -c $f.cu -o "$OUT/$f.o" 2>/dev/null &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT/libgl_b200.so" "$OUT"/*.o -lcudart -Xcompiler -fsanitize=address
cat > "$OUT/run.py" <<PY
import sys
sys.path.insert(0, "$ROOT")
from gadget_leicester_b200 import lib
lib.LIB_PATH = "$OUT/libgl_b200.so"
import pytest
T = "$ROOT/tests/"
sys.exit(pytest.main(["-x", "-q", T + "test_snapshot.py", T + "test_timestep.py", T + "test_drift.py", T + "test_domain.py", T + "test_abi.py",
                      "-p", "no:cacheprovider", "-k", "not glue and not reference_host and not integration_loop"]))
PY
LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0:protect_shadow_gap=0 python "$OUT/run.py"
rm -rf "$OUT"
