#!/bin/bash
# Memory / undefined-behaviour check of the KERNEL SOURCE on the CPU: the emulated build (csrc/emu.cpp; device memory = malloc) compiled
# with AddressSanitizer or UBSan, and the emulator parity tests run against it.  Out-of-bounds reads / writes of the global arrays
# (halo rows, pad columns, chunk edges of the marching kernels), misaligned 128-bit accesses and signed overflow in the index arithmetic
# show up here without a GPU (compute-sanitizer needs one).  Usage: tools/sanitize_emu.sh [address|undefined] [pytest args...]
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
KIND=${1:-address}; shift || true
OUT=/tmp/libimpec2d_emu_$KIND.so
cd "$ROOT/impec2d_b200/csrc"
g++ -O1 -g -fsanitize=$KIND -fno-sanitize-recover=all -fno-omit-frame-pointer -ffp-contract=off -march=x86-64-v3 -std=c++17 -fPIC -DIMPEC_EMU \
    -Wno-unknown-pragmas -shared -o $OUT -x c++ impec2d_b200.cu emu.cpp -pthread
cat > /tmp/impec_sanitize_run.py <<PY
import sys
sys.path.insert(0, "$ROOT")
import impec2d_b200._lib as L
_orig = L.library_path
L.library_path = lambda emulated=False: "$OUT" if emulated else _orig(emulated)
import pytest
sys.exit(pytest.main(sys.argv[1:]))
PY
cd "$ROOT"
RT=$(gcc -print-file-name=lib$([ $KIND = address ] && echo asan || echo ubsan).so)
ARGS=("$@"); [ ${#ARGS[@]} -eq 0 ] && ARGS=(tests/test_emu_parity.py tests/test_ref_golden.py tests/test_golden.py -q -m "not gpu")
LD_PRELOAD=$RT ASAN_OPTIONS=detect_leaks=0:halt_on_error=1:detect_stack_use_after_return=0 UBSAN_OPTIONS=print_stacktrace=1:halt_on_error=1 \
    python /tmp/impec_sanitize_run.py "${ARGS[@]}"
