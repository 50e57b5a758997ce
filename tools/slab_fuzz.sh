#!/bin/bash
# Slab decomposition on awkward extents, without a GPU: tests/multirank_worker.py (emulated kernels, gloo ranks; every rank compares its rows
# with a one-rank run of the same problem) on grids whose extents are 2 x odd, 4 x odd, ..., with the marching kernels forced, dispersion,
# the horizontal front, sequential splittings and both self tests.  Odd y_cells must be REFUSED at context creation (one multigrid level only).
# Last run: 15 accepted cases rc 0, 6 odd-extent cases refused with the documented message.   Usage: tools/slab_fuzz.sh
ROOT=$(cd "$(dirname "$0")/.." && pwd)
fails=0
run() { local expect=$1 w=$2 n1=$3 n2=$4 st=$5 d=$6 s=$7 port=$8; shift 8
  env OMP_NUM_THREADS=1 IMPEC_DIST_MIN_CELLS=0 "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $w --master-addr 127.0.0.1 --master-port $port \
     "$ROOT/tests/multirank_worker.py" $n1 $n2 $st $d $s > /tmp/slab_fuzz_$port.log 2>&1
  local rc=$? what=ok
  if [ $expect = refuse ]; then grep -q "no coarser level to replicate" /tmp/slab_fuzz_$port.log || what=FAIL
  else [ $rc -eq 0 ] || what=FAIL; fi
  echo "$what ($expect) ranks=$w ${n1}x${n2} stype=$st diffusion=$d steps=$s $*"; [ $what = ok ]; }
batch() { local pids=(); for c in "$@"; do eval "run $c" & pids+=($!); done; for p in "${pids[@]}"; do wait $p || fails=$((fails+1)); done; }
batch "pass 2 44 26 1 0 2 29821" "pass 4 88 54 1 1 1 29822" "pass 2 36 70 1 1 2 29823 IMPEC_MARCH_MIN_CELLS=0" "pass 2 60 22 2 1 2 29824"
batch "pass 4 104 46 1 1 2 29825 IMPEC_TEST_ALPHA_T=0.2" "pass 8 144 38 1 0 1 29826" "pass 2 52 34 1 0 2 29827 IMPEC_TEST_SPLITTING=3"
batch "pass 2 28 28 1 0 4 29828 IMPEC_TEST_PERIODIC=muscl" "pass 2 36 54 1 0 1 29829 IMPEC_TEST_ELLIPTIC=2 IMPEC_TEST_NU=0.5" "pass 4 72 72 0 0 2 29830 IMPEC_MARCH_MIN_CELLS=0 IMPEC_MARCH_ROWS=8"
batch "pass 4 96 50 1 1 1 29802" "pass 4 64 90 1 0 2 29804 IMPEC_MARCH_MIN_CELLS=0 IMPEC_MARCH_ROWS=8" "pass 2 36 36 0 0 2 29805" "pass 4 128 128 0 1 1 29806 IMPEC_MARCH_MIN_CELLS=0"
batch "pass 2 48 38 1 0 2 29801" "refuse 2 80 23 1 1 2 29803" "refuse 2 52 31 2 1 2 29807" "refuse 2 44 29 1 1 2 29808"
batch "refuse 4 72 41 1 0 2 29809" "refuse 2 20 77 1 0 2 29810" "refuse 8 128 33 1 0 1 29811"
echo "failures: $fails"; exit $fails
