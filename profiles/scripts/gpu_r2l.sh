#!/bin/bash
# round 2, call L (1 GPU): L2 prefetch on/off under SUSTAINED load (whole bench steps, power cap active), back to back in one session
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
run() { echo "== $1" >> gpurun_out/r2l_sustained.txt; env $1 timeout 600 python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ms_per_step %.2f  clocks %s  kernel_ms %s' % (d['ms_per_step'], d['clocks'], {k: round(v, 3) for k, v in d['roofline']['kernel_ms'].items()}))" >> gpurun_out/r2l_sustained.txt; }
run "IMPEC_L2_PREFETCH=4"
run "IMPEC_L2_PREFETCH=0 IMPEC_SWEEP_PREFETCH=0"
run "IMPEC_L2_PREFETCH=4,0,0,0"
run "IMPEC_L2_PREFETCH=4,4,0,4"
run "IMPEC_L2_PREFETCH=4,0,4,4"
run "IMPEC_L2_PREFETCH=0 IMPEC_SWEEP_PREFETCH=0"
run "IMPEC_L2_PREFETCH=4"
cat gpurun_out/r2l_sustained.txt
