#!/bin/bash
# round 2, call V (1 GPU): deferred solution update on/off under sustained load; its kernels alone; the new GPU tests
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tools/ncu_kernels.py 16384 update_xr,update_r,update_xxr 10 > gpurun_out/r2v_defer.txt 2>&1
step() { echo "== sustained $1" >> gpurun_out/r2v_defer.txt; env $1 timeout 600 python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ms_per_step %.2f  its %s clocks %s launches %d' % (d['ms_per_step'], d['run']['pcg_iterations'][:2], d['clocks']['sm_mhz'], d['gpu_launches']))" >> gpurun_out/r2v_defer.txt; }
step IMPEC_DEFER_X=1
step IMPEC_DEFER_X=0
step IMPEC_DEFER_X=1
step IMPEC_DEFER_X=0
cat gpurun_out/r2v_defer.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "deferred or ragged or tiny or anisotropic" 2>&1 | tail -3
