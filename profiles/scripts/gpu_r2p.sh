#!/bin/bash
# round 2, call P (1 GPU): rotating-register-slot marching kernels (tools/studies/ab/librot.so, -DIMPEC_MARCH_ROT) against the shifting pipelines,
# with the CTA-per-SM variants that avoid their spills; kernels alone and whole steps under sustained load
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
cp impec2d_b200/libimpec2d_b200.so /tmp/main.so
K=vcycle,fused_pre,fused_post,fused_plain
one() { echo "== lib=$1 IMPEC_MARCH_VARIANT=$2" >> gpurun_out/r2p_rot.txt; cp $3 impec2d_b200/libimpec2d_b200.so; IMPEC_MARCH_VARIANT=$2 timeout 300 python tools/ncu_kernels.py 16384 $K 10 >> gpurun_out/r2p_rot.txt 2>&1; }
one main 1 /tmp/main.so
one main 3 /tmp/main.so
one main 5 /tmp/main.so
one rot 1 tools/studies/ab/librot.so
one rot 3 tools/studies/ab/librot.so
one rot 5 tools/studies/ab/librot.so
one rot 7 tools/studies/ab/librot.so
step() { echo "== sustained lib=$1 IMPEC_MARCH_VARIANT=$2" >> gpurun_out/r2p_rot.txt; cp $3 impec2d_b200/libimpec2d_b200.so; IMPEC_MARCH_VARIANT=$2 timeout 600 python bench.py --steps 8 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ms_per_step %.2f  its %s clocks %s  kernel_ms %s' % (d['ms_per_step'], d['run']['pcg_iterations'][:2], d['clocks']['sm_mhz'], {k: round(v, 3) for k, v in d['roofline']['kernel_ms'].items() if k in ('vcycle','fused_pre','fused_post','fused_plain')}))" >> gpurun_out/r2p_rot.txt; }
step main 1 /tmp/main.so
step rot 3 tools/studies/ab/librot.so
step rot 7 tools/studies/ab/librot.so
step main 1 /tmp/main.so
step rot 3 tools/studies/ab/librot.so
cp /tmp/main.so impec2d_b200/libimpec2d_b200.so
cat gpurun_out/r2p_rot.txt
