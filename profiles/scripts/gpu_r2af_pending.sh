#!/bin/bash
# NOT RUN (no GPU minutes were left): A/B of the host binding of bench.py's N > 1 arm (impec_bind_host_to_device), the one change of the
# closing session that can move a number the driver reports (e2e at N = 2 / 4 / 8; the device-timed value does not depend on it).
# Needs 8 GPUs:  gpurun --gpus 8 --timeout 900 -- bash profiles/scripts/gpu_r2af_pending.sh
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for bind in 1 0; do
    IMPEC_BENCH_BIND=$bind timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29700 \
        bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2af_bench_parallel_n8_bind$bind.json 2> gpurun_out/r2af_bench_parallel_n8_bind$bind.err
    python - <<PY
import json
l = json.loads(open("gpurun_out/r2af_bench_parallel_n8_bind$bind.json").read().strip().splitlines()[-1])
print("bind=$bind", "ms/step", round(l["ms_per_step"], 2), "e2e ms/step", round(l["e2e"]["ms_per_step"], 2), l["run"].get("host_binding_rank0"))
PY
done
