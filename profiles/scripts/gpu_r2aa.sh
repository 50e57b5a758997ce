#!/bin/bash
# round 2, call AA (1 GPU): host scan of q with dynamic 16-row chunks + pinned staging: e2e breakdown, scattered-source parity on the GPU, bench line
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tools/e2e_breakdown.py 16384 > gpurun_out/r2aa_e2e_breakdown.txt 2>&1; cat gpurun_out/r2aa_e2e_breakdown.txt
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "source or impes_step or one_cell" 2>&1 | tail -3
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r2aa_bench_parallel.json 2> gpurun_out/r2aa_bench_parallel.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2aa_bench_parallel.json'))
print("ms_per_step", d["ms_per_step"], "e2e", d["e2e"])
PY
tail -3 gpurun_out/r2aa_bench_parallel.err
