#!/bin/bash
# NOT RUN (the round's GPU minutes were spent before this session): the GPU-side confirmation of what the last CPU-only session added.
# Everything below passed on CPU ranks (emulated kernels, gloo / file communicator); the NCCL / peer-memory build uses the same
# exchange_rows / comm_allreduce calls.  Needs 2 GPUs:  gpurun --gpus 2 --timeout 1500 -- bash profiles/scripts/gpu_r2ae_pending.sh
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
# 1. slabs: sequential splitting modes, both self tests, parallel.cfg through the C++ driver with the collective projection
timeout 1200 python -m pytest tests/test_gpu_multirank.py -x -q -m gpu > gpurun_out/r2ae_pytest_multirank_n2.txt 2>&1
tail -3 gpurun_out/r2ae_pytest_multirank_n2.txt
# 2. one GPU: the whole parity suite (k_project_parallel and k_cn9_apply changed their argument lists)
timeout 1500 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_multirank.py > gpurun_out/r2ae_pytest_gpu.txt 2>&1
tail -3 gpurun_out/r2ae_pytest_gpu.txt
# 3. compute-sanitizer (memcheck + racecheck) on a small IMPES step with dispersion through the C ABI: never run on the GPU so far; the CPU
#    twin is tools/sanitize_emu.sh (AddressSanitizer / UBSan on the emulated kernels, profiles/r2ae_sanitizers.txt)
for tool in memcheck racecheck; do
    timeout 900 compute-sanitizer --tool $tool --error-exitcode 1 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "impes_step and not production" > gpurun_out/r2ae_sanitizer_$tool.txt 2>&1
    tail -4 gpurun_out/r2ae_sanitizer_$tool.txt
done
