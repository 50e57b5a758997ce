mkdir -p gpurun_out
timeout 600 python scratch/gpu_levels.py 16384 2>&1 | tail -14
