mkdir -p gpurun_out
timeout 600 python scratch/gpu_cycle_shapes.py 16384 2>&1 | grep -v "^$" | tail -12
