mkdir -p gpurun_out
n=$(nvidia-smi -L | wc -l); echo gpus $n
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29701 scratch/gpu_comm_diag.py 16384 2>&1 | grep -v "^$\|OMP_NUM\|\*\*\*\*" | tail -20
