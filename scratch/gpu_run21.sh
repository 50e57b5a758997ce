mkdir -p gpurun_out
n=$(nvidia-smi -L | wc -l); echo gpus $n
timeout 600 python -m pytest tests/test_gpu_multirank.py -x -q 2>&1 | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29701 scratch/gpu_comm_diag.py 16384 2>&1 | grep -v "^$\|OMP_NUM\|\*\*\*\*" | tail -16
IMPEC_P2P=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29702 scratch/gpu_comm_diag.py 16384 2>&1 | grep -v "^$\|OMP_NUM\|\*\*\*\*" | tail -16
