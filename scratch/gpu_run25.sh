mkdir -p gpurun_out
n=$(nvidia-smi -L | wc -l); echo gpus $n
timeout 600 python -m pytest tests/test_gpu_multirank.py -x -q 2>&1 | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29701 scratch/gpu_comm_diag.py 16384 2>&1 | grep -v "^$\|OMP_NUM\|\*\*\*\*" | tail -28
for k in 8 4 2; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $k --master-addr 127.0.0.1 --master-port 2971$k bench.py --gpus $k --steps 5 --warmup 3 > gpurun_out/bench_r1m_n$k.json 2> gpurun_out/bench_r1m_n$k.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_r1m_n$k.json').read().strip().splitlines()[-1])
print('N=$k value',d['value'],'ms/step',d['ms_per_step'],'its',d['config']['pcg_iterations'],'e2e',d['e2e']['value'],'vcycle',d['vcycle_ms'],'launches',d['gpu_launches'])
print({k:round(v,3) for k,v in d['roofline']['kernel_ms'].items()})
PY
done
