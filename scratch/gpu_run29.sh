mkdir -p gpurun_out
n=$(nvidia-smi -L | wc -l); echo gpus $n
timeout 600 python -m pytest tests/test_gpu_multirank.py -x -q 2>&1 | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_r1s_n2.json 2> gpurun_out/bench_r1s_n2.err
tail -c 400 gpurun_out/bench_r1s_n2.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_r1s_n2.json').read().strip().splitlines()[-1])
print('N=2 value',d['value'],'ms/step',d['ms_per_step'],'its',d['config']['pcg_iterations'],'e2e',d['e2e']['value'],'vcycle',d['vcycle_ms'],'launches',d['gpu_launches'])
print({k:round(v,3) for k,v in d['roofline']['kernel_ms'].items()})
PY
