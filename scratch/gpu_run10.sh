mkdir -p gpurun_out
nvidia-smi -L | wc -l
python -m pytest tests/test_gpu_multirank.py -x -q 2>&1 | tail -4
for n in 8 4; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2964$n bench.py --gpus $n --steps 3 --warmup 3 > gpurun_out/bench_r1f_n$n.json 2> gpurun_out/bench_r1f_n$n.err
  tail -c 400 gpurun_out/bench_r1f_n$n.err | grep -v -i "omp\|\*\*\*"
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_r1f_n$n.json').read().strip().splitlines()[-1])
    print('N=$n value',d['value'],'ms/step',d['ms_per_step'],'its',d['config']['pcg_iterations'],'e2e',d['e2e']['value'] if d['e2e'] else None,'vcycle',d['vcycle_ms'],'launches',d['gpu_launches'])
    print({k:round(v,3) for k,v in d['roofline']['kernel_ms'].items()})
except Exception as e: print('parse fail', e)
PY
done
