mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for n in 1 2; do
  if [ $n -eq 1 ]; then python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1d_n$n.json 2> gpurun_out/bench_r1d_n$n.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus $n --steps 3 --warmup 3 > gpurun_out/bench_r1d_n$n.json 2> gpurun_out/bench_r1d_n$n.err; fi
  tail -c 300 gpurun_out/bench_r1d_n$n.err | grep -v OMP
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_r1d_n$n.json').read().strip().splitlines()[-1])
    print('N=$n value',d['value'],'ms/step',d['ms_per_step'],'its',d['config']['pcg_iterations'],'e2e',d['e2e']['value'] if d['e2e'] else None,'vcycle',d['vcycle_ms'],'launches',d['gpu_launches'])
    print({k:round(v,3) for k,v in d['roofline']['kernel_ms'].items()})
except Exception as e: print('parse fail', e)
PY
done
