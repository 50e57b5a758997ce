mkdir -p gpurun_out
for pol in "6e6 6e6" "1e30 6e6"; do
  set -- $pol
  echo "== TH32_BELOW=$1 TH16_BELOW=$2"
  IMPEC_TH32_BELOW=$1 IMPEC_TH16_BELOW=$2 timeout 600 python scratch/gpu_levels.py 16384 2>&1 | head -5
done
python - <<PY
import sys; sys.path.insert(0,'.')
import impec2d_b200 as I
N=16384
P = I.default_params(N1=N, N2=N, x_ub=1.41421356237, y_ub=1.41421356237, stype=0, nu=2/3, eps=1e-14)
c = I.synth_state(P); q = I.init_source(P)
ctx = I.Context(P); ctx.upload_c(c); ctx.upload_q(q); ctx.impes_step(0.0)
print(" ".join("%s %.3f ms" % (k, ctx.profile_kernel(k, 10)) for k in ("pcg_matvec", "update_xr", "fused_pre", "fused_post", "vcycle")))
PY
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r1j_n1.json 2> gpurun_out/bench_r1j_n1.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_r1j_n1.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms/step',d['ms_per_step'],'its',d['config']['pcg_iterations'],'e2e',d['e2e']['value'],'vcycle',d['vcycle_ms'],'launches',d['gpu_launches'])
PY
