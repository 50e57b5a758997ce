mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
for pol in "0 0" "1e30 0" "1e30 1e30"; do
  set -- $pol
  echo "== TH32_BELOW=$1 TH16_BELOW=$2"
  IMPEC_TH32_BELOW=$1 IMPEC_TH16_BELOW=$2 timeout 600 python scratch/gpu_levels.py 16384 2>&1 | tail -12
done
