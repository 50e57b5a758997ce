set -x
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1a.json 2> gpurun_out/bench_r1a.err; tail -c 600 gpurun_out/bench_r1a.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r1a_ref.json 2>&1
nproc; lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket" 
# launch list of the same command (1 warm-up + 1 step)
ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/launches_r1a.csv python bench.py --steps 1 --warmup 1 --profile-mode > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_launch.log
# full capture of the two top kernels
ncu --set full --clock-control none --import-source on -k regex:k_smooth_color -s 40 -c 3 -o gpurun_out/prof_smooth_r1a python bench.py --steps 1 --warmup 0 --profile-mode > gpurun_out/ncu_full1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 4 -c 2 -o gpurun_out/prof_sweep_r1a python bench.py --steps 1 --warmup 1 --profile-mode > gpurun_out/ncu_full2.log 2>&1
ls -la gpurun_out
