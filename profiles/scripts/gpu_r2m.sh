#!/bin/bash
# round 2, call M (1 GPU): GPU tests with the anisotropic-dispersion path, smoke(), default bench line, ncu captures of the kernels missing
# from r2j_traffic.json (k_update_xr, k_muscl_x, k_muscl_y2)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -rA 2>&1 | tail -95 ) > gpurun_out/r2m_pytest.txt 2>&1
tail -4 gpurun_out/r2m_pytest.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2m_smoke.txt 2>&1; tail -2 gpurun_out/r2m_smoke.txt
( time timeout 900 python bench.py ) > gpurun_out/r2m_bench_default.json 2> gpurun_out/r2m_bench_default.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_update_xr|k_muscl_x|k_muscl_y2" -c 4 -f -o /tmp/r2m_hot python tools/ncu_kernels.py 16384 update_xr,sweep_x,sweep_y > gpurun_out/r2m_ncu.log 2>&1
python tools/ncu_summary.py /tmp/r2m_hot.ncu-rep > gpurun_out/r2m_ncu_summary.txt 2>&1
cut -c1-300 gpurun_out/r2m_bench_default.json; tail -5 gpurun_out/r2m_bench_default.err; grep -n "^==\|gpu__time\|dram__bytes" gpurun_out/r2m_ncu_summary.txt | cut -c1-160
