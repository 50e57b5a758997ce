mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -k "fingering" 2>&1 | tail -3
ncu --set full --clock-control none --import-source on -k regex:"k_mg_fused|k_matvec_dot|k_update_xr" -c 10 -o gpurun_out/prof_fused_fine_r1e python scratch/ncu_kernels.py 16384 fused_pre,fused_post,fused_plain,matvec_dot,update_xr > gpurun_out/ncu_full5.log 2>&1
tail -3 gpurun_out/ncu_full5.log
