python tests/gpu_checks/kbench.py > gpurun_out/kbench_v1.txt 2>&1; cat gpurun_out/kbench_v1.txt
ncu --set full --clock-control none --import-source on -k regex:"sdw_fwd|sdw_bwd|bn_bwd_apply|tdw_fwd" -o gpurun_out/prof_dw_v1 python tests/gpu_checks/kbench.py blk0 blk1 --ncu > gpurun_out/ncu_v1.log 2>&1; tail -3 gpurun_out/ncu_v1.log; ls -la gpurun_out/
