#!/bin/bash
# round-2 evidence after the TMA-staged sdw_bwd: ncu --set full of the spatial depth-wise kernels at the six C2 block shapes
# (kbench --ncu: one launch each), the ncu launch list of one graph-replayed bench step
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"sdw_" -o gpurun_out/r2_sdw_tma -f \
  python tests/gpu_checks/kbench.py sdw_fwd sdw_bwd --ncu > gpurun_out/ncu_sdw_tma.log 2>&1; echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_sdw_tma.log
ncu -i gpurun_out/r2_sdw_tma.ncu-rep --page raw --csv > gpurun_out/r2_sdw_tma_raw.csv 2>/dev/null
python tests/gpu_checks/ncu_summary.py gpurun_out/r2_sdw_tma_raw.csv gpurun_out/r2_ncu_full_sdw_tma.csv \
  "ncu --set full --clock-control none on tests/gpu_checks/kbench.py sdw_fwd sdw_bwd --ncu (C2 shapes, batch 32): TMA-staged sdw_bwd_v6 / v7"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/launches_bench_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --e2e-steps 1 --cuda-profiler-step \
  > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu list rc=$?"; wc -l gpurun_out/launches_bench_final.csv
python tests/gpu_checks/kbench.py sdw_fwd sdw_bwd tdw > gpurun_out/kbench_r2_tma.txt 2>&1; cat gpurun_out/kbench_r2_tma.txt
rm -f gpurun_out/*.ncu-rep
