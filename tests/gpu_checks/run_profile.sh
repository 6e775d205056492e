#!/bin/bash
# round-1 final evidence: ncu launch list of one bench step + full capture of the dominant kernel (sdw_bwd, 9 launches)
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/launches_bench_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --cuda-profiler-step \
  > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu list rc=$?"; wc -l gpurun_out/launches_bench_final.csv
timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"sdw_bwd_v" \
  -o gpurun_out/prof_sdw_bwd_r1final2 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --cuda-profiler-step \
  > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_full.log
ncu -i gpurun_out/prof_sdw_bwd_r1final2.ncu-rep --page raw --csv > gpurun_out/prof_sdw_bwd_r1final2_raw.csv 2>/dev/null
ls -la gpurun_out/*.ncu-rep; du -sh gpurun_out
