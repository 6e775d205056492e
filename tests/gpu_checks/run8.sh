#!/bin/bash
# round-1 final evidence: GPU tests, eager-on-B200 yardstick, ncu launch list + full capture of the dominant kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --tb=short 2>&1 | grep -E "^E  [^ +]|^tests|Error|passed|failed" | cut -c1-250 | head -30
timeout 600 python tests/gpu_checks/eager_b200.py > gpurun_out/eager_b200.log 2>&1; echo "eager rc=$?"; grep -v "^\*\|OMP" gpurun_out/eager_b200.log | tail -14
# launch list of the bench command: exactly one steady-state train step (cudaProfilerStart/Stop range)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --cuda-profiler-step \
  > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu list rc=$?"; wc -l gpurun_out/launches_bench.csv
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:sdw_bwd_v3 \
  -o gpurun_out/prof_sdw_bwd_r1final -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --cuda-profiler-step \
  > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_full.log
