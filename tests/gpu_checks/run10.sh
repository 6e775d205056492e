#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:"block_out|block_in_bwd|block_bwd_reduce|sdw_fwd_v3|tdw_fwd|tdw_bwd_bulk" \
  -o gpurun_out/prof_trunk_r1 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --cuda-profiler-step \
  > gpurun_out/ncu_trunk.log 2>&1; echo "ncu trunk rc=$?"; tail -2 gpurun_out/ncu_trunk.log
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:gemm_tc -c 16 \
  -o gpurun_out/prof_gemm_r1 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --cuda-profiler-step \
  > gpurun_out/ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"; tail -2 gpurun_out/ncu_gemm.log
ls -la gpurun_out/*.ncu-rep
