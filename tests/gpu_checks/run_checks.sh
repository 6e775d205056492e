#!/bin/bash
# one-GPU regression after a kernel change: per-layer / per-kernel checks against torch, the GPU parity tests, an
# optional micro-benchmark (KBENCH="sdw_bwd tdw_bwd") and a short bench with the per-kernel table.
#   gpurun --timeout 900 -- 'bash tests/gpu_checks/run_checks.sh'
mkdir -p gpurun_out
python tests/gpu_checks/check_gemm.py > gpurun_out/gemm.log 2>&1; echo "gemm rc=$?"; grep -E "FAIL|GEMM CHECK" gpurun_out/gemm.log | head
python tests/gpu_checks/check_layers.py > gpurun_out/layers.log 2>&1; echo "layers rc=$?"; grep -E "FAIL|LAYER" gpurun_out/layers.log | head -20
python tests/gpu_checks/check_backward.py > gpurun_out/bwd.log 2>&1; echo "bwd rc=$?"; grep -E "FAIL|BACKWARD|Error|error" gpurun_out/bwd.log | head -20
timeout 900 python -m pytest tests -q -m gpu --tb=short 2>&1 | grep -E "^E  [^ +]|^tests|Error|passed|failed" | cut -c1-250 | head -30
if [ -n "$KBENCH" ]; then python tests/gpu_checks/kbench.py $KBENCH 2>&1 | grep -v "^\*\|OMP" | tee gpurun_out/kbench.txt; fi
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/kernels.csv > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value']); print(d['kernel_table_ms_per_step'])"
