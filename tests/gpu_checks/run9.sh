#!/bin/bash
mkdir -p gpurun_out
python tests/gpu_checks/check_layers.py > gpurun_out/layers.log 2>&1; echo "layers rc=$?"; grep -E "FAIL|LAYER" gpurun_out/layers.log | head -20
python tests/gpu_checks/check_backward.py > gpurun_out/bwd.log 2>&1; echo "bwd rc=$?"; grep -E "FAIL|BACKWARD|Error|error" gpurun_out/bwd.log | head -20
timeout 900 python -m pytest tests -q -m gpu --tb=short 2>&1 | grep -E "^E  [^ +]|^tests|Error|passed|failed" | cut -c1-250 | head -30
python tests/gpu_checks/kbench.py sdw_fwd 2>&1 | grep -v "^\*\|OMP" | tee gpurun_out/kbench_v10.txt
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/kernels_r1m.csv > gpurun_out/bench_r1m.json 2> gpurun_out/bench_r1m.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_r1m.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r1m.json')); print(d['value'], d['ms_per_step'], d['e2e']['value']); print(d['kernel_table_ms_per_step']); print(d['roofline'])"
