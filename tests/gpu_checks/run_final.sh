#!/bin/bash
# end-of-round validation on one GPU: smoke, GPU tests, default bench (both arms), ncu launch list of one bench step
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python -m pytest tests -q -m gpu --tb=short 2>&1 | grep -E "^E  [^ +]|^tests|Error|passed|failed" | cut -c1-250 | head -30
( time python bench.py ) > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "rc=$?"; tail -4 gpurun_out/bench_default.err; python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_default.json').readline()); print({k:d[k] for k in ['value','ms_per_step','steps','warmup','gpu_launches','clocks','cpu_baseline']}); print(d['e2e']); print(d['roofline'])"
( time python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -4 gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/launches_bench_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --cuda-profiler-step \
  > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu list rc=$?"; wc -l gpurun_out/launches_bench_final.csv
