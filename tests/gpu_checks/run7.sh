( time python bench.py ) > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "rc=$?"; tail -6 gpurun_out/bench_default.err; cat gpurun_out/bench_default.json | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print({k:d[k] for k in ['value','ms_per_step','steps','warmup','gpu_launches','clocks','cpu_baseline']}); print(d['e2e']); print(d['roofline'])"
( time python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -4 gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
nproc; lscpu | grep -E "Model name|^CPU\(s\)"
