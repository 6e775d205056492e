set -x
python __graft_entry__.py smoke 2>&1 | tail -5
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/kernels_r1a.csv > gpurun_out/bench_r1a.json 2> gpurun_out/bench_r1a.err; echo "bench rc=$?"
tail -5 gpurun_out/bench_r1a.err
cat gpurun_out/bench_r1a.json
cat gpurun_out/kernels_r1a.csv
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
