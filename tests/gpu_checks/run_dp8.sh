#!/bin/bash
# 8-GPU weak-scaling bench line (own short timeout: a hang must not burn the GPU budget)
mkdir -p gpurun_out
N=${1:-8}
timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 \
  bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_dp$N.json 2> gpurun_out/bench_dp$N.err
echo "bench dp$N rc=$?"; tail -c 1200 gpurun_out/bench_dp$N.json
