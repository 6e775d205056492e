#!/bin/bash
# 2-GPU checks of the data-parallel path; every command under its own short timeout.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 150 $TR --master-port 29521 tests/gpu_checks/check_dp.py > gpurun_out/check_dp.log 2>&1
echo "check_dp rc=$?"; grep -v "^\*\|OMP_NUM" gpurun_out/check_dp.log | tail -8
# rank 0's batch (seed 1004) has no sample of mouse 2, rank 1's has all ten mice
timeout 200 $TR --master-port 29522 bench.py --gpus 2 --steps 10 --warmup 3 --seed-base 1004 --no-cpu-baseline \
  > gpurun_out/bench_dp2_absent.json 2> gpurun_out/bench_dp2_absent.err
echo "bench dp2 (absent mouse) rc=$?"; echo "stdout lines: $(wc -l < gpurun_out/bench_dp2_absent.json)"; head -c 300 gpurun_out/bench_dp2_absent.json; echo
python -c "
import json; d=json.loads(open('gpurun_out/bench_dp2_absent.json').read()); print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'])"
