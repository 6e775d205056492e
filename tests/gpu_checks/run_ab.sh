#!/bin/bash
# A/B timing of two builds of libdwn_b200.so on ONE box, interleaved (box-to-box and power-cap noise is +-0.15 ms per step,
# more than most single changes).  tests/gpu_checks/build/libdwn_old.so = the revision to compare against.
mkdir -p gpurun_out
OLD=tests/gpu_checks/build/libdwn_old.so
run() { # tag, env...
  tag=$1; shift
  env "$@" python bench.py --no-cpu-baseline --no-extras --steps 20 --warmup 5 --e2e-steps 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline()); print('$tag', round(d['ms_per_step'],3), 'ms', round(d['value'],1), 'clips/s', d['clocks']['reasons'])"
}
for round in 1 2 3; do
  run old DWN_LIB=$OLD
  run new X=1
done
