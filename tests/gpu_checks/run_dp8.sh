for N in 8 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_dp$N.json 2> gpurun_out/bench_dp$N.err; echo "N=$N rc=$?"; tail -2 gpurun_out/bench_dp$N.err; python -c "
import json
d=json.load(open('gpurun_out/bench_dp$N.json')); print($N, d['value'], d['ms_per_step'], d['e2e']['value'])"
done
