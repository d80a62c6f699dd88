set -x
export TQ_BENCH_EXTRAS=none
N=${NGPU:-8}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 1 --workload c5 > gpurun_out/bench_n${N}_c5.json 2> gpurun_out/bench_n${N}_c5.err; tail -c 300 gpurun_out/bench_n${N}_c5.err; grep "^{" gpurun_out/bench_n${N}_c5.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['value'], d['ms_per_step'], d['amplitude'])"
