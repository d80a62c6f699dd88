set -x
export TQ_BENCH_EXTRAS=none
N=${NGPU:-8}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 1 --workload c5 > gpurun_out/bench_n${N}_c5.json 2> gpurun_out/bench_n${N}_c5.err; tail -c 300 gpurun_out/bench_n${N}_c5.err; tail -1 gpurun_out/bench_n${N}_c5.json | cut -c 1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n${N}_c2.json 2> gpurun_out/bench_n${N}_c2.err; tail -c 300 gpurun_out/bench_n${N}_c2.err; tail -1 gpurun_out/bench_n${N}_c2.json | cut -c 1-200
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 3 --warmup 3 --workload c3 > gpurun_out/bench_n${N}_c3.json 2> gpurun_out/bench_n${N}_c3.err; tail -c 300 gpurun_out/bench_n${N}_c3.err; tail -1 gpurun_out/bench_n${N}_c3.json | cut -c 1-200
