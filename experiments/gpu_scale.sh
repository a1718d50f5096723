#!/bin/bash
# multi-GPU bench line at N = $1 GPUs (torchrun, one rank per GPU), default workload / mode
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps ${STEPS:-3} --warmup 3 ${EXTRA} > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -2 gpurun_out/bench_n$N.err; cat gpurun_out/bench_n$N.json
