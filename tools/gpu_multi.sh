#!/bin/bash
# usage: tools/gpu_multi.sh N [timeout_s=240] [extra bench args...]; N GPUs are charged N x the wall time: keep the timeout short
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi -L > gpurun_out/multi_env.txt
timeout ${2:-240} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 3 --warmup 2 --no-cpu-baseline --watchdog ${2:-240} ${@:3} > gpurun_out/bench_n$N.txt 2> gpurun_out/bench_n$N.err
echo "rc=$?" >> gpurun_out/bench_n$N.err
tail -n 5 gpurun_out/bench_n$N.err; tail -c 1500 gpurun_out/bench_n$N.txt
