#!/bin/bash
# the default bench line on N GPUs (torchrun), as the driver launches it
N=${1:-2}; tag=${2:-r2v}
mkdir -p gpurun_out
SECONDS=0
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N > gpurun_out/${tag}_bench_${N}gpu.json 2> gpurun_out/${tag}_bench_${N}gpu.err; echo "bench exit $? wall ${SECONDS}s"
tail -c 2500 gpurun_out/${tag}_bench_${N}gpu.json; tail -5 gpurun_out/${tag}_bench_${N}gpu.err
