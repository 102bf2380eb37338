#!/bin/bash
# Multi-GPU bench line exactly as the driver launches it: bash tools/gpu_scale.sh <N> <tag>
n=$1; tag=${2:-r01_v8}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/bench_${tag}_${n}gpu.json 2> gpurun_out/bench_${tag}_${n}gpu.err
tail -c 600 gpurun_out/bench_${tag}_${n}gpu.json
