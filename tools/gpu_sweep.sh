#!/bin/bash
# parity tests + short bench at a few settings of one environment switch: bash tools/gpu_sweep.sh VAR v1 v2 ...
var=$1; shift
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/gputests.log
for v in "$@"; do
    env $var=$v python bench.py --nwavel 4000 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_sweep_$v.json 2> gpurun_out/bench_sweep_$v.err
done
tail -3 gpurun_out/gputests.log
