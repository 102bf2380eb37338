#!/bin/bash
# short bench at several workspace budgets: bash tools/gpu_ws.sh 48 96 140
mkdir -p gpurun_out
for g in "$@"; do
  python bench.py --workspace-gb $g --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ws_$g.json 2> gpurun_out/bench_ws_$g.err
done
