#!/bin/bash
# parity + short bench + one `ncu --set full` capture of the kernels matching a regex: bash tools/gpu_prof_k.sh <regex> <tag>
bash tools/gpu_check.sh
ncu --set full --clock-control none --import-source on -k regex:"$1" -c 2 -f -o gpurun_out/prof_$2 \
    python bench.py --nwavel 600 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_$2.log 2>&1
tail -2 gpurun_out/ncu_$2.log
