#!/bin/bash
# One `ncu --set full` capture of every kernel of a 600-wavelength C5 step and of the two-stream kernel (C3 shape).
# Run as: gpurun --timeout 1500 -- 'bash tools/gpu_ncu_full.sh <tag>'
tag=${1:-r02_v1}
mkdir -p gpurun_out
python bench.py --nwavel 600 --steps 1 --warmup 1 --no-cpu-baseline --no-other-configs > gpurun_out/bench_600_$tag.json 2> gpurun_out/bench_600_$tag.err || exit 1
ncu --set full --clock-control none --import-source on -c 18 -f -o gpurun_out/prof_${tag}_all \
    python bench.py --nwavel 600 --steps 1 --warmup 1 --no-cpu-baseline --no-other-configs > gpurun_out/ncu_full_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_twostream -c 1 -f -o gpurun_out/prof_${tag}_c3 \
    python bench.py --config c3 --nwavel 200000 --steps 1 --warmup 1 --no-cpu-baseline --no-other-configs > gpurun_out/ncu_c3_$tag.log 2>&1
tail -2 gpurun_out/ncu_full_$tag.log gpurun_out/ncu_c3_$tag.log
