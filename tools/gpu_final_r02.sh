#!/bin/bash
# Round-2 measurement pass on one B200: parity tests, smoke, bench lines (own arm + reference arm), single-line benches of
# configs 3 and 4, the ncu launch list of a short bench and `ncu --set full` captures (C5 step, limb kernels, two-stream).
# Run as: gpurun --timeout 1800 -- 'bash tools/gpu_final_r02.sh <tag>'
tag=${1:-r02_final}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/gputests_$tag.log
python __graft_entry__.py smoke > gpurun_out/smoke_$tag.txt 2>&1
python bench.py > gpurun_out/bench_${tag}_default.json 2> gpurun_out/bench_${tag}_default.err
python bench.py --impl reference > gpurun_out/bench_${tag}_reference_arm.json 2> gpurun_out/bench_${tag}_reference_arm.err
python bench.py --config c4 > gpurun_out/bench_${tag}_c4.json 2> gpurun_out/bench_${tag}_c4.err
python bench.py --config c3 > gpurun_out/bench_${tag}_c3.json 2> gpurun_out/bench_${tag}_c3.err
python bench.py --nwavel 2000 --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/bench_${tag}_2k.json 2>/dev/null && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 48 -c 64 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --nwavel 2000 --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/ncu_launch_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -c 19 -f -o gpurun_out/prof_${tag}_all \
    python bench.py --nwavel 600 --steps 1 --warmup 1 --no-cpu-baseline --no-other-configs > gpurun_out/ncu_full_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_limb|k_bvp_multi" -c 7 -f -o gpurun_out/prof_${tag}_c4 \
    python bench.py --config c4 --nwavel 1000 --steps 1 --warmup 1 > gpurun_out/ncu_c4_$tag.log 2>&1
tail -2 gpurun_out/gputests_$tag.log; tail -1 gpurun_out/smoke_$tag.txt
