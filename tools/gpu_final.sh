#!/bin/bash
# Round-end measurement pass on one B200: parity tests, smoke, bench lines (own arm + reference arm), the ncu launch
# list of a short bench and one `ncu --set full` capture of every kernel of a 600-wavelength step.
# Run as: gpurun --timeout 1500 -- 'bash tools/gpu_final.sh <tag>'
tag=${1:-r01_v8}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/gputests_$tag.log
python __graft_entry__.py smoke > gpurun_out/smoke_$tag.txt 2>&1
python bench.py > gpurun_out/bench_${tag}_default.json 2> gpurun_out/bench_${tag}_default.err
python bench.py --impl reference > gpurun_out/bench_${tag}_reference_arm.json 2> gpurun_out/bench_${tag}_reference_arm.err
python bench.py --nwavel 2000 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${tag}_2k.json 2>/dev/null && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 64 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --nwavel 2000 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -c 18 -f -o gpurun_out/prof_${tag}_all \
    python bench.py --nwavel 600 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_$tag.log 2>&1
tail -2 gpurun_out/gputests_$tag.log; cat gpurun_out/smoke_$tag.txt | tail -1
