#!/bin/bash
# Round-2 first GPU pass: parity tests, the default (C5) bench, the reference arm, a launch list.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/gputests.log
tail -3 gpurun_out/gputests.log
( time python bench.py ) > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
tail -c 600 gpurun_out/bench_default.err
( time python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
python bench.py --nwavel 2000 --steps 2 --warmup 1 --no-cpu-baseline --no-other-configs > gpurun_out/bench_2k.json 2> gpurun_out/bench_2k.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02_v1.csv \
  python bench.py --nwavel 2000 --steps 2 --warmup 1 --no-cpu-baseline --no-other-configs > gpurun_out/ncu_launch.log 2>&1
head -c 1500 gpurun_out/bench_default.json
