#!/bin/bash
# One GPU-box pass used during development: parity tests, a short headline bench, a C3-shape (two-stream) bench.
# Run as: gpurun --timeout 900 -- 'bash tools/gpu_check.sh'
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/gputests.log
python bench.py --nwavel 4000 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err
python bench.py --nstr 2 --layers 60 --nlos 2 --wf 0 --nwavel 200000 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err
tail -3 gpurun_out/gputests.log
