#!/bin/bash
# parity tests + the default bench line (end-to-end figure) with and without copy/compute overlap
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/gputests.log
python bench.py --no-cpu-baseline > gpurun_out/bench_overlap1.json 2> gpurun_out/bench_overlap1.err
SK_B200_OVERLAP=0 python bench.py --no-cpu-baseline > gpurun_out/bench_overlap0.json 2> gpurun_out/bench_overlap0.err
tail -2 gpurun_out/gputests.log
