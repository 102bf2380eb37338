#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_thermal.py -m gpu -x -q -s > gpurun_out/thermal_tests.log 2>&1; echo "thermal rc=$?"
tail -15 gpurun_out/thermal_tests.log
python -m pytest tests -m gpu -x -q > gpurun_out/gpu_tests_after_thermal.log 2>&1; echo "all rc=$?"
tail -4 gpurun_out/gpu_tests_after_thermal.log
python bench.py --steps 2 --warmup 1 --no-other-configs 2>&1 | tail -1 > gpurun_out/bench_default_thermal.json; cut -c1-400 gpurun_out/bench_default_thermal.json
