#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_brdf.py -m gpu -x -q -s > gpurun_out/brdfwf_tests.log 2>&1; echo "rc=$?"
tail -25 gpurun_out/brdfwf_tests.log | cut -c1-300
