#!/bin/bash
# quick regression: GPU tests + smoke + a 10 000-wavelength C5-shape step
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/quick_tests.log 2>&1; echo "tests rc=$?"
tail -4 gpurun_out/quick_tests.log
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py --nwavel 10000 --steps 3 --warmup 1 --no-other-configs --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_quick.json
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_quick.json').read())
print(round(d['value']), {k:round(x,1) for k,x in d['kernel_ms_per_step'].items()})
PY
