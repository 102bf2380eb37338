#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_thermal.py tests/test_limb.py -m gpu -x -q > gpurun_out/jacobi_tests.log 2>&1; echo "tests rc=$?"
tail -4 gpurun_out/jacobi_tests.log
for v in rolled unrolled; do
  SK_B200_JACOBI=$v python bench.py --nwavel 10000 --steps 3 --warmup 1 --no-other-configs --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_jacobi_$v.json
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_jacobi_$v.json').read())
print('$v', round(d['value']), {k:round(x,1) for k,x in d['kernel_ms_per_step'].items()})
PY
done
