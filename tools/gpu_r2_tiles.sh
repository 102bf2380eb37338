#!/bin/bash
# LOS-tiled k_wf_layer_fast + padded limb transposes: parity, then timings
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_limb.py -m gpu -x -q > gpurun_out/tiles_tests.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/tiles_tests.log
for n in 10 24 40 100; do
  python tools/gpu_wf_many_los.py $n 1000 2>&1 | tail -1 >> gpurun_out/wf_many_los.jsonl
  SK_B200_WF_TILE=-1 python tools/gpu_wf_many_los.py $n 1000 2>&1 | tail -1 >> gpurun_out/wf_many_los.jsonl
done
cat gpurun_out/wf_many_los.jsonl | cut -c1-600
python bench.py --config c4 --steps 2 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_c4_tiles.json; cut -c1-900 gpurun_out/bench_c4_tiles.json
python bench.py --steps 2 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_default_tiles.json; cut -c1-1200 gpurun_out/bench_default_tiles.json
