#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/los_skip.jsonl
python tools/gpu_wf_many_los.py 10 2000 2>&1 | tail -1 >> gpurun_out/los_skip.jsonl
SK_B200_LOS_SKIP=0 python tools/gpu_wf_many_los.py 10 2000 2>&1 | tail -1 >> gpurun_out/los_skip.jsonl
cut -c1-420 gpurun_out/los_skip.jsonl
