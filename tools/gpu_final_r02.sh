#!/bin/bash
# Round-2 measurement pass on one B200: parity tests, smoke, bench lines (own arm + reference arm), single-line benches of
# configs 3 and 4, the ncu launch list of a short bench and `ncu --set full` captures (C5 step, limb kernels, two-stream).
# Run as: gpurun --timeout 1800 -- 'bash tools/gpu_final_r02.sh <tag>'
tag=${1:-r02_final}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/gputests_$tag.log
python __graft_entry__.py smoke > gpurun_out/smoke_$tag.txt 2>&1
python bench.py > gpurun_out/bench_${tag}_default.json 2> gpurun_out/bench_${tag}_default.err
python bench.py --impl reference > gpurun_out/bench_${tag}_reference_arm.json 2> gpurun_out/bench_${tag}_reference_arm.err
python bench.py --config c4 > gpurun_out/bench_${tag}_c4.json 2> gpurun_out/bench_${tag}_c4.err
python bench.py --config c3 > gpurun_out/bench_${tag}_c3.json 2> gpurun_out/bench_${tag}_c3.err
python bench.py --nwavel 2000 --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/bench_${tag}_2k.json 2>/dev/null && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 48 -c 64 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --nwavel 2000 --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/ncu_launch_$tag.log 2>&1
# the reports themselves stay on the box (gpurun brings back at most 64 MiB): digest them here
ncu --set full --clock-control none --import-source on -c 19 -f -o /tmp/prof_${tag}_all \
    python bench.py --nwavel 600 --steps 1 --warmup 1 --no-cpu-baseline --no-other-configs > gpurun_out/ncu_full_$tag.log 2>&1
python tools/ncu_summary.py /tmp/prof_${tag}_all.ncu-rep > gpurun_out/ncu_${tag}_summary.csv
for k in k_wf_layer_fast k_bvp_v2 k_bvp_tsolve k_layer_post k_eig_jacobi; do
    python tools/ncu_source_top.py /tmp/prof_${tag}_all.ncu-rep $k 12 >> gpurun_out/ncu_${tag}_source_top_lines.txt 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:"k_limb|k_bvp_multi" -c 7 -f -o /tmp/prof_${tag}_c4 \
    python bench.py --config c4 --nwavel 1000 --steps 1 --warmup 1 > gpurun_out/ncu_c4_$tag.log 2>&1
python tools/ncu_summary.py /tmp/prof_${tag}_c4.ncu-rep > gpurun_out/ncu_${tag}_c4_summary.csv
ncu --set full --clock-control none -k regex:k_twostream -c 1 -f -o /tmp/prof_${tag}_c3 \
    python bench.py --config c3 --nwavel 200000 --steps 1 --warmup 1 > gpurun_out/ncu_c3_$tag.log 2>&1
python tools/ncu_summary.py /tmp/prof_${tag}_c3.ncu-rep > gpurun_out/ncu_${tag}_c3_summary.csv
tail -2 gpurun_out/gputests_$tag.log; tail -1 gpurun_out/smoke_$tag.txt
