#!/bin/bash
# Run on the GPU box (gpurun -- 'bash tools/gpu_profile.sh r01'): launch lists and
# ncu --set full captures of every libbrcnn kernel for the three bench modes.
# Outputs go to gpurun_out/; tools/summarize_profiles.py turns them into profiles/.
tag=${1:-r01}
K='regex:rpn_|nms_|roi_|transpose_|rcnn_|bbox2roi|boost_'
out=gpurun_out
# launch lists (device time per launch; cold-cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 80 --csv \
    --log-file $out/${tag}_launches_infer.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline > $out/${tag}_infer_under_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $out/${tag}_launches_train.csv python bench.py --mode train --cfg coco --batch 2 --steps 1 --warmup 3 > $out/${tag}_train_under_ncu.log 2>&1
# full captures of our kernels, one step after 3 warm-up steps
ncu --set full --clock-control none --import-source on -k "$K" -s 30 -c 10 -o $out/${tag}_full_infer \
    python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k "$K" -s 36 -c 12 -o $out/${tag}_full_train \
    python bench.py --mode train --cfg coco --batch 2 --steps 1 --warmup 3 > /dev/null 2>&1
# clean bench lines (not under a profiler)
python bench.py --steps 100 --warmup 5 > $out/${tag}_bench_infer.json 2> $out/${tag}_bench_infer.err
python bench.py --mode train --cfg coco --batch 2 --steps 20 --warmup 3 > $out/${tag}_bench_train.json 2>/dev/null
python bench.py --mode stress --batch 4 --steps 5 --warmup 3 > $out/${tag}_bench_stress.json 2>/dev/null
python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_reference.json 2>/dev/null
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $out/${tag}_nvidia_smi.csv
