#!/bin/bash
# Run on the GPU box (gpurun -- 'bash tools/gpu_profile.sh r02'): launch lists and
# ncu --set full captures of every libbrcnn kernel for the bench modes, plus clean bench lines.
# Outputs go to gpurun_out/; tools/summarize_profiles.py turns them into profiles/.
tag=${1:-r02}
K='regex:rpn_|nms_|roi_|transpose_|rcnn_|bbox2roi|boost_'
out=gpurun_out
mkdir -p $out
FAST="--no-train-record --no-cpu-baseline"
# launch lists (device time per launch; cold-cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 80 --csv \
    --log-file $out/${tag}_launches_infer.csv python bench.py --steps 2 --warmup 3 --no-graph $FAST > $out/${tag}_infer_under_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv \
    --log-file $out/${tag}_launches_train.csv python bench.py --mode train --train-eager --cfg coco --batch 2 --steps 1 --warmup 3 > $out/${tag}_train_under_ncu.log 2>&1
# full captures of our kernels, one step after 3 warm-up steps
ncu --set full --clock-control none --import-source on -k "$K" -s 30 -c 10 -o $out/${tag}_full_infer \
    python bench.py --steps 1 --warmup 3 --no-graph $FAST > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k "$K" -s 54 -c 18 -o $out/${tag}_full_train \
    python bench.py --mode train --train-eager --cfg coco --batch 2 --steps 1 --warmup 3 > /dev/null 2>&1
# clean bench lines (not under a profiler)
python bench.py --steps 100 --warmup 5 > $out/${tag}_bench_infer.json 2> $out/${tag}_bench_infer.err
python bench.py --mode train --cfg coco --batch 2 --steps 20 --warmup 3 > $out/${tag}_bench_train.json 2>/dev/null
python bench.py --mode stress --batch 4 --steps 10 --warmup 3 > $out/${tag}_bench_stress.json 2>/dev/null
python bench.py --mode stress --stress-input clustered --batch 4 --steps 10 --warmup 3 > $out/${tag}_bench_stress_clustered.json 2>/dev/null
python bench.py --mode stress --stress-input anchors --batch 4 --steps 10 --warmup 3 > $out/${tag}_bench_stress_anchors.json 2>/dev/null
python bench.py --cfg voc --rpn-max-per-img 1000 --batch 4 --lean --steps 20 --warmup 5 $FAST > $out/${tag}_bench_voc1000.json 2>/dev/null
python bench.py --cfg coco --steps 20 --warmup 5 $FAST > $out/${tag}_bench_coco_infer.json 2>/dev/null
python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_reference.json 2>/dev/null
python tools/nms_microbench.py > $out/${tag}_nms_microbench.txt 2>&1
python tools/roi_microbench.py --cfg utdac --batch 16 > $out/${tag}_roi_microbench_infer.json 2>/dev/null
python tools/roi_microbench.py --cfg coco --batch 2 --train > $out/${tag}_roi_microbench_train.json 2>/dev/null
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $out/${tag}_nvidia_smi.csv
