python -m pytest tests/test_gpu_roi.py tests/test_gpu_fuzz.py tests/test_gpu_rpn.py -x -q 2>&1 | tail -2
python tools/roi_microbench.py --cfg utdac --batch 16 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['fwd_hwc_ms'], d['fwd_hwc_frac'])"
python tools/roi_microbench.py --cfg coco --batch 2 --train | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['fwd_hwc_ms'], d['fwd_hwc_frac'], d['bwd_hwc_ms'])"
