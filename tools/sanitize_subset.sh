#!/bin/bash
# compute-sanitizer over a small-but-covering subset of the GPU tests (run under gpurun).
T="tests/test_gpu_roi.py::test_roi_forward_small tests/test_gpu_roi.py::test_roi_forward_edge_rois tests/test_gpu_roi.py::test_roi_backward tests/test_gpu_roi.py::test_roi_forward_channel_slabs tests/test_gpu_roi.py::test_pyramid_transposes_one_launch tests/test_gpu_rpn.py::test_rpn_small_unique tests/test_gpu_rpn.py::test_rpn_small_duplicates tests/test_gpu_rpn.py::test_rpn_degenerate_boxes_filtered tests/test_gpu_rcnn.py::test_rcnn_utdac tests/test_gpu_rcnn.py::test_rcnn_ragged_and_empty_images tests/test_gpu_loss.py::test_boost_loss_no_positives_and_agnostic tests/test_gpu_nms.py::test_batched_nms_class_agnostic_and_empty"
for tool in memcheck racecheck synccheck initcheck; do
  echo "=== $tool"
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 99 --print-limit 5 python -m pytest $T -x -q -m gpu 2>&1 | grep -v "^$" | tail -12
  echo "rc=$?"
done
