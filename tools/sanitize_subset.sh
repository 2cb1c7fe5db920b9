#!/bin/bash
# compute-sanitizer over a small-but-covering subset of the GPU tests (run under gpurun).
T="tests/test_gpu_roi.py::test_roi_forward_small tests/test_gpu_roi.py::test_roi_forward_edge_rois tests/test_gpu_roi.py::test_roi_backward tests/test_gpu_roi.py::test_roi_forward_channel_slabs tests/test_gpu_roi.py::test_pyramid_transposes_one_launch tests/test_gpu_roi.py::test_roi_forward_hwc_other_pooled_sizes tests/test_gpu_roi.py::test_roi_backward_hwc_deterministic tests/test_gpu_roi.py::test_roi_forward_schedule_is_invisible tests/test_gpu_roi.py::test_roi_forward_wide_footprints_multi_pass tests/test_gpu_rpn.py::test_rpn_small_unique tests/test_gpu_rpn.py::test_rpn_small_duplicates tests/test_gpu_rpn.py::test_rpn_degenerate_boxes_filtered tests/test_gpu_rpn.py::test_rpn_scores_clustered_on_histogram_bin_edges tests/test_gpu_rpn.py::test_rpn_concentrated_scores_slow_path tests/test_gpu_rcnn.py::test_rcnn_utdac tests/test_gpu_rcnn.py::test_rcnn_ragged_and_empty_images tests/test_gpu_loss.py::test_boost_loss_no_positives_and_agnostic tests/test_gpu_nms.py::test_batched_nms_class_agnostic_and_empty tests/test_gpu_rpn_loss.py::test_forward_train_returns_reference_keys"
TOOLS=${1:-"memcheck synccheck racecheck initcheck"}
for tool in $TOOLS; do
  echo "=== $tool"
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 99 --print-limit 5 python -m pytest $T -x -q -m gpu 2>&1 | grep -v "^$" | tail -14
  echo "rc=$?"
done
