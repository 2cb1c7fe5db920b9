/* Plain-C consumer of include/brcnn.h: no torch, no Python, only the CUDA runtime for
 * device memory.  Shows that the drop-in boundary is a C ABI with raw pointers:
 *   gcc abi_smoke.c -I../../include -L../../boosting_rcnn_b200 -lbrcnn -lcudart
 * Runs mmcv-style nms on 6 boxes and delta2bbox on 2, checks the known answers
 * (delta_xywh_bbox_coder.py:191-204 doctest), prints "abi_smoke ok".            */
#include <cuda_runtime_api.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "brcnn.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "cuda error %d at %s:%d\n", (int)e_, __FILE__, __LINE__); return 2; } } while (0)

int main(void) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    printf("abi_smoke: no CUDA device (library loaded, version %s)\n", brcnn_version());
    return 77;
  }
  /* ---- nms: boxes 0/1 overlap (IoU 0.90), 2/3 overlap (0.65 < 0.7), 4, 5 apart ---- */
  const float boxes[6][4] = {{0, 0, 10, 10}, {1, 0, 10, 10}, {20, 20, 30, 30},
                             {23, 20, 30, 31}, {50, 50, 60, 60}, {100, 0, 110, 5}};
  const float scores[6] = {0.9f, 0.8f, 0.7f, 0.95f, 0.1f, 0.5f};
  float *d_boxes, *d_scores, *d_dets;
  int64_t* d_keep;
  int32_t* d_num;
  void* d_ws;
  const size_t ws_bytes = brcnn_nms_workspace_bytes(6);
  CK(cudaMalloc((void**)&d_boxes, sizeof(boxes)));
  CK(cudaMalloc((void**)&d_scores, sizeof(scores)));
  CK(cudaMalloc((void**)&d_dets, 6 * 5 * sizeof(float)));
  CK(cudaMalloc((void**)&d_keep, 6 * sizeof(int64_t)));
  CK(cudaMalloc((void**)&d_num, sizeof(int32_t)));
  CK(cudaMalloc(&d_ws, ws_bytes));
  CK(cudaMemcpy(d_boxes, boxes, sizeof(boxes), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_scores, scores, sizeof(scores), cudaMemcpyHostToDevice));
  int rc = brcnn_batched_nms(d_boxes, d_scores, NULL, 6, 1, 0.7f, 0, -1, d_keep, d_dets, d_num, d_ws,
                             ws_bytes, NULL /* default stream */);
  if (rc != BRCNN_OK) { fprintf(stderr, "brcnn_batched_nms rc=%d\n", rc); return 1; }
  CK(cudaDeviceSynchronize());
  int64_t keep[6];
  int32_t num = 0;
  CK(cudaMemcpy(&num, d_num, sizeof(num), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(keep, d_keep, sizeof(keep), cudaMemcpyDeviceToHost));
  const int64_t expect[5] = {3, 0, 2, 5, 4};   /* score order, box 1 suppressed by box 0 */
  if (num != 5 || memcmp(keep, expect, sizeof(expect)) != 0) {
    fprintf(stderr, "nms keep list wrong: n=%d [%lld %lld %lld %lld %lld]\n", num,
            (long long)keep[0], (long long)keep[1], (long long)keep[2], (long long)keep[3],
            (long long)keep[4]);
    return 1;
  }
  /* argument errors are return codes, not crashes */
  if (brcnn_batched_nms(NULL, d_scores, NULL, 6, 1, 0.7f, 0, -1, d_keep, d_dets, d_num, d_ws,
                        ws_bytes, NULL) != BRCNN_ERR_ARG) return 1;
  if (brcnn_batched_nms(d_boxes, d_scores, NULL, 6, 1, 0.7f, 0, -1, d_keep, d_dets, d_num, d_ws, 16,
                        NULL) != BRCNN_ERR_WORKSPACE) return 1;
  /* ---- delta2bbox doctest: rois [0,0,1,1] / [5,5,5,5], zero deltas, max_shape (32,32) ---- */
  const float rois[2][4] = {{0, 0, 1, 1}, {5, 5, 5, 5}};
  const float deltas[2][4] = {{0, 0, 0, 0}, {0.f, 0.f, 1.f, 1.f}};
  const float means[4] = {0, 0, 0, 0}, stds[4] = {1, 1, 1, 1};
  float *d_rois, *d_deltas, *d_out, out[2][4];
  CK(cudaMalloc((void**)&d_rois, sizeof(rois)));
  CK(cudaMalloc((void**)&d_deltas, sizeof(deltas)));
  CK(cudaMalloc((void**)&d_out, sizeof(out)));
  CK(cudaMemcpy(d_rois, rois, sizeof(rois), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_deltas, deltas, sizeof(deltas), cudaMemcpyHostToDevice));
  rc = brcnn_delta2bbox(d_rois, d_deltas, 2, 1, means, stds, (float)fabs(log(16.0 / 1000.0)), 32.f,
                        32.f, d_out, NULL);
  if (rc != BRCNN_OK) { fprintf(stderr, "brcnn_delta2bbox rc=%d\n", rc); return 1; }
  CK(cudaMemcpy(out, d_out, sizeof(out), cudaMemcpyDeviceToHost));
  const float want[2][4] = {{0, 0, 1, 1}, {5, 5, 5, 5}};   /* zero-size roi stays a point */
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 4; ++j)
      if (fabsf(out[i][j] - want[i][j]) > 1e-6f) {
        fprintf(stderr, "delta2bbox[%d][%d] = %f\n", i, j, out[i][j]);
        return 1;
      }
  printf("abi_smoke ok (%s, %lld kernel launches)\n", brcnn_version(),
         (long long)brcnn_launch_count());
  return 0;
}
