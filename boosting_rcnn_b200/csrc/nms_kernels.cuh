// Segmented bitmask NMS building blocks shared by the RPN path (segments =
// image x pyramid level), the R-CNN path (segments = image x class) and the
// generic mmcv-style nms operator (one segment).
//
// Semantics follow mmcv 1.4.0 `nms_cpu`: candidates are
// visited in descending score order (ties: lower original index first, which
// is what the descending u64 `cand_key` encodes), a candidate is dropped when
// inter/(area_i+area_j-inter) > thr against an earlier kept candidate.
//
// Layout: S segments of uniform capacity `cap`; segment s owns
//   boxes[s*cap .. s*cap+count[s])      float4, already sorted, already offset
//   mask [s][row][W]                    u64, W = ceil(cap/64); only words
//                                       col_tile >= row_tile are produced
#pragma once
#include "common.cuh"

namespace brcnn {

typedef unsigned long long u64;

extern int64_t g_launch_count_add(int n);

// grid (T, T, S) with T = ceil(cap/64); block 64 threads.
// If img_maxc != nullptr the boxes are raw and the mmcv batched_nms offset
//   id * (max_coordinate + 1),  id = s % Sg,  max_coordinate = img_maxc[s / Sg]
// is added here in fp32 (boxes + offsets[:, None]), before any IoU arithmetic.
__device__ __forceinline__ float4 add_seg_offset(float4 b, float o) {
  b.x = b.x + o; b.y = b.y + o; b.z = b.z + o; b.w = b.w + o;
  return b;
}

__global__ void __launch_bounds__(64)
nms_mask_kernel(const float4* __restrict__ boxes,
                const int32_t* __restrict__ count, int cap, int W, float thr,
                float off, const float* __restrict__ img_maxc, int Sg,
                u64* __restrict__ mask) {
  const int ct = blockIdx.x, rt = blockIdx.y, s = blockIdx.z;
  if (ct < rt) return;
  const int n = min(count[s], cap);
  if (rt * 64 >= n || ct * 64 >= n) return;
  __shared__ float4 cb[64];
  __shared__ float ca[64];
  const float4* seg = boxes + (size_t)s * cap;
  const bool has_off = (img_maxc != nullptr);
  float segoff = 0.f;
  if (has_off) segoff = (float)(s % Sg) * (img_maxc[s / Sg] + 1.0f);
  const int i = threadIdx.x;
  const int col = ct * 64 + i;
  if (col < n) {
    float4 b = seg[col];
    if (has_off) b = add_seg_offset(b, segoff);
    cb[i] = b;
    ca[i] = (b.z - b.x + off) * (b.w - b.y + off);
  }
  __syncthreads();
  const int row = rt * 64 + i;
  if (row >= n) return;
  float4 a = seg[row];
  if (has_off) a = add_seg_offset(a, segoff);
  const float aa = (a.z - a.x + off) * (a.w - a.y + off);
  const int ncol = min(64, n - ct * 64);
  const int jstart = (rt == ct) ? i + 1 : 0;
  u64 bits = 0;
  for (int j = jstart; j < ncol; ++j) {
    const float4 b = cb[j];
    float xx1 = fmaxf(a.x, b.x);
    float yy1 = fmaxf(a.y, b.y);
    float xx2 = fminf(a.z, b.z);
    float yy2 = fminf(a.w, b.w);
    float w = fmaxf(0.f, xx2 - xx1 + off);
    float h = fmaxf(0.f, yy2 - yy1 + off);
    float inter = w * h;
    float ovr = inter / (aa + ca[j] - inter);
    if (ovr > thr) bits |= (1ull << j);
  }
  mask[((size_t)s * cap + row) * W + ct] = bits;
}

// One CTA per segment: greedy sweep over the bitmask.  Dynamic smem: W u64.
// valid (optional): uint8 [S][cap]; invalid candidates are never kept and
// never suppress.  Writes kept_pos[s][0..kept_count[s]) (candidate ranks in
// visiting order) and, if cand_key != nullptr, kept_key likewise.
__global__ void __launch_bounds__(128)
nms_sweep_kernel(const u64* __restrict__ mask,
                 const uint8_t* __restrict__ valid,
                 const int32_t* __restrict__ count, int cap, int W,
                 const u64* __restrict__ cand_key, int32_t* __restrict__ kept_pos,
                 u64* __restrict__ kept_key, int32_t* __restrict__ kept_count,
                 int keep_cap, int max_keep) {
  extern __shared__ u64 remv[];
  __shared__ u64 s_keepbits;
  __shared__ int s_nkept;
  const int s = blockIdx.x;
  const int n = min(count[s], cap);
  const int Wn = (n + 63) >> 6;
  const u64* segmask = mask + (size_t)s * cap * W;
  for (int w = threadIdx.x; w < Wn; w += blockDim.x) {
    u64 r = 0;
    const int base = w * 64;
    if (valid != nullptr) {
      const uint8_t* v = valid + (size_t)s * cap + base;
      const int m = min(64, n - base);
      for (int j = 0; j < m; ++j)
        if (!v[j]) r |= (1ull << j);
    }
    if (n - base < 64) r |= ~((1ull << (n - base)) - 1ull);  // beyond count
    remv[w] = r;
  }
  if (threadIdx.x == 0) s_nkept = 0;
  for (int t = 0; t < Wn; ++t) {
    __syncthreads();
    if (threadIdx.x < 32) {
      const int lane = threadIdx.x;
      const int r0 = t * 64 + lane, r1 = r0 + 32;
      u64 d0 = (r0 < n) ? segmask[(size_t)r0 * W + t] : 0ull;
      u64 d1 = (r1 < n) ? segmask[(size_t)r1 * W + t] : 0ull;
      u64 alive = ~remv[t];
      u64 keep = 0;
#pragma unroll 8
      for (int r = 0; r < 64; ++r) {
        u64 d = __shfl_sync(0xffffffffu, (r < 32) ? d0 : d1, r & 31);
        if ((alive >> r) & 1ull) {
          keep |= (1ull << r);
          alive &= ~d;
        }
      }
      const int base = s_nkept;
      // lane writes bits `lane` and `lane+32`
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int r = lane + 32 * h;
        if ((keep >> r) & 1ull) {
          const int idx = base + __popcll(keep & ((1ull << r) - 1ull));
          if (idx < keep_cap) {
            kept_pos[(size_t)s * keep_cap + idx] = t * 64 + r;
            if (cand_key != nullptr)
              kept_key[(size_t)s * keep_cap + idx] =
                  cand_key[(size_t)s * cap + t * 64 + r];
          }
        }
      }
      __syncwarp();
      if (lane == 0) {
        s_keepbits = keep;
        s_nkept = base + __popcll(keep);
      }
    }
    __syncthreads();
    const u64 kb = s_keepbits;
    if (s_nkept >= max_keep) break;
    for (int w = t + 1 + threadIdx.x; w < Wn; w += blockDim.x) {
      u64 acc = 0;
      u64 bitsleft = kb;
      while (bitsleft) {
        const int r = __ffsll((long long)bitsleft) - 1;
        bitsleft &= bitsleft - 1;
        acc |= segmask[(size_t)(t * 64 + r) * W + w];
      }
      remv[w] |= acc;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) kept_count[s] = min(min(s_nkept, keep_cap), max_keep);
}

// Count of keys strictly greater than `key` in a descending-sorted list.
__device__ __forceinline__ int count_greater_desc(const u64* __restrict__ list,
                                                  int n, u64 key) {
  int lo = 0, hi = n;  // first position with list[pos] <= key
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (list[mid] > key) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// One CTA per image: merge the Sg kept lists of image b (each descending in
// kept_key) and hand the global rank of every kept candidate to the
// epilogue.  Epilogue::operator()(b, rank, seg, pos_in_seg, key) is called
// for rank < max_out; Epilogue::pad(b, rank) for total <= rank < max_out.
template <class Epilogue>
__global__ void __launch_bounds__(256)
nms_merge_kernel(const int32_t* __restrict__ kept_pos,
                 const u64* __restrict__ kept_key,
                 const int32_t* __restrict__ kept_count, int Sg, int keep_cap,
                 int max_out, int32_t* __restrict__ num_out, Epilogue ep) {
  const int b = blockIdx.x;
  __shared__ int s_total;
  if (threadIdx.x == 0) {
    int tot = 0;
    for (int g = 0; g < Sg; ++g) tot += min(kept_count[b * Sg + g], max_out);
    s_total = tot;
    num_out[b] = min(tot, max_out);
  }
  __syncthreads();
  const int total = s_total;
  for (int g = 0; g < Sg; ++g) {
    const int seg = b * Sg + g;
    const int ng = min(kept_count[seg], max_out);
    for (int j = threadIdx.x; j < ng; j += blockDim.x) {
      const u64 key = kept_key[(size_t)seg * keep_cap + j];
      int rank = j;
      for (int g2 = 0; g2 < Sg && rank < max_out; ++g2) {
        if (g2 == g) continue;
        const int seg2 = b * Sg + g2;
        rank += count_greater_desc(kept_key + (size_t)seg2 * keep_cap,
                                   min(kept_count[seg2], max_out), key);
      }
      if (rank < max_out)
        ep(b, rank, seg, kept_pos[(size_t)seg * keep_cap + j], key);
    }
  }
  for (int r = total + threadIdx.x; r < max_out; r += blockDim.x) ep.pad(b, r);
}

// host-side launcher for mask + sweep over S uniform segments
inline int launch_nms_segments(const float4* boxes, const uint8_t* valid,
                               const int32_t* count, int S, int cap, float thr,
                               float off, const float* img_maxc, int Sg,
                               u64* mask, const u64* cand_key,
                               int32_t* kept_pos, u64* kept_key,
                               int32_t* kept_count, int keep_cap, int max_keep,
                               cudaStream_t stream) {
  if (S <= 0 || cap <= 0) return BRCNN_OK;
  if (S > 65535) return BRCNN_ERR_UNSUPPORTED;
  const int W = (cap + 63) / 64;
  const int T = W;
  {
    dim3 grid(T, T, S);
    nms_mask_kernel<<<grid, 64, 0, stream>>>(boxes, count, cap, W, thr, off,
                                             img_maxc, Sg, mask);
    g_launch_count_add(1);
    BRCNN_CUDA_CHECK_LAST();
  }
  const size_t smem = (size_t)W * sizeof(u64);
  if (smem > 48 * 1024) return BRCNN_ERR_UNSUPPORTED;
  nms_sweep_kernel<<<S, 128, smem, stream>>>(mask, valid, count, cap, W,
                                             cand_key, kept_pos, kept_key,
                                             kept_count, keep_cap, max_keep);
  g_launch_count_add(1);
  BRCNN_CUDA_CHECK_LAST();
  return BRCNN_OK;
}

}  // namespace brcnn
