// Segmented NMS building blocks shared by the RPN path (segments = image x
// pyramid level), the R-CNN path (segments = image x class) and the generic
// mmcv-style nms operator (one segment).
//
// Semantics follow mmcv 1.4.0 `nms_cpu`: candidates are
// visited in descending score order (ties: lower original index first, which
// is what the descending u64 `cand_key` encodes), a candidate is dropped when
// inter/(area_i+area_j-inter) > thr against an earlier kept candidate.
//
// Layout: S segments of uniform capacity `cap`; segment s owns
//   boxes[s*cap .. s*cap+count[s])      float4, already sorted
// Two implementations, same results:
//   nms_fused_kernel  (cap*20 B fits in shared memory): one CTA per segment,
//       boxes resident in smem, 64-candidate tiles resolved in order; IoUs are
//       evaluated lazily — only kept candidates are ever tested against later
//       ones — and the sweep stops at max_keep.  No global bitmask at all.
//   nms_mask_kernel + nms_sweep_kernel (large segments): the classic K x K/64
//       bitmask (upper-triangular tiles only) built by the whole GPU, then one
//       CTA per segment sweeps it on the device (no D2H copy, no host loop).
#pragma once
#include "common.cuh"

namespace brcnn {

typedef unsigned long long u64;

extern int64_t g_launch_count_add(int n);

// If img_maxc != nullptr the boxes are raw and the mmcv batched_nms offset
//   id * (max_coordinate + 1),  id = s % Sg,  max_coordinate = img_maxc[s / Sg]
// is added here in fp32 (boxes + offsets[:, None]), before any IoU arithmetic.
__device__ __forceinline__ float4 add_seg_offset(float4 b, float o) {
  b.x = b.x + o; b.y = b.y + o; b.z = b.z + o; b.w = b.w + o;
  return b;
}

// mmcv nms_cpu test.  When the boxes do not intersect, inter == 0 and
// 0/(a+b) > thr is false for every thr >= 0 (NaN for a+b == 0 is false too),
// so the division is skipped without changing any result.
__device__ __forceinline__ bool nms_suppresses(const float4 a, const float aa,
                                               const float4 b, const float ba,
                                               const float thr, const float off) {
  const float w = fminf(a.z, b.z) - fmaxf(a.x, b.x) + off;
  const float h = fminf(a.w, b.w) - fmaxf(a.y, b.y) + off;
  if (!(w > 0.f) || !(h > 0.f)) return false;
  const float inter = w * h;
  return inter / (aa + ba - inter) > thr;
}

// ---------------------------------------------------------------------------
// fused small-segment NMS.  dynamic smem: cap_pad*(16+4) + W*8 + 64*8 bytes
// ---------------------------------------------------------------------------
constexpr int NMS_FUSED_THREADS = 512;

__global__ void __launch_bounds__(NMS_FUSED_THREADS)
nms_fused_kernel(const float4* __restrict__ boxes, const uint8_t* __restrict__ valid,
                 const int32_t* __restrict__ count, int cap, float thr, float off,
                 const float* __restrict__ img_maxc, int Sg,
                 const u64* __restrict__ cand_key, int32_t* __restrict__ kept_pos,
                 u64* __restrict__ kept_key, int32_t* __restrict__ kept_count,
                 int keep_cap, int max_keep) {
  extern __shared__ __align__(16) unsigned char nms_smem[];
  const int W = (cap + 63) >> 6;
  const int cap_pad = W * 64;
  float4* sb = reinterpret_cast<float4*>(nms_smem);
  float* sa = reinterpret_cast<float*>(sb + cap_pad);
  u64* remv = reinterpret_cast<u64*>(sa + cap_pad);
  u64* diag = remv + W;
  __shared__ u64 s_keepbits;
  __shared__ int s_nkept;

  const int s = blockIdx.x;
  const int n = min(count[s], cap);
  const int Wn = (n + 63) >> 6;
  const int tid = threadIdx.x, lane = tid & 31;
  const float4* seg = boxes + (size_t)s * cap;
  float segoff = 0.f;
  const bool has_off = (img_maxc != nullptr);
  if (has_off) segoff = (float)(s % Sg) * (img_maxc[s / Sg] + 1.0f);
  for (int i = tid; i < n; i += NMS_FUSED_THREADS) {
    float4 b = seg[i];
    if (has_off) b = add_seg_offset(b, segoff);
    sb[i] = b;
    sa[i] = (b.z - b.x + off) * (b.w - b.y + off);
  }
  for (int w = tid; w < Wn; w += NMS_FUSED_THREADS) {
    u64 r = 0;
    const int base = w * 64;
    if (valid != nullptr) {
      const uint8_t* v = valid + (size_t)s * cap + base;
      const int m = min(64, n - base);
      for (int j = 0; j < m; ++j)
        if (!v[j]) r |= (1ull << j);
    }
    if (n - base < 64) r |= ~((1ull << (n - base)) - 1ull);
    remv[w] = r;
  }
  if (tid == 0) s_nkept = 0;
  __syncthreads();

  for (int t = 0; t < Wn; ++t) {
    const u64 alive = ~remv[t];
    if (alive == 0ull) continue;  // block-uniform
    // (b) 64x64 diagonal block: thread (r, g) tests row r against cols 8g..8g+7
    {
      const int r = tid >> 3, g = tid & 7;
      const int row = t * 64 + r;
      unsigned bits8 = 0;
      if ((alive >> r) & 1ull) {
        const float4 a = sb[row];
        const float aa = sa[row];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const int cc = g * 8 + c;
          if (cc > r && ((alive >> cc) & 1ull) &&
              nms_suppresses(a, aa, sb[t * 64 + cc], sa[t * 64 + cc], thr, off))
            bits8 |= (1u << c);
        }
      }
      u64 word = (u64)bits8 << (8 * g);
      word |= __shfl_xor_sync(0xffffffffu, word, 1);
      word |= __shfl_xor_sync(0xffffffffu, word, 2);
      word |= __shfl_xor_sync(0xffffffffu, word, 4);
      if (g == 0) diag[r] = word;
    }
    __syncthreads();
    // (c) warp 0 resolves the tile in order
    if (tid < 32) {
      const u64 d0 = diag[lane], d1 = diag[lane + 32];
      u64 al = alive, keep = 0;
      while (al) {
        const int r = __ffsll((long long)al) - 1;
        const u64 d = __shfl_sync(0xffffffffu, (r < 32) ? d0 : d1, r & 31);
        keep |= (1ull << r);
        al &= ~(d | (1ull << r));
      }
      const int base = s_nkept;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int r = lane + 32 * h;
        if ((keep >> r) & 1ull) {
          const int idx = base + __popcll(keep & ((1ull << r) - 1ull));
          if (idx < keep_cap) kept_pos[(size_t)s * keep_cap + idx] = t * 64 + r;
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
    // (d) kept candidates of this tile suppress later ones (lazy IoUs)
    const int jend = Wn * 64;
    for (int j = (t + 1) * 64 + tid; j < jend; j += NMS_FUSED_THREADS) {
      bool newly = false;
      if (!((remv[j >> 6] >> (j & 63)) & 1ull)) {
        const float4 b = sb[j];
        const float ba = sa[j];
        u64 bl = kb;
        while (bl) {
          const int r = __ffsll((long long)bl) - 1;
          bl &= bl - 1;
          if (nms_suppresses(sb[t * 64 + r], sa[t * 64 + r], b, ba, thr, off)) {
            newly = true;
            break;
          }
        }
      }
      const unsigned bal = __ballot_sync(0xffffffffu, newly);
      // each warp owns one 32-bit half of a remv word in this round
      if (lane == 0 && bal) reinterpret_cast<unsigned*>(remv)[j >> 5] |= bal;
    }
    __syncthreads();
  }
  __syncthreads();
  const int nk = min(min(s_nkept, keep_cap), max_keep);
  if (tid == 0) kept_count[s] = nk;
  if (cand_key != nullptr)
    for (int i = tid; i < nk; i += NMS_FUSED_THREADS)
      kept_key[(size_t)s * keep_cap + i] =
          cand_key[(size_t)s * cap + kept_pos[(size_t)s * keep_cap + i]];
}

// ---------------------------------------------------------------------------
// large segments: bitmask + device sweep
// ---------------------------------------------------------------------------
// grid (T, T, S) with T = ceil(cap/64); block 64 threads; mask [S][cap][W].
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
  for (int j = jstart; j < ncol; ++j)
    if (nms_suppresses(a, aa, cb[j], ca[j], thr, off)) bits |= (1ull << j);
  mask[((size_t)s * cap + row) * W + ct] = bits;
}

// One CTA per segment: greedy sweep over the bitmask.  Dynamic smem: W u64.
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
  const int lane = threadIdx.x & 31;
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
  // diagonal words of tile 0 (prefetched one tile ahead below)
  u64 d0 = 0, d1 = 0;
  if (threadIdx.x < 32 && Wn > 0) {
    if (lane < n) d0 = segmask[(size_t)lane * W];
    if (lane + 32 < n) d1 = segmask[(size_t)(lane + 32) * W];
  }
  for (int t = 0; t < Wn; ++t) {
    __syncthreads();
    if (threadIdx.x < 32) {
      u64 al = ~remv[t], keep = 0;
      while (al) {
        const int r = __ffsll((long long)al) - 1;
        const u64 d = __shfl_sync(0xffffffffu, (r < 32) ? d0 : d1, r & 31);
        keep |= (1ull << r);
        al &= ~(d | (1ull << r));
      }
      const int base = s_nkept;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int r = lane + 32 * h;
        if ((keep >> r) & 1ull) {
          const int idx = base + __popcll(keep & ((1ull << r) - 1ull));
          if (idx < keep_cap) kept_pos[(size_t)s * keep_cap + idx] = t * 64 + r;
        }
      }
      if (lane == 0) {
        s_keepbits = keep;
        s_nkept = base + __popcll(keep);
      }
      // prefetch the next diagonal block while the other warps push
      if (t + 1 < Wn) {
        const int r0 = (t + 1) * 64 + lane, r1 = r0 + 32;
        d0 = (r0 < n) ? segmask[(size_t)r0 * W + t + 1] : 0ull;
        d1 = (r1 < n) ? segmask[(size_t)r1 * W + t + 1] : 0ull;
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
  const int nk = min(min(s_nkept, keep_cap), max_keep);
  if (threadIdx.x == 0) kept_count[s] = nk;
  if (cand_key != nullptr)
    for (int i = threadIdx.x; i < nk; i += blockDim.x)
      kept_key[(size_t)s * keep_cap + i] =
          cand_key[(size_t)s * cap + kept_pos[(size_t)s * keep_cap + i]];
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
// Dynamic smem (optional, smem_lists != 0): Sg*lcap u64 — the first
// min(count, max_out) keys of every list, so the binary searches run on-chip.
template <class Epilogue>
__global__ void __launch_bounds__(256)
nms_merge_kernel(const int32_t* __restrict__ kept_pos,
                 const u64* __restrict__ kept_key,
                 const int32_t* __restrict__ kept_count, int Sg, int keep_cap,
                 int max_out, int lcap, int smem_lists,
                 int32_t* __restrict__ num_out, Epilogue ep) {
  extern __shared__ __align__(16) u64 s_lists[];
  const int b = blockIdx.x;
  __shared__ int s_total;
  if (threadIdx.x == 0) {
    int tot = 0;
    for (int g = 0; g < Sg; ++g) tot += min(kept_count[b * Sg + g], max_out);
    s_total = tot;
    num_out[b] = min(tot, max_out);
  }
  if (smem_lists) {
    for (int g = 0; g < Sg; ++g) {
      const int seg = b * Sg + g;
      const int ng = min(min(kept_count[seg], max_out), lcap);
      for (int j = threadIdx.x; j < ng; j += blockDim.x)
        s_lists[(size_t)g * lcap + j] = kept_key[(size_t)seg * keep_cap + j];
    }
  }
  __syncthreads();
  const int total = s_total;
  for (int g = 0; g < Sg; ++g) {
    const int seg = b * Sg + g;
    const int ng = min(kept_count[seg], max_out);
    for (int j = threadIdx.x; j < ng; j += blockDim.x) {
      const u64 key = smem_lists ? s_lists[(size_t)g * lcap + j]
                                 : kept_key[(size_t)seg * keep_cap + j];
      int rank = j;
      for (int g2 = 0; g2 < Sg && rank < max_out; ++g2) {
        if (g2 == g) continue;
        const int seg2 = b * Sg + g2;
        const int n2 = min(kept_count[seg2], max_out);
        if (n2 == 0) continue;
        const u64* list = smem_lists ? s_lists + (size_t)g2 * lcap
                                     : kept_key + (size_t)seg2 * keep_cap;
        rank += count_greater_desc(list, n2, key);
      }
      if (rank < max_out)
        ep(b, rank, seg, kept_pos[(size_t)seg * keep_cap + j], key);
    }
  }
  for (int r = total + threadIdx.x; r < max_out; r += blockDim.x) ep.pad(b, r);
}

template <class Epilogue>
inline int launch_nms_merge(const int32_t* kept_pos, const u64* kept_key,
                            const int32_t* kept_count, int B, int Sg, int keep_cap,
                            int max_out, int32_t* num_out, Epilogue ep,
                            cudaStream_t stream) {
  const int lcap = keep_cap < max_out ? keep_cap : max_out;
  size_t smem = (size_t)Sg * lcap * sizeof(u64);
  int use_smem = 1;
  if (smem > 200 * 1024) { use_smem = 0; smem = 0; }
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(nms_merge_kernel<Epilogue>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  nms_merge_kernel<Epilogue><<<B, 256, smem, stream>>>(
      kept_pos, kept_key, kept_count, Sg, keep_cap, max_out, lcap, use_smem, num_out, ep);
  g_launch_count_add(1);
  BRCNN_CUDA_CHECK_LAST();
  return BRCNN_OK;
}

// bytes of bitmask workspace needed for S segments of capacity cap (0 when the
// fused kernel handles them)
inline bool nms_use_fused(int cap) {
  const int W = (cap + 63) / 64;
  const size_t smem = (size_t)W * 64 * 20 + (size_t)W * 8 + 64 * 8;
  return smem <= 160 * 1024 && cap <= 2048;
}

// host-side launcher for S uniform segments
inline int launch_nms_segments(const float4* boxes, const uint8_t* valid,
                               const int32_t* count, int S, int cap, float thr,
                               float off, const float* img_maxc, int Sg,
                               u64* mask, const u64* cand_key,
                               int32_t* kept_pos, u64* kept_key,
                               int32_t* kept_count, int keep_cap, int max_keep,
                               cudaStream_t stream) {
  if (S <= 0 || cap <= 0) return BRCNN_OK;
  const int W = (cap + 63) / 64;
  if (nms_use_fused(cap)) {
    const size_t smem = (size_t)W * 64 * 20 + (size_t)W * 8 + 64 * 8;
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(nms_fused_kernel,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)smem);
      if (e != cudaSuccess) return (int)e;
    }
    nms_fused_kernel<<<S, NMS_FUSED_THREADS, smem, stream>>>(
        boxes, valid, count, cap, thr, off, img_maxc, Sg, cand_key, kept_pos, kept_key,
        kept_count, keep_cap, max_keep);
    g_launch_count_add(1);
    BRCNN_CUDA_CHECK_LAST();
    return BRCNN_OK;
  }
  if (S > 65535) return BRCNN_ERR_UNSUPPORTED;
  if (mask == nullptr) return BRCNN_ERR_WORKSPACE;
  {
    dim3 grid(W, W, S);
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
