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
  const float u = aa + ba - inter;
  // fl(inter/u) > thr decided without the division when inter is more than
  // ~16 ulp away from thr*u (the IEEE quotient is within 1 ulp of the real
  // ratio); the exact division only runs inside that band.
  const float p = thr * u;
  if (thr > 1e-30f && p > 1e-30f && p < 1e30f) {  // normal range, u > 0
    if (inter > p * 1.000001f) return true;
    if (inter < p * 0.999999f) return false;
  }
  return inter / u > thr;
}

// ---------------------------------------------------------------------------
// fused NMS, "pull" form.  One CTA per segment walks the sorted candidates in
// tiles of 64:
//   1. pull   : the tile's candidates are tested against every box kept so
//               far (kept list lives in shared memory) -> dead bits;
//   2. diag   : 64x64 intra-tile suppression bits for the survivors;
//   3. resolve: warp 0 walks the tile in order (ffs over the alive bits);
//   4. append : kept boxes join the shared-memory list.
// IoUs are only ever evaluated between a candidate and an earlier KEPT box,
// tiles after the max_keep-th keep are never touched, and there is no global
// bitmask.  dynamic smem: keep_pad*(16+4) bytes.
// ---------------------------------------------------------------------------
constexpr int NMS_FUSED_THREADS = 512;

__global__ void __launch_bounds__(NMS_FUSED_THREADS)
nms_fused_kernel(const float4* __restrict__ boxes, const uint8_t* __restrict__ valid,
                 const int32_t* __restrict__ count, int cap, float thr, float off,
                 const float* __restrict__ img_maxc, int Sg,
                 const u64* __restrict__ cand_key, int32_t* __restrict__ kept_pos,
                 u64* __restrict__ kept_key, int32_t* __restrict__ kept_count,
                 int keep_cap, int max_keep, int keep_pad,
                 const int32_t* __restrict__ seg_start) {
  extern __shared__ __align__(16) unsigned char nms_smem[];
  float4* kbox = reinterpret_cast<float4*>(nms_smem);        // [keep_pad]
  float* karea = reinterpret_cast<float*>(kbox + keep_pad);  // [keep_pad]
  __shared__ float4 tb[64];
  __shared__ float ta[64];
  __shared__ u64 diag[64];
  __shared__ unsigned s_dead[2];
  __shared__ int s_nkept;

  const int s = blockIdx.x;
  const int n = min(count[s], cap);
  const int Wn = (n + 63) >> 6;
  const int tid = threadIdx.x, lane = tid & 31;
  // uniform capacity `cap` per segment, or (seg_start != nullptr) variable-size segments
  // packed back to back; kept lists then share the segment's offset (kept <= count)
  const size_t sbase = seg_start != nullptr ? (size_t)seg_start[s] : (size_t)s * cap;
  const size_t kbase = seg_start != nullptr ? (size_t)seg_start[s] : (size_t)s * keep_cap;
  const float4* seg = boxes + sbase;
  float segoff = 0.f;
  const bool has_off = (img_maxc != nullptr);
  if (has_off) segoff = (float)(s % Sg) * (img_maxc[s / Sg] + 1.0f);
  if (tid == 0) s_nkept = 0;
  // first tile prefetched into registers by threads 0..63
  float4 nb = make_float4(0.f, 0.f, 0.f, 0.f);
  bool ndead = true;
  if (tid < 64 && tid < n) {
    nb = seg[tid];
    ndead = (valid != nullptr && valid[sbase + tid] == 0);
  }
  __syncthreads();

  for (int t = 0; t < Wn; ++t) {
    // ---- stage tile t, prefetch tile t+1 ----
    if (tid < 64) {
      float4 b = nb;
      if (has_off) b = add_seg_offset(b, segoff);
      tb[tid] = b;
      ta[tid] = (b.z - b.x + off) * (b.w - b.y + off);
      const unsigned bal = __ballot_sync(0xffffffffu, ndead);
      if (lane == 0) s_dead[tid >> 5] = bal;
      const int i = (t + 1) * 64 + tid;
      ndead = true;
      if (i < n) {
        nb = seg[i];
        ndead = (valid != nullptr && valid[sbase + i] == 0);
      }
    }
    __syncthreads();
    const int nkept = s_nkept;
    // ---- 1. pull: candidate c vs kept boxes q = g, g+8, ... ----
    {
      const int c = tid & 63, g = tid >> 6;
      bool hit = false;
      if (!((s_dead[c >> 5] >> (c & 31)) & 1u)) {
        const float4 b = tb[c];
        const float ba = ta[c];
        int q = g;
        const int ns = min(nkept, keep_pad);          // kept boxes resident in shared memory
        for (; q + 24 < ns && !hit; q += 32) {
          const bool h0 = nms_suppresses(kbox[q], karea[q], b, ba, thr, off);
          const bool h1 = nms_suppresses(kbox[q + 8], karea[q + 8], b, ba, thr, off);
          const bool h2 = nms_suppresses(kbox[q + 16], karea[q + 16], b, ba, thr, off);
          const bool h3 = nms_suppresses(kbox[q + 24], karea[q + 24], b, ba, thr, off);
          hit = h0 | h1 | h2 | h3;
        }
        for (; q < ns && !hit; q += 8)
          hit = nms_suppresses(kbox[q], karea[q], b, ba, thr, off);
        // kept list longer than the shared-memory budget (operator path, one huge id):
        // the tail is re-read from the sorted boxes through kept_pos (rare, slow, correct)
        for (; q < nkept && !hit; q += 8) {
          float4 kb = seg[kept_pos[kbase + q]];
          if (has_off) kb = add_seg_offset(kb, segoff);
          hit = nms_suppresses(kb, (kb.z - kb.x + off) * (kb.w - kb.y + off), b, ba, thr, off);
        }
      }
      const unsigned bal = __ballot_sync(0xffffffffu, hit);
      if (lane == 0 && bal) atomicOr(&s_dead[(tid >> 5) & 1], bal);
    }
    __syncthreads();
    const u64 alive = ~(((u64)s_dead[1] << 32) | (u64)s_dead[0]);
    if (alive == 0ull) {  // block-uniform
      __syncthreads();  // s_dead is rewritten by the next tile's staging
      continue;
    }
    // ---- 2. diag: thread (r, g) tests row r against cols 8g..8g+7 ----
    {
      const int r = tid >> 3, g = tid & 7;
      unsigned bits8 = 0;
      if ((alive >> r) & 1ull) {
        const float4 a = tb[r];
        const float aa = ta[r];
        // column form: which EARLIER alive candidates of the tile suppress r
        // (the IoU test is symmetric)
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const int cc = g * 8 + c;
          if (cc < r && ((alive >> cc) & 1ull) &&
              nms_suppresses(tb[cc], ta[cc], a, aa, thr, off))
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
    // ---- 3./4. warp 0 resolves the tile and appends the keeps ----
    // Greedy NMS inside the tile is the unique fixpoint of
    //   K = { r alive : no earlier r' in K suppresses r };
    // iterating from K = alive fixes the first i candidates after i rounds, so
    // it converges in (suppression-chain depth + 1) ballot rounds instead of
    // one serial step per kept box.
    if (tid < 32) {
      const u64 c0 = diag[lane], c1 = diag[lane + 32];
      const bool a0 = (alive >> lane) & 1ull, a1 = (alive >> (lane + 32)) & 1ull;
      u64 keep = alive;
      for (int it = 0; it < 64; ++it) {
        const unsigned k0 = __ballot_sync(0xffffffffu, a0 && !(c0 & keep));
        const unsigned k1 = __ballot_sync(0xffffffffu, a1 && !(c1 & keep));
        const u64 kn = ((u64)k1 << 32) | (u64)k0;
        if (kn == keep) break;
        keep = kn;
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int r = lane + 32 * h;
        if ((keep >> r) & 1ull) {
          const int q = nkept + __popcll(keep & ((1ull << r) - 1ull));
          if (q < keep_cap && q < max_keep) {
            kept_pos[kbase + q] = t * 64 + r;
            if (q < keep_pad) {
              kbox[q] = tb[r];
              karea[q] = ta[r];
            }
          }
        }
      }
      if (lane == 0) s_nkept = nkept + __popcll(keep);
    }
    __syncthreads();
    if (s_nkept >= max_keep) break;
  }
  __syncthreads();
  const int nk = min(min(s_nkept, keep_cap), max_keep);
  if (tid == 0) kept_count[s] = nk;
  if (cand_key != nullptr)
    for (int i = tid; i < nk; i += NMS_FUSED_THREADS)
      kept_key[kbase + i] = cand_key[sbase + kept_pos[kbase + i]];
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
  for (int i = threadIdx.x; i < Wn * 64; i += blockDim.x) {
    const bool dead = (i >= n) || (valid != nullptr && valid[(size_t)s * cap + i] == 0);
    const unsigned bal = __ballot_sync(0xffffffffu, dead);
    if (lane == 0) reinterpret_cast<unsigned*>(remv)[i >> 5] = bal;
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

// Operator path with many ids and more kept keys than one CTA can sort in shared memory: the
// kept lists (each descending, packed at seg_start[g]) are merged by rank counting over the
// whole GPU: thread (j, g) finds the global rank of kept key j of list g with one binary search
// per other list and writes keep[rank] = original index (low half of the key is ~index).
// grid (ceil(max_len / 256), Sg).
__global__ void __launch_bounds__(256)
nms_op_merge_rank_kernel(const u64* __restrict__ kept_key, const int32_t* __restrict__ kept_count,
                         const int32_t* __restrict__ seg_start, int Sg, int max_out,
                         int64_t* __restrict__ keep, int32_t* __restrict__ num_keep) {
  const int g = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (g == 0 && j == 0) {
    int tot = 0;
    for (int i = 0; i < Sg; ++i) tot += kept_count[i];
    num_keep[0] = min(tot, max_out);
  }
  if (j >= kept_count[g]) return;
  const u64 key = kept_key[(size_t)seg_start[g] + j];
  int rank = j;
  for (int g2 = 0; g2 < Sg; ++g2) {
    if (g2 == g) continue;
    const int n2 = kept_count[g2];
    if (n2 > 0 && rank < max_out)
      rank += count_greater_desc(kept_key + (size_t)seg_start[g2], n2, key);
  }
  if (rank < max_out) keep[rank] = (int64_t)(0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFull));
}

// Same contract, for epilogues that identify the element from its key alone
// (Epilogue::kNeedsPos == false): the <= Sg * min(keep_cap, max_out) kept keys of
// image b are gathered into shared memory and bitonic-sorted; the first max_out are
// the result.  O(n log^2 n) on-chip instead of Sg binary searches per element in
// global memory -- 80 class lists (COCO) merge in ~20 us instead of ~700 us.
// Dynamic smem: np2 u64.  Sg <= 1024.
template <class Epilogue>
__global__ void __launch_bounds__(1024)
nms_merge_sort_kernel(const u64* __restrict__ kept_key, const int32_t* __restrict__ kept_count,
                      int Sg, int keep_cap, int max_out, int np2,
                      int32_t* __restrict__ num_out, Epilogue ep,
                      const int32_t* __restrict__ seg_start = nullptr) {
  extern __shared__ __align__(16) u64 s_keys[];
  __shared__ int s_off[1025];
  const int b = blockIdx.x, tid = threadIdx.x;
  if (tid == 0) {
    int tot = 0;
    for (int g = 0; g < Sg; ++g) {
      s_off[g] = tot;
      tot += min(min(kept_count[b * Sg + g], max_out), keep_cap);
    }
    s_off[Sg] = tot;
    num_out[b] = min(tot, max_out);
  }
  __syncthreads();
  const int total = s_off[Sg];
  // one warp per list: coalesced copy of its kept keys
  for (int g = tid >> 5; g < Sg; g += blockDim.x >> 5) {
    const int o = s_off[g], ng = s_off[g + 1] - o;
    const u64* src = kept_key + (seg_start != nullptr ? (size_t)seg_start[b * Sg + g]
                                                      : (size_t)(b * Sg + g) * keep_cap);
    for (int j = tid & 31; j < ng; j += 32) s_keys[o + j] = src[j];
  }
  int np = 1;
  while (np < total) np <<= 1;          // <= np2
  for (int i = total + tid; i < np; i += blockDim.x) s_keys[i] = 0ull;
  if (total > 1) bitonic_sort_desc_u64(s_keys, np);
  else __syncthreads();
  const int nout = min(total, max_out);
  for (int r = tid; r < nout; r += blockDim.x) ep(b, r, -1, -1, s_keys[r]);
  for (int r = nout + tid; r < max_out; r += blockDim.x) ep.pad(b, r);
  (void)np2;
}

template <class Epilogue>
inline int launch_nms_merge(const int32_t* kept_pos, const u64* kept_key,
                            const int32_t* kept_count, int B, int Sg, int keep_cap,
                            int max_out, int32_t* num_out, Epilogue ep,
                            cudaStream_t stream) {
  const int lcap = keep_cap < max_out ? keep_cap : max_out;
  if (!Epilogue::kNeedsPos && Sg <= 1024) {
    int np2 = 1;
    while (np2 < Sg * lcap) np2 <<= 1;
    const size_t sm = (size_t)np2 * sizeof(u64);
    if (sm <= 200 * 1024) {
      if (sm > 32 * 1024) {
        cudaError_t e = ensure_dyn_smem((const void*)nms_merge_sort_kernel<Epilogue>, sm);
        if (e != cudaSuccess) return (int)e;
      }
      nms_merge_sort_kernel<Epilogue><<<B, 1024, sm, stream>>>(
          kept_key, kept_count, Sg, keep_cap, max_out, np2, num_out, ep);
      g_launch_count_add(1);
      BRCNN_CUDA_CHECK_LAST();
      return BRCNN_OK;
    }
  }
  size_t smem = (size_t)Sg * lcap * sizeof(u64);
  int use_smem = 1;
  if (smem > 200 * 1024) { use_smem = 0; smem = 0; }
  if (smem > 48 * 1024) {
    cudaError_t e = ensure_dyn_smem((const void*)nms_merge_kernel<Epilogue>, smem);
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
// the fused kernel keeps min(keep_cap, max_keep) boxes (20 B each) in smem
inline int nms_keep_pad(int keep_cap, int max_keep) {
  const int k = keep_cap < max_keep ? keep_cap : max_keep;
  return ((k > 0 ? k : 1) + 7) & ~7;
}
inline bool nms_use_fused(int keep) { return (size_t)nms_keep_pad(keep, keep) * 20 <= 160 * 1024; }

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
  const int keep_pad = nms_keep_pad(keep_cap, max_keep);
  if (nms_use_fused(keep_pad)) {
    const size_t smem = (size_t)keep_pad * 20;
    if (smem > 48 * 1024) {
      cudaError_t e = ensure_dyn_smem((const void*)nms_fused_kernel, smem);
      if (e != cudaSuccess) return (int)e;
    }
    nms_fused_kernel<<<S, NMS_FUSED_THREADS, smem, stream>>>(
        boxes, valid, count, cap, thr, off, img_maxc, Sg, cand_key, kept_pos, kept_key,
        kept_count, keep_cap, max_keep, keep_pad, (const int32_t*)nullptr);
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
