// K1: RPN proposal generation for a whole batch (every image, every level).
//
//   rpn_score_kernel        whole-GPU pass over the RPN outputs: an APPROXIMATE score
//       sqrt(sigmoid(cls) * sigmoid(iou)) from the special-function unit (|error| < 2e-5,
//       25 instructions instead of the 130 of the pinned IEEE chain that made this kernel
//       ALU-bound at 29 % of the HBM roofline) written once (plane order, coalesced) plus a
//       2048-bin histogram of the score VALUE (bin = floor(s * 2048), uniform over [0,1]) per
//       (image, level) segment (shared-memory privatised, flushed with integer atomics).
//   rpn_collect_kernel      whole-GPU pass over the L2-resident approximate scores: every CTA
//       re-derives its segment's threshold bin d (#(bins > d) < k <= #(bins >= d)) from the
//       histogram, keeps the elements of bins >= d - 1 -- a superset of the EXACT top-k,
//       because the approximation error is a small fraction of a bin -- and recomputes their
//       scores with the pinned arithmetic: 64-bit composites (exact_score_bits << 32 |
//       ~anchor_index) go to the segment's candidate buffer (CTA-aggregated: one global atomic
//       per CTA).  Only ~2 % of the anchors pay for the exact arithmetic.  Segments that go to
//       the radix path below get their keys rewritten with exact scores here.
//   rpn_topk_decode_kernel  one CTA per segment: bitonic sort of the collected
//       composites in shared memory = (score desc, index asc) with no ties,
//       first k decoded (anchor built on the fly + delta2bbox + clip +
//       min-size flag).  If the threshold bin is over-populated (degenerate
//       score distributions) the CTA falls back to an exact radix selection
//       on the composite by re-scanning the keys itself.
//   nms_fused / nms_mask+sweep / nms_merge (nms_kernels.cuh) finish the job.
//
// Reference behaviour: atss_rpn_head.py:688-760.
#pragma once
#include "common.cuh"
#include "nms_kernels.cuh"

namespace brcnn {

constexpr int RPN_SCORE_THREADS = 256;
constexpr int RPN_SCORE_CHUNK = 2048;   // elements per CTA of the score kernel
constexpr int RPN_BINS = 2048;          // 11-bit digits
constexpr int RPN_TOPK_THREADS = 1024;

struct RpnLevel {
  const float* cls;   // (B, A, H, W)
  const float* bbox;  // (B, 4A, H, W)
  const float* iou;   // (B, A, H, W)
  int H, W, stride_w, stride_h;
  int n;           // H*W*A
  int k;           // candidates kept = min(n, nms_pre) (all if nms_pre <= 0)
  int idx_base;    // index of this level's first anchor in the concatenation
  int chunk_base;  // first score-kernel chunk of this level (per image)
  int key_off;     // offset of this level's keys in an image's key row (x4 aligned)
  int pad;
};

struct RpnArgs {
  RpnLevel lv[BRCNN_MAX_LEVELS];
  int A, L, B, Kc;
  int key_stride; // keys per image (levels padded to multiples of 4)
  int cand_cap;   // smem candidate slots of the top-k kernel (pow2)
  float means[4], stds[4];
  float max_ratio, min_size;
};

// histogram bin of a score: monotone in s on [0,1]; NaN (largest key bits) -> top bin
__device__ __forceinline__ uint32_t rpn_value_bin(float s) {
  if (!(s == s)) return RPN_BINS - 1;
  const int bin = (int)(s * (float)RPN_BINS);   // exact scaling by a power of two
  return (uint32_t)min(max(bin, 0), RPN_BINS - 1);
}
// key bits of the lower edge of bin d: bin(s) >= d  <=>  key_bits(s) >= this (s >= 0 or NaN)
__device__ __forceinline__ uint32_t rpn_bin_floor_bits(int d) {
  return __float_as_uint((float)d / (float)RPN_BINS);
}

__device__ __forceinline__ u64 rpn_make_key(uint32_t score_bits, uint32_t concat_idx) {
  return ((u64)score_bits << 32) | (u64)(0xFFFFFFFFu - concat_idx);
}

// Approximate sqrt(sigmoid(c) * sigmoid(i)) from the special-function unit (ex2 / rcp / sqrt
// .approx, ~25 instructions instead of ~130 for the pinned IEEE chain).  Absolute error
// < 2e-5 on [0, 1] (relative error of each approximate op <= 2^-21, argument rounding of the
// exponential <= |x| * 2^-23 where the sigmoid's slope is <= 1/4), i.e. far below one histogram
// bin (1 / 2048).  Only the per-level CANDIDATE SELECTION uses it: every candidate's key is
// recomputed with the pinned arithmetic (rpn_exact_key) before anything is ranked.
__device__ __forceinline__ float rpn_fast_score(float c, float i) {
  const float p = __fdividef(1.0f, 1.0f + __expf(-c)) * __fdividef(1.0f, 1.0f + __expf(-i));
  float s;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(p));
  return s;
}
// the reference's score, op for op (atss_rpn_head.py:722: (cls.sigmoid() * iou.sigmoid()).sqrt())
__device__ __forceinline__ uint32_t rpn_exact_key(float c, float i) {
  return __float_as_uint(sqrtf(pinned_sigmoid(c) * pinned_sigmoid(i)));
}

// grid (chunks_per_image, B)
__global__ void __launch_bounds__(RPN_SCORE_THREADS)
rpn_score_kernel(const __grid_constant__ RpnArgs a, uint32_t* __restrict__ keys,
                 uint32_t* __restrict__ ghist) {
  __shared__ uint32_t sh[RPN_BINS];
  const int b = blockIdx.y;
  int l = 0;
  while (l + 1 < a.L && (int)blockIdx.x >= a.lv[l + 1].chunk_base) ++l;
  const RpnLevel& lv = a.lv[l];
  const int e0 = ((int)blockIdx.x - lv.chunk_base) * RPN_SCORE_CHUNK;
  for (int i = threadIdx.x; i < RPN_BINS; i += RPN_SCORE_THREADS) sh[i] = 0;
  __syncthreads();
  const float* cls = lv.cls + (size_t)b * lv.n;
  const float* iou = lv.iou + (size_t)b * lv.n;
  uint32_t* kout = keys + (size_t)b * a.key_stride + lv.key_off;
  // all 16 loads of the thread in flight before the arithmetic
  constexpr int NJ = RPN_SCORE_CHUNK / RPN_SCORE_THREADS;
  float vc[NJ], vi[NJ];
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const int e = e0 + j * RPN_SCORE_THREADS + threadIdx.x;
    vc[j] = 0.f; vi[j] = 0.f;
    if (e < lv.n) { vc[j] = __ldcs(cls + e); vi[j] = __ldcs(iou + e); }
  }
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const int e = e0 + j * RPN_SCORE_THREADS + threadIdx.x;
    if (e < lv.n) {
      const float s = rpn_fast_score(vc[j], vi[j]);
      kout[e] = __float_as_uint(s);
      atomicAdd(&sh[rpn_value_bin(s)], 1u);
    }
  }
  __syncthreads();
  uint32_t* gh = ghist + ((size_t)b * a.L + l) * RPN_BINS;
  for (int i = threadIdx.x; i < RPN_BINS; i += RPN_SCORE_THREADS) {
    const uint32_t c = sh[i];
    if (c) atomicAdd(gh + i, c);
  }
}

// Block-wide search over a histogram of nb (<= 2048) bins for the digit d with
//   #(digits > d) < krem <= #(digits >= d).  Returns via smem out[0]=d,
// out[1]=#(digits > d), out[2]=hist[d].  All RPN_TOPK_THREADS threads call it.
__device__ __forceinline__ void rpn_find_digit(const uint32_t* sh, int nb, uint32_t krem,
                                               uint32_t* s_warp, int* out) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  // thread t owns digits d1 = nb-1-2t and d2 = nb-2-2t (descending order)
  const int d1 = nb - 1 - 2 * tid, d2 = d1 - 1;
  const uint32_t c1 = (d1 >= 0) ? sh[d1] : 0u, c2 = (d2 >= 0) ? sh[d2] : 0u;
  uint32_t incl = c1 + c2;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) s_warp[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    uint32_t w = s_warp[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t v = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += v;
    }
    s_warp[lane] = w;  // inclusive over warps
  }
  __syncthreads();
  const uint32_t wbase = (wid == 0) ? 0u : s_warp[wid - 1];
  incl += wbase;
  const uint32_t excl = incl - (c1 + c2);
  if (excl < krem && krem <= excl + c1) {
    out[0] = d1; out[1] = (int)excl; out[2] = (int)c1;
  } else if (excl + c1 < krem && krem <= incl) {
    out[0] = d2; out[1] = (int)(excl + c1); out[2] = (int)c2;
  }
  __syncthreads();
}

// Same search with 256 threads (8 digits per thread, descending).
__device__ __forceinline__ void rpn_find_digit_256(const uint32_t* __restrict__ hist, int nb,
                                                   uint32_t krem, uint32_t* s_warp, int* out) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  uint32_t c[8];
  uint32_t sum = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int d = nb - 1 - 8 * tid - j;
    c[j] = (d >= 0) ? hist[d] : 0u;
    sum += c[j];
  }
  uint32_t incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) s_warp[wid] = incl;
  __syncthreads();
  uint32_t wbase = 0;
  for (int w = 0; w < wid; ++w) wbase += s_warp[w];
  uint32_t run = incl - sum + wbase;   // #(digits above this thread's first digit)
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (run < krem && krem <= run + c[j]) {
      out[0] = nb - 1 - 8 * tid - j; out[1] = (int)run; out[2] = (int)c[j];
    }
    run += c[j];
  }
  __syncthreads();
}

// grid (chunks_per_image, B), same chunking as the score kernel.
// `keys` hold the APPROXIMATE scores of rpn_score_kernel.  With d = the histogram bin of the
// k-th best approximate score, every element of the exact top-k has an approximate score in
// bin >= d - 1 (the approximation error is a small fraction of a bin), so the chunk keeps the
// elements of bins >= d - 1 and recomputes THEIR keys with the pinned arithmetic; the top-k
// kernel ranks exact keys only.  A segment whose candidates would not fit (degenerate score
// distribution) goes to the top-k kernel's exact radix path instead: this kernel then rewrites
// the segment's keys with exact scores in place.
__global__ void __launch_bounds__(RPN_SCORE_THREADS)
rpn_collect_kernel(const __grid_constant__ RpnArgs a, uint32_t* __restrict__ keys,
                   const uint32_t* __restrict__ ghist, u64* __restrict__ cand_raw,
                   int32_t* __restrict__ cand_n) {
  __shared__ u64 s_buf[RPN_SCORE_CHUNK];
  __shared__ uint32_t s_warp[RPN_SCORE_THREADS / 32];
  __shared__ int s_out[3];
  __shared__ int s_cnt, s_base;
  const int b = blockIdx.y;
  int l = 0;
  while (l + 1 < a.L && (int)blockIdx.x >= a.lv[l + 1].chunk_base) ++l;
  const RpnLevel& lv = a.lv[l];
  const int seg = b * a.L + l;
  const int n = lv.n, k = lv.k, P = lv.H * lv.W, A = a.A;
  const int tid = threadIdx.x, lane = tid & 31;
  // this thread's keys: loads issued before the (latency-bound) threshold search
  constexpr int NJ = RPN_SCORE_CHUNK / RPN_SCORE_THREADS;
  const int e0 = ((int)blockIdx.x - lv.chunk_base) * RPN_SCORE_CHUNK;
  uint32_t* kseg = keys + (size_t)b * a.key_stride + lv.key_off;
  const float* cls = lv.cls + (size_t)b * lv.n;
  const float* iou = lv.iou + (size_t)b * lv.n;
  uint32_t kreg[NJ];
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const int e = e0 + j * RPN_SCORE_THREADS + tid;
    kreg[j] = (e < n) ? __ldg(kseg + e) : 0u;
  }
  uint32_t thr = 0u;
  bool exact_path = false;           // block-uniform
  if (k < n) {
    const uint32_t* hist = ghist + (size_t)seg * RPN_BINS;
    rpn_find_digit_256(hist, RPN_BINS, (uint32_t)k, s_warp, s_out);
    const int d = s_out[0];
    const int margin = d > 0 ? (int)hist[d - 1] : 0;
    if (s_out[1] + s_out[2] + margin > a.cand_cap) exact_path = true;   // same rule as the top-k kernel
    else thr = rpn_bin_floor_bits(d > 0 ? d - 1 : 0);
  } else if (n > a.cand_cap) {
    exact_path = true;
  }
  if (exact_path) {
#pragma unroll 1
    for (int j = 0; j < NJ; ++j) {
      const int e = e0 + j * RPN_SCORE_THREADS + tid;
      if (e < n) kseg[e] = rpn_exact_key(__ldg(cls + e), __ldg(iou + e));
    }
    return;
  }
  if (tid == 0) s_cnt = 0;
  __syncthreads();
  // ---- compact the chunk's candidates (element numbers) ----
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const int e = e0 + j * RPN_SCORE_THREADS + tid;
    const bool take = (e < n) && (kreg[j] >= thr);
    const unsigned m = __ballot_sync(0xffffffffu, take);
    if (m) {
      const int leader = __ffs(m) - 1;
      int base = 0;
      if (lane == leader) base = atomicAdd(&s_cnt, __popc(m));
      base = __shfl_sync(0xffffffffu, base, leader);
      if (take) s_buf[base + __popc(m & ((1u << lane) - 1u))] = (u64)(uint32_t)e;
    }
  }
  __syncthreads();
  const int cnt = s_cnt;
  if (cnt == 0) return;
  if (tid == 0) s_base = atomicAdd(cand_n + seg, cnt);
  // ---- exact keys of the candidates, all lanes busy ----
  const float inv_P = 1.0f / (float)P;
  for (int i = tid; i < cnt; i += RPN_SCORE_THREADS) {
    const int e = (int)(uint32_t)s_buf[i];     // slot i is read and rewritten by this thread only
    const uint32_t key = rpn_exact_key(__ldg(cls + e), __ldg(iou + e));
    // e = an*P + p -> concatenated anchor index idx_base + p*A + an
    int an = (n < (1 << 24)) ? __float2int_rz(__int2float_rn(e) * inv_P) : e / P;
    int rem = e - an * P;
    an += (rem >= P) - (rem < 0);
    const int p = e - an * P;
    s_buf[i] = rpn_make_key(key, (uint32_t)(lv.idx_base + p * A + an));
  }
  __syncthreads();
  u64* dst = cand_raw + (size_t)seg * a.cand_cap + s_base;
  for (int i = tid; i < cnt; i += RPN_SCORE_THREADS) dst[i] = s_buf[i];
}

// grid (B, L): x = image so that the heavy level-0 CTAs are scheduled first.
// dynamic smem: cand_cap u64
__global__ void __launch_bounds__(RPN_TOPK_THREADS)
rpn_topk_decode_kernel(const __grid_constant__ RpnArgs a,
                       const uint32_t* __restrict__ keys,
                       const uint32_t* __restrict__ ghist,
                       const float* __restrict__ base_anchors,
                       const float* __restrict__ img_hw,
                       float4* __restrict__ cand_boxes, u64* __restrict__ cand_key,
                       uint8_t* __restrict__ cand_valid, int32_t* __restrict__ cand_count,
                       int* __restrict__ img_maxc_bits, const u64* __restrict__ cand_raw,
                       const int32_t* __restrict__ cand_n) {
  extern __shared__ __align__(16) u64 cand[];
  __shared__ uint32_t sh[RPN_BINS];
  __shared__ uint32_t s_warp[32];
  __shared__ int s_out[3];
  __shared__ int s_ncand;
  const int b = blockIdx.x, l = blockIdx.y;
  const RpnLevel& lv = a.lv[l];
  const int A = a.A, P = lv.H * lv.W, n = lv.n, k = lv.k;
  const int tid = threadIdx.x, lane = tid & 31;
  const uint32_t* kseg = keys + (size_t)b * a.key_stride + lv.key_off;  // 16 B aligned
  const int seg = b * a.L + l;
  const int total = n;  // plane order: e = an*P + p

  // Visits every key whose score half can still matter: f(e, key) is called
  // for elements with (key & pm_hi) >= pr_hi.  The keys are L2-resident;
  // 128-bit loads, two in flight per thread, one cheap reject per key.
  auto scan = [&](uint32_t pm_hi, uint32_t pr_hi, auto&& f) {
    const uint4* k4 = reinterpret_cast<const uint4*>(kseg);
    const int nvec = (total + 3) >> 2;
    for (int v0 = tid; v0 < nvec; v0 += 2 * RPN_TOPK_THREADS) {
      const int v1 = v0 + RPN_TOPK_THREADS;
      const uint4 q0 = __ldg(k4 + v0);
      uint4 q1 = make_uint4(0u, 0u, 0u, 0u);
      if (v1 < nvec) q1 = __ldg(k4 + v1);
      const uint32_t kk[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int e = ((j < 4) ? v0 : v1) * 4 + (j & 3);
        if ((kk[j] & pm_hi) >= pr_hi && e < total && (j < 4 || v1 < nvec)) f(e, kk[j]);
      }
    }
  };
  // e = an*P + p  ->  concatenated anchor index idx_base + p*A + an.  The
  // division by the per-level constant P is a float reciprocal + fix-up while
  // e is exactly representable in fp32 (e < 2^24), an integer division otherwise.
  const float inv_P = 1.0f / (float)P;
  const bool small_e = total < (1 << 24);
  auto composite = [&](int e, uint32_t key) {
    int an;
    if (small_e) {
      an = __float2int_rz(__int2float_rn(e) * inv_P);
      const int rem = e - an * P;
      an += (rem >= P) - (rem < 0);
    } else {
      an = e / P;
    }
    const int p = e - an * P;
    return rpn_make_key(key, (uint32_t)(lv.idx_base + p * A + an));
  };

  // ---- fast path: the collect kernel already gathered a superset of the top-k ----
  bool fast;
  if (k < n) {
    for (int i = tid; i < RPN_BINS; i += RPN_TOPK_THREADS) sh[i] = ghist[(size_t)seg * RPN_BINS + i];
    __syncthreads();
    rpn_find_digit(sh, RPN_BINS, (uint32_t)k, s_warp, s_out);
    // same rule as rpn_collect_kernel (bins >= d - 1 were collected)
    fast = (s_out[1] + s_out[2] + (s_out[0] > 0 ? (int)sh[s_out[0] - 1] : 0) <= a.cand_cap);
    __syncthreads();
  } else {
    fast = (n <= a.cand_cap);
  }
  if (fast) {
    // The collected set is bins >= d - 1 of the approximate score.  If at least k of its
    // EXACT scores lie in bins >= d, nothing below bin d can be in the top-k: drop it, so the
    // margin bin does not push the bitonic sort over the next power of two (k = 1000 sits
    // just under 1024).
    const int nc = cand_n[seg];
    const u64* src = cand_raw + (size_t)seg * a.cand_cap;
    const uint32_t cut = (k < n && s_out[0] > 0) ? rpn_bin_floor_bits(s_out[0]) : 0u;
    if (tid == 0) { s_ncand = 0; s_out[1] = 0; }
    __syncthreads();
    int mine = 0;
    for (int i = tid; i < nc; i += RPN_TOPK_THREADS) mine += ((uint32_t)(src[i] >> 32) >= cut);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
    if (lane == 0 && mine) atomicAdd(&s_out[1], mine);
    __syncthreads();
    const uint32_t keep_from = (s_out[1] >= k) ? cut : 0u;
    for (int i0 = 0; i0 < nc; i0 += RPN_TOPK_THREADS) {
      const int i = i0 + tid;
      const u64 c = (i < nc) ? src[i] : 0ull;
      const bool take = (i < nc) && ((uint32_t)(c >> 32) >= keep_from);
      const unsigned m = __ballot_sync(0xffffffffu, take);
      if (m) {
        const int leader = __ffs(m) - 1;
        int base = 0;
        if (lane == leader) base = atomicAdd(&s_ncand, __popc(m));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (take) cand[base + __popc(m & ((1u << lane) - 1u))] = c;
      }
    }
    __syncthreads();
  } else {
  // ---- slow path: exact selection threshold on the 64-bit composite ----
  // stop refining once the survivors fit the sort we would do anyway
  // (the bitonic sort pads to a power of two: refining until the survivors fit
  // next_pow2(k) halves the sort compared with stopping at 2*next_pow2(k))
  int sort_cap = 2;
  while (sort_cap < k) sort_cap <<= 1;
  if (sort_cap - k < 64) sort_cap <<= 1;   // too little slack: an extra key scan costs more
  if (sort_cap > a.cand_cap) sort_cap = a.cand_cap;
  u64 prefix = 0, pmask = 0;
  if (k < n) {
    uint32_t krem = (uint32_t)k;
    int above = 0;
    for (int pass = 0; pass < 6; ++pass) {
      const int shift = (pass < 5) ? 53 - 11 * pass : 0;
      const int nb = (pass < 5) ? RPN_BINS : 512;
      for (int i = tid; i < RPN_BINS; i += RPN_TOPK_THREADS) sh[i] = 0;
      __syncthreads();
      scan((uint32_t)(pmask >> 32), (uint32_t)(prefix >> 32), [&](int e, uint32_t key) {
        const u64 c = composite(e, key);
        if ((c & pmask) == prefix) atomicAdd(&sh[(uint32_t)(c >> shift) & (nb - 1)], 1u);
      });
      __syncthreads();
      rpn_find_digit(sh, nb, krem, s_warp, s_out);
      const int d = s_out[0], excl = s_out[1], cnt = s_out[2];
      prefix |= (u64)d << shift;
      pmask |= (u64)(nb - 1) << shift;
      above += excl;
      krem -= (uint32_t)excl;
      __syncthreads();
      if (above + cnt <= sort_cap) break;
    }
  }

  // ---- collect every composite >= the threshold bin ----
  if (tid == 0) s_ncand = 0;
  __syncthreads();
  {
    // warp-converged scan with ONE shared-memory atomic per warp per key slot
    // (a per-key atomicAdd on a single counter serialises ~2k survivors)
    const uint32_t pm_hi = (uint32_t)(pmask >> 32), pr_hi = (uint32_t)(prefix >> 32);
    const uint4* k4 = reinterpret_cast<const uint4*>(kseg);
    const int nvec = (total + 3) >> 2;
    for (int vb = tid - lane; vb < nvec; vb += 2 * RPN_TOPK_THREADS) {
      const int v0 = vb + lane, v1 = v0 + RPN_TOPK_THREADS;
      uint4 q0 = make_uint4(0u, 0u, 0u, 0u), q1 = q0;
      if (v0 < nvec) q0 = __ldg(k4 + v0);
      if (v1 < nvec) q1 = __ldg(k4 + v1);
      const uint32_t kk[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int e = ((j < 4) ? v0 : v1) * 4 + (j & 3);
        u64 c = 0ull;
        bool take = false;
        if (e < total && (kk[j] & pm_hi) >= pr_hi) {
          c = composite(e, kk[j]);
          take = (c & pmask) >= prefix;
        }
        const unsigned m = __ballot_sync(0xffffffffu, take);
        if (m) {
          const int leader = __ffs(m) - 1;
          int base = 0;
          if (lane == leader) base = atomicAdd(&s_ncand, __popc(m));
          base = __shfl_sync(0xffffffffu, base, leader);
          if (take) cand[base + __popc(m & ((1u << lane) - 1u))] = c;
        }
      }
    }
  }
  __syncthreads();
  }  // slow path
  const int ncand = s_ncand;
  int np = 1;
  while (np < ncand) np <<= 1;
  for (int i = ncand + tid; i < np; i += RPN_TOPK_THREADS) cand[i] = 0ull;
  bitonic_sort_desc_u64(cand, np);
  if (tid == 0) cand_count[seg] = k;

  // ---- decode the k best (already in (score desc, index asc) order) ----
  const float max_h = img_hw[b * 2 + 0], max_w = img_hw[b * 2 + 1];
  const size_t sbase = (size_t)seg * a.Kc;
  const float* bbox = lv.bbox + (size_t)b * 4 * A * P;
  float local_max = 0.f;
  for (int j = tid; j < k; j += RPN_TOPK_THREADS) {
    const u64 ck = cand[j];
    const uint32_t cidx = 0xFFFFFFFFu - (uint32_t)(ck & 0xFFFFFFFFull);
    const int idx = (int)(cidx - (uint32_t)lv.idx_base);
    const int p = idx / A, an = idx - p * A;
    const int y = p / lv.W, x = p - y * lv.W;
    const float* d = bbox + (size_t)(an * 4) * P + p;
    const float d0 = __ldg(d), d1 = __ldg(d + P), d2 = __ldg(d + 2 * (size_t)P),
                d3 = __ldg(d + 3 * (size_t)P);
    const float4 ba = __ldg(reinterpret_cast<const float4*>(base_anchors) + l * A + an);
    const float sx = (float)(x * lv.stride_w), sy = (float)(y * lv.stride_h);
    Box4 roi;
    roi.x1 = ba.x + sx; roi.y1 = ba.y + sy; roi.x2 = ba.z + sx; roi.y2 = ba.w + sy;
    const Box4 o = delta2bbox_one(roi, d0, d1, d2, d3, a.means, a.stds, a.max_ratio, 1,
                                  max_w, max_h);
    bool valid = true;
    if (a.min_size >= 0.f) {
      const float w = o.x2 - o.x1, h = o.y2 - o.y1;
      valid = (w > a.min_size) && (h > a.min_size);
    }
    cand_boxes[sbase + j] = make_float4(o.x1, o.y1, o.x2, o.y2);
    cand_key[sbase + j] = ck;
    cand_valid[sbase + j] = valid ? 1 : 0;
    if (valid) local_max = fmaxf(local_max, fmaxf(fmaxf(o.x1, o.y1), fmaxf(o.x2, o.y2)));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    local_max = fmaxf(local_max, __shfl_xor_sync(0xffffffffu, local_max, o));
  // boxes are clipped to >= 0, so int ordering == float ordering
  if (lane == 0 && local_max > 0.f) atomicMax(img_maxc_bits + b, __float_as_int(local_max));
}

// epilogue of the per-image merge: proposals[b][rank] = (box, score)
struct RpnMergeEpilogue {
  static constexpr bool kNeedsPos = true;   // operator() uses (seg, pos)
  const float4* cand_boxes;
  float* proposals;  // (B, max_out, 5)
  int Kc, max_out;
  __device__ void operator()(int b, int rank, int seg, int pos, u64 key) const {
    const float4 bx = cand_boxes[(size_t)seg * Kc + pos];
    float* o = proposals + ((size_t)b * max_out + rank) * 5;
    o[0] = bx.x; o[1] = bx.y; o[2] = bx.z; o[3] = bx.w;
    o[4] = __uint_as_float((uint32_t)(key >> 32));
  }
  __device__ void pad(int b, int rank) const {
    float* o = proposals + ((size_t)b * max_out + rank) * 5;
    o[0] = 0.f; o[1] = 0.f; o[2] = 0.f; o[3] = 0.f; o[4] = 0.f;
  }
};

}  // namespace brcnn
