// K2 (RPN): per-image batched NMS in GLOBAL score order with early stop.
//
// Reference: mmcv.ops.batched_nms(ids = pyramid level) + slice [:max_per_img]
// (atss_rpn_head.py:756-760).  The final proposals are the first max_per_img
// kept boxes in (score desc, index asc) order over all levels, and a box's
// keep status depends only on higher-scored kept boxes of ITS OWN level
// (offset boxes of different levels never intersect).  So instead of running
// every level's NMS to completion (up to max_per_img keeps PER LEVEL) and then
// merging, one CTA per image walks the L sorted candidate lists as a lazy
// L-way merge, 64 candidates per round, and stops as soon as max_per_img boxes
// are kept: ~max_per_img/keep_rate candidates are ever touched instead of
// L*nms_pre, and each is tested against its own level's kept list only.
//
// Per round:
//   1. window : next <=64 keys (+boxes, valid flags) of every level -> smem
//   2. rank   : window-rank of every key by binary search in the other
//               levels' windows; rank < 64 <=> member of the global next-64
//   3. pull   : candidate vs kept boxes of its level (kept list in smem)
//   4. diag   : 64x64 same-level suppression bits among the survivors
//   5. resolve: greedy order as a ballot fix-point (see nms_fused_kernel),
//               append keeps, emit proposals rows directly in final order.
// IoU arithmetic: mmcv nms_cpu division form on boxes + level*(max_coord+1)
// added in fp32 (DESIGN.md "Pinned arithmetic"), identical to the segment
// kernels, so results are bit-identical to them.
#pragma once
#include "common.cuh"
#include "nms_kernels.cuh"

namespace brcnn {

constexpr int RNI_THREADS = 512;
constexpr int RNI_TILE = 64;

struct RpnNmsImageSmem {
  // byte offsets into dynamic smem
  int kbox, karea, lidx, total;
};
inline RpnNmsImageSmem rpn_nms_image_smem(int L, int max_out) {
  RpnNmsImageSmem s;
  const int kp = (max_out + 7) & ~7;
  int o = 0;
  s.kbox = o;  o += kp * 16;
  s.karea = o; o += kp * 4;
  s.lidx = o;  o += L * kp * 2;
  s.total = (o + 15) & ~15;
  return s;
}

// grid B, block RNI_THREADS.  L <= BRCNN_MAX_LEVELS.
__global__ void __launch_bounds__(RNI_THREADS)
rpn_nms_image_kernel(const float4* __restrict__ cand_boxes, const u64* __restrict__ cand_key,
                     const uint8_t* __restrict__ cand_valid,
                     const int32_t* __restrict__ cand_count, int L, int Kc, float thr,
                     const float* __restrict__ img_maxc, int max_out,
                     float* __restrict__ proposals, int32_t* __restrict__ num_proposals,
                     RpnNmsImageSmem lay) {
  extern __shared__ __align__(16) unsigned char rni_smem[];
  float4* kbox = reinterpret_cast<float4*>(rni_smem + lay.kbox);
  float* karea = reinterpret_cast<float*>(rni_smem + lay.karea);
  unsigned short* lidx = reinterpret_cast<unsigned short*>(rni_smem + lay.lidx);
  const int kp = (max_out + 7) & ~7;

  __shared__ u64 w_key[BRCNN_MAX_LEVELS][RNI_TILE];      // level windows
  __shared__ float4 w_box[BRCNN_MAX_LEVELS][RNI_TILE];
  __shared__ uint8_t w_valid[BRCNN_MAX_LEVELS][RNI_TILE];
  __shared__ float4 tb[RNI_TILE], traw[RNI_TILE];         // tile: offset / raw boxes
  __shared__ float ta[RNI_TILE];
  __shared__ u64 tkey[RNI_TILE];
  __shared__ int tlvl[RNI_TILE];
  __shared__ u64 diag[RNI_TILE];
  __shared__ u64 s_lmask[BRCNN_MAX_LEVELS];
  __shared__ unsigned s_dead[2];
  __shared__ int s_cursor[BRCNN_MAX_LEVELS], s_count[BRCNN_MAX_LEVELS], s_wn[BRCNN_MAX_LEVELS];
  __shared__ int s_lcnt[BRCNN_MAX_LEVELS], s_taken[BRCNN_MAX_LEVELS];
  __shared__ int s_nkept;

  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31;
  const float maxc1 = img_maxc[b] + 1.0f;
  if (tid < BRCNN_MAX_LEVELS) {
    s_cursor[tid] = 0;
    s_count[tid] = (tid < L) ? min(cand_count[b * L + tid], Kc) : 0;
    s_lcnt[tid] = 0;
  }
  if (tid == 0) s_nkept = 0;
  __syncthreads();

  while (true) {
    // ---- 1. windows ----
    {
      const int l = tid >> 6, i = tid & 63;
      if (l < L) {
        const int pos = s_cursor[l] + i;
        const bool present = pos < s_count[l];
        if (present) {
          const size_t g = ((size_t)b * L + l) * Kc + pos;
          w_key[l][i] = cand_key[g];
          w_box[l][i] = cand_boxes[g];
          w_valid[l][i] = cand_valid != nullptr ? cand_valid[g] : 1;
        }
        if (i == 0) {
          s_wn[l] = max(0, min(RNI_TILE, s_count[l] - s_cursor[l]));
          s_taken[l] = 0;
        }
      } else if (l < BRCNN_MAX_LEVELS && i == 0) {
        s_wn[l] = 0; s_taken[l] = 0;
      }
      if (tid < 2) s_dead[tid] = 0xffffffffu;
      if (tid < BRCNN_MAX_LEVELS) s_lmask[tid] = 0ull;
    }
    __syncthreads();
    int remaining = 0;
    for (int l = 0; l < L; ++l) remaining += s_wn[l];
    if (remaining == 0) break;                      // block-uniform
    const int ntile = min(RNI_TILE, remaining);
    // ---- 2. rank inside the union of the windows ----
    {
      const int l = tid >> 6, i = tid & 63;
      if (l < L && i < s_wn[l]) {
        const u64 key = w_key[l][i];
        int rank = i;
        for (int l2 = 0; l2 < L && rank < RNI_TILE; ++l2) {
          if (l2 == l || s_wn[l2] == 0) continue;
          rank += count_greater_desc(w_key[l2], s_wn[l2], key);
        }
        if (rank < RNI_TILE) {
          const float4 raw = w_box[l][i];
          const float4 ob = add_seg_offset(raw, (float)l * maxc1);
          traw[rank] = raw;
          tb[rank] = ob;
          ta[rank] = (ob.z - ob.x) * (ob.w - ob.y);
          tkey[rank] = key;
          tlvl[rank] = l;
          atomicAdd(&s_taken[l], 1);
          atomicOr(&s_lmask[l], 1ull << rank);
          if (w_valid[l][i]) atomicAnd(&s_dead[rank >> 5], ~(1u << (rank & 31)));
        }
      }
    }
    __syncthreads();
    const int nkept = s_nkept;
    // ---- 3. pull: candidate c vs the kept boxes of its level ----
    {
      const int c = tid & 63, g = tid >> 6;
      bool hit = false;
      if (c < ntile && !((s_dead[c >> 5] >> (c & 31)) & 1u)) {
        const int l = tlvl[c];
        const float4 bx = tb[c];
        const float ba = ta[c];
        const unsigned short* li = lidx + (size_t)l * kp;
        const int n = s_lcnt[l];
        for (int q = g; q < n && !hit; q += 8) {
          const int k = li[q];
          hit = nms_suppresses(kbox[k], karea[k], bx, ba, thr, 0.f);
        }
      }
      const unsigned bal = __ballot_sync(0xffffffffu, hit);
      if (lane == 0 && bal) atomicOr(&s_dead[(tid >> 5) & 1], bal);
    }
    __syncthreads();
    const u64 alive = ~(((u64)s_dead[1] << 32) | (u64)s_dead[0]);
    u64 keep = 0ull;
    if (alive != 0ull) {   // block-uniform
      // ---- 4. diag: row r vs earlier same-level alive candidates ----
      {
        const int r = tid >> 3, g = tid & 7;
        unsigned bits8 = 0;
        if ((alive >> r) & 1ull) {
          const float4 a4 = tb[r];
          const float aa = ta[r];
          const u64 same = s_lmask[tlvl[r]] & alive;
#pragma unroll
          for (int cc8 = 0; cc8 < 8; ++cc8) {
            const int cc = g * 8 + cc8;
            if (cc < r && ((same >> cc) & 1ull) &&
                nms_suppresses(tb[cc], ta[cc], a4, aa, thr, 0.f))
              bits8 |= (1u << cc8);
          }
        }
        u64 word = (u64)bits8 << (8 * g);
        word |= __shfl_xor_sync(0xffffffffu, word, 1);
        word |= __shfl_xor_sync(0xffffffffu, word, 2);
        word |= __shfl_xor_sync(0xffffffffu, word, 4);
        if (g == 0) diag[r] = word;
      }
      __syncthreads();
      // ---- 5. resolve (warp 0) ----
      if (tid < 32) {
        const u64 c0 = diag[lane], c1 = diag[lane + 32];
        const bool a0 = (alive >> lane) & 1ull, a1 = (alive >> (lane + 32)) & 1ull;
        keep = alive;
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
            const u64 below = keep & ((1ull << r) - 1ull);
            const int q = nkept + __popcll(below);
            if (q < max_out) {
              const int l = tlvl[r];
              kbox[q] = tb[r];
              karea[q] = ta[r];
              lidx[(size_t)l * kp + s_lcnt[l] + __popcll(below & s_lmask[l])] =
                  (unsigned short)q;
              const float4 raw = traw[r];
              float* o = proposals + ((size_t)b * max_out + q) * 5;
              o[0] = raw.x; o[1] = raw.y; o[2] = raw.z; o[3] = raw.w;
              o[4] = __uint_as_float((uint32_t)(tkey[r] >> 32));
            }
          }
        }
        __syncwarp();
        if (lane < L) s_lcnt[lane] += __popcll(keep & s_lmask[lane]);
        if (lane == 0) s_nkept = nkept + __popcll(keep);
      }
    }
    if (tid < L) s_cursor[tid] += s_taken[tid];
    __syncthreads();
    if (s_nkept >= max_out) break;
  }
  __syncthreads();
  const int nk = min(s_nkept, max_out);
  if (tid == 0) num_proposals[b] = nk;
  for (int i = nk * 5 + tid; i < max_out * 5; i += RNI_THREADS)
    proposals[(size_t)b * max_out * 5 + i] = 0.f;
}

}  // namespace brcnn
