// K1/K2: RPN proposal generation for a whole batch.
//
//   rpn_select_decode_kernel   one 8-CTA thread-block cluster per (image,
//       level): scores sqrt(sigmoid(cls)*sigmoid(iou)) are computed once from
//       HBM and kept in the cluster's distributed shared memory; an exact
//       radix select (4 x 8-bit digits, cluster-wide histograms over DSMEM)
//       finds the nms_pre-th largest key, ties resolved by lower anchor
//       index; the survivors are gathered into CTA 0, bitonic-sorted, and the
//       cluster decodes them (anchor built on the fly + delta2bbox + clip +
//       min-size flag).
//   nms_mask / nms_sweep / nms_merge (nms_kernels.cuh) finish the job.
//
// Reference behaviour: atss_rpn_head.py:688-760.
#pragma once
#include <cooperative_groups.h>
#include "common.cuh"
#include "nms_kernels.cuh"

namespace brcnn {
namespace cg = cooperative_groups;

constexpr int RPN_CS = 8;         // CTAs per cluster
constexpr int RPN_THREADS = 512;  // threads per CTA

struct RpnLevel {
  const float* cls;   // (B, A, H, W)
  const float* bbox;  // (B, 4A, H, W)
  const float* iou;   // (B, A, H, W)
  int H, W, stride_w, stride_h;
  int n;         // H*W*A
  int k;         // candidates kept = min(n, nms_pre) (all if nms_pre <= 0)
  int idx_base;  // index of this level's first anchor in the concatenation
  int pad;
};

struct RpnArgs {
  RpnLevel lv[BRCNN_MAX_LEVELS];
  int A, L, B, Kc;
  int kpow2;      // smem slots of the gather/sort buffer (pow2 >= max k)
  int slice_cap;  // smem slots of the per-CTA key slice
  float means[4], stds[4];
  float max_ratio, min_size;
};

__device__ __forceinline__ u64 rpn_make_key(uint32_t score_bits,
                                            uint32_t concat_idx) {
  return ((u64)score_bits << 32) | (u64)(0xFFFFFFFFu - concat_idx);
}

__global__ void __cluster_dims__(RPN_CS, 1, 1) __launch_bounds__(RPN_THREADS, 1)
rpn_select_decode_kernel(const __grid_constant__ RpnArgs a,
                         const float* __restrict__ base_anchors,
                         const float* __restrict__ img_hw,
                         float4* __restrict__ cand_boxes,
                         u64* __restrict__ cand_key,
                         uint8_t* __restrict__ cand_valid,
                         int32_t* __restrict__ cand_count,
                         int* __restrict__ img_maxc_bits) {
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int b = blockIdx.y, l = blockIdx.z;
  const RpnLevel& lv = a.lv[l];
  const int A = a.A;
  const int P = lv.H * lv.W;
  const int pp = (P + RPN_CS - 1) / RPN_CS;
  const int p0 = min(rank * pp, P), p1 = min(p0 + pp, P);
  const int nloc = (p1 - p0) * A;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  constexpr int NW = RPN_THREADS / 32;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  u64* sel = reinterpret_cast<u64*>(smem_raw);                  // [kpow2]
  uint32_t* keys = reinterpret_cast<uint32_t*>(sel + a.kpow2);  // [slice_cap]
  __shared__ uint32_t hist[2][256];
  __shared__ uint32_t s_warp[NW];
  __shared__ int s_digit, s_krem, s_slot;
  __shared__ int s_cnt[2];     // n_gt, n_eq of this CTA (read remotely)
  __shared__ int s_plan[3];    // sel_before, take_eq, n_gt

  // ---- phase 0: scores -> keys in shared memory (k-order: (p*A + a)) ----
  {
    const float* cls = lv.cls + (size_t)b * A * P;
    const float* iou = lv.iou + (size_t)b * A * P;
    for (int an = 0; an < A; ++an) {
      const float* c = cls + (size_t)an * P;
      const float* u = iou + (size_t)an * P;
      for (int p = p0 + tid; p < p1; p += RPN_THREADS) {
        float s = sqrtf(pinned_sigmoid(__ldg(c + p)) * pinned_sigmoid(__ldg(u + p)));
        keys[(p - p0) * A + an] = __float_as_uint(s);
      }
    }
  }
  __syncthreads();

  // ---- phase 1: exact k-th largest key over the cluster ----
  const int k = lv.k, n = lv.n;
  uint32_t prefix = 0, pmask = 0;
  int krem = n;  // number of keys == threshold to take (k == n: take all)
  if (k < n) {
    krem = k;
    const int nround = (nloc + 31) & ~31;
    for (int pass = 0; pass < 4; ++pass) {
      const int shift = 24 - 8 * pass;
      uint32_t* h = hist[pass & 1];
      for (int i = tid; i < 256; i += RPN_THREADS) h[i] = 0;
      __syncthreads();
      for (int i = tid; i < nround; i += RPN_THREADS) {
        const uint32_t key = (i < nloc) ? keys[i] : 0u;
        const bool act = (i < nloc) && ((key & pmask) == prefix);
        const uint32_t d = (key >> shift) & 255u;
        const unsigned am = __ballot_sync(0xffffffffu, act);
        if (act) {
          const unsigned peers = __match_any_sync(am, d);
          if (lane == __ffs(peers) - 1) atomicAdd(&h[d], (uint32_t)__popc(peers));
        }
      }
      cluster.sync();
      // cluster-wide histogram, thread t owns digit 255 - t
      uint32_t tot = 0;
      if (tid < 256) {
        const int d = 255 - tid;
#pragma unroll
        for (int r = 0; r < RPN_CS; ++r) tot += cluster.map_shared_rank(h, r)[d];
      }
      // inclusive scan over tid (i.e. over digits descending)
      uint32_t incl = tot;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      if (lane == 31) s_warp[wid] = incl;
      __syncthreads();
      if (tid < 256) {
        uint32_t wbase = 0;
        for (int w = 0; w < wid; ++w) wbase += s_warp[w];
        incl += wbase;
        const uint32_t excl = incl - tot;
        if (excl < (uint32_t)krem && (uint32_t)krem <= incl) {
          s_digit = 255 - tid;
          s_krem = krem - (int)excl;
        }
      }
      __syncthreads();
      prefix |= ((uint32_t)s_digit) << shift;
      pmask |= 255u << shift;
      krem = s_krem;
      __syncthreads();
    }
  }
  const uint32_t T = prefix;

  // ---- phase 2: local counts, cluster plan, gather into CTA 0 ----
  if (tid == 0) { s_cnt[0] = 0; s_cnt[1] = 0; s_slot = 0; }
  __syncthreads();
  {
    int ngt = 0, neq = 0;
    for (int i = tid; i < nloc; i += RPN_THREADS) {
      const uint32_t key = keys[i];
      ngt += (key > T);
      neq += (key == T);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      ngt += __shfl_xor_sync(0xffffffffu, ngt, o);
      neq += __shfl_xor_sync(0xffffffffu, neq, o);
    }
    if (lane == 0) { atomicAdd(&s_cnt[0], ngt); atomicAdd(&s_cnt[1], neq); }
  }
  cluster.sync();
  if (tid == 0) {
    int eq_before = 0, sel_before = 0;
    int my_take = 0, my_gt = 0;
    for (int r = 0; r <= rank; ++r) {
      const int* rc = cluster.map_shared_rank(s_cnt, r);
      const int g = rc[0], e = rc[1];
      int take = krem - eq_before;
      take = take < 0 ? 0 : (take > e ? e : take);
      if (r == rank) { my_take = take; my_gt = g; }
      else { sel_before += g + take; eq_before += e; }
    }
    s_plan[0] = sel_before; s_plan[1] = my_take; s_plan[2] = my_gt;
  }
  __syncthreads();
  {
    const int sel_before = s_plan[0], take = s_plan[1], ngt = s_plan[2];
    const int neq = s_cnt[1];
    u64* sel0 = cluster.map_shared_rank(sel, 0);
    const uint32_t idx0 = (uint32_t)(lv.idx_base + p0 * A);
    const bool take_all_eq = (take == neq);
    const int nround = (nloc + 31) & ~31;
    // unordered part: key > T, and key == T when every tie is taken
    for (int i = tid; i < nround; i += RPN_THREADS) {
      const uint32_t key = (i < nloc) ? keys[i] : 0u;
      const bool pick = (i < nloc) && (key > T || (take_all_eq && key == T));
      const unsigned bm = __ballot_sync(0xffffffffu, pick);
      if (bm) {
        int base = 0;
        if (lane == 0) base = atomicAdd(&s_slot, __popc(bm));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (pick) {
          const int slot = sel_before + base + __popc(bm & ((1u << lane) - 1u));
          sel0[slot] = rpn_make_key(key, idx0 + (uint32_t)i);
        }
      }
    }
    // ordered part: only the first `take` ties in index order
    if (!take_all_eq && take > 0) {
      int running = 0;
      for (int base_i = 0; base_i < nloc && running < take; base_i += RPN_THREADS) {
        const int i = base_i + tid;
        const bool f = (i < nloc) && (keys[i] == T);
        const unsigned bm = __ballot_sync(0xffffffffu, f);
        if (lane == 0) s_warp[wid] = __popc(bm);
        __syncthreads();
        int wbase = 0, chunk = 0;
        for (int w = 0; w < NW; ++w) {
          const int c = (int)s_warp[w];
          if (w < wid) wbase += c;
          chunk += c;
        }
        const int trank = running + wbase + __popc(bm & ((1u << lane) - 1u));
        if (f && trank < take)
          sel0[sel_before + ngt + trank] = rpn_make_key(T, idx0 + (uint32_t)i);
        running += chunk;
        __syncthreads();
      }
    }
  }
  cluster.sync();

  // ---- phase 3: CTA 0 sorts the k survivors (descending key) ----
  int kp = 1;
  while (kp < k) kp <<= 1;
  if (rank == 0) {
    for (int i = k + tid; i < kp; i += RPN_THREADS) sel[i] = 0ull;
    bitonic_sort_desc_u64(sel, kp);
    if (tid == 0) cand_count[b * a.L + l] = k;
  }
  cluster.sync();

  // ---- phase 4: decode the sorted candidates (whole cluster) ----
  {
    const u64* sel0 = cluster.map_shared_rank(sel, 0);
    const float max_h = img_hw[b * 2 + 0], max_w = img_hw[b * 2 + 1];
    const size_t seg = ((size_t)b * a.L + l) * a.Kc;
    const float* bbox = lv.bbox + (size_t)b * 4 * A * P;
    float local_max = 0.f;
    for (int j = rank * RPN_THREADS + tid; j < k; j += RPN_CS * RPN_THREADS) {
      const u64 ck = sel0[j];
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
      Box4 o = delta2bbox_one(roi, d0, d1, d2, d3, a.means, a.stds, a.max_ratio,
                              1, max_w, max_h);
      bool valid = true;
      if (a.min_size >= 0.f) {
        const float w = o.x2 - o.x1, h = o.y2 - o.y1;
        valid = (w > a.min_size) && (h > a.min_size);
      }
      cand_boxes[seg + j] = make_float4(o.x1, o.y1, o.x2, o.y2);
      cand_key[seg + j] = ck;
      cand_valid[seg + j] = valid ? 1 : 0;
      if (valid) local_max = fmaxf(local_max, fmaxf(fmaxf(o.x1, o.y1), fmaxf(o.x2, o.y2)));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
      local_max = fmaxf(local_max, __shfl_xor_sync(0xffffffffu, local_max, o));
    // boxes are clipped to >= 0, so int ordering == float ordering
    if (lane == 0 && local_max > 0.f) atomicMax(img_maxc_bits + b, __float_as_int(local_max));
  }
  cluster.sync();  // keep CTA 0's shared memory alive until all reads are done
}

// epilogue of the per-image merge: proposals[b][rank] = (box, score)
struct RpnMergeEpilogue {
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
