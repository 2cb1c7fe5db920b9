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
// A thread-block CLUSTER of CS CTAs works on one image: the kept list is
// distributed round-robin over the CTAs' shared memories, every CTA pulls the
// tile against its share, the 64-bit dead masks are OR-combined through
// distributed shared memory (one cluster barrier per round, masks double
// buffered), and the cheap, deterministic steps (windows, ranks, diag,
// resolve) are replicated in every CTA so no result ever has to be broadcast.
// IoU arithmetic: mmcv nms_cpu division form on boxes + level*(max_coord+1)
// added in fp32 (DESIGN.md "Pinned arithmetic"), identical to the segment
// kernels, so results are bit-identical to them.
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"
#include "nms_kernels.cuh"

namespace brcnn {

namespace cg = cooperative_groups;

constexpr int RNI_THREADS = 512;
constexpr int RNI_TILE = 64;
constexpr int RNI_CLUSTER = 8;

struct RpnNmsImageSmem {
  int kp;                     // local kept capacity (multiple of 8)
  int kbox, lidx, total;      // byte offsets into dynamic smem
};
inline RpnNmsImageSmem rpn_nms_image_smem(int L, int max_out, int cs) {
  RpnNmsImageSmem s;
  s.kp = (((max_out + cs - 1) / cs) + 7) & ~7;
  int o = 0;
  s.kbox = o;  o += s.kp * 16;
  s.lidx = o;  o += L * s.kp * 2;
  s.total = (o + 15) & ~15;
  return s;
}

// grid B*CS (cluster dims CS x 1 x 1), block RNI_THREADS.  L <= BRCNN_MAX_LEVELS.
template <int CS>
__global__ void __launch_bounds__(RNI_THREADS)
rpn_nms_image_kernel(const float4* __restrict__ cand_boxes, const u64* __restrict__ cand_key,
                     const uint8_t* __restrict__ cand_valid,
                     const int32_t* __restrict__ cand_count, int L, int Kc, float thr,
                     const float* __restrict__ img_maxc, int max_out,
                     float* __restrict__ proposals, int32_t* __restrict__ num_proposals,
                     RpnNmsImageSmem lay, long long* __restrict__ dbg,
                     const int32_t* __restrict__ seg_start, float off,
                     int64_t* __restrict__ keep_idx) {
  extern __shared__ __align__(16) unsigned char rni_smem[];
  float4* kbox = reinterpret_cast<float4*>(rni_smem + lay.kbox);            // local share
  unsigned short* lidx = reinterpret_cast<unsigned short*>(rni_smem + lay.lidx);
  const int kp = lay.kp;

  __shared__ u64 w_key[BRCNN_MAX_LEVELS][RNI_TILE];      // level windows
  __shared__ float4 w_box[BRCNN_MAX_LEVELS][RNI_TILE];
  __shared__ uint8_t w_valid[BRCNN_MAX_LEVELS][RNI_TILE];
  __shared__ float4 tb[RNI_TILE], traw[RNI_TILE];         // tile: offset / raw boxes
  __shared__ float ta[RNI_TILE];
  __shared__ u64 tkey[RNI_TILE];
  __shared__ int tlvl[RNI_TILE];
  __shared__ u64 diag[RNI_TILE];
  __shared__ u64 s_lmask[BRCNN_MAX_LEVELS];
  __shared__ unsigned s_lmask32[BRCNN_MAX_LEVELS][2];
  __shared__ unsigned s_dslice[2][RNI_TILE / CS][2];   // this CTA's diag rows (double buffered)
  __shared__ u64 s_hit[2];          // this CTA's pull hits, double buffered by round parity
  __shared__ u64 s_alive;
  __shared__ unsigned s_dead[2], s_hitw[2];
  __shared__ int s_cursor[BRCNN_MAX_LEVELS], s_count[BRCNN_MAX_LEVELS], s_wn[BRCNN_MAX_LEVELS];
  __shared__ int s_lcnt[BRCNN_MAX_LEVELS];   // LOCAL kept count per level
  __shared__ int s_taken[BRCNN_MAX_LEVELS];
  __shared__ int s_nkept;                    // GLOBAL kept count (same in every CTA)
  __shared__ int s_lkept;                    // LOCAL kept count (all levels)

  const long long t_entry = dbg != nullptr ? clock64() : 0;
  const int crank = (CS > 1) ? (int)cg::this_cluster().block_rank() : 0;
  const int b = blockIdx.x / CS;
  const int tid = threadIdx.x, lane = tid & 31;
  const float maxc1 = img_maxc != nullptr ? img_maxc[b] + 1.0f : 0.0f;   // no ids: no offset
  if (tid < BRCNN_MAX_LEVELS) {
    s_cursor[tid] = 0;
    s_count[tid] = (tid < L) ? min(cand_count[b * L + tid], Kc) : 0;
    s_lcnt[tid] = 0;
  }
  if (tid == 0) { s_nkept = 0; s_lkept = 0; }
  __syncthreads();

  // window element (l, i) of this thread, prefetched one round ahead
  const int wl = tid >> 6, wi = tid & 63;
  u64 pk = 0ull;
  float4 pb = make_float4(0.f, 0.f, 0.f, 0.f);
  uint8_t pv = 0;
  auto prefetch = [&](int cursor) {
    const int pos = cursor + wi;
    if (wl < L && pos < s_count[wl]) {
      const size_t g = (seg_start != nullptr ? (size_t)seg_start[b * L + wl]
                                             : ((size_t)b * L + wl) * Kc) + pos;
      pk = cand_key[g];
      pb = cand_boxes[g];
      pv = cand_valid != nullptr ? cand_valid[g] : 1;
    }
  };
  if (wl < BRCNN_MAX_LEVELS) prefetch(0);

  long long t_prev = 0, t_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int n_rounds = 0;
#define RNI_MARK(i) do { if (dbg != nullptr && tid == 0) { const long long t_now = clock64(); t_acc[i] += t_now - t_prev; t_prev = t_now; } } while (0)
  if (dbg != nullptr && tid == 0) { t_prev = clock64(); t_acc[7] = t_prev - t_entry; }
  for (int round = 0;; ++round) {
    ++n_rounds;
    // ---- 1. windows: registers -> smem ----
    if (wl < BRCNN_MAX_LEVELS) {
      w_key[wl][wi] = pk; w_box[wl][wi] = pb; w_valid[wl][wi] = pv;
      if (wi == 0) {
        s_wn[wl] = (wl < L) ? max(0, min(RNI_TILE, s_count[wl] - s_cursor[wl])) : 0;
        s_taken[wl] = 0;
      }
    }
    if (tid < 2) { s_dead[tid] = 0xffffffffu; s_hitw[tid] = 0u; }
    if (tid < RNI_TILE) tlvl[tid] = -1;
    __syncthreads();
    RNI_MARK(0);
    int remaining = 0;
    for (int l = 0; l < L; ++l) remaining += s_wn[l];
    if (remaining == 0) break;                      // cluster-uniform
    const int ntile = min(RNI_TILE, remaining);
    // ---- 2. rank inside the union of the windows ----
    if (wl < L && wi < s_wn[wl]) {
      const u64 key = w_key[wl][wi];
      // interleaved binary searches over the other levels' windows (independent
      // chains -> ILP); lo[j] ends as #keys of that window greater than `key`
      int lo[BRCNN_MAX_LEVELS], hi[BRCNN_MAX_LEVELS];
#pragma unroll
      for (int l2 = 0; l2 < BRCNN_MAX_LEVELS; ++l2) {
        lo[l2] = 0;
        hi[l2] = (l2 < L && l2 != wl) ? s_wn[l2] : 0;
      }
      // branch-free steps: a finished search (lo == hi) re-reads a clamped slot and
      // keeps its bounds
#define RNI_SEARCH(NL)                                                        \
      _Pragma("unroll") for (int step = 0; step < 7; ++step) {                \
        _Pragma("unroll") for (int l2 = 0; l2 < NL; ++l2) {                   \
          const int mid = (lo[l2] + hi[l2]) >> 1;                             \
          const bool open = lo[l2] < hi[l2];                                  \
          const bool gt = w_key[l2][min(mid, RNI_TILE - 1)] > key;            \
          lo[l2] = (open && gt) ? mid + 1 : lo[l2];                           \
          hi[l2] = (open && !gt) ? mid : hi[l2];                              \
        }                                                                     \
      }
      if (L <= 5) { RNI_SEARCH(5) } else { RNI_SEARCH(BRCNN_MAX_LEVELS) }
#undef RNI_SEARCH
      int rank = wi;
#pragma unroll
      for (int l2 = 0; l2 < BRCNN_MAX_LEVELS; ++l2) rank += lo[l2];
      if (rank < RNI_TILE) {
        const float4 raw = w_box[wl][wi];
        const float4 ob = add_seg_offset(raw, (float)wl * maxc1);
        traw[rank] = raw;
        tb[rank] = ob;
        ta[rank] = (ob.z - ob.x + off) * (ob.w - ob.y + off);
        tkey[rank] = key;
        tlvl[rank] = wl;
        atomicAdd(&s_taken[wl], 1);
        if (w_valid[wl][wi]) atomicAnd(&s_dead[rank >> 5], ~(1u << (rank & 31)));
      }
    }
    __syncthreads();
    RNI_MARK(1);
    // next round's windows start at cursor + taken: issue the global loads now
    // so that their latency hides behind pull / diag / resolve
    if (wl < L) prefetch(s_cursor[wl] + s_taken[wl]);
    const int nkept = s_nkept;
    // ---- 3. pull: candidate c vs this CTA's share of its level's kept boxes ----
    {
      const int c = tid & 63, g = tid >> 6;   // 8 groups
      bool hit = false;
      if (c < ntile && !((s_dead[c >> 5] >> (c & 31)) & 1u)) {
        const int l = tlvl[c];
        const float4 bx = tb[c];
        const float ba = ta[c];
        const unsigned short* li = lidx + (size_t)l * kp;
        const int n = s_lcnt[l];
        int q = g;
        for (; q + 24 < n && !hit; q += 32) {
          const float4 k0 = kbox[li[q]], k1 = kbox[li[q + 8]], k2 = kbox[li[q + 16]],
                       k3 = kbox[li[q + 24]];
          const bool h0 = nms_suppresses(k0, (k0.z - k0.x + off) * (k0.w - k0.y + off), bx, ba, thr, off);
          const bool h1 = nms_suppresses(k1, (k1.z - k1.x + off) * (k1.w - k1.y + off), bx, ba, thr, off);
          const bool h2 = nms_suppresses(k2, (k2.z - k2.x + off) * (k2.w - k2.y + off), bx, ba, thr, off);
          const bool h3 = nms_suppresses(k3, (k3.z - k3.x + off) * (k3.w - k3.y + off), bx, ba, thr, off);
          hit = h0 | h1 | h2 | h3;
        }
        for (; q < n && !hit; q += 8) {
          const float4 k0 = kbox[li[q]];
          hit = nms_suppresses(k0, (k0.z - k0.x + off) * (k0.w - k0.y + off), bx, ba, thr, off);
        }
      }
      const unsigned bal = __ballot_sync(0xffffffffu, hit);
      if (lane == 0 && bal) atomicOr(&s_hitw[(tid >> 5) & 1], bal);
    }
    // ---- 4. diag slice: rows [crank*RPC, crank*RPC + RPC) of the 64x64 same-level
    // suppression matrix, one pair test per thread (all rows when CS == 1).  It does
    // not depend on the pull result: bits of candidates that turn out dead are
    // simply ignored by the resolve step.
    {
      constexpr int RPC = RNI_TILE / CS;               // rows per CTA
      constexpr int PASSES = (RPC * RNI_TILE) / RNI_THREADS > 0 ? (RPC * RNI_TILE) / RNI_THREADS : 1;
#pragma unroll
      for (int ps = 0; ps < PASSES; ++ps) {
        const int rl = (tid >> 6) + ps * (RNI_THREADS / RNI_TILE);   // local row
        const int r = crank * RPC + rl, cc = tid & 63;
        bool sup = false;
        if (rl < RPC && r < ntile && cc < r && tlvl[cc] == tlvl[r] &&
            !((s_dead[cc >> 5] >> (cc & 31)) & 1u) && !((s_dead[r >> 5] >> (r & 31)) & 1u))
          sup = nms_suppresses(tb[cc], ta[cc], tb[r], ta[r], thr, off);
        const unsigned bal = __ballot_sync(0xffffffffu, sup);
        if (lane == 0 && rl < RPC) s_dslice[round & 1][rl][(tid >> 5) & 1] = bal;
      }
    }
    __syncthreads();
    RNI_MARK(2);
    // ---- combine pull hits and diag rows of the cluster (one cluster barrier) ----
    if (CS > 1) {
      if (tid == 0) s_hit[round & 1] = ((u64)s_hitw[1] << 32) | (u64)s_hitw[0];
      cg::this_cluster().sync();
      if (tid < 32) {
        u64 h = 0ull;
        if (lane < CS) h = *cg::this_cluster().map_shared_rank(&s_hit[round & 1], lane);
#pragma unroll
        for (int o = 1; o < CS; o <<= 1) h |= __shfl_xor_sync(0xffffffffu, h, o);
        if (lane == 0) s_alive = ~((((u64)s_dead[1] << 32) | (u64)s_dead[0]) | h);
      } else if (tid >= 64 && tid < 64 + RNI_TILE) {
        constexpr int RPC = RNI_TILE / CS;
        const int r = tid - 64;
        const unsigned* src = cg::this_cluster().map_shared_rank(
            &s_dslice[round & 1][r % RPC][0], r / RPC);
        diag[r] = ((u64)src[1] << 32) | (u64)src[0];
      }
    } else {
      if (tid == 0)
        s_alive = ~((((u64)s_dead[1] << 32) | (u64)s_dead[0]) |
                    (((u64)s_hitw[1] << 32) | (u64)s_hitw[0]));
      if (tid >= 64 && tid < 64 + RNI_TILE) {
        const int r = tid - 64;
        diag[r] = ((u64)s_dslice[round & 1][r][1] << 32) | (u64)s_dslice[round & 1][r][0];
      }
    }
    if (tid >= 128 && tid < 128 + RNI_TILE) {
      // per-level membership masks of the tile
      const int l = tlvl[tid - 128];
      for (int l2 = 0; l2 < L; ++l2) {
        const unsigned m = __ballot_sync(0xffffffffu, l == l2);
        if (lane == 0) s_lmask32[l2][(tid >> 5) & 1] = m;
      }
    }
    __syncthreads();
    RNI_MARK(3);
    const u64 alive = s_alive;
    if (alive != 0ull) {   // cluster-uniform
      // ---- 5. resolve (warp 0; identical in every CTA) ----
      if (tid < 32) {
        if (lane < L) s_lmask[lane] = ((u64)s_lmask32[lane][1] << 32) | (u64)s_lmask32[lane][0];
        __syncwarp();
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
        // kept box with global rank q lives in CTA q % CS
        u64 mine = 0ull;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int r = lane + 32 * h;
          const bool kept = (keep >> r) & 1ull;
          const int q = nkept + __popcll(keep & ((1ull << r) - 1ull));
          const bool own = kept && (q < max_out) && ((CS == 1) || (q % CS == crank));
          const unsigned ob = __ballot_sync(0xffffffffu, own);
          mine |= (u64)ob << (32 * h);
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int r = lane + 32 * h;
          if ((mine >> r) & 1ull) {
            const u64 below = mine & ((1ull << r) - 1ull);
            const int l = tlvl[r];
            const int slot = s_lkept + __popcll(below);      // local kept so far
            kbox[slot] = tb[r];
            lidx[(size_t)l * kp + s_lcnt[l] + __popcll(below & s_lmask[l])] =
                (unsigned short)slot;
          }
          if (crank == 0 && ((keep >> r) & 1ull)) {
            const int q = nkept + __popcll(keep & ((1ull << r) - 1ull));
            if (q < max_out) {
              if (proposals != nullptr) {
                const float4 raw = traw[r];
                float* o = proposals + ((size_t)b * max_out + q) * 5;
                o[0] = raw.x; o[1] = raw.y; o[2] = raw.z; o[3] = raw.w;
                o[4] = __uint_as_float((uint32_t)(tkey[r] >> 32));
              }
              if (keep_idx != nullptr)   // low key half = ~original index
                keep_idx[(size_t)b * max_out + q] =
                    (int64_t)(0xFFFFFFFFu - (uint32_t)(tkey[r] & 0xFFFFFFFFull));
            }
          }
        }
        __syncwarp();
        if (lane < L) s_lcnt[lane] += __popcll(mine & s_lmask[lane]);
        if (lane == 0) { s_nkept = nkept + __popcll(keep); s_lkept += __popcll(mine); }
      }
    }
    if (tid < L) s_cursor[tid] += s_taken[tid];
    __syncthreads();
    RNI_MARK(4);
    if (s_nkept >= max_out) break;                  // cluster-uniform
  }
  __syncthreads();
  if (crank == 0) {
    const int nk = min(s_nkept, max_out);
    if (tid == 0) num_proposals[b] = nk;
    if (proposals != nullptr)
      for (int i = nk * 5 + tid; i < max_out * 5; i += RNI_THREADS)
        proposals[(size_t)b * max_out * 5 + i] = 0.f;
  }
  if (dbg != nullptr && tid == 0 && blockIdx.x == 0) {
    for (int i = 0; i < 5; ++i) dbg[i] = t_acc[i];
    dbg[5] = n_rounds;
    dbg[6] = clock64() - t_entry;     // kernel entry -> here (CTA 0)
    dbg[7] = t_acc[7];                // prologue
  }
  // nobody leaves while a peer may still read its hit masks
  if (CS > 1) cg::this_cluster().sync();
}

}  // namespace brcnn
