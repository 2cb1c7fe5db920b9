// K4 (v5): RoIAlign backward gather with warp-private asynchronous prefetch rings.
//
// Same arithmetic as roi_align_bwd3.cuh / roi_align_bwd4.cuh (deterministic gather over the
// per-tile RoI lists of roi_bwd_prep_kernel, summation in RoI-index order):
//     G[y][x][c] = sum_roi sum_ph Wy[y][ph] * ( sum_pw Wx[x][pw] * g[roi][ph][pw][c] ).
// What the v3 / v4 profiles showed: the gather is bound by per-RoI latency chains, not by bytes
// (walk = 3.2 k cycles per listed RoI with one producer warp feeding 8 consumer warps through
// one mbarrier ring: table fetch -> band -> staged copy -> fold, all serialised per RoI).  v5
// removes the shared producer and every barrier from the walk:
//   * one CTA per (8x8-pixel tile, 128-channel slab), 8 warps = the 8 tile columns, lane =
//     channel quad, 8 row accumulators in registers (as before);
//   * per chunk of <= 16 listed RoIs ALL threads fetch the tile's rows / columns of the Wy / Wx
//     tables in one parallel round (thread = (RoI, table row)) into shared memory, Wy
//     transposed to [ph][tile row] so a fold reads two broadcast 128-bit words;
//   * each warp then builds ITS OWN step list (lane = RoI: column band [qa, qb] of the warp's
//     pixel column, pooled-row band of the tile; one step = one pooled row x <= 2 pooled
//     columns) with a warp scan, and walks it with a private ring of cp.async (LDGSTS) slots:
//     a lane copies exactly the 16 bytes it will read back, so `cp.async.wait_group` is the
//     only synchronisation — no mbarrier, no __syncwarp, no producer warp.  Three steps are in
//     flight per warp (24 warps per SM -> up to 150 KB of gradient loads outstanding per SM);
//   * a warp whose column lies outside a RoI's footprint never touches that RoI.
// Tiles whose list overflowed the prep kernel's per-tile capacity scan the (image, level)
// bucket in windows of 128 RoI indices instead (same order, same sums).
#pragma once
#include "common.cuh"
#include "roi_align.cuh"
#include "roi_align_bwd2.cuh"
#include "roi_align_bwd3.cuh"

namespace brcnn {

constexpr int B5_TS = B3_TS;        // 8x8-pixel tiles (the prep kernel's tile lists)
constexpr int B5_CS = B3_CS;        // 128 channels per CTA: one quad per lane
constexpr int B5_THREADS = 256;     // 8 warps = tile columns
constexpr int B5_WIN = 128;         // list window (== tile list capacity)
constexpr int B5_CHUNK = 12;        // RoIs per table round (<= 16: thread = (RoI, row) in one pass)
constexpr int B5_SLOT_BINS = 2;     // pooled columns per step (a pixel column sees ~2 bins)
constexpr int B5_NSLOT = 4;         // ring slots per warp (3 steps in flight)
constexpr int B5_SLOT_BYTES = B5_SLOT_BINS * B5_CS * 4;
static_assert(B5_SLOT_BINS == 2, "the walk handles exactly two bins per step");
constexpr int B5_MAX_STEPS = B5_CHUNK * 7 * ((7 + B5_SLOT_BINS - 1) / B5_SLOT_BINS);   // per warp and chunk
constexpr int B5_RING_BYTES = 8 * B5_NSLOT * B5_SLOT_BYTES;

__device__ __forceinline__ void cp_async_commit() {
  asm volatile("cp.async.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void cp_async16_addr(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16_addr_if(uint32_t dst, const void* src, unsigned on) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t"
               "@p cp.async.ca.shared.global [%0], [%1], 16;\n\t}"
               ::"r"(dst), "l"(src), "r"(on) : "memory");
}
// (t0, t1) += w * v when `on` (predicated packed FMAs: v may be stale shared memory otherwise)
__device__ __forceinline__ void ffma2_if(float2& t0, float2& t1, const float2 w, const float4 v,
                                         unsigned on) {
  const float2 va = make_float2(v.x, v.y), vb = make_float2(v.z, v.w);
  asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %5, 0;\n\t"
      "@p fma.rn.f32x2 %0, %2, %3, %0;\n\t@p fma.rn.f32x2 %1, %2, %4, %1;\n\t}"
      : "+l"(reinterpret_cast<unsigned long long&>(t0)),
        "+l"(reinterpret_cast<unsigned long long&>(t1))
      : "l"(reinterpret_cast<const unsigned long long&>(w)),
        "l"(reinterpret_cast<const unsigned long long&>(va)),
        "l"(reinterpret_cast<const unsigned long long&>(vb)), "r"(on));
}
__device__ __forceinline__ uint2 lds64u(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ float lds32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}

// grid (total_tiles, ceil(C / B5_CS)), coarse levels first; dynamic smem: B5_RING_BYTES
__global__ void __launch_bounds__(B5_THREADS, 3)
roi_bwd_gather5_kernel(const __grid_constant__ RoiBwd3Args ba,
                       const int32_t* __restrict__ tile_r,        // [tiles][B5_WIN]
                       const RoiBwdRec* __restrict__ tile_rec,    // [tiles][B5_WIN]
                       const int32_t* __restrict__ tile_cnt,      // [tiles]
                       const RoiBwdRec* __restrict__ bucket_rec, const int32_t* __restrict__ bucket,
                       const int32_t* __restrict__ bucket_cnt, int R,
                       const float* __restrict__ tab,
                       const float* __restrict__ gt /* (R, nbins, C) */) {
  extern __shared__ __align__(128) unsigned char b5_ring[];
  __shared__ int s_id[2][B5_WIN];                   // unsorted / sorted RoI ids
  __shared__ int4 s_box[2][B5_WIN];                 // their footprint boxes
  __shared__ __align__(16) float s_ytT[B5_CHUNK][8][8];   // Wy [roi][ph][tile row]
  __shared__ __align__(16) int s_yband[B5_CHUNK][8];      // packed band of each tile row
  __shared__ __align__(16) float s_xt[B5_CHUNK][8][8];    // Wx [roi][tile column][pw | band]
  __shared__ __align__(8) uint2 s_steps[8][B5_MAX_STEPS + B5_NSLOT];
  __shared__ int s_warp[8];
  __shared__ int s_n;

  const RoiArgs& a = ba.a;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int C = a.C, nbins = a.PH * a.PW, PW = a.PW;
  const int tile = blockIdx.x;
  const int cnt = __ldg(tile_cnt + tile);
  int lvl = a.L - 1;
  while (lvl > 0 && tile >= ba.tile_first[lvl - 1]) --lvl;
  int t = tile - ba.tile_first[lvl];
  const int tpi = ba.tiles_x[lvl] * ba.tiles_y[lvl];
  const int b = t / tpi; t -= b * tpi;
  const int ty = t / ba.tiles_x[lvl], tx = t - ty * ba.tiles_x[lvl];
  const int y0 = ty * B5_TS, x0 = tx * B5_TS;
  const int H = a.H[lvl], W = a.W[lvl];
  const int c0 = blockIdx.y * B5_CS;
  const bool lane_ok = c0 + lane * 4 < C;
  const bool col_ok = x0 + wid < W;

  float2 acc[B5_TS][2];
#pragma unroll
  for (int r = 0; r < B5_TS; ++r) acc[r][0] = acc[r][1] = make_float2(0.f, 0.f);

  if (cnt > 0) {
    const bool overflow = cnt > B5_WIN;
    const int key = b * a.L + lvl;
    const int nb = overflow ? __ldg(bucket_cnt + key) : 0;
    const uint32_t ring = smem_u32(b5_ring) + (uint32_t)(wid * B5_NSLOT * B5_SLOT_BYTES + lane * 16);
    for (int w0 = 0; w0 < (overflow ? R : 1); w0 += B5_WIN) {
      // ---- the window's (unsorted) list -> s_id[0] / s_box[0] ----
      int n;
      if (!overflow) {
        n = cnt;
        if (tid < n) {
          s_id[0][tid] = __ldg(tile_r + (size_t)tile * B5_WIN + tid);
          s_box[0][tid] = __ldg(reinterpret_cast<const int4*>(tile_rec) + (size_t)tile * B5_WIN + tid);
        }
        __syncthreads();
      } else {
        if (tid == 0) s_n = 0;
        __syncthreads();
        const int32_t* bk = bucket + (size_t)key * ba.bucket_cap;
        const int4* bkr = reinterpret_cast<const int4*>(bucket_rec) + (size_t)key * ba.bucket_cap;
        const int w1 = min(R, w0 + B5_WIN);
        for (int i0 = 0; i0 < nb; i0 += B5_THREADS) {
          const int i = i0 + tid;
          bool f = false;
          int r = -1;
          int4 q = make_int4(1, 0, 1, 0);
          if (i < nb) {
            r = bk[i];
            q = bkr[i];
            f = (r >= w0 && r < w1) && (q.x <= y0 + B5_TS - 1) && (q.y >= y0) &&
                (q.z <= x0 + B5_TS - 1) && (q.w >= x0);
          }
          const unsigned bm = __ballot_sync(0xffffffffu, f);
          if (lane == 0) s_warp[wid] = __popc(bm);
          __syncthreads();
          int base = s_n, tot = 0;
#pragma unroll
          for (int w = 0; w < 8; ++w) {
            const int c = s_warp[w];
            if (w < wid) base += c;
            tot += c;
          }
          if (f) {
            const int slot = base + __popc(bm & ((1u << lane) - 1u));
            s_id[0][slot] = r;
            s_box[0][slot] = q;
          }
          __syncthreads();
          if (tid == 0) s_n += tot;
          __syncthreads();
        }
        n = s_n;
      }
      // ---- sort by RoI index (rank by counting): fixed summation order ----
      if (tid < n) {
        const int v = s_id[0][tid];
        int rk = 0;
        for (int k = 0; k < n; ++k) rk += (s_id[0][k] < v);
        s_id[1][rk] = v;
        s_box[1][rk] = s_box[0][tid];
      }
      __syncthreads();
      for (int ch0 = 0; ch0 < n; ch0 += B5_CHUNK) {
        const int m = min(B5_CHUNK, n - ch0);
        // tables / step lists of the previous chunk are still being read by slower warps
        if (ch0 > 0 || w0 > 0) __syncthreads();
        // ---- table round: thread = (RoI, tile row | tile column) ----
        {
          const int li = tid >> 4, row = tid & 15;
          const bool isy = row < 8;
          const int j = row & 7;
          if (li < m) {
            const int r = s_id[1][ch0 + li];
            const int4 box = s_box[1][ch0 + li];
            const int pos = (isy ? y0 : x0) + j;
            const bool in = isy ? (pos >= box.x && pos <= box.y) : (pos >= box.z && pos <= box.w);
            float4 wa = make_float4(0.f, 0.f, 0.f, 0.f), wb = wa;
            wb.w = __int_as_float(1);                      // empty band (pa = 1 > pb = 0)
            if (in) {
              const float4* src = reinterpret_cast<const float4*>(
                  tab + ((size_t)r * ba.TR + (isy ? (pos - box.x) : (a.max_h + pos - box.z))) * 8);
              wa = __ldg(src);
              wb = __ldg(src + 1);
            }
            if (isy) {
              s_ytT[li][0][j] = wa.x; s_ytT[li][1][j] = wa.y; s_ytT[li][2][j] = wa.z;
              s_ytT[li][3][j] = wa.w; s_ytT[li][4][j] = wb.x; s_ytT[li][5][j] = wb.y;
              s_ytT[li][6][j] = wb.z;
              s_yband[li][j] = __float_as_int(wb.w);
            } else {
              float4* d = reinterpret_cast<float4*>(&s_xt[li][j][0]);
              d[0] = wa; d[1] = wb;
            }
          }
        }
        __syncthreads();
        if (col_ok) {
          // ---- this warp's step list ----
          // per-RoI bands (lane = RoI): pooled columns [qa, qb] of the warp's pixel column and
          // pooled rows [pa, pb] of the tile, packed qa | qb << 4 | pa << 8 | pb << 12
          int bands = 1;                                   // qa = 1 > qb = 0: nothing
          if (lane < m) {
            const int qk = __float_as_int(s_xt[lane][wid][7]);
            const int qa = qk & 0xff, qb = qk >> 8;
            const int4 ya = *reinterpret_cast<const int4*>(&s_yband[lane][0]);
            const int4 yb = *reinterpret_cast<const int4*>(&s_yband[lane][4]);
            const int yy[8] = {ya.x, ya.y, ya.z, ya.w, yb.x, yb.y, yb.z, yb.w};
            int pa = 8, pb = 0;
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
              const int lo = yy[jj] & 0xff, hi = yy[jj] >> 8;
              if (lo <= hi) { pa = min(pa, lo); pb = max(pb, hi); }
            }
            if (qa <= qb && pa <= pb) bands = qa | (qb << 4) | (pa << 8) | (pb << 12);
          }
          // steps (lane = (RoI, pooled row) pair): record = {gradient offset in float4 units,
          // Wy offset | Wx offset << 10 | bins << 20}
          int total = 0;
          for (int p0 = 0; p0 < m * 7; p0 += 32) {
            const int pr = p0 + lane;
            const int li = pr / 7, ph = pr - li * 7;
            const int bd = __shfl_sync(0xffffffffu, bands, li & 31);
            const int qa = bd & 15, qb = (bd >> 4) & 15, pa = (bd >> 8) & 15, pb = bd >> 12;
            const bool on = (li < m) && (qa <= qb) && (ph >= pa) && (ph <= pb);
            const int nst = on ? (qb - qa + B5_SLOT_BINS) / B5_SLOT_BINS : 0;
            int off = nst;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
              const int v = __shfl_up_sync(0xffffffffu, off, o);
              if (lane >= o) off += v;
            }
            const int tot = __shfl_sync(0xffffffffu, off, 31);
            if (on) {
              uint2* sp = &s_steps[wid][total + off - nst];
              const unsigned g0 = ((unsigned)s_id[1][ch0 + li] * nbins + ph * PW) * (unsigned)(C >> 2);
              for (int q0 = qa; q0 <= qb; q0 += B5_SLOT_BINS)
                *sp++ = make_uint2(g0 + q0 * (C >> 2),
                                   (li * 64 + ph * 8) | ((li * 64 + q0) << 10) |
                                       (min(B5_SLOT_BINS, qb - q0 + 1) << 20));
            }
            total += tot;
          }
          const float4* g4 = reinterpret_cast<const float4*>(gt) + (c0 >> 2) + (lane_ok ? lane : 0);
          const uint32_t yt0 = smem_u32(&s_ytT[0][0][0]);
          const uint32_t xt0 = smem_u32(&s_xt[0][wid][0]);
          const uint32_t cq = (uint32_t)(C >> 2);
          // B5_NSLOT - 1 trailing copies of the last step: the prefetch below never needs a
          // bounds test (the copies land in slots nobody reads any more)
          __syncwarp();                      // the last step was written by another lane
          if (total > 0 && lane < B5_NSLOT - 1) s_steps[wid][total + lane] = s_steps[wid][total - 1];
          __syncwarp();
          // ---- walk: private cp.async ring, B5_NSLOT - 1 steps in flight ----
          if (total > 0) {
            const uint32_t st0 = smem_u32(&s_steps[wid][0]);
            auto issue = [&](int si) {
              const uint2 rec = lds64u(st0 + 8u * si);
              const float4* src = g4 + rec.x;
              const uint32_t dst = ring + (uint32_t)((si & (B5_NSLOT - 1)) * B5_SLOT_BYTES);
              cp_async16_addr(dst, src);
              cp_async16_addr_if(dst + B5_CS * 4, src + cq, rec.y >> 21);     // second bin
              cp_async_commit();
            };
#pragma unroll
            for (int si = 0; si < B5_NSLOT - 1; ++si) issue(si);
#pragma unroll 1
            for (int si = 0; si < total; ++si) {
              issue(si + B5_NSLOT - 1);
              const unsigned rec = lds64u(st0 + 8u * si).y;
              const uint32_t ya = yt0 + (rec & 0x3ffu) * 4u;
              const uint32_t xa = xt0 + ((rec >> 10) & 0x3ffu) * 4u;
              const float4 wa = lds128(ya);
              const float4 wb = lds128(ya + 16);
              const float wx0 = lds32(xa), wx1 = lds32(xa + 4);   // in-table: 8 floats per row
              cp_async_wait<B5_NSLOT - 1>();
              const uint32_t slot = ring + (uint32_t)((si & (B5_NSLOT - 1)) * B5_SLOT_BYTES);
              const float4 v0 = lds128(slot);
              const float4 v1 = lds128(slot + B5_CS * 4);          // stale unless the step has 2 bins
              float2 t0 = make_float2(wx0 * v0.x, wx0 * v0.y), t1 = make_float2(wx0 * v0.z, wx0 * v0.w);
              ffma2_if(t0, t1, make_float2(wx1, wx1), v1, rec >> 21);
              const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
              for (int r = 0; r < B5_TS; ++r) {
                const float2 w2 = make_float2(wv[r], wv[r]);
                ffma2(acc[r][0], w2, t0);
                ffma2(acc[r][1], w2, t1);
              }
            }
            cp_async_wait<0>();
          }
        }
      }
    }
  }
  // ---- every element of the tile is written exactly once ----
  const int x = x0 + wid;
  if (x < W && lane_ok) {
    float* gout = ba.grad[lvl] + (((size_t)b * H + y0) * W + x) * C + c0 + lane * 4;
    const size_t pitch = (size_t)W * C;
#pragma unroll
    for (int r = 0; r < B5_TS; ++r) {
      if (y0 + r < H)
        __stcs(reinterpret_cast<float4*>(gout + r * pitch),
               make_float4(acc[r][0].x, acc[r][0].y, acc[r][1].x, acc[r][1].y));
    }
  }
}

}  // namespace brcnn
