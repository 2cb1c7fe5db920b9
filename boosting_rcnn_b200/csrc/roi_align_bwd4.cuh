// K4 (v4): persistent RoIAlign backward gather.  Same arithmetic, stage format and consumer
// code as roi_align_bwd3.cuh (deterministic gather, gradient blocks staged through shared
// memory with cp.async + mbarriers, FFMA2 folds); what changes is the schedule, because the v3
// profile showed a busy tile CTA spending a third of its life in per-CTA fixed latencies
// (bucket scan 5 k cycles, first-stage latency 6 k, of 32 k) and half of its walk waiting for
// data with only 5 RoIs per list:
//   * the prep kernel appends every RoI (id + footprint box) to the list of each TILE its box
//     touches (capacity B4_CAP entries per tile; longer lists fall back to the (image, level)
//     bucket scan), so a list is one count + one parallel read away;
//   * CTAs are persistent (3 per SM) and pull (tile, channel-slab) work items from an atomic
//     counter; the producer warp builds the next item's list (warp-level sort by RoI index ->
//     fixed summation order) and streams its stages while the consumer warps are still folding
//     the previous item: list building, table fetches and the first-stage latency of item k+1
//     hide behind the arithmetic of item k.  An "item end" stage tells the consumers where to
//     store their accumulators (empty tiles are just an item-end stage: zero fill).
#pragma once
#include "common.cuh"
#include "roi_align.cuh"
#include "roi_align_bwd2.cuh"
#include "roi_align_bwd3.cuh"
#include "roi_align_tma.cuh"

namespace brcnn {

constexpr int B4_CAP = 128;     // entries per tile list

// grid: persistent CTAs; dynamic smem: B3_NS * B3_STAGE
__global__ void __launch_bounds__(B3_THREADS, 3)
roi_bwd_gather4_kernel(const __grid_constant__ RoiBwd3Args ba,
                       const int32_t* __restrict__ tile_r,        // [tiles][B4_CAP]
                       const RoiBwdRec* __restrict__ tile_rec,    // [tiles][B4_CAP]
                       const int32_t* __restrict__ tile_cnt,      // [tiles]
                       const RoiBwdRec* __restrict__ bucket_rec, const int32_t* __restrict__ bucket,
                       const int32_t* __restrict__ bucket_cnt, int32_t* __restrict__ work_counter,
                       int total_tiles, int nslab, int R, const float* __restrict__ tab,
                       const float* __restrict__ gt /* (R, nbins, C) */) {
  extern __shared__ __align__(128) unsigned char b3_smem[];
  __shared__ int s_list[2][B4_CAP];          // producer-private: unsorted / sorted RoI ids
  __shared__ int4 s_rec[2][B4_CAP];          // their footprint boxes
  __shared__ __align__(8) uint64_t full_bar[B3_NS];
  __shared__ __align__(8) uint64_t empty_bar[B3_NS];

  const RoiArgs& a = ba.a;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int C = a.C, nbins = a.PH * a.PW, PW = a.PW;
  if (tid == 0) {
    for (int s = 0; s < B3_NS; ++s) {
      mbar_init(&full_bar[s], 33);     // 32 cp.async lanes + the header writer
      mbar_init(&empty_bar[s], B3_CONS_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);
  const int total_items = total_tiles * nslab;

  if (wid == B3_CONS_WARPS) {
    // =============================== producer warp ===============================
    int p_stage = 0, p_round = 0;
    auto stage_acquire = [&]() -> unsigned char* {
      if (p_round > 0) mbar_wait_addr(empty0 + 8u * p_stage, (uint32_t)((p_round - 1) & 1));
      return b3_smem + (size_t)p_stage * B3_STAGE;
    };
    auto stage_advance = [&]() { if (++p_stage == B3_NS) { p_stage = 0; ++p_round; } };
    while (true) {
      int item = 0;
      if (lane == 0) item = atomicAdd(work_counter, 1);
      item = __shfl_sync(0xffffffffu, item, 0);
      if (item >= total_items) {
        unsigned char* st = stage_acquire();
        int* meta = reinterpret_cast<int*>(st) + B3_OFF_META;
        if (lane == 0) { meta[0] = 2; mbar_arrive(&full_bar[p_stage]); }
        cp_async_arrive_noinc(&full_bar[p_stage]);
        stage_advance();
        break;
      }
      const int tile = item / nslab, slab = item - tile * nslab;
      int lvl = a.L - 1;
      while (lvl > 0 && tile >= ba.tile_first[lvl - 1]) --lvl;
      int t = tile - ba.tile_first[lvl];
      const int tpi = ba.tiles_x[lvl] * ba.tiles_y[lvl];
      const int b = t / tpi; t -= b * tpi;
      const int ty = t / ba.tiles_x[lvl], tx = t - ty * ba.tiles_x[lvl];
      const int y0 = ty * B3_TS, x0 = tx * B3_TS;
      const int c0 = slab * B3_CS;
      const int cs = min(B3_CS, C - c0);
      const int cnt = __ldg(tile_cnt + tile);
      const bool isy = lane < 8;
      const int j = lane & 7;
      const int pos = (isy ? y0 : x0) + j;

      // windows over the RoI index range: one window unless the tile list overflowed
      const bool overflow = cnt > B4_CAP;
      const int key = b * a.L + lvl;
      const int nb = overflow ? __ldg(bucket_cnt + key) : 0;
      for (int w0 = 0; w0 < (overflow ? R : 1); w0 += B4_CAP) {
        // ---- gather the (unsorted) list into s_list[0] / s_rec[0] ----
        int n = 0;
        if (!overflow) {
          n = cnt;
          for (int i = lane; i < n; i += 32) {
            s_list[0][i] = __ldg(tile_r + (size_t)tile * B4_CAP + i);
            s_rec[0][i] = __ldg(reinterpret_cast<const int4*>(tile_rec) + (size_t)tile * B4_CAP + i);
          }
        } else {
          const int32_t* bk = bucket + (size_t)key * ba.bucket_cap;
          const int4* bkr = reinterpret_cast<const int4*>(bucket_rec) + (size_t)key * ba.bucket_cap;
          const int w1 = min(R, w0 + B4_CAP);
          for (int i0 = 0; i0 < nb; i0 += 32) {
            const int i = i0 + lane;
            bool f = false;
            int r = -1;
            int4 q = make_int4(1, 0, 1, 0);
            if (i < nb) {
              r = bk[i];
              q = bkr[i];
              f = (r >= w0 && r < w1) && (q.x <= y0 + B3_TS - 1) && (q.y >= y0) &&
                  (q.z <= x0 + B3_TS - 1) && (q.w >= x0);
            }
            const unsigned bm = __ballot_sync(0xffffffffu, f);
            if (f) {
              const int slot = n + __popc(bm & ((1u << lane) - 1u));
              s_list[0][slot] = r;
              s_rec[0][slot] = q;
            }
            n += __popc(bm);
          }
        }
        __syncwarp();
        // ---- sort by RoI index (rank by counting) into s_list[1] / s_rec[1] ----
        for (int i = lane; i < n; i += 32) {
          const int v = s_list[0][i];
          int rk = 0;
          for (int k = 0; k < n; ++k) rk += (s_list[0][k] < v);
          s_list[1][rk] = v;
          s_rec[1][rk] = s_rec[0][i];
        }
        __syncwarp();
        // ---- stream the list: table rows fetched one entry ahead ----
        auto fetch = [&](int li, float4& wa, float4& wb, bool& in) {
          wa = make_float4(0.f, 0.f, 0.f, 0.f);
          wb = wa;
          in = false;
          if (li < n && lane < 16) {
            const int4 rec = s_rec[1][li];
            in = isy ? (pos >= rec.x && pos <= rec.y) : (pos >= rec.z && pos <= rec.w);
            if (in) {
              const float4* src = reinterpret_cast<const float4*>(
                  tab + ((size_t)s_list[1][li] * ba.TR +
                         (isy ? (pos - rec.x) : (a.max_h + pos - rec.z))) * 8);
              wa = __ldg(src);
              wb = __ldg(src + 1);
            }
          }
        };
        float4 nwa, nwb;
        bool nin;
        fetch(0, nwa, nwb, nin);
        for (int li = 0; li < n; ++li) {
          float4 wa = nwa, wb = nwb;
          const bool in = nin;
          fetch(li + 1, nwa, nwb, nin);
          const int r = s_list[1][li];
          int lo = 8, hi = -1;
          if (lane < 16 && in) {
            const int pk = __float_as_int(wb.w);
            lo = pk & 0xff; hi = pk >> 8;
            if (lo > hi) { lo = 8; hi = -1; }
          }
          if (!isy) wb.w = __int_as_float(lo | (hi << 8));
          int ylo_b = lane < 8 ? lo : 8, yhi_b = lane < 8 ? hi : -1;
          int xlo_b = (lane >= 8 && lane < 16) ? lo : 8, xhi_b = (lane >= 8 && lane < 16) ? hi : -1;
#pragma unroll
          for (int o = 8; o > 0; o >>= 1) {
            ylo_b = min(ylo_b, __shfl_xor_sync(0xffffffffu, ylo_b, o));
            yhi_b = max(yhi_b, __shfl_xor_sync(0xffffffffu, yhi_b, o));
            xlo_b = min(xlo_b, __shfl_xor_sync(0xffffffffu, xlo_b, o));
            xhi_b = max(xhi_b, __shfl_xor_sync(0xffffffffu, xhi_b, o));
          }
          ylo_b = __shfl_sync(0xffffffffu, ylo_b, 0); yhi_b = __shfl_sync(0xffffffffu, yhi_b, 0);
          xlo_b = __shfl_sync(0xffffffffu, xlo_b, 8); xhi_b = __shfl_sync(0xffffffffu, xhi_b, 8);
          const int nph = yhi_b - ylo_b + 1, npw = xhi_b - xlo_b + 1;
          if (nph <= 0 || npw <= 0) continue;          // nothing of this RoI lands on the tile
          const float* gr = gt + (size_t)r * nbins * C + c0;
          const int rows_per = B3_BINS / npw;           // pooled rows per stage (npw <= 7: >= 4)
          for (int p0 = ylo_b; p0 <= yhi_b; p0 += rows_per) {
            const int p1 = min(yhi_b, p0 + rows_per - 1);
            unsigned char* st = stage_acquire();
            float* hdr = reinterpret_cast<float*>(st);
            int* meta = reinterpret_cast<int*>(hdr + B3_OFF_META);
            if (lane < 16) {
              if (isy) {
                hdr[0 * 8 + j] = wa.x; hdr[1 * 8 + j] = wa.y; hdr[2 * 8 + j] = wa.z;
                hdr[3 * 8 + j] = wa.w; hdr[4 * 8 + j] = wb.x; hdr[5 * 8 + j] = wb.y;
                hdr[6 * 8 + j] = wb.z;
              } else {
                float4* d = reinterpret_cast<float4*>(hdr + B3_OFF_WX + j * 8);
                d[0] = wa; d[1] = wb;
              }
            }
            if (lane == 0) {
              meta[0] = 0; meta[1] = p0; meta[2] = p1; meta[3] = xlo_b; meta[4] = npw;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&full_bar[p_stage]);     // header visible (release)
            if (lane * 4 < cs) {
              unsigned char* dstp = st + B3_HDR + lane * 16;
              for (int ph = p0; ph <= p1; ++ph) {
                const float* srcp = gr + (size_t)(ph * PW + xlo_b) * C + lane * 4;
                for (int q = 0; q < npw; ++q) {
                  cp_async16(dstp, srcp);
                  dstp += B3_BIN_BYTES;
                  srcp += C;
                }
              }
            }
            cp_async_arrive_noinc(&full_bar[p_stage]);
            stage_advance();
          }
        }
        __syncwarp();
      }
      // ---- item end: the consumers store their accumulators for this (tile, slab) ----
      {
        unsigned char* st = stage_acquire();
        int* meta = reinterpret_cast<int*>(st) + B3_OFF_META;
        if (lane == 0) {
          meta[0] = 1; meta[5] = lvl; meta[6] = b; meta[7] = y0; meta[8] = x0; meta[9] = c0;
          mbar_arrive(&full_bar[p_stage]);
        }
        cp_async_arrive_noinc(&full_bar[p_stage]);
        stage_advance();
      }
    }
  } else {
    // =============================== consumer warps ==============================
    float2 acc[B3_TS][2];
#pragma unroll
    for (int r = 0; r < B3_TS; ++r) acc[r][0] = acc[r][1] = make_float2(0.f, 0.f);
    int c_stage = 0, c_round = 0;
    while (true) {
      const unsigned char* st = b3_smem + (size_t)c_stage * B3_STAGE;
      const float* hdr = reinterpret_cast<const float*>(st);
      mbar_wait_addr_hint(full0 + 8u * c_stage, (uint32_t)(c_round & 1), 2000u);
      const int4 m = *reinterpret_cast<const int4*>(hdr + B3_OFF_META);
      const int4 m2 = *reinterpret_cast<const int4*>(hdr + B3_OFF_META + 4);   // npw, lvl, b, y0
      const int2 m3 = *reinterpret_cast<const int2*>(hdr + B3_OFF_META + 8);   // x0, c0
      const int kind = m.x;
      if (kind == 0) {
        const int npw = m2.x;
        const float* wxr = hdr + B3_OFF_WX + wid * 8;
        const int qk = __float_as_int(wxr[7]);
        const int qa = qk & 0xff, qb = qk >> 8;         // this column's pw band
        if (qa <= qb) {
          const float4* g4 = reinterpret_cast<const float4*>(st + B3_HDR) + lane;
          for (int ph = m.y; ph <= m.z; ++ph) {
            float2 t0 = make_float2(0.f, 0.f), t1 = t0;
            const float4* gp = g4 + (size_t)((ph - m.y) * npw + (qa - m.w)) * (B3_CS / 4);
            for (int pw = qa; pw <= qb; ++pw) {
              const float w = wxr[pw];
              const float4 v = *gp;
              gp += B3_CS / 4;
              const float2 w2 = make_float2(w, w);
              ffma2(t0, w2, make_float2(v.x, v.y));
              ffma2(t1, w2, make_float2(v.z, v.w));
            }
            const float4 wa = *reinterpret_cast<const float4*>(hdr + ph * 8);
            const float4 wb = *reinterpret_cast<const float4*>(hdr + ph * 8 + 4);
            if (wa.x != 0.f || wa.y != 0.f || wa.z != 0.f || wa.w != 0.f) {   // CTA-uniform
              const float wv[4] = {wa.x, wa.y, wa.z, wa.w};
#pragma unroll
              for (int r = 0; r < 4; ++r) {
                const float2 w2 = make_float2(wv[r], wv[r]);
                ffma2(acc[r][0], w2, t0);
                ffma2(acc[r][1], w2, t1);
              }
            }
            if (wb.x != 0.f || wb.y != 0.f || wb.z != 0.f || wb.w != 0.f) {
              const float wv[4] = {wb.x, wb.y, wb.z, wb.w};
#pragma unroll
              for (int r = 0; r < 4; ++r) {
                const float2 w2 = make_float2(wv[r], wv[r]);
                ffma2(acc[4 + r][0], w2, t0);
                ffma2(acc[4 + r][1], w2, t1);
              }
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive_addr(empty0 + 8u * c_stage);
      if (++c_stage == B3_NS) { c_stage = 0; ++c_round; }
      if (kind == 2) break;
      if (kind == 1) {
        const int lvl = m2.y, b = m2.z, y0 = m2.w, x0 = m3.x, c0 = m3.y;
        const int H = a.H[lvl], W = a.W[lvl];
        const int x = x0 + wid;
        if (x < W && c0 + lane * 4 < C) {
#pragma unroll
          for (int r = 0; r < B3_TS; ++r) {
            const int y = y0 + r;
            if (y < H) {
              float* gout = ba.grad[lvl] + (((size_t)b * H + y) * W + x) * C + c0 + lane * 4;
              __stcs(reinterpret_cast<float4*>(gout),
                     make_float4(acc[r][0].x, acc[r][0].y, acc[r][1].x, acc[r][1].y));
            }
          }
        }
#pragma unroll
        for (int r = 0; r < B3_TS; ++r) acc[r][0] = acc[r][1] = make_float2(0.f, 0.f);
      }
    }
  }
}

}  // namespace brcnn
