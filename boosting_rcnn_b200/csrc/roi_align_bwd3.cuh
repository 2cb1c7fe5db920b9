// K4 (v3): multi-level RoIAlign backward as a deterministic gather whose gradient
// blocks are staged through shared memory by an asynchronous-copy producer warp
// (cp.async + mbarrier ring).
// Reference: autograd of mmcv RoIAlign (roi_align_backward, 4 atomicAdd per sample)
// reached from single_level_roi_extractor.py:79,103; levels without RoIs still receive a
// zero gradient like :105-114.
//
// The gradient one RoI sends to feature pixel (y,x) is separable,
//     G[y][x][c] = sum_ph Wy[y][ph] * ( sum_pw Wx[x][pw] * g[ph][pw][c] ),
// with the Wy (1/count folded in) / Wx tables of roi_bwd_prep_kernel (roi_align_bwd2.cuh).
// grad_out is consumed bin-major, (R, PH*PW, C): that is the layout the (R,7,7,C) feature
// hand-off produces (no transpose), and a (bin, 128-channel) piece is one contiguous 512 B run.
//
//   roi_bwd_prep_kernel   (roi_align_bwd2.cuh) one CTA per RoI: footprint box, tables and an
//                         append of the RoI to the bucket of its (image, level).
//   roi_bwd_gather3_kernel one CTA per 8x8-pixel tile x 128-channel slab (coarse levels first).
//     - all threads scan the tile's (image, level) bucket only, keep the RoIs whose footprint
//       touches the tile and sort that list by RoI index (fixed summation order ->
//       bit-reproducible although the bucket was filled with atomics);
//     - a producer warp walks the list: it transposes the tile's rows of Wy / copies its columns
//       of Wx into the stage header and copies the needed (ph, pw) bins (only the sub-rectangle
//       the tile can see, <= 28 bins per stage) with one warp-wide 512 B `cp.async` each,
//       completion on the stage's mbarrier (cp.async.mbarrier.arrive); table rows are fetched
//       one list entry ahead;
//     - 8 consumer warps = the 8 tile columns, lane = channel quad, 8 row accumulators in
//       registers.  Per RoI and ph a thread forms t = sum_pw Wx[x][pw] * g[ph][pw] once
//       (conflict-free 128-bit LDS, warp-uniform weights) and folds it into its 8 rows with
//       packed FFMA2 (fma.rn.f32x2).  DRAM/L2 latency is paid once per (RoI, tile) and hidden
//       behind the previous RoI's arithmetic; no block barrier inside the walk.
//   Every output element is written exactly once (empty tiles write zeros).
#pragma once
#include "common.cuh"
#include "roi_align.cuh"
#include "roi_align_bwd2.cuh"
#include "roi_align_tma.cuh"

namespace brcnn {

constexpr int B3_TS = 8;                         // tile side in pixels
constexpr int B3_CS = 128;                       // channels per CTA (one quad per lane)
constexpr int B3_CONS_WARPS = B3_TS;             // consumer warp = tile column
constexpr int B3_THREADS = (B3_CONS_WARPS + 1) * 32;
constexpr int B3_NS = 4;                         // stages
constexpr int B3_LIST = 512;                     // RoIs per round
constexpr int B3_HDR = 1024;                     // stage header bytes
constexpr int B3_BIN_BYTES = B3_CS * 4;
constexpr int B3_P = 7;                          // max pooled side
constexpr int B3_BINS = 28;                      // bins per stage (4 pooled rows x 7 columns);
                                                 // larger (RoI, tile) blocks span several stages
constexpr int B3_STAGE = B3_HDR + B3_BINS * B3_BIN_BYTES;
// header layout (floats): wy_t[8 ph][8 rows] | wx[8 cols][8] | meta[16 ints]
constexpr int B3_OFF_WX = 64;
constexpr int B3_OFF_META = 128;

struct RoiBwd3Args {
  RoiArgs a;
  float* grad[BRCNN_MAX_LEVELS];  // NHWC (B,H,W,C)
  int tiles_x[BRCNN_MAX_LEVELS], tiles_y[BRCNN_MAX_LEVELS];
  int tile_first[BRCNN_MAX_LEVELS];  // first CTA (x) of each level; coarse levels first
  int TR;                            // table rows per RoI = max_h + max_w
  int bucket_cap;                    // entries per (image, level) bucket = R
};

// 16-byte asynchronous copy global -> shared (LDGSTS, L2 only) and the matching
// "arrive on the mbarrier once all my earlier cp.async have landed" (pending count unchanged:
// the barrier is initialised with one arrival per issuing lane)
__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;"
               ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];"
               ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void ffma2(float2& d, const float2 a, const float2 b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;"
      : "+l"(reinterpret_cast<unsigned long long&>(d))
      : "l"(reinterpret_cast<const unsigned long long&>(a)),
        "l"(reinterpret_cast<const unsigned long long&>(b)));
}

// grid (total_tiles, ceil(C / B3_CS)); dynamic smem: B3_NS * B3_STAGE
__global__ void __launch_bounds__(B3_THREADS, 3)
roi_bwd_gather3_kernel(const __grid_constant__ RoiBwd3Args ba,
                       const RoiBwdRec* __restrict__ bucket_rec,
                       const int32_t* __restrict__ bucket,
                       const int32_t* __restrict__ bucket_cnt,
                       const int32_t* __restrict__ tile_cnt, int R,
                       const float* __restrict__ tab,
                       const float* __restrict__ gt /* (R, nbins, C) */
#ifdef BRCNN_DEBUG_TIMING
                       , unsigned long long* __restrict__ dbg
#endif
                       ) {
#ifdef BRCNN_DEBUG_TIMING
  long long dt0 = clock64(), dt1 = dt0, dt2 = dt0, dt3 = dt0, dwait = 0, dstages = 0;
  int dn = 0, dfirst = 1;
#endif
  extern __shared__ __align__(128) unsigned char b3_smem[];
  __shared__ int s_list[B3_LIST];
  __shared__ int4 s_rec[B3_LIST];     // footprint boxes of the listed RoIs (sorted order)
  __shared__ __align__(8) uint64_t full_bar[B3_NS];
  __shared__ __align__(8) uint64_t empty_bar[B3_NS];
  __shared__ int s_n;

  const RoiArgs& a = ba.a;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  // coarse levels own the first CTAs (their tiles have the longest RoI lists)
  int lvl = a.L - 1;
  while (lvl > 0 && (int)blockIdx.x >= ba.tile_first[lvl - 1]) --lvl;
  int t = blockIdx.x - ba.tile_first[lvl];
  const int tpi = ba.tiles_x[lvl] * ba.tiles_y[lvl];
  const int b = t / tpi; t -= b * tpi;
  const int ty = t / ba.tiles_x[lvl], tx = t - ty * ba.tiles_x[lvl];
  const int y0 = ty * B3_TS, x0 = tx * B3_TS;
  const int H = a.H[lvl], W = a.W[lvl], C = a.C;
  const int c0 = blockIdx.y * B3_CS;
  const int cs = min(B3_CS, C - c0);          // channels of this slab
  const int nbins = a.PH * a.PW, PW = a.PW;
  const int key = b * a.L + lvl;
  // tiles no RoI touches (most of the fine levels) skip the bucket scan: one L2 read, then
  // the zero fill
  const int nb = tile_cnt[blockIdx.x] > 0 ? bucket_cnt[key] : 0;
  const int32_t* bk = bucket + (size_t)key * ba.bucket_cap;
  const int4* bkr = reinterpret_cast<const int4*>(bucket_rec) + (size_t)key * ba.bucket_cap;

  float2 acc[B3_TS][2];
#pragma unroll
  for (int r = 0; r < B3_TS; ++r) acc[r][0] = acc[r][1] = make_float2(0.f, 0.f);

  if (nb > 0) {
    const bool windowed = nb > B3_LIST;
    bool bars_ready = false;
    int p_stage = 0, p_round = 0;     // producer ring position
    int c_stage = 0, c_round = 0;     // consumer ring position
    for (int w0 = 0; w0 < R; w0 += B3_LIST) {
      const int w1 = windowed ? min(R, w0 + B3_LIST) : R;
      // ---- list of the bucket's RoIs (index window [w0, w1)) touching this tile ----
      if (tid == 0) s_n = 0;
      __syncthreads();
      for (int i0 = 0; i0 < nb; i0 += B3_THREADS) {
        const int i = i0 + tid;
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
        int base = 0;
        if (lane == 0 && bm) base = atomicAdd(&s_n, __popc(bm));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (f) {
          const int slot = base + __popc(bm & ((1u << lane) - 1u));
          s_list[slot] = r;
          s_rec[slot] = q;
        }
      }
      __syncthreads();
      const int n = s_n;
#ifdef BRCNN_DEBUG_TIMING
      dn += n;
#endif
      if (n > 1) {   // sort by RoI index (rank by counting; lists are short), boxes follow
        int v[2], rk[2];
        int4 vq[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const int i = tid + k * B3_THREADS;
          v[k] = i < n ? s_list[i] : 0x7fffffff;
          vq[k] = i < n ? s_rec[i] : make_int4(1, 0, 1, 0);
          rk[k] = 0;
        }
        for (int j = 0; j < n; ++j) {
          const int o = s_list[j];
          rk[0] += (o < v[0]);
          rk[1] += (o < v[1]);
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 2; ++k)
          if (tid + k * B3_THREADS < n) { s_list[rk[k]] = v[k]; s_rec[rk[k]] = vq[k]; }
      }
      if (n > 0 && !bars_ready) {
        if (tid == 0) {
          for (int s = 0; s < B3_NS; ++s) {
            mbar_init(&full_bar[s], 33);     // 32 cp.async lanes + the header writer
            mbar_init(&empty_bar[s], B3_CONS_WARPS);
          }
          asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        bars_ready = true;
      }
      __syncthreads();
#ifdef BRCNN_DEBUG_TIMING
      dt1 = clock64();
#endif
      if (n > 0) {
        const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);
        if (wid == B3_CONS_WARPS) {
          // =========================== producer warp ===========================
          // table rows of list entry `li` for this lane (lanes 0-7: tile rows, 8-15: tile
          // columns); issued one entry ahead so that their L2 latency hides behind the
          // previous entry's header stores and bulk-copy issue
          const bool isy = lane < 8;
          const int j = lane & 7;
          const int pos = (isy ? y0 : x0) + j;
          auto fetch = [&](int li, float4& wa, float4& wb, bool& in) {
            wa = make_float4(0.f, 0.f, 0.f, 0.f);
            wb = wa;
            in = false;
            if (li < n && lane < 16) {
              const int4 rec = s_rec[li];
              in = isy ? (pos >= rec.x && pos <= rec.y) : (pos >= rec.z && pos <= rec.w);
              if (in) {
                const float4* src = reinterpret_cast<const float4*>(
                    tab + ((size_t)s_list[li] * ba.TR +
                           (isy ? (pos - rec.x) : (a.max_h + pos - rec.z))) * 8);
                wa = __ldg(src);
                wb = __ldg(src + 1);
              }
            }
          };
          float4 nwa, nwb;
          bool nin;
          fetch(0, nwa, nwb, nin);
          for (int li = 0; li <= n; ++li) {
            float4 wa = nwa, wb = nwb;
            const bool in = nin;
            fetch(li + 1, nwa, nwb, nin);
            if (li == n) {               // end-of-list sentinel
              if (p_round > 0)
                mbar_wait_addr(empty0 + 8u * p_stage, (uint32_t)((p_round - 1) & 1));
              int* meta = reinterpret_cast<int*>(b3_smem + (size_t)p_stage * B3_STAGE) + B3_OFF_META;
              if (lane == 0) { meta[0] = 1; mbar_arrive(&full_bar[p_stage]); }
              cp_async_arrive_noinc(&full_bar[p_stage]);
              if (++p_stage == B3_NS) { p_stage = 0; ++p_round; }
              break;
            }
            const int r = s_list[li];
            int lo = 8, hi = -1;
            if (lane < 16 && in) {
              const int pk = __float_as_int(wb.w);
              lo = pk & 0xff; hi = pk >> 8;
              if (lo > hi) { lo = 8; hi = -1; }
            }
            if (!isy) wb.w = __int_as_float(lo | (hi << 8));
            // union of the bands over the tile's rows (lanes 0-7) and columns (lanes 8-15)
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
              unsigned char* st = b3_smem + (size_t)p_stage * B3_STAGE;
              float* hdr = reinterpret_cast<float*>(st);
              int* meta = reinterpret_cast<int*>(hdr + B3_OFF_META);
              if (p_round > 0)
                mbar_wait_addr(empty0 + 8u * p_stage, (uint32_t)((p_round - 1) & 1));
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
              // one warp-wide 512 B cp.async per bin (a 512 B `cp.async.bulk` per bin keeps the
              // TMA unit busy ~46 cycles per request: measured TMA-issue bound)
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
              if (++p_stage == B3_NS) { p_stage = 0; ++p_round; }
            }
          }
        } else {
          // =========================== consumer warps ==========================
          const bool q_ok = lane * 4 < cs;
          while (true) {
            const unsigned char* st = b3_smem + (size_t)c_stage * B3_STAGE;
            const float* hdr = reinterpret_cast<const float*>(st);
#ifdef BRCNN_DEBUG_TIMING
            const long long dw0 = clock64();
#endif
            mbar_wait_addr(full0 + 8u * c_stage, (uint32_t)(c_round & 1));
#ifdef BRCNN_DEBUG_TIMING
            if (dfirst) { dt2 = clock64(); dfirst = 0; }
            else { dwait += clock64() - dw0; }
            ++dstages;
#endif
            const int4 m = *reinterpret_cast<const int4*>(hdr + B3_OFF_META);
            const int npw = *reinterpret_cast<const int*>(hdr + B3_OFF_META + 4);
            const bool end = m.x != 0;
            if (!end && q_ok) {
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
                  // Wy[ph][tile rows]: two broadcast 128-bit loads; a half whose four rows are
                  // all outside the band is skipped (CTA-uniform).  (A per-row band + switch
                  // variant issued fewer instructions but ran 17 % slower: the walk is bound by
                  // the dependent LDS -> FFMA2 chain, not by issue slots.)
                  const float4 wa = *reinterpret_cast<const float4*>(hdr + ph * 8);
                  const float4 wb = *reinterpret_cast<const float4*>(hdr + ph * 8 + 4);
                  if (wa.x != 0.f || wa.y != 0.f || wa.z != 0.f || wa.w != 0.f) {
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
            if (end) break;
          }
        }
      }
      if (!windowed) break;
      // both roles advanced their own ring position by n + 1 stages; the list may be rebuilt
      // once every warp has left the walk
      __syncthreads();
    }
  }
#ifdef BRCNN_DEBUG_TIMING
  dt3 = clock64();
  if (dfirst) dt2 = dt3;
#endif
  if (wid < B3_CONS_WARPS) {
    const int x = x0 + wid;
    if (x < W && lane * 4 < cs) {
#pragma unroll
      for (int r = 0; r < B3_TS; ++r) {
        const int y = y0 + r;
        if (y < H) {
          float* gout = ba.grad[lvl] + (((size_t)b * H + y) * W + x) * C + c0 + lane * 4;
          *reinterpret_cast<float4*>(gout) =
              make_float4(acc[r][0].x, acc[r][0].y, acc[r][1].x, acc[r][1].y);
        }
      }
    }
  }
#ifdef BRCNN_DEBUG_TIMING
  if (tid == 0 && dbg != nullptr) {
    const long long dt4 = clock64();
    const int cls = dn > 0 ? 1 : 0;            // 0: empty tile, 1: tile with RoIs
    atomicAdd(dbg + cls * 8 + 0, 1ull);
    atomicAdd(dbg + cls * 8 + 1, (unsigned long long)(dt1 - dt0));   // list build
    atomicAdd(dbg + cls * 8 + 2, (unsigned long long)(dt2 - dt1));   // first stage latency
    atomicAdd(dbg + cls * 8 + 3, (unsigned long long)(dt3 - dt2));   // walk
    atomicAdd(dbg + cls * 8 + 4, (unsigned long long)(dt4 - dt3));   // stores
    atomicAdd(dbg + cls * 8 + 5, (unsigned long long)dn);
    atomicAdd(dbg + cls * 8 + 6, (unsigned long long)dwait);     // warp 0: waits after the first
    atomicAdd(dbg + cls * 8 + 7, (unsigned long long)dstages);
  }
#endif
}

struct RoiBwd3Ws {
  size_t recs, keys, tab, gt, bucket_cnt, tile_cnt, work_counter, zero_bytes, bucket, bucket_rec,
      tile_r, tile_rec, total;
};
constexpr int B4_TILE_CAP = 128;   // == B4_CAP (roi_align_bwd4.cuh)
static inline long long roi_bwd3_tiles(const RoiArgs& a) {
  long long t = 0;
  for (int l = 0; l < a.L; ++l)
    t += (long long)((a.W[l] + B3_TS - 1) / B3_TS) * ((a.H[l] + B3_TS - 1) / B3_TS) * a.B;
  return t;
}
static inline RoiBwd3Ws roi_bwd3_ws(const RoiArgs& a, int R, bool need_transpose) {
  RoiBwd3Ws w;
  size_t o = 0;
  const size_t Rn = (size_t)(R > 0 ? R : 1);
  int max_h = 0, max_w = 0;
  for (int l = 0; l < a.L; ++l) { max_h = a.H[l] > max_h ? a.H[l] : max_h; max_w = a.W[l] > max_w ? a.W[l] : max_w; }
  auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
  w.recs = o; o = al(o + Rn * sizeof(RoiBwdRec));
  w.keys = o; o = al(o + Rn * 2);
  w.tab = o;  o = al(o + Rn * (size_t)(max_h + max_w) * 8 * 4);
  w.gt = o;   o = al(o + (need_transpose ? Rn * (size_t)a.PH * a.PW * a.C * 4 : 0));
  w.bucket_cnt = o; o = al(o + (size_t)a.B * a.L * 4);
  w.tile_cnt = o;   o = al(o + (size_t)roi_bwd3_tiles(a) * 4);
  w.work_counter = o; o = al(o + 4);
  w.zero_bytes = o - w.bucket_cnt;                 // bucket_cnt + tile_cnt + counter: one memset
  w.bucket = o;     o = al(o + (size_t)a.B * a.L * Rn * 4);
  w.bucket_rec = o; o = al(o + (size_t)a.B * a.L * Rn * sizeof(RoiBwdRec));
  w.tile_r = o;     o = al(o + (size_t)roi_bwd3_tiles(a) * B4_TILE_CAP * 4);
  w.tile_rec = o;   o = al(o + (size_t)roi_bwd3_tiles(a) * B4_TILE_CAP * sizeof(RoiBwdRec));
  w.total = o;
  return w;
}

}  // namespace brcnn
