// K3 (v3): persistent, cross-RoI pipelined RoIAlign forward with the (R, PH, PW, C)
// feature hand-off.
//
// Reference behaviour: single_level_roi_extractor.py:36-115 + mmcv RoIAlign (aligned=True,
// pool_mode='avg', sampling_ratio=0), SURVEY.md App. A5/A6.  Same separable formulation as
// roi_align_tma.cuh,
//     out[ph][pw][c] = sum_y Wy[ph][y] * ( sum_x Wx[pw][x] * F[y][x][c] ),
// evaluated x-first one footprint row at a time; what changes is the schedule and the
// instruction count (the v2 kernel was issue-bound: 61 M warp instructions per 4096 RoIs):
//   * one CTA per SM slot (2 per SM) loops over RoIs pulled from an atomic counter (static
//     stride without a scheduling scratch).  The ring of
//     fixed-size row slots and its mbarriers live across RoIs: the producer warp streams RoI
//     i+1's footprint rows (1-D `cp.async.bulk`) while the consumer warps are still folding /
//     storing RoI i — no per-RoI barrier init, table build or output staging bubble;
//   * the producer warp does ALL per-RoI scalar work once (geometry, level map, footprint box,
//     separable weight tables, per-row / per-column non-zero bands) and publishes it through a
//     double-buffered descriptor + table area (tab_full / tab_empty mbarriers); consumers read
//     a 48-byte descriptor instead of redoing the geometry in every thread;
//   * arithmetic is packed: fma.rn.f32x2 (FFMA2) halves the FMA issue slots of both passes, and
//     the y-fold jumps straight to the row's band [pa, pb] of pooled rows (switch on pa) instead
//     of seven predicated 8-FMA groups;
//   * measured dead ends (B200, configs[1], 4096 RoIs): carving rows from a byte ring instead of
//     fixed 16-pixel slots (1.8x more rows in flight) ran 0.158 ms vs 0.119 ms — the extra
//     bookkeeping in the single producer warp costs more than the deeper ring gains; slot sizes
//     of 8 / 10 / 12 / 20 pixels: 0.127 / 0.126 / 0.122 / 0.119 ms;
//   * the output is written (R, PH, PW, C): a consumer lane holds a channel quad for one pooled
//     column, so each pooled row leaves as a fully coalesced 512 B streaming store per warp —
//     no shared-memory staging, no bank conflicts.  ConvFCBBoxHead's first FC consumes that
//     order directly (its weight columns are permuted once, convfc_bbox_head.py:164).
#pragma once
#include "common.cuh"
#include "roi_align.cuh"
#include "roi_align_tma.cuh"

namespace brcnn {

constexpr int R3_MAX_STAGES = 8;
constexpr int R3_DESC = 32;   // descriptor ints at the head of a table buffer
constexpr int R3_TABS = 2;    // descriptor + table buffers (publisher runs R3_TABS - 1 ahead)
constexpr int R3_THREADS = (RT_CONS_WARPS + 2) * 32;

__device__ __forceinline__ void mbar_expect_tx_addr(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void tma_load_1d_addr(uint32_t dst, const void* src, uint32_t bytes,
                                                 uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

struct Roi3Smem {
  int slot_bytes, ns, tab_floats, total;
};

__device__ __forceinline__ void r3_ffma2(float2& d, const float2 a, const float2 b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;"
      : "+l"(reinterpret_cast<unsigned long long&>(d))
      : "l"(reinterpret_cast<const unsigned long long&>(a)),
        "l"(reinterpret_cast<const unsigned long long&>(b)));
}

__device__ __forceinline__ float2 r3_fmul2(const float2 a, const float2 b) {
  float2 d;
  asm("mul.rn.f32x2 %0, %1, %2;"
      : "=l"(reinterpret_cast<unsigned long long&>(d))
      : "l"(reinterpret_cast<const unsigned long long&>(a)),
        "l"(reinterpret_cast<const unsigned long long&>(b)));
  return d;
}
__device__ __forceinline__ float4 r3_lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
template <int N> struct R3Int { static constexpr int value = N; };

// acc[PH] += w[PH] * h (static pooled-row index)
template <int PH>
__device__ __forceinline__ void r3_fold1(float2 (&a0)[RT_P][2], float2 (&a1)[RT_P][2],
                                         const float (&wv)[8], const float2 (&h0)[2],
                                         const float2 (&h1)[2]) {
  if constexpr (PH < RT_P) {
    const float2 w2 = make_float2(wv[PH], wv[PH]);
    r3_ffma2(a0[PH][0], w2, h0[0]);
    r3_ffma2(a0[PH][1], w2, h0[1]);
    r3_ffma2(a1[PH][0], w2, h1[0]);
    r3_ffma2(a1[PH][1], w2, h1[1]);
  }
}
// the row's band starts at pooled row P and is at most 3 long: fold rows P .. min(P+2, pb)
template <int P>
__device__ __forceinline__ void r3_fold3(float2 (&a0)[RT_P][2], float2 (&a1)[RT_P][2],
                                         const float (&wv)[8], int pb, const float2 (&h0)[2],
                                         const float2 (&h1)[2]) {
  r3_fold1<P>(a0, a1, wv, h0, h1);
  if (P + 1 <= pb) r3_fold1<P + 1>(a0, a1, wv, h0, h1);
  if (P + 2 <= pb) r3_fold1<P + 2>(a0, a1, wv, h0, h1);
}

// dynamic smem: ring ns*slot_bytes | tab[R3_TABS], each: desc[32 ints] | wy[max_h][8] | wx[8][max_w]
//   desc: 0 state (0 live, 1 dead, 2 end), 1 ylo(unused by consumers), 2 fh, 3 fw, 4 npass, 5 cw,
//         6 RoI index, 7 row pitch (floats), 8..15 xs[pw], 16..23 xe[pw],
//         24/25 global address of the footprint origin (channel c0)
//   wy row: 7 weights (1/count folded in) + packed non-zero band pa | pb << 8
// warps: 0..6 consumers (pooled column pw), 7 row issuer, 8 table publisher (SPLIT); without
// SPLIT warp 7 does both jobs (publishes RoI k+1 between the first ring-full of RoI k's rows
// and the rest) and the kernel keeps 128 registers per thread
template <bool SPLIT>
__global__ void __launch_bounds__(SPLIT ? R3_THREADS : RT_THREADS, 2)
roi_align_fwd3_kernel(const __grid_constant__ RoiArgs a, const float* __restrict__ rois, int R,
                      float* __restrict__ out, int32_t* __restrict__ roi_levels,
                      const Roi3Smem lay, unsigned int* __restrict__ sched
#ifdef BRCNN_DEBUG_TIMING
                      , unsigned long long* __restrict__ dbg
#endif
                      ) {
#ifdef BRCNN_DEBUG_TIMING
  const long long d_t0 = clock64();
  long long d_a = 0, d_b = 0, d_c = 0;   // role-specific phase sums
  long long d_x;
#define R3_TIC() d_x = clock64()
#define R3_TOC(v) v += clock64() - d_x
#else
#define R3_TIC()
#define R3_TOC(v)
#endif
  extern __shared__ __align__(128) unsigned char r3_smem[];
  float* ring = reinterpret_cast<float*>(r3_smem);
  float* tabs = reinterpret_cast<float*>(r3_smem + (size_t)lay.ns * lay.slot_bytes);
  __shared__ __align__(8) uint64_t ring_bar[2 * R3_MAX_STAGES];   // full[i] | empty[i] 64 B apart
  uint64_t* full_bar = ring_bar;
  uint64_t* empty_bar = ring_bar + R3_MAX_STAGES;
  __shared__ __align__(8) uint64_t tab_full[R3_TABS];
  __shared__ __align__(8) uint64_t tab_empty[R3_TABS];

  const int c0 = blockIdx.y * a.chunk_c;
  const int cc = min(a.chunk_c, a.C - c0);
  const int ncq = cc >> 2;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int NS = lay.ns;
  const int slot_floats = lay.slot_bytes >> 2;
  const int px_bytes = cc * 4;
  const int cw_max = max(1, lay.slot_bytes / px_bytes);

  if (tid == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], RT_CONS_WARPS);
    }
    for (int s = 0; s < R3_TABS; ++s) {
      mbar_init(&tab_full[s], 1);
      mbar_init(&tab_empty[s], RT_CONS_WARPS + (SPLIT ? 1 : 0));   // consumers (+ issuer)
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const uint32_t full0 = smem_u32(ring_bar), empty0 = full0 + 8u * R3_MAX_STAGES;
  const uint32_t tfull0 = smem_u32(tab_full), tempty0 = smem_u32(tab_empty);

  int s = 0, round = 0;   // ring position (issuer and consumers advance identically)
  int k = 0;              // RoIs processed so far by this CTA
  int tb_i = 0, tb_ph = 0;   // table buffer of RoI k and its barrier phase

  if (warp >= RT_CONS_WARPS) {
    // ============================= table publisher ===========================
    // Claims RoIs (atomic counter, or a static stride without the scheduling scratch), does
    // ALL per-RoI scalar work once and publishes descriptor + separable weight tables into
    // the next free table buffer (at most R3_TABS - 1 RoIs ahead of the consumers).
    unsigned int* ctr = sched ? sched + 2 * blockIdx.y : nullptr;
    int r_static = blockIdx.x;
    int pk = 0, pb_i = 0, pb_ph = 0;      // RoIs published, next buffer and its phase
    auto publish_next = [&]() -> bool {
      int r;
      if (ctr != nullptr) {
        r = 0;
        if (lane == 0) r = (int)atomicAdd(ctr, 1u);
        r = __shfl_sync(0xffffffffu, r, 0);
      } else {
        r = r_static;
        r_static += gridDim.x;
      }
      const bool end = r >= R;
      float rv[5];
#pragma unroll
      for (int i = 0; i < 5; ++i) rv[i] = end ? -1.f : __ldg(rois + (size_t)r * 5 + i);
      const bool padding = rv[0] < 0.f;
      RoiGeom g;
      int ylo = 1, yhi = 0, xlo = 1, xhi = 0;
      g.b = -1; g.lvl = 0; g.W = 0;
      if (!padding) {
        g = roi_geometry(a, rv);
        roi_axis_range(g.start_h, g.bin_h, a.PH, g.gh, g.H, ylo, yhi);
        roi_axis_range(g.start_w, g.bin_w, a.PW, g.gw, g.W, xlo, xhi);
      }
      const int fh = yhi - ylo + 1, fw = xhi - xlo + 1;
      const bool dead = padding || fh <= 0 || fw <= 0 || g.b < 0 || g.b >= a.B;
      int npass = 1, cw = fw;
      if (!dead && fw > cw_max) {
        npass = (fw + cw_max - 1) / cw_max;
        cw = (fw + npass - 1) / npass;
      }
      if (!end && roi_levels != nullptr && blockIdx.y == 0 && lane == 0)
        roi_levels[r] = padding ? -1 : g.lvl;
      float* tb = tabs + (size_t)pb_i * lay.tab_floats;
      int* desc = reinterpret_cast<int*>(tb);
      float* wy = tb + R3_DESC;                        // [fh][8]
      float* wx = wy + (size_t)a.max_h * 8;            // [pw][max_w]
      R3_TIC();
      if (pk >= R3_TABS) mbar_wait_addr_hint(tempty0 + 8u * pb_i, (uint32_t)(pb_ph ^ 1), 4000u);
      R3_TOC(d_b);
      R3_TIC();
      if (lane == 0) {
        desc[0] = end ? 2 : (dead ? 1 : 0); desc[1] = ylo; desc[2] = fh; desc[3] = fw;
        desc[4] = npass; desc[5] = cw; desc[6] = r; desc[7] = g.W * a.C;
        if (!dead) {
          const float* org = a.feat[g.lvl] +
                             (((size_t)g.b * g.H + ylo) * g.W + xlo) * a.C + c0;
          *reinterpret_cast<const float**>(desc + 24) = org;
        }
      }
      if (lane < 8) { desc[8 + lane] = fw; desc[16 + lane] = -1; }
      __syncwarp();
      if (!dead) {
        // Wy: lane = (row, pooled row); the band of a row comes from a ballot over its 8 lanes
        for (int i0 = 0; i0 < 8 * fh; i0 += 32) {
          const int i = i0 + lane, dy = i >> 3, ph = i & 7;
          float v = 0.f;
          if (dy < fh && ph < a.PH)
            v = roi_axis_weight(g.start_h, g.bin_h, g.gh, g.H, ph, ylo + dy) * g.inv_count;
          const unsigned nz = (__ballot_sync(0xffffffffu, v != 0.f) >> (lane & 24)) & 0x7fu;
          if (ph == 7) v = __int_as_float(nz ? ((__ffs(nz) - 1) | ((31 - __clz(nz)) << 8)) : 1);
          if (dy < fh) wy[i] = v;
        }
        // Wx: lane = (column, pooled column); bands through shared-memory min / max
        for (int i0 = 0; i0 < 8 * fw; i0 += 32) {
          const int i = i0 + lane, dx = i >> 3, pw = i & 7;
          if (dx < fw && pw < a.PW) {
            const float v = roi_axis_weight(g.start_w, g.bin_w, g.gw, g.W, pw, xlo + dx);
            wx[pw * a.max_w + dx] = v;
            if (v != 0.f) { atomicMin(&desc[8 + pw], dx); atomicMax(&desc[16 + pw], dx); }
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&tab_full[pb_i]);
      R3_TOC(d_a);
      if (++pb_i == R3_TABS) { pb_i = 0; pb_ph ^= 1; }
      ++pk;
      return end;
    };
    auto rearm = [&]() {
      // the last CTA to run dry re-arms the counters for the next launch
      if (ctr != nullptr && lane == 0) {
        __threadfence();
        if (atomicAdd(ctr + 1, 1u) == gridDim.x - 1) {
          atomicExch(ctr, 0u);
          atomicExch(ctr + 1, 0u);
        }
      }
    };
    if (SPLIT && warp == RT_CONS_WARPS + 1) {
      while (!publish_next()) {}
      rearm();
    } else {
      // ============================== row issuer =============================
      // One 1-D bulk copy per footprint row (per x-chunk pass) into the next ring slot; the
      // loop carries pointers and slot / barrier addresses, nothing is recomputed per row.
      const bool contiguous = (cc == a.C);
      const uint32_t ring0 = smem_u32(ring);
      uint32_t slot_addr = ring0;
      bool more = true;                     // !SPLIT: RoIs left to publish
      if (!SPLIT) more = !publish_next();
      for (;; ++k) {
        const int* desc = reinterpret_cast<const int*>(tabs + (size_t)tb_i * lay.tab_floats);
        if (SPLIT) {
          R3_TIC();
          mbar_wait_addr_hint(tfull0 + 8u * tb_i, (uint32_t)tb_ph, 4000u);
          R3_TOC(d_a);
        }
        const int4 d0 = *reinterpret_cast<const int4*>(desc);        // state, ylo, fh, fw
        if (d0.x == 2) break;
        const int4 d1 = *reinterpret_cast<const int4*>(desc + 4);    // npass, cw, r, pitch
        const float* org = *reinterpret_cast<const float* const*>(desc + 24);
        const int fh = d0.z, fw = d0.w;
        int rows_left = d0.x == 0 ? d1.x * fh : 0;
        int budget = SPLIT ? rows_left : min(rows_left, NS);          // rows before the publish
        int pass = 0, dy = 0;
        int cwe = min(d1.y, fw);
        uint32_t bytes = (uint32_t)(cwe * px_bytes);
        const float* src = org;
        for (int phase = 0; phase < 2; ++phase) {
          for (; budget > 0; --budget, --rows_left) {
            R3_TIC();
            if (round > 0) mbar_wait_addr_hint(empty0 + 8u * s, (uint32_t)((round - 1) & 1), 4000u);
            R3_TOC(d_b);
            const uint32_t fb = full0 + 8u * s;
            if (contiguous) {
              if (lane == 0) {
                mbar_expect_tx_addr(fb, bytes);
                tma_load_1d_addr(slot_addr, src, bytes, fb);
              }
            } else {
              if (lane == 0) mbar_expect_tx_addr(fb, bytes);
              __syncwarp();
              for (int px = lane; px < cwe; px += 32)
                tma_load_1d_addr(slot_addr + (uint32_t)(px * px_bytes), src + (size_t)px * a.C,
                                 (uint32_t)px_bytes, fb);
            }
            src += d1.w;
            slot_addr += (uint32_t)lay.slot_bytes;
            if (++s == NS) { s = 0; ++round; slot_addr = ring0; }
            if (++dy == fh) {                 // next x-chunk pass
              dy = 0;
              ++pass;
              const int x0 = pass * d1.y;
              cwe = min(d1.y, fw - x0);
              bytes = (uint32_t)(cwe * px_bytes);
              src = org + (size_t)x0 * a.C;
            }
          }
          if (phase == 0) {
            if (!SPLIT && more) {
              R3_TIC();
              more = !publish_next();
              R3_TOC(d_c);
            }
            budget = rows_left;
          }
        }
        if (SPLIT) {
          __syncwarp();
          if (lane == 0) mbar_arrive_addr(tempty0 + 8u * tb_i);
        }
        if (++tb_i == R3_TABS) { tb_i = 0; tb_ph ^= 1; }
      }
      if (!SPLIT) rearm();
    }
  } else {
    // ============================= consumer warps ============================
    const int pw = warp;
    const bool act0 = (pw < a.PW) && (lane < ncq);
    const bool act1 = (pw < a.PW) && (lane + 32 < ncq);
    const int q1 = act1 ? lane + 32 : lane;   // inactive second quad re-reads the first
    const size_t bin_stride = (size_t)a.C;
    for (;; ++k) {
      const float* tb = tabs + (size_t)tb_i * lay.tab_floats;
      const int* desc = reinterpret_cast<const int*>(tb);
      const float* wy = tb + R3_DESC;
      const float* wxp = wy + (size_t)a.max_h * 8 + (size_t)(act0 ? pw : 0) * a.max_w;
      R3_TIC();
      mbar_wait_addr_hint(tfull0 + 8u * tb_i, (uint32_t)tb_ph, 2000u);
      R3_TOC(d_a);
      const int4 d0 = *reinterpret_cast<const int4*>(desc);        // state, ylo, fh, fw
      if (d0.x == 2) break;                                        // no RoI left
      const int npass = desc[4], cw = desc[5];
      float* dst = out + (size_t)desc[6] * a.PH * a.PW * a.C + c0 + (size_t)pw * bin_stride;
      const int fh = d0.z, fw = d0.w;
      float2 acc0[RT_P][2], acc1[RT_P][2];
#pragma unroll
      for (int i = 0; i < RT_P; ++i) {
        acc0[i][0] = acc0[i][1] = make_float2(0.f, 0.f);
        acc1[i][0] = acc1[i][1] = make_float2(0.f, 0.f);
      }
      if (d0.x == 0) {
        const int bxs = act0 ? desc[8 + pw] : 1, bxe = act0 ? desc[16 + pw] : 0;
        // y-fold of one x-folded footprint row into the pooled rows of its band
        auto fold = [&](int dy, const float2 (&h0)[2], const float2 (&h1)[2]) {
              const float4 wa = *reinterpret_cast<const float4*>(wy + dy * 8);
              const float4 wb = *reinterpret_cast<const float4*>(wy + dy * 8 + 4);
              const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, 0.f};
              const int pk = __float_as_int(wb.w);
              const int pa = pk & 0xff, pb = pk >> 8;       // warp-uniform band of this row
              if (pa <= pb) {
                if (pb - pa <= 2) {
                  switch (pa) {
                    case 0: r3_fold3<0>(acc0, acc1, wv, pb, h0, h1); break;
                    case 1: r3_fold3<1>(acc0, acc1, wv, pb, h0, h1); break;
                    case 2: r3_fold3<2>(acc0, acc1, wv, pb, h0, h1); break;
                    case 3: r3_fold3<3>(acc0, acc1, wv, pb, h0, h1); break;
                    case 4: r3_fold3<4>(acc0, acc1, wv, pb, h0, h1); break;
                    case 5: r3_fold3<5>(acc0, acc1, wv, pb, h0, h1); break;
                    default: r3_fold3<6>(acc0, acc1, wv, pb, h0, h1); break;
                  }
                } else {
#pragma unroll
                  for (int ph = 0; ph < RT_P; ++ph) {
                    const float w = wv[ph];
                    if (w == 0.f) continue;   // zero outside the band (skip: 0 * inf stays out)
                    const float2 w2 = make_float2(w, w);
                    r3_ffma2(acc0[ph][0], w2, h0[0]);
                    r3_ffma2(acc0[ph][1], w2, h0[1]);
                    r3_ffma2(acc1[ph][0], w2, h1[0]);
                    r3_ffma2(acc1[ph][1], w2, h1[1]);
                  }
                }
              }
        };
        const int nb = bxe - bxs + 1;                  // this pooled column's pixel band
        if (npass == 1 && nb <= 4) {
          // ---- common case: one x-chunk pass, band of <= 4 pixels.  The band's weights are
          // per-RoI constants of the warp (registers), the row loop is specialised on the band
          // length: no inner loop, no per-row weight loads, slot / barrier addresses carried ----
          float wq[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) wq[i] = (i < nb) ? wxp[bxs + i] : 0.f;
          const uint32_t ps = (uint32_t)ncq * 16u;                       // bytes per pixel
          const uint32_t q1off = (uint32_t)(q1 - lane) * 16u;
          const uint32_t ring_end = smem_u32(ring) + (uint32_t)(NS * lay.slot_bytes);
          uint32_t c_slot = smem_u32(ring) + (uint32_t)(s * lay.slot_bytes) +
                            (nb > 0 ? (uint32_t)(bxs * ncq + lane) * 16u : 0u);
          uint32_t c_fb = full0 + 8u * s;
          uint32_t c_par = (uint32_t)(round & 1);
          auto rows = [&](auto nbc) {
            constexpr int NB = decltype(nbc)::value;
#pragma unroll 1
            for (int dy = 0; dy < fh; ++dy) {
              R3_TIC();
              mbar_wait_addr(c_fb, c_par);
              R3_TOC(d_b);
              float2 h0[2], h1[2];
              if constexpr (NB > 0) {
                float4 v0 = r3_lds128(c_slot), v1 = r3_lds128(c_slot + q1off);
                float2 w2 = make_float2(wq[0], wq[0]);
                h0[0] = r3_fmul2(w2, make_float2(v0.x, v0.y));
                h0[1] = r3_fmul2(w2, make_float2(v0.z, v0.w));
                h1[0] = r3_fmul2(w2, make_float2(v1.x, v1.y));
                h1[1] = r3_fmul2(w2, make_float2(v1.z, v1.w));
#pragma unroll
                for (int i = 1; i < NB; ++i) {
                  v0 = r3_lds128(c_slot + i * ps);
                  v1 = r3_lds128(c_slot + i * ps + q1off);
                  w2 = make_float2(wq[i], wq[i]);
                  r3_ffma2(h0[0], w2, make_float2(v0.x, v0.y));
                  r3_ffma2(h0[1], w2, make_float2(v0.z, v0.w));
                  r3_ffma2(h1[0], w2, make_float2(v1.x, v1.y));
                  r3_ffma2(h1[1], w2, make_float2(v1.z, v1.w));
                }
              }
              __syncwarp();
              if (lane == 0) mbar_arrive_addr(c_fb + 8u * R3_MAX_STAGES);
              if constexpr (NB > 0) fold(dy, h0, h1);
              c_slot += (uint32_t)lay.slot_bytes;
              c_fb += 8u;
              if (++s == NS) {
                s = 0; ++round; c_par ^= 1u;
                c_slot -= (uint32_t)(NS * lay.slot_bytes);
                c_fb = full0;
              }
            }
          };
          switch (nb) {
            case 1: rows(R3Int<1>()); break;
            case 2: rows(R3Int<2>()); break;
            case 3: rows(R3Int<3>()); break;
            case 4: rows(R3Int<4>()); break;
            default: rows(R3Int<0>()); break;       // nothing of this column (or inactive lanes)
          }
          (void)ring_end;
        } else {
        for (int pass = 0; pass < npass; ++pass) {
          const int x0 = pass * cw;
          const int cwe = min(cw, fw - x0);
          const int xs = max(bxs, x0), xe = min(bxe, x0 + cwe - 1);
          const bool work = act0 && (xs <= xe);
          for (int dy = 0; dy < fh; ++dy) {
            R3_TIC();
            mbar_wait_addr(full0 + 8u * s, (uint32_t)(round & 1));
            R3_TOC(d_b);
            float2 h0[2], h1[2];
            h0[0] = h0[1] = h1[0] = h1[1] = make_float2(0.f, 0.f);
            if (work) {
              const float4* px = reinterpret_cast<const float4*>(ring + (size_t)s * slot_floats) +
                                 (xs - x0) * ncq;
              for (int x = xs; x <= xe; ++x) {
                const float4 v0 = px[lane];
                const float4 v1 = px[q1];
                const float w = wxp[x];
                px += ncq;
                const float2 w2 = make_float2(w, w);
                r3_ffma2(h0[0], w2, make_float2(v0.x, v0.y));
                r3_ffma2(h0[1], w2, make_float2(v0.z, v0.w));
                r3_ffma2(h1[0], w2, make_float2(v1.x, v1.y));
                r3_ffma2(h1[1], w2, make_float2(v1.z, v1.w));
              }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive_addr(empty0 + 8u * s);
            if (work) fold(dy, h0, h1);
            if (++s == NS) { s = 0; ++round; }
          }
        }
        }
      }
      // tables no longer needed: the publisher may overwrite this buffer
      __syncwarp();
      if (lane == 0) mbar_arrive_addr(tempty0 + 8u * tb_i);
      if (++tb_i == R3_TABS) { tb_i = 0; tb_ph ^= 1; }
      R3_TIC();
      if (pw < a.PW) {
#pragma unroll
        for (int ph = 0; ph < RT_P; ++ph) {
          if (ph < a.PH) {
            float4* o = reinterpret_cast<float4*>(dst + (size_t)(ph * a.PW) * bin_stride);
            if (act0) __stcs(o + lane, make_float4(acc0[ph][0].x, acc0[ph][0].y, acc0[ph][1].x,
                                                   acc0[ph][1].y));
            if (act1) __stcs(o + lane + 32, make_float4(acc1[ph][0].x, acc1[ph][0].y,
                                                        acc1[ph][1].x, acc1[ph][1].y));
          }
        }
      }
      R3_TOC(d_c);
    }
  }
#ifdef BRCNN_DEBUG_TIMING
  if (dbg != nullptr && lane == 0 && (warp == 0 || warp >= RT_CONS_WARPS)) {
    const int o = warp == 0 ? 0 : (warp == RT_CONS_WARPS ? 8 : 16);
    atomicAdd(dbg + o + 0, 1ull);
    atomicAdd(dbg + o + 1, (unsigned long long)(clock64() - d_t0));
    atomicAdd(dbg + o + 2, (unsigned long long)d_a);   // consumer / issuer: table waits | publisher: tables
    atomicAdd(dbg + o + 3, (unsigned long long)d_b);   // consumer: row waits | issuer: slot waits | publisher: buffer waits
    atomicAdd(dbg + o + 4, (unsigned long long)d_c);   // consumer: stores
    atomicAdd(dbg + o + 5, (unsigned long long)k);
  }
#endif
}

}  // namespace brcnn
