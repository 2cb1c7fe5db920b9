// K3 (v2): fused FPN level mapping + multi-level RoIAlign forward, footprint
// rows streamed through shared memory by the TMA engine.
//
// Reference behaviour: single_level_roi_extractor.py:36-115 + mmcv RoIAlign
// (aligned=True, pool_mode='avg', sampling_ratio=0), SURVEY.md App. A5/A6.
//
// Same separable formulation as roi_align.cuh,
//     out[c][ph][pw] = sum_y Wy[ph][y] * ( sum_x Wx[pw][x] * F[y][x][c] ),
// but evaluated x-first, one footprint row at a time:
//   * In NHWC a footprint row (fw pixels x C channels) is ONE contiguous run of
//     fw*C*4 bytes.  A dedicated producer warp streams the rows into a ring of
//     shared-memory slots with `cp.async.bulk` (1-D TMA, UBLKCP in SASS),
//     completion signalled on a per-slot mbarrier.  Up to 96 KB of loads are in
//     flight per CTA with zero registers spent on them; every footprint pixel
//     crosses L2->SM exactly once.
//   * 7 consumer warps = the 7 pooled columns (pw); a lane owns two channel
//     quads (lane, lane+32).  Per row a thread forms
//     h = sum_{x in band(pw)} Wx[pw][x]*F[y][x] (128-bit conflict-free LDS,
//     warp-uniform weights), releases the slot, then folds h into the 7 pooled
//     rows with the row's Wy[.][y] (zero outside the band; branch-free, two
//     128-bit weight loads) in register accumulators.
//   * The (c, ph, pw) result is staged in the (now idle) ring and leaves the
//     SM as one `cp.async.bulk` shared->global store of C*49*4 bytes.
// Wide footprints (fw*C*4 too large for >= 3 slots) are processed in several
// x-chunk passes over the rows; the accumulators simply carry across passes.
#pragma once
#include "common.cuh"
#include "roi_align.cuh"

namespace brcnn {

constexpr int RT_P = 7;                       // max pooled side on this path
constexpr int RT_CONS_WARPS = RT_P;           // one consumer warp per pooled column
constexpr int RT_THREADS = (RT_CONS_WARPS + 1) * 32;
constexpr int RT_MAX_STAGES = 8;
constexpr int RT_SLAB_Q = 64;                 // channel quads per CTA (2 per lane)

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
               ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void mbar_arrive_addr(uint32_t addr) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_addr(uint32_t addr, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
  } while (!ok);
}
// same, with a suspend-time hint: an idle waiter is parked by the hardware for up to `ns`
// nanoseconds per try instead of spinning through the issue slots the busy warps need
__device__ __forceinline__ void mbar_wait_addr_hint(uint32_t addr, uint32_t parity, uint32_t ns) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(addr), "r"(parity), "r"(ns) : "memory");
  } while (!ok);
}
// 1-D bulk copy global -> shared, completion bytes on `bar`
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src, uint32_t bytes,
                                            uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// 1-D bulk copy shared -> global (bulk async-group)
__device__ __forceinline__ void tma_store_1d(void* dst, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
               ::"l"(dst), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// dynamic smem: ring_bytes | wy [max_h][8] | wx [8][max_w]
__global__ void __launch_bounds__(RT_THREADS, 2)
roi_align_fwd_tma_kernel(const __grid_constant__ RoiArgs a, const float* __restrict__ rois,
                         int R, float* __restrict__ out, int32_t* __restrict__ roi_levels,
                         int ring_bytes) {
  extern __shared__ __align__(128) unsigned char rt_smem[];
  float* ring = reinterpret_cast<float*>(rt_smem);
  float* wy = reinterpret_cast<float*>(rt_smem + ring_bytes);   // [fh][8], 1/count folded in,
                                                                 // zero outside the row's band
  float* wx = wy + (size_t)a.max_h * 8;                          // [pw][max_w]
  __shared__ __align__(8) uint64_t full_bar[RT_MAX_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[RT_MAX_STAGES];
  __shared__ int s_xs[8], s_xe[8];

  const int r = blockIdx.x;
  const int c0 = blockIdx.y * a.chunk_c;
  const int cc = min(a.chunk_c, a.C - c0);
  const int ncq = cc >> 2;
  const int nbins = a.PH * a.PW;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* roi = rois + (size_t)r * 5;
  float* dst = out + ((size_t)r * a.C + c0) * nbins;
  const int total = cc * nbins;

  const bool padding = roi[0] < 0.f;
  RoiGeom g;
  int ylo = 1, yhi = 0, xlo = 1, xhi = 0;
  if (!padding) {
    g = roi_geometry(a, roi);
    roi_axis_range(g.start_h, g.bin_h, a.PH, g.gh, g.H, ylo, yhi);
    roi_axis_range(g.start_w, g.bin_w, a.PW, g.gw, g.W, xlo, xhi);
  }
  if (roi_levels != nullptr && blockIdx.y == 0 && tid == 0) roi_levels[r] = padding ? -1 : g.lvl;
  const int fh = yhi - ylo + 1, fw = xhi - xlo + 1;
  if (padding || fh <= 0 || fw <= 0 || g.b < 0 || g.b >= a.B) {  // block-uniform
    for (int i = tid; i < total; i += RT_THREADS) dst[i] = 0.f;
    return;
  }

  // ---- ring geometry (block-uniform) ----
  const int px_bytes = cc * 4;
  const int cw_max = max(1, (ring_bytes / 3) / px_bytes);   // >= 3 slots
  int npass = 1, cw = fw;
  if (fw > cw_max) {
    npass = (fw + cw_max - 1) / cw_max;
    cw = (fw + npass - 1) / npass;                          // pixels per slot
  }
  const int slot_bytes = ((cw * px_bytes) + 127) & ~127;
  int NS = RT_MAX_STAGES;
  while (NS * slot_bytes > ring_bytes) --NS;
  const int slot_floats = slot_bytes >> 2;
  const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar);

  if (tid == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], RT_CONS_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < 8) { s_xs[tid] = fw; s_xe[tid] = -1; }
  __syncthreads();

  if (warp == RT_CONS_WARPS) {
    // =========================== producer warp ===========================
    const float* fbase = a.feat[g.lvl] + ((size_t)g.b * g.H * g.W) * a.C + c0;
    const bool contiguous = (cc == a.C);
    int s = 0, round = 0;
    for (int pass = 0; pass < npass; ++pass) {
      const int x0 = pass * cw;
      const int cwe = min(cw, fw - x0);
      for (int dy = 0; dy < fh; ++dy) {
        if (round > 0) mbar_wait_addr(empty0 + 8u * s, (uint32_t)((round - 1) & 1));
        const float* src = fbase + ((size_t)(ylo + dy) * g.W + xlo + x0) * a.C;
        float* slot = ring + (size_t)s * slot_floats;
        if (lane == 0) mbar_expect_tx(&full_bar[s], (uint32_t)(cwe * px_bytes));
        if (contiguous) {
          if (lane == 0) tma_load_1d(slot, src, (uint32_t)(cwe * px_bytes), &full_bar[s]);
        } else {
          __syncwarp();
          for (int px = lane; px < cwe; px += 32)
            tma_load_1d(slot + (size_t)px * cc, src + (size_t)px * a.C, (uint32_t)px_bytes,
                        &full_bar[s]);
        }
        if (++s == NS) { s = 0; ++round; }
      }
    }
  } else {
    // =========================== consumer warps ==========================
    constexpr int NCT = RT_CONS_WARPS * 32;
    // separable weight tables (built while the first rows are in flight)
    for (int i = tid; i < 8 * fh; i += NCT) {
      const int dy = i >> 3, ph = i & 7;
      wy[i] = (ph < a.PH)
          ? roi_axis_weight(g.start_h, g.bin_h, g.gh, g.H, ph, ylo + dy) * g.inv_count : 0.f;
    }
    for (int i = tid; i < a.PW * fw; i += NCT) {
      const int pw = i / fw, dx = i - pw * fw;
      const float w = roi_axis_weight(g.start_w, g.bin_w, g.gw, g.W, pw, xlo + dx);
      wx[pw * a.max_w + dx] = w;
      if (w != 0.f) {  // non-zero band of pooled column pw
        atomicMin(&s_xs[pw], dx);
        atomicMax(&s_xe[pw], dx);
      }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(NCT) : "memory");

    const int pw = warp;
    const bool act0 = (pw < a.PW) && (lane < ncq);
    const bool act1 = (pw < a.PW) && (lane + 32 < ncq);
    float4 acc0[RT_P], acc1[RT_P];
#pragma unroll
    for (int i = 0; i < RT_P; ++i) {
      acc0[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      acc1[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const int bxs = act0 ? s_xs[pw] : 1, bxe = act0 ? s_xe[pw] : 0;
    const float* wxp = wx + (size_t)(act0 ? pw : 0) * a.max_w;
    const int q1 = act1 ? lane + 32 : lane;   // inactive second quad re-reads the first

    int s = 0, round = 0;
    for (int pass = 0; pass < npass; ++pass) {
      const int x0 = pass * cw;
      const int cwe = min(cw, fw - x0);
      const int xs = max(bxs, x0), xe = min(bxe, x0 + cwe - 1);
      const bool work = act0 && (xs <= xe);
      for (int dy = 0; dy < fh; ++dy) {
        mbar_wait_addr(full0 + 8u * s, (uint32_t)(round & 1));
        float4 h0 = make_float4(0.f, 0.f, 0.f, 0.f), h1 = h0;
        if (work) {
          const float4* rowq = reinterpret_cast<const float4*>(ring + (size_t)s * slot_floats);
          for (int x = xs; x <= xe; ++x) {
            const float4* px = rowq + (x - x0) * ncq;
            const float4 v0 = px[lane];
            const float4 v1 = px[q1];
            const float w = wxp[x];
            h0.x = fmaf(w, v0.x, h0.x); h0.y = fmaf(w, v0.y, h0.y);
            h0.z = fmaf(w, v0.z, h0.z); h0.w = fmaf(w, v0.w, h0.w);
            h1.x = fmaf(w, v1.x, h1.x); h1.y = fmaf(w, v1.y, h1.y);
            h1.z = fmaf(w, v1.z, h1.z); h1.w = fmaf(w, v1.w, h1.w);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive_addr(empty0 + 8u * s);
        if (work) {
          const float4 wa = *reinterpret_cast<const float4*>(wy + dy * 8);
          const float4 wb = *reinterpret_cast<const float4*>(wy + dy * 8 + 4);
          const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
          for (int ph = 0; ph < RT_P; ++ph) {
            const float w = wv[ph];
            if (w == 0.f) continue;   // warp-uniform: Wy[ph][y] is zero outside the band
            acc0[ph].x = fmaf(w, h0.x, acc0[ph].x); acc0[ph].y = fmaf(w, h0.y, acc0[ph].y);
            acc0[ph].z = fmaf(w, h0.z, acc0[ph].z); acc0[ph].w = fmaf(w, h0.w, acc0[ph].w);
            acc1[ph].x = fmaf(w, h1.x, acc1[ph].x); acc1[ph].y = fmaf(w, h1.y, acc1[ph].y);
            acc1[ph].z = fmaf(w, h1.z, acc1[ph].z); acc1[ph].w = fmaf(w, h1.w, acc1[ph].w);
          }
        }
        if (++s == NS) { s = 0; ++round; }
      }
    }
    // every slot has been consumed by this warp; wait for the other consumers
    // before the ring is reused as the output stage
    asm volatile("bar.sync 1, %0;" ::"n"(NCT) : "memory");
    if (act0) {
      float* st = ring + (size_t)(lane * 4) * nbins + pw;
#pragma unroll
      for (int ph = 0; ph < RT_P; ++ph) {
        if (ph < a.PH) {
          st[ph * a.PW] = acc0[ph].x;
          st[nbins + ph * a.PW] = acc0[ph].y;
          st[2 * nbins + ph * a.PW] = acc0[ph].z;
          st[3 * nbins + ph * a.PW] = acc0[ph].w;
        }
      }
    }
    if (act1) {
      float* st = ring + (size_t)((lane + 32) * 4) * nbins + pw;
#pragma unroll
      for (int ph = 0; ph < RT_P; ++ph) {
        if (ph < a.PH) {
          st[ph * a.PW] = acc1[ph].x;
          st[nbins + ph * a.PW] = acc1[ph].y;
          st[2 * nbins + ph * a.PW] = acc1[ph].z;
          st[3 * nbins + ph * a.PW] = acc1[ph].w;
        }
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    tma_store_1d(dst, ring, (uint32_t)(total * 4));
    tma_store_wait_read();
  }
}

}  // namespace brcnn
