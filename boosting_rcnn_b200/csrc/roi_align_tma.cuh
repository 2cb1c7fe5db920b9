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
//   * 14 consumer warps = 7 pooled columns (pw) x 2 half-slabs of 32 channel
//     quads.  Per row a thread forms h = sum_{x in band(pw)} Wx[pw][x]*F[y][x]
//     (128-bit conflict-free LDS, warp-uniform weights), releases the slot,
//     then folds h into the 1-3 pooled rows whose Wy[.][y] is non-zero
//     (register accumulators acc[7]).
//   * The (c, ph, pw) result is staged in the (now idle) ring and leaves the
//     SM as one `cp.async.bulk` shared->global store of C*49*4 bytes.
// Wide footprints (fw*C*4 too large for >= 3 slots) are processed in several
// x-chunk passes over the rows; the accumulators simply carry across passes.
#pragma once
#include "common.cuh"
#include "roi_align.cuh"

namespace brcnn {

constexpr int RT_P = 7;                       // max pooled side on this path
constexpr int RT_CONS_WARPS = 2 * RT_P;       // (pw, half-slab)
constexpr int RT_THREADS = (RT_CONS_WARPS + 1) * 32;
constexpr int RT_MAX_STAGES = 8;
constexpr int RT_SLAB_Q = 64;                 // channel quads per CTA (256 channels)

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

// dynamic smem: ring_bytes | wy [max_h][8] | wx [8][max_w] | rowpk [max_h]
__global__ void __launch_bounds__(RT_THREADS, 2)
roi_align_fwd_tma_kernel(const __grid_constant__ RoiArgs a, const float* __restrict__ rois,
                         int R, float* __restrict__ out, int32_t* __restrict__ roi_levels,
                         int ring_bytes) {
  extern __shared__ __align__(128) unsigned char rt_smem[];
  float* ring = reinterpret_cast<float*>(rt_smem);
  float* wy = reinterpret_cast<float*>(rt_smem + ring_bytes);   // [fh][8]   (1/count folded in)
  float* wx = wy + (size_t)a.max_h * 8;                          // [pw][max_w]
  int* rowpk = reinterpret_cast<int*>(wx + (size_t)8 * a.max_w); // [fh] first | last<<8 pooled row
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
  const int npass = (fw + cw_max - 1) / cw_max;
  const int cw = (fw + npass - 1) / npass;                  // pixels per slot
  const int slot_bytes = ((cw * px_bytes) + 127) & ~127;
  const int NS = min(RT_MAX_STAGES, ring_bytes / slot_bytes);
  const int slot_floats = slot_bytes >> 2;
  const int niter = npass * fh;

  if (tid == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], RT_CONS_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == RT_CONS_WARPS) {
    // =========================== producer warp ===========================
    const float* fbase = a.feat[g.lvl] + ((size_t)g.b * g.H * g.W) * a.C + c0;
    const bool contiguous = (cc == a.C);
    int s = 0, round = 0, it = 0;
    for (int pass = 0; pass < npass; ++pass) {
      const int x0 = pass * cw;
      const int cwe = min(cw, fw - x0);
      for (int dy = 0; dy < fh; ++dy, ++it) {
        if (round > 0) mbar_wait(&empty_bar[s], (uint32_t)((round - 1) & 1));
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
    const int ctid = tid;  // 0 .. 447
    // separable weight tables (built while the first rows are in flight)
    for (int i = ctid; i < a.PH * fh; i += RT_CONS_WARPS * 32) {
      const int dy = i / a.PH, ph = i - dy * a.PH;
      wy[dy * 8 + ph] =
          roi_axis_weight(g.start_h, g.bin_h, g.gh, g.H, ph, ylo + dy) * g.inv_count;
    }
    for (int i = ctid; i < a.PW * fw; i += RT_CONS_WARPS * 32) {
      const int pw = i / fw, dx = i - pw * fw;
      wx[pw * a.max_w + dx] = roi_axis_weight(g.start_w, g.bin_w, g.gw, g.W, pw, xlo + dx);
    }
    asm volatile("bar.sync 1, %0;" ::"n"(RT_CONS_WARPS * 32) : "memory");
    for (int dy = ctid; dy < fh; dy += RT_CONS_WARPS * 32) {
      int pa = a.PH, pb = -1;
      for (int ph = 0; ph < a.PH; ++ph)
        if (wy[dy * 8 + ph] != 0.f) { pa = min(pa, ph); pb = ph; }
      rowpk[dy] = (pb < 0) ? (1 | (0 << 8)) : (pa | (pb << 8));   // empty: pa=1 > pb=0
    }
    if (ctid >= 64 && ctid < 64 + a.PW) {
      const int pw = ctid - 64;
      int xs = fw, xe = -1;
      for (int dx = 0; dx < fw; ++dx)
        if (wx[pw * a.max_w + dx] != 0.f) { xs = min(xs, dx); xe = dx; }
      s_xs[pw] = xs; s_xe[pw] = xe;
    }
    asm volatile("bar.sync 1, %0;" ::"n"(RT_CONS_WARPS * 32) : "memory");

    const int pw = warp >> 1;
    const int q = ((warp & 1) << 5) | lane;
    const bool active = (pw < a.PW) && (q < ncq);
    float4 acc[RT_P];
#pragma unroll
    for (int i = 0; i < RT_P; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int bxs = active ? s_xs[pw] : 1, bxe = active ? s_xe[pw] : 0;
    const float* wxp = wx + (size_t)(active ? pw : 0) * a.max_w;

    int s = 0, round = 0;
    for (int pass = 0; pass < npass; ++pass) {
      const int x0 = pass * cw;
      const int cwe = min(cw, fw - x0);
      const int xs = max(bxs, x0), xe = min(bxe, x0 + cwe - 1);
      const bool work = active && (xs <= xe);
      for (int dy = 0; dy < fh; ++dy) {
        mbar_wait(&full_bar[s], (uint32_t)(round & 1));
        float4 h = make_float4(0.f, 0.f, 0.f, 0.f);
        if (work) {
          const float4* rowq =
              reinterpret_cast<const float4*>(ring + (size_t)s * slot_floats) + q;
          for (int x = xs; x <= xe; ++x) {
            const float4 v = rowq[(size_t)(x - x0) * ncq];
            const float w = wxp[x];
            h.x = fmaf(w, v.x, h.x);
            h.y = fmaf(w, v.y, h.y);
            h.z = fmaf(w, v.z, h.z);
            h.w = fmaf(w, v.w, h.w);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[s]);
        if (work) {
          const int pk = rowpk[dy];
          const int pa = pk & 0xff, pb = pk >> 8;
          const float* wrow = wy + dy * 8;
#pragma unroll
          for (int ph = 0; ph < RT_P; ++ph) {
            if (ph >= pa && ph <= pb) {
              const float w = wrow[ph];
              acc[ph].x = fmaf(w, h.x, acc[ph].x);
              acc[ph].y = fmaf(w, h.y, acc[ph].y);
              acc[ph].z = fmaf(w, h.z, acc[ph].z);
              acc[ph].w = fmaf(w, h.w, acc[ph].w);
            }
          }
        }
        if (++s == NS) { s = 0; ++round; }
      }
    }
    // every slot has been consumed by this warp; wait for the other consumers
    // before the ring is reused as the output stage
    asm volatile("bar.sync 1, %0;" ::"n"(RT_CONS_WARPS * 32) : "memory");
    if (active) {
      float* st = ring + (size_t)(q * 4) * nbins + pw;
#pragma unroll
      for (int ph = 0; ph < RT_P; ++ph) {
        if (ph < a.PH) {
          st[ph * a.PW] = acc[ph].x;
          st[nbins + ph * a.PW] = acc[ph].y;
          st[2 * nbins + ph * a.PW] = acc[ph].z;
          st[3 * nbins + ph * a.PW] = acc[ph].w;
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
  (void)niter;
}

}  // namespace brcnn
