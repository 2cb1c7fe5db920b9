// libbrcnn.so — C ABI (include/brcnn.h) over the sm_100a kernels.
// Single translation unit: nvcc -gencode arch=compute_100a,code=sm_100a
//   -fmad=false -O3 -lineinfo -shared -Xcompiler -fPIC
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <utility>

#include "../../include/brcnn.h"
#include "boost_loss.cuh"
#include "common.cuh"
#include "nms_kernels.cuh"
#include "rcnn_post.cuh"
#include "rcnn_train_prep.cuh"
#include "roi_align.cuh"
#include "roi_align_bwd.cuh"
#include "roi_align_bwd2.cuh"
#include "roi_align_bwd3.cuh"
#include "roi_align_bwd5.cuh"
#include "roi_align_fwd3.cuh"
#include "roi_align_tma.cuh"
#include "rpn.cuh"
#include "rpn_loss.cuh"
#include "rpn_nms.cuh"

namespace brcnn {
static std::atomic<int64_t> g_launches{0};
int64_t g_launch_count_add(int n) { return g_launches.fetch_add(n) + n; }

static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

// cudaFuncSetAttribute is issued once per (kernel, device): the largest dynamic
// shared-memory size requested so far is remembered, later calls with a size that
// already fits are free.  (Thread-safe; the only process-wide state of the library
// besides the launch counter.)
static std::mutex g_attr_mu;
static std::map<std::pair<const void*, int>, int> g_attr_smem;
static std::map<int, int> g_sm_count;
cudaError_t ensure_dyn_smem(const void* fn, size_t bytes, bool max_carveout) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lk(g_attr_mu);
  const auto key = std::make_pair(fn, dev);
  const auto it = g_attr_smem.find(key);
  if (it != g_attr_smem.end() && it->second >= (int)bytes) return cudaSuccess;
  e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) return e;
  if (max_carveout) {
    e = cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout,
                             cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
  }
  g_attr_smem[key] = (int)bytes;
  return cudaSuccess;
}
static int sm_count() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  std::lock_guard<std::mutex> lk(g_attr_mu);
  const auto it = g_sm_count.find(dev);
  if (it != g_sm_count.end()) return it->second;
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
    n = 148;
  g_sm_count[dev] = n;
  return n;
}
static inline bool misaligned16(const void* p) { return ((uintptr_t)p & 15) != 0; }
static inline int next_pow2(int x) { int p = 1; while (p < x) p <<= 1; return p; }

// --------------------------------------------------------------------------
// generic nms operator kernels (single segment)
// --------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ordered_bits(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// max over all 4K coordinates (handles negatives).  Multi-CTA: every block reduces its slice
// and does one atomicMax on the order-preserving bit pattern in `bits` (zeroed by the caller:
// 0 orders below every float); nms_maxcoord_finish turns the pattern back into out[0].
__global__ void __launch_bounds__(256)
nms_maxcoord_kernel(const float* __restrict__ boxes, int K, unsigned int* __restrict__ bits) {
  __shared__ float s[8];
  float m = -INFINITY;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 4 * K; i += gridDim.x * blockDim.x)
    m = fmaxf(m, boxes[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) m = fmaxf(m, s[w]);
    atomicMax(bits, ordered_bits(m));
  }
}
__global__ void nms_maxcoord_finish_kernel(const unsigned int* __restrict__ bits, float* out) {
  const unsigned int u = bits[0];
  out[0] = __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// rank by counting: rank_i = #{j : key_j > key_i}; scatter sorted arrays.
__global__ void __launch_bounds__(256)
nms_rank_scatter_kernel(const float* __restrict__ boxes,
                        const float* __restrict__ scores,
                        const int64_t* __restrict__ idxs, int K,
                        const float* __restrict__ maxc,
                        float4* __restrict__ sorted_boxes,
                        u64* __restrict__ sorted_key, int32_t* __restrict__ order,
                        int32_t* __restrict__ count) {
  __shared__ u64 tile[256];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  u64 mine = 0;
  if (i < K) mine = ((u64)ordered_bits(scores[i]) << 32) | (u64)(0xFFFFFFFFu - (uint32_t)i);
  int rank = 0;
  for (int j0 = 0; j0 < K; j0 += 256) {
    const int j = j0 + threadIdx.x;
    __syncthreads();
    tile[threadIdx.x] = (j < K)
        ? (((u64)ordered_bits(scores[j]) << 32) | (u64)(0xFFFFFFFFu - (uint32_t)j))
        : 0ull;
    __syncthreads();
    const int m = min(256, K - j0);
    for (int t = 0; t < m; ++t) rank += (tile[t] > mine);
  }
  if (i < K) {
    float4 b = reinterpret_cast<const float4*>(boxes)[i];
    if (idxs != nullptr) {
      const float o = (float)idxs[i] * (maxc[0] + 1.0f);
      b = add_seg_offset(b, o);
    }
    sorted_boxes[rank] = b;
    sorted_key[rank] = mine;
    order[rank] = i;
  }
  if (i == 0) count[0] = K;
}

// Sort by (id asc, score desc, index asc) by counting: position of box i =
// #{j : id_j < id_i} + #{j : id_j == id_i, key_j > key_i}.  The K x K comparison is split over
// a 2-D grid (x: 256 boxes i, y: a slice of the boxes j) with integer atomics into lt / rank
// (order independent, deterministic); nms_id_scatter_kernel then writes the raw boxes and keys
// in that order plus the start / size of every id's list (ids in [0, num_ids); idxs == nullptr:
// one list).  (One CTA column per 256 boxes looping over all K took 0.36 ms at K = 20 000.)
constexpr int NMS_RANK_JSPLIT = 2048;   // boxes j per CTA row
__global__ void __launch_bounds__(256)
nms_id_rank_count_kernel(const float* __restrict__ scores, const int64_t* __restrict__ idxs, int K,
                         int32_t* __restrict__ lt_out, int32_t* __restrict__ rank_out,
                         const int32_t* __restrict__ gate) {
  __shared__ u64 t_key[256];
  __shared__ int t_id[256];
  if (gate != nullptr && *gate == 0) return;     // the fast id sort already did the job
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  u64 mine = 0;
  int my_id = 0;
  if (i < K) {
    mine = ((u64)ordered_bits(scores[i]) << 32) | (u64)(0xFFFFFFFFu - (uint32_t)i);
    if (idxs != nullptr) my_id = (int)idxs[i];
  }
  int lt = 0, rank = 0;
  const int j_begin = blockIdx.y * NMS_RANK_JSPLIT, j_end = min(K, j_begin + NMS_RANK_JSPLIT);
  for (int j0 = j_begin; j0 < j_end; j0 += 256) {
    const int j = j0 + threadIdx.x;
    __syncthreads();
    t_key[threadIdx.x] = (j < j_end)
        ? (((u64)ordered_bits(scores[j]) << 32) | (u64)(0xFFFFFFFFu - (uint32_t)j)) : 0ull;
    t_id[threadIdx.x] = (j < j_end) ? (idxs != nullptr ? (int)idxs[j] : 0) : 0x7fffffff;
    __syncthreads();
    const int m = min(256, j_end - j0);
    for (int t = 0; t < m; ++t) {
      const int id = t_id[t];
      lt += (id < my_id);
      rank += (id == my_id) && (t_key[t] > mine);
    }
  }
  if (i < K) {
    if (lt) atomicAdd(lt_out + i, lt);
    if (rank) atomicAdd(rank_out + i, rank);
  }
}
__global__ void __launch_bounds__(256)
nms_id_scatter_kernel(const float* __restrict__ boxes, const float* __restrict__ scores,
                      const int64_t* __restrict__ idxs, int K, int num_ids,
                      const int32_t* __restrict__ lt_in, const int32_t* __restrict__ rank_in,
                      float4* __restrict__ sorted_boxes, u64* __restrict__ sorted_key,
                      int32_t* __restrict__ seg_start, int32_t* __restrict__ seg_count,
                      const int32_t* __restrict__ gate) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= K) return;
  if (gate != nullptr && *gate == 0) return;     // the fast id sort already did the job
  const int my_id = idxs != nullptr ? (int)idxs[i] : 0;
  if (my_id < 0 || my_id >= num_ids) return;
  const int lt = lt_in[i], rank = rank_in[i];
  const int pos = lt + rank;
  sorted_boxes[pos] = reinterpret_cast<const float4*>(boxes)[i];
  sorted_key[pos] = ((u64)ordered_bits(scores[i]) << 32) | (u64)(0xFFFFFFFFu - (uint32_t)i);
  if (rank == 0) seg_start[my_id] = lt;     // the best box of the id
  if (gate == nullptr) atomicAdd(seg_count + my_id, 1);   // (gated: counted by nms_id_count_kernel)
}

// Fast form of the same sort when every id's list fits one CTA's shared memory: the ids are
// counted (nms_id_count_kernel), then one CTA per id gathers its boxes' keys, sorts them with the
// shared-memory bitonic network and writes its segment at the prefix sum of the counts.  An id
// with more than NMS_IDSORT_CAP boxes raises `overflow`; the counting kernels above then redo
// the whole sort (they return at once when the flag is clear).  K = 20 000 / 5 ids: 183 us of
// rank counting -> ~20 us.
constexpr int NMS_IDSORT_CAP = 8192;
__global__ void __launch_bounds__(256)
nms_id_count_kernel(const int64_t* __restrict__ idxs, int K, int num_ids,
                    int32_t* __restrict__ seg_count) {
  extern __shared__ int s_hist[];
  for (int i = threadIdx.x; i < num_ids; i += blockDim.x) s_hist[i] = 0;
  __syncthreads();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < K; i += gridDim.x * blockDim.x) {
    const int id = (int)idxs[i];
    if (id >= 0 && id < num_ids) atomicAdd(&s_hist[id], 1);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < num_ids; i += blockDim.x)
    if (s_hist[i]) atomicAdd(seg_count + i, s_hist[i]);
}
// grid num_ids, block 1024, dynamic smem NMS_IDSORT_CAP * 8
__global__ void __launch_bounds__(1024)
nms_id_sort_kernel(const float* __restrict__ boxes, const float* __restrict__ scores,
                   const int64_t* __restrict__ idxs, int K, int num_ids,
                   const int32_t* __restrict__ seg_count, int32_t* __restrict__ seg_start,
                   float4* __restrict__ sorted_boxes, u64* __restrict__ sorted_key,
                   int32_t* __restrict__ overflow) {
  extern __shared__ __align__(16) u64 s_keys[];
  __shared__ int s_part[32];
  __shared__ int s_n;
  const int id = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  // start of this id's segment: sum of the counts of the smaller ids
  int part = 0;
  for (int j = tid; j < id; j += blockDim.x) part += seg_count[j];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if (lane == 0) s_part[wid] = part;
  if (tid == 0) s_n = 0;
  __syncthreads();
  int start = 0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) start += s_part[w];
  const int n = seg_count[id];
  if (tid == 0) seg_start[id] = start;
  if (n == 0) return;
  if (n > NMS_IDSORT_CAP) {
    if (tid == 0) atomicExch(overflow, 1);
    return;
  }
  // gather (arbitrary order: sorted next), warp-aggregated slots
  for (int i0 = 0; i0 < K; i0 += blockDim.x) {
    const int i = i0 + tid;
    const bool mine = i < K && (int)idxs[i] == id;
    const unsigned m = __ballot_sync(0xffffffffu, mine);
    if (m) {
      int base = 0;
      const int leader = __ffs(m) - 1;
      if (lane == leader) base = atomicAdd(&s_n, __popc(m));
      base = __shfl_sync(0xffffffffu, base, leader);
      if (mine)
        s_keys[base + __popc(m & ((1u << lane) - 1u))] =
            ((u64)ordered_bits(scores[i]) << 32) | (u64)(0xFFFFFFFFu - (uint32_t)i);
    }
  }
  __syncthreads();
  int np = 1;
  while (np < n) np <<= 1;
  for (int i = n + tid; i < np; i += blockDim.x) s_keys[i] = 0ull;
  bitonic_sort_desc_u64(s_keys, np);
  for (int r = tid; r < n; r += blockDim.x) {
    const u64 key = s_keys[r];
    const uint32_t src = 0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFull);
    sorted_key[start + r] = key;
    sorted_boxes[start + r] = reinterpret_cast<const float4*>(boxes)[src];
  }
}

__global__ void nms_finalize_kernel(const float* __restrict__ boxes,
                                    const float* __restrict__ scores,
                                    const int32_t* __restrict__ order,
                                    const int32_t* __restrict__ kept_pos,
                                    const int32_t* __restrict__ kept_count,
                                    int64_t* __restrict__ keep,
                                    float* __restrict__ dets,
                                    int32_t* __restrict__ num_keep) {
  const int n = kept_count[0];
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j == 0) num_keep[0] = n;
  if (j >= n) return;
  const int src = order[kept_pos[j]];
  keep[j] = src;
  if (dets != nullptr) {
    const float4 b = reinterpret_cast<const float4*>(boxes)[src];
    float* o = dets + (size_t)j * 5;
    o[0] = b.x; o[1] = b.y; o[2] = b.z; o[3] = b.w; o[4] = scores[src];
  }
}

struct DecodeArgs { float means[4], stds[4], max_ratio, max_h, max_w; int n, ncls; };
__global__ void delta2bbox_kernel(const __grid_constant__ DecodeArgs a,
                                  const float* __restrict__ rois,
                                  const float* __restrict__ deltas,
                                  float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n * a.ncls) return;
  const int r = i / a.ncls;
  const float4 q = reinterpret_cast<const float4*>(rois)[r];
  const float4 d = reinterpret_cast<const float4*>(deltas)[i];
  Box4 rb; rb.x1 = q.x; rb.y1 = q.y; rb.x2 = q.z; rb.y2 = q.w;
  const Box4 o = delta2bbox_one(rb, d.x, d.y, d.z, d.w, a.means, a.stds, a.max_ratio,
                                a.max_h >= 0.f, a.max_w, a.max_h);
  reinterpret_cast<float4*>(out)[i] = make_float4(o.x1, o.y1, o.x2, o.y2);
}

constexpr int NMS_MAX_IDS = 4096;   // ids of the segmented operator path
struct NmsWs {
  size_t sorted_boxes, sorted_key, order, count, maxc, mask, kept_pos, kept_count, seg, total;
};
static NmsWs nms_ws(int K);
static inline int32_t* kept_count_ids(char* ws, const NmsWs& w);
static NmsWs nms_ws(int K) {
  NmsWs w; size_t o = 0;
  const size_t W = (K + 63) / 64;
  w.sorted_boxes = o; o = align256(o + (size_t)K * 16);
  w.sorted_key = o;   o = align256(o + (size_t)K * 8);
  w.order = o;        o = align256(o + (size_t)K * 4);
  w.count = o;        o = align256(o + 4);
  w.maxc = o;         o = align256(o + 4);
  w.mask = o;         o = align256(o + (nms_use_fused(K) ? (size_t)K * 8 : (size_t)K * W * 8));
  w.kept_pos = o;     o = align256(o + (size_t)K * 4);
  w.kept_count = o;   o = align256(o + 4);
  w.seg = o;          o = align256(o + 3 * NMS_MAX_IDS * 4);   // seg_start | seg_count | kept_count
  w.total = o;
  return w;
}

static inline int32_t* kept_count_ids(char* ws, const NmsWs& w) {
  return (int32_t*)(ws + w.seg) + 2 * NMS_MAX_IDS;
}

// --------------------------------------------------------------------------
// RPN workspace layout
// --------------------------------------------------------------------------
struct RpnDerived {
  int Kc, keep_cap, W, n_total, key_stride, cand_cap;
  int level_n[BRCNN_MAX_LEVELS], level_k[BRCNN_MAX_LEVELS];
};
static int rpn_derive(const brcnn_rpn_params* p, RpnDerived* d) {
  if (!p || p->batch <= 0 || p->num_levels <= 0 || p->num_levels > BRCNN_MAX_LEVELS ||
      p->num_anchors <= 0 || p->num_anchors > BRCNN_MAX_ANCHORS || p->max_per_img <= 0)
    return BRCNN_ERR_ARG;
  int Kc = 0;
  long long n_total = 0, key_stride = 0;
  for (int l = 0; l < p->num_levels; ++l) {
    if (p->feat_h[l] <= 0 || p->feat_w[l] <= 0) return BRCNN_ERR_ARG;
    const long long n = (long long)p->feat_h[l] * p->feat_w[l] * p->num_anchors;
    if (n > 0x3fffffff) return BRCNN_ERR_UNSUPPORTED;
    n_total += n; key_stride += (n + 3) & ~3LL;
    d->level_n[l] = (int)n;
    d->level_k[l] = (p->nms_pre > 0 && n > p->nms_pre) ? p->nms_pre : (int)n;
    if (d->level_k[l] > Kc) Kc = d->level_k[l];
  }
  if (n_total > 0x3fffffff) return BRCNN_ERR_UNSUPPORTED;
  d->n_total = (int)n_total; d->key_stride = (int)key_stride;
  // smem candidate slots of the top-k kernel: >= 2x the largest k, >= 4096
  d->cand_cap = 2 * next_pow2(Kc) > 4096 ? 2 * next_pow2(Kc) : 4096;
  d->Kc = Kc;
  d->keep_cap = Kc < p->max_per_img ? Kc : p->max_per_img;
  d->W = (Kc + 63) / 64;
  return BRCNN_OK;
}

struct RpnWsInternal {
  brcnn_rpn_ws_layout pub;
  size_t mask, kept_key, keys, ghist, cand_n, cand_raw, zero_bytes;
};
static int rpn_ws(const brcnn_rpn_params* p, RpnWsInternal* w) {
  RpnDerived d;
  int rc = rpn_derive(p, &d);
  if (rc) return rc;
  const size_t S = (size_t)p->batch * p->num_levels;
  size_t o = 0;
  w->pub.cand_cap = d.Kc;
  w->pub.keep_cap = d.keep_cap;
  w->pub.cand_boxes = o; o = align256(o + S * d.Kc * 16);
  w->pub.cand_key = o;   o = align256(o + S * d.Kc * 8);
  w->pub.cand_valid = o; o = align256(o + S * d.Kc);
  w->pub.cand_count = o; o = align256(o + S * 4);
  w->pub.img_maxc = o;   o = align256(o + (size_t)p->batch * 4);
  w->ghist = o;          o = align256(o + S * RPN_BINS * 4);
  w->cand_n = o;         o = align256(o + S * 4);
  w->zero_bytes = o - (size_t)w->pub.img_maxc;  // img_maxc + ghist + cand_n, zeroed per call
  w->cand_raw = o;       o = align256(o + S * (size_t)d.cand_cap * 8);
  w->keys = o;           o = align256(o + (size_t)p->batch * d.key_stride * 4);
  w->pub.kept_pos = o;   o = align256(o + S * d.keep_cap * 4);
  w->pub.kept_count = o; o = align256(o + S * 4);
  w->kept_key = o;       o = align256(o + S * d.keep_cap * 8);
  w->mask = o;           o = align256(o + (nms_use_fused(d.keep_cap) ? 0 : S * d.Kc * d.W * 8));
  w->pub.total_bytes = (int64_t)o;
  return BRCNN_OK;
}

// --------------------------------------------------------------------------
// RCNN workspace layout
// --------------------------------------------------------------------------
struct RcnnWsInternal {
  brcnn_rcnn_ws_layout pub;
  size_t seg_boxes, mask, kept_key;
  int keep_cap, W;
};
static int rcnn_ws(const brcnn_rcnn_params* p, RcnnWsInternal* w) {
  if (!p || p->batch <= 0 || p->rois_per_img <= 0 || p->num_classes <= 0 ||
      p->max_per_img <= 0)
    return BRCNN_ERR_ARG;
  if (p->rois_per_img > 4096) return BRCNN_ERR_UNSUPPORTED;
  const size_t B = p->batch, Rc = p->rois_per_img, C = p->num_classes;
  const size_t nbox = p->reg_class_agnostic ? 1 : C;
  w->keep_cap = (int)(Rc < (size_t)p->max_per_img ? Rc : (size_t)p->max_per_img);
  w->W = (int)((Rc + 63) / 64);
  size_t o = 0;
  w->pub.scores = o;     o = align256(o + B * Rc * (C + 1) * 4);
  w->pub.bboxes = o;     o = align256(o + B * Rc * nbox * 16);
  w->pub.img_maxc = o;   o = align256(o + B * 4);
  w->pub.seg_count = o;  o = align256(o + B * C * 4);
  w->pub.seg_key = o;    o = align256(o + B * C * Rc * 8);
  w->seg_boxes = o;      o = align256(o + B * C * Rc * 16);
  w->pub.kept_pos = o;   o = align256(o + B * C * w->keep_cap * 4);
  w->pub.kept_count = o; o = align256(o + B * C * 4);
  w->kept_key = o;       o = align256(o + B * C * w->keep_cap * 8);
  w->mask = o;           o = align256(o + (nms_use_fused(w->keep_cap) ? 0 : B * C * Rc * w->W * 8));
  w->pub.total_bytes = (int64_t)o;
  return BRCNN_OK;
}

}  // namespace brcnn

using namespace brcnn;

// clusters of CS rpn_nms_image_kernel CTAs that can be resident at once on this device
// (cudaOccupancyMaxActiveClusters; cached per cluster size, shared-memory size and device)
template <int CS>
static int rpn_nms_max_clusters(int L, int max_out) {
  static std::mutex mu;
  static std::map<std::pair<int, int>, int> cache;
  const RpnNmsImageSmem lay = rpn_nms_image_smem(L, max_out, CS);
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> g(mu);
  auto it = cache.find({dev, lay.total});
  if (it != cache.end()) return it->second;
  int n = 0;
  if (lay.total <= 160 * 1024) {
    if (lay.total > 32 * 1024 &&
        ensure_dyn_smem((const void*)rpn_nms_image_kernel<CS>, lay.total) != cudaSuccess)
      n = 0;
    else {
      cudaLaunchConfig_t cfg;
      memset(&cfg, 0, sizeof(cfg));
      cfg.gridDim = dim3(CS * 64);
      cfg.blockDim = dim3(RNI_THREADS);
      cfg.dynamicSmemBytes = (size_t)lay.total;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      if (cudaOccupancyMaxActiveClusters(&n, rpn_nms_image_kernel<CS>, &cfg) != cudaSuccess) {
        (void)cudaGetLastError();
        n = 0;
      }
    }
  }
  cache[{dev, lay.total}] = n;
  return n;
}

extern "C" {

const char* brcnn_version(void) { return "libbrcnn 0.1 sm_100a"; }
int64_t brcnn_launch_count(void) { return g_launches.load(); }

// ------------------------------- RPN --------------------------------------
int brcnn_rpn_workspace_layout(const brcnn_rpn_params* p, brcnn_rpn_ws_layout* out) {
  RpnWsInternal w;
  int rc = rpn_ws(p, &w);
  if (rc) return rc;
  if (out) *out = w.pub;
  return BRCNN_OK;
}

size_t brcnn_rpn_workspace_bytes(const brcnn_rpn_params* p) {
  RpnWsInternal w;
  if (rpn_ws(p, &w)) return 0;
  return (size_t)w.pub.total_bytes;
}

int brcnn_rpn_get_bboxes(const brcnn_rpn_params* p,
                         const float* const* cls_scores_host,
                         const float* const* bbox_preds_host,
                         const float* const* iou_preds_host,
                         const float* base_anchors, const float* img_hw,
                         float* proposals, int32_t* num_proposals,
                         void* workspace, size_t workspace_bytes,
                         brcnn_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  RpnDerived d;
  int rc = rpn_derive(p, &d);
  if (rc) return rc;
  RpnWsInternal w;
  rc = rpn_ws(p, &w);
  if (rc) return rc;
  if (!cls_scores_host || !bbox_preds_host || !iou_preds_host || !base_anchors ||
      !img_hw || !proposals || !num_proposals || !workspace)
    return BRCNN_ERR_ARG;
  if (misaligned16(base_anchors) || misaligned16(workspace)) return BRCNN_ERR_ARG;
  if (workspace_bytes < (size_t)w.pub.total_bytes) return BRCNN_ERR_WORKSPACE;
  if (p->batch > 65535) return BRCNN_ERR_UNSUPPORTED;

  RpnArgs a;
  memset(&a, 0, sizeof(a));
  a.A = p->num_anchors; a.L = p->num_levels; a.B = p->batch; a.Kc = d.Kc;
  int idx_base = 0, chunk_base = 0, key_off = 0;
  for (int l = 0; l < p->num_levels; ++l) {
    RpnLevel& lv = a.lv[l];
    lv.cls = cls_scores_host[l]; lv.bbox = bbox_preds_host[l]; lv.iou = iou_preds_host[l];
    if (!lv.cls || !lv.bbox || !lv.iou) return BRCNN_ERR_ARG;
    lv.H = p->feat_h[l]; lv.W = p->feat_w[l];
    lv.stride_w = p->stride_w[l]; lv.stride_h = p->stride_h[l];
    lv.n = d.level_n[l]; lv.k = d.level_k[l]; lv.idx_base = idx_base;
    lv.chunk_base = chunk_base;
    lv.key_off = key_off;
    idx_base += lv.n;
    key_off += (lv.n + 3) & ~3;
    chunk_base += (lv.n + RPN_SCORE_CHUNK - 1) / RPN_SCORE_CHUNK;
  }
  a.key_stride = d.key_stride;
  a.cand_cap = d.cand_cap;
  for (int i = 0; i < 4; ++i) { a.means[i] = p->means[i]; a.stds[i] = p->stds[i]; }
  a.max_ratio = p->max_ratio; a.min_size = p->min_bbox_size;
  const size_t smem = (size_t)a.cand_cap * 8;
  if (smem > 200 * 1024) return BRCNN_ERR_UNSUPPORTED;

  char* ws = (char*)workspace;
  float4* cand_boxes = (float4*)(ws + w.pub.cand_boxes);
  u64* cand_key = (u64*)(ws + w.pub.cand_key);
  uint8_t* cand_valid = (uint8_t*)(ws + w.pub.cand_valid);
  int32_t* cand_count = (int32_t*)(ws + w.pub.cand_count);
  int* img_maxc = (int*)(ws + w.pub.img_maxc);
  int32_t* kept_pos = (int32_t*)(ws + w.pub.kept_pos);
  int32_t* kept_count = (int32_t*)(ws + w.pub.kept_count);
  u64* kept_key = (u64*)(ws + w.kept_key);
  u64* mask = (u64*)(ws + w.mask);
  uint32_t* keys = (uint32_t*)(ws + w.keys);
  uint32_t* ghist = (uint32_t*)(ws + w.ghist);
  int32_t* cand_n = (int32_t*)(ws + w.cand_n);
  u64* cand_raw = (u64*)(ws + w.cand_raw);
  const int S = p->batch * p->num_levels;

  // img_maxc and ghist are adjacent in the workspace: one memset
  cudaError_t e = cudaMemsetAsync(ws + w.pub.img_maxc, 0, w.zero_bytes, stream);
  if (e != cudaSuccess) return (int)e;
  {
    dim3 grid(chunk_base, p->batch);
    rpn_score_kernel<<<grid, RPN_SCORE_THREADS, 0, stream>>>(a, keys, ghist);
    g_launch_count_add(1);
    BRCNN_CUDA_CHECK_LAST();
    rpn_collect_kernel<<<grid, RPN_SCORE_THREADS, 0, stream>>>(a, keys, ghist, cand_raw, cand_n);
    g_launch_count_add(1);
    BRCNN_CUDA_CHECK_LAST();
  }
  if (smem > 48 * 1024) {
    e = ensure_dyn_smem((const void*)rpn_topk_decode_kernel, smem);
    if (e != cudaSuccess) return (int)e;
  }
  {
    // x = image: the heavy level-0 CTAs of every image are scheduled first
    dim3 grid(p->batch, p->num_levels);
    rpn_topk_decode_kernel<<<grid, RPN_TOPK_THREADS, smem, stream>>>(
        a, keys, ghist, base_anchors, img_hw, cand_boxes, cand_key, cand_valid, cand_count,
        img_maxc, cand_raw, cand_n);
    g_launch_count_add(1);
    BRCNN_CUDA_CHECK_LAST();
  }

  // per-image NMS in global score order with early stop (rpn_nms.cuh), one
  // cluster of 8 CTAs per image; the per-(image, level) segment kernels + merge
  // remain as the fallback (BRCNN_RPN_NMS=segments | image1 select them for A/B)
  {
    static const int mode = [] {
      const char* e = getenv("BRCNN_RPN_NMS");
      if (e && e[0] == 's') return 0;                       // segments + merge
      if (e && e[0] == 'i' && strchr(e, '1')) return 1;     // single CTA per image
      return 2;                                             // cluster
    }();
    // Cluster size: 8 CTAs per image unless the batch's clusters would not be co-resident (a
    // 512-thread, 123-register CTA fills an SM and a cluster must sit inside one GPC: B200 holds
    // fewer than 16 clusters of 8 at once, so batch 16 ran as two waves = twice the latency);
    // then 4 CTAs per image, all resident.
    int cs = 1;
    if (mode == 2) {
      cs = RNI_CLUSTER;
      if (p->batch > rpn_nms_max_clusters<RNI_CLUSTER>(p->num_levels, p->max_per_img) &&
          p->batch <= rpn_nms_max_clusters<4>(p->num_levels, p->max_per_img))
        cs = 4;
      static const int forced_cs = [] {
        const char* e = getenv("BRCNN_RNI_CS");       // developer knob: 1 | 2 | 4 | 8
        return e ? atoi(e) : 0;
      }();
      if (forced_cs == 1 || forced_cs == 2 || forced_cs == 4 || forced_cs == 8) cs = forced_cs;
    }
    const RpnNmsImageSmem lay = rpn_nms_image_smem(p->num_levels, p->max_per_img, cs);
    if (mode != 0 && lay.total <= 160 * 1024 && lay.kp <= 65535) {
      cudaLaunchConfig_t cfg;
      memset(&cfg, 0, sizeof(cfg));
      cfg.gridDim = dim3((unsigned)(p->batch * cs));
      cfg.blockDim = dim3(RNI_THREADS);
      cfg.dynamicSmemBytes = (size_t)lay.total;
      cfg.stream = stream;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = (unsigned)cs;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      const float* maxc_f = (const float*)img_maxc;
      // developer builds (-DBRCNN_DEBUG_TIMING) add per-phase cycle counters of CTA 0; the
      // shipped library never allocates or synchronises here
#ifdef BRCNN_DEBUG_TIMING
      static long long* dbg_buf = [] {
        long long* pbuf = nullptr;
        cudaMalloc(&pbuf, 64);
        return pbuf;
      }();
      long long* dbg = dbg_buf;
#else
      long long* dbg = nullptr;
#endif
#define BRCNN_LAUNCH_RNI(CSV)                                                                     \
      do {                                                                                        \
        if (lay.total > 32 * 1024) {                                                              \
          e = ensure_dyn_smem((const void*)rpn_nms_image_kernel<CSV>, lay.total);                 \
          if (e != cudaSuccess) return (int)e;                                                    \
        }                                                                                         \
        e = cudaLaunchKernelEx(&cfg, rpn_nms_image_kernel<CSV>, (const float4*)cand_boxes,        \
                               (const u64*)cand_key, (const uint8_t*)cand_valid,                  \
                               (const int32_t*)cand_count, (int)p->num_levels, (int)d.Kc,         \
                               p->iou_threshold, maxc_f, (int)p->max_per_img, proposals,          \
                               num_proposals, lay, dbg, (const int32_t*)nullptr, 0.0f,            \
                               (int64_t*)nullptr);                                                \
      } while (0)
      if (cs == RNI_CLUSTER) BRCNN_LAUNCH_RNI(RNI_CLUSTER);
      else if (cs == 4) BRCNN_LAUNCH_RNI(4);
      else if (cs == 2) BRCNN_LAUNCH_RNI(2);
      else BRCNN_LAUNCH_RNI(1);
#undef BRCNN_LAUNCH_RNI
      if (e != cudaSuccess) return (int)e;
#ifdef BRCNN_DEBUG_TIMING
      if (dbg != nullptr) {
        long long h[8];
        cudaStreamSynchronize(stream);
        cudaMemcpy(h, dbg, 64, cudaMemcpyDeviceToHost);
        fprintf(stderr, "[rpn_nms_image cs=%d] rounds=%lld cycles: windows=%lld rank=%lld pull+diag=%lld "
                "combine=%lld resolve=%lld | prologue=%lld total=%lld\n", cs, h[5], h[0], h[1], h[2],
                h[3], h[4], h[7], h[6]);
      }
#endif
      g_launch_count_add(1);
      BRCNN_CUDA_CHECK_LAST();
      return BRCNN_OK;
    }
  }
  rc = launch_nms_segments(cand_boxes, cand_valid, cand_count, S, d.Kc,
                           p->iou_threshold, 0.f, (const float*)img_maxc,
                           p->num_levels, mask, cand_key, kept_pos, kept_key,
                           kept_count, d.keep_cap, p->max_per_img, stream);
  if (rc) return rc;

  RpnMergeEpilogue ep{cand_boxes, proposals, d.Kc, p->max_per_img};
  return launch_nms_merge(kept_pos, kept_key, kept_count, p->batch, p->num_levels,
                          d.keep_cap, p->max_per_img, num_proposals, ep, stream);
}

int brcnn_delta2bbox(const float* rois, const float* deltas, int32_t n, int32_t ncls,
                     const float* means_host, const float* stds_host, float max_ratio,
                     float max_h, float max_w, float* out, brcnn_stream_t stream_) {
  if (n < 0 || ncls <= 0 || !means_host || !stds_host) return BRCNN_ERR_ARG;
  if (n == 0) return BRCNN_OK;
  if (!rois || !deltas || !out) return BRCNN_ERR_ARG;
  if (misaligned16(rois) || misaligned16(deltas) || misaligned16(out)) return BRCNN_ERR_ARG;
  DecodeArgs a;
  for (int i = 0; i < 4; ++i) { a.means[i] = means_host[i]; a.stds[i] = stds_host[i]; }
  a.max_ratio = max_ratio; a.max_h = max_h; a.max_w = max_w; a.n = n; a.ncls = ncls;
  const long long tot = (long long)n * ncls;
  if (tot > 0x7fffffff) return BRCNN_ERR_UNSUPPORTED;
  delta2bbox_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(
      a, rois, deltas, out);
  g_launch_count_add(1);
  BRCNN_CUDA_CHECK_LAST();
  return BRCNN_OK;
}

// ------------------------------- RPN loss ---------------------------------
static int rpn_loss_args(const brcnn_rpn_loss_params* p, RpnLossArgs* a) {
  if (!p || p->batch <= 0 || p->num_levels <= 0 || p->num_levels > BRCNN_MAX_LEVELS ||
      p->num_anchors <= 0 || p->num_anchors > BRCNN_MAX_ANCHORS || p->max_gts < 0)
    return BRCNN_ERR_ARG;
  if (p->max_gts > 1024 || p->batch > 65535) return BRCNN_ERR_UNSUPPORTED;
  memset(a, 0, sizeof(*a));
  a->A = p->num_anchors; a->L = p->num_levels; a->B = p->batch;
  a->Gmax = p->max_gts > 0 ? p->max_gts : 1;
  long long blocks = 0;
  for (int l = 0; l < p->num_levels; ++l) {
    if (p->feat_h[l] <= 0 || p->feat_w[l] <= 0 || p->stride_w[l] <= 0 || p->stride_h[l] <= 0)
      return BRCNN_ERR_ARG;
    const long long n = (long long)p->feat_h[l] * p->feat_w[l] * p->num_anchors;
    if (n > 0x3fffffff) return BRCNN_ERR_UNSUPPORTED;
    RpnLossLevel& lv = a->lv[l];
    lv.H = p->feat_h[l]; lv.W = p->feat_w[l];
    lv.stride_w = p->stride_w[l]; lv.stride_h = p->stride_h[l];
    lv.n = (int)n; lv.block_base = (int)blocks;
    blocks += (n + RL_THREADS - 1) / RL_THREADS;
  }
  if (blocks > 0x7fffffff) return BRCNN_ERR_UNSUPPORTED;
  a->blocks_per_img = (int)blocks;
  a->pos_iou_thr = p->pos_iou_thr; a->neg_iou_thr = p->neg_iou_thr;
  a->min_pos_iou = p->min_pos_iou; a->gamma = p->gamma;
  a->focal_gamma = p->focal_gamma; a->focal_alpha = p->focal_alpha;
  if (p->cls_loss_type != 0 && p->cls_loss_type != 1) return BRCNN_ERR_ARG;
  a->cls_loss_type = p->cls_loss_type;
  a->w_cls = p->loss_cls_weight; a->w_bbox = p->loss_bbox_weight;
  a->w_iou = p->loss_iou_weight; a->w_aug = p->loss_aug_weight;
  a->max_ratio = p->max_ratio;
  return BRCNN_OK;
}

size_t brcnn_rpn_loss_workspace_bytes(const brcnn_rpn_loss_params* p) {
  RpnLossArgs a;
  if (rpn_loss_args(p, &a)) return 0;
  return align256((size_t)a.B * a.Gmax * 4) +
         align256((size_t)a.B * a.blocks_per_img * RL_SUMS * 4);
}

int brcnn_rpn_loss_forward(const brcnn_rpn_loss_params* p, const float* const* cls_scores_host,
                           const float* const* bbox_preds_host, const float* const* iou_preds_host,
                           const float* base_anchors, const float* gt_boxes,
                           const int32_t* num_gt, const float* pad_hw, float* sums,
                           float* const* grad_cls_host, float* const* grad_bbox_host,
                           float* const* grad_iou_host, void* workspace, size_t workspace_bytes,
                           brcnn_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  RpnLossArgs a;
  int rc = rpn_loss_args(p, &a);
  if (rc) return rc;
  if (!cls_scores_host || !bbox_preds_host || !iou_preds_host || !base_anchors || !num_gt ||
      !pad_hw || !sums || !grad_cls_host || !grad_bbox_host || !grad_iou_host || !workspace)
    return BRCNN_ERR_ARG;
  if (p->max_gts > 0 && (!gt_boxes || misaligned16(gt_boxes))) return BRCNN_ERR_ARG;
  if (misaligned16(base_anchors) || misaligned16(workspace)) return BRCNN_ERR_ARG;
  if (workspace_bytes < brcnn_rpn_loss_workspace_bytes(p)) return BRCNN_ERR_WORKSPACE;
  for (int l = 0; l < a.L; ++l) {
    RpnLossLevel& lv = a.lv[l];
    lv.cls = cls_scores_host[l]; lv.bbox = bbox_preds_host[l]; lv.iou = iou_preds_host[l];
    lv.g_cls = grad_cls_host[l]; lv.g_bbox = grad_bbox_host[l]; lv.g_iou = grad_iou_host[l];
    if (!lv.cls || !lv.bbox || !lv.iou || !lv.g_cls || !lv.g_bbox || !lv.g_iou)
      return BRCNN_ERR_ARG;
  }
  char* ws = (char*)workspace;
  unsigned int* gt_max = (unsigned int*)ws;
  float* partials = (float*)(ws + align256((size_t)a.B * a.Gmax * 4));
  cudaError_t e = cudaMemsetAsync(gt_max, 0, (size_t)a.B * a.Gmax * 4, stream);
  if (e != cudaSuccess) return (int)e;
  const size_t smem = (size_t)a.Gmax * 20;
  dim3 grid(a.blocks_per_img, a.B);
  if (p->max_gts > 0) {
    rpn_loss_gtmax_kernel<<<grid, RL_THREADS, smem, stream>>>(a, base_anchors, gt_boxes, num_gt,
                                                             pad_hw, gt_max);
    g_launch_count_add(1);
    BRCNN_CUDA_CHECK_LAST();
  }
  rpn_loss_main_kernel<<<grid, RL_THREADS, smem, stream>>>(a, base_anchors, gt_boxes, num_gt,
                                                          pad_hw, gt_max, partials);
  g_launch_count_add(1);
  BRCNN_CUDA_CHECK_LAST();
  rpn_loss_reduce_kernel<<<1, RL_THREADS, 0, stream>>>(a, partials, sums);
  g_launch_count_add(1);
  BRCNN_CUDA_CHECK_LAST();
  return BRCNN_OK;
}

int brcnn_rpn_loss_scale(const brcnn_rpn_loss_params* p, const float* const* raw_cls_host,
                         const float* const* raw_bbox_host, const float* const* raw_iou_host,
                         const float* scale, float* const* out_cls_host,
                         float* const* out_bbox_host, float* const* out_iou_host,
                         brcnn_stream_t stream_) {
  RpnLossArgs a;
  int rc = rpn_loss_args(p, &a);
  if (rc) return rc;
  if (!raw_cls_host || !raw_bbox_host || !raw_iou_host || !scale || !out_cls_host ||
      !out_bbox_host || !out_iou_host)
    return BRCNN_ERR_ARG;
  RpnLossScaleArgs s;
  memset(&s, 0, sizeof(s));
  s.L = a.L;
  long long base = 0;
  for (int kind = 0; kind < 3; ++kind) {
    const float* const* raw = kind == 0 ? raw_cls_host : (kind == 1 ? raw_bbox_host : raw_iou_host);
    float* const* out = kind == 0 ? out_cls_host : (kind == 1 ? out_bbox_host : out_iou_host);
    for (int l = 0; l < a.L; ++l) {
      if (!raw[l] || !out[l]) return BRCNN_ERR_ARG;
      const long long n = (long long)a.lv[l].n * a.B * (kind == 1 ? 4 : 1);
      if (n > 0x7fffffff) return BRCNN_ERR_UNSUPPORTED;
      s.raw[kind][l] = raw[l]; s.out[kind][l] = out[l]; s.n[kind][l] = (int)n;
      s.block_base[kind * a.L + l] = (int)base;
      base += (n + RL_THREADS * 4 - 1) / (RL_THREADS * 4);
    }
  }
  s.block_base[3 * a.L] = (int)base;
  if (base > 0x7fffffff) return BRCNN_ERR_UNSUPPORTED;
  rpn_loss_scale_kernel<<<(unsigned)base, RL_THREADS, 0, (cudaStream_t)stream_>>>(s, scale);
  g_launch_count_add(1);
  BRCNN_CUDA_CHECK_LAST();
  return BRCNN_OK;
}

// ------------------------------- NMS --------------------------------------
size_t brcnn_nms_workspace_bytes(int32_t num_boxes) {
  if (num_boxes <= 0) return 256;
  return nms_ws(num_boxes).total;
}

// device-side tail of the clustered path: dets rows from the keep list
__global__ void nms_gather_dets_kernel(const float* __restrict__ boxes,
                                       const float* __restrict__ scores,
                                       const int64_t* __restrict__ keep,
                                       const int32_t* __restrict__ num_keep,
                                       float* __restrict__ dets) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= num_keep[0]) return;
  const int64_t src = keep[j];
  const float4 b = reinterpret_cast<const float4*>(boxes)[src];
  float* o = dets + (size_t)j * 5;
  o[0] = b.x; o[1] = b.y; o[2] = b.z; o[3] = b.w; o[4] = scores[src];
}

// epilogue of the many-ids operator path: the key's low half is ~original index
struct OpMergeEpilogue {
  static constexpr bool kNeedsPos = false;
  int64_t* keep;
  __device__ void operator()(int, int rank, int, int, u64 key) const {
    keep[rank] = (int64_t)(0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFull));
  }
  __device__ void pad(int, int) const {}
};

// few ids (<= BRCNN_MAX_LEVELS): clustered global-order walk when an early stop can pay off,
// independent per-id CTAs otherwise
static inline bool nms_prefer_per_id(bool has_ids, int num_ids, int max_keep, int K) {
  return has_ids && num_ids >= 2 && max_keep >= K / 2 && K >= 2048;
}

// (id asc, score desc, index asc) counting sort; lt / rank scratch = order + kept_pos arrays
static int launch_id_sort(const float* boxes, const float* scores, const int64_t* idxs, int K,
                          int num_ids, int32_t* lt, int32_t* rank, float4* sboxes, u64* skey,
                          int32_t* seg_start, int32_t* seg_count, int32_t* gate,
                          cudaStream_t stream) {
  cudaError_t e = cudaMemsetAsync(lt, 0, (size_t)K * 4, stream);
  if (e != cudaSuccess) return (int)e;
  e = cudaMemsetAsync(rank, 0, (size_t)K * 4, stream);
  if (e != cudaSuccess) return (int)e;
  // fast path: count the ids, one sorting CTA per id (seg_count must be zero on entry); the
  // rank-counting kernels below only run when an id's list overflowed a CTA's shared memory
  // (below ~8k boxes the K x K counting costs less than the two extra launches)
  const bool fast = idxs != nullptr && gate != nullptr && num_ids >= 1 && num_ids <= NMS_MAX_IDS &&
                    K >= 8192;
  if (fast) {
    e = cudaMemsetAsync(gate, 0, 4, stream);
    if (e != cudaSuccess) return (int)e;
    int cb = (K + 256 * 8 - 1) / (256 * 8);
    if (cb > 2 * sm_count()) cb = 2 * sm_count();
    nms_id_count_kernel<<<cb, 256, (size_t)num_ids * 4, stream>>>(idxs, K, num_ids, seg_count);
    g_launch_count_add(1);
    BRCNN_CUDA_CHECK_LAST();
    const size_t sm = (size_t)NMS_IDSORT_CAP * 8;
    e = ensure_dyn_smem((const void*)nms_id_sort_kernel, sm);
    if (e != cudaSuccess) return (int)e;
    nms_id_sort_kernel<<<num_ids, 1024, sm, stream>>>(boxes, scores, idxs, K, num_ids, seg_count,
                                                      seg_start, sboxes, skey, gate);
    g_launch_count_add(1);
    BRCNN_CUDA_CHECK_LAST();
  }
  dim3 grid((K + 255) / 256, (K + NMS_RANK_JSPLIT - 1) / NMS_RANK_JSPLIT);
  nms_id_rank_count_kernel<<<grid, 256, 0, stream>>>(scores, idxs, K, lt, rank,
                                                    fast ? gate : nullptr);
  g_launch_count_add(1);
  BRCNN_CUDA_CHECK_LAST();
  nms_id_scatter_kernel<<<(K + 255) / 256, 256, 0, stream>>>(boxes, scores, idxs, K, num_ids, lt,
                                                            rank, sboxes, skey, seg_start,
                                                            seg_count, fast ? gate : nullptr);
  g_launch_count_add(1);
  BRCNN_CUDA_CHECK_LAST();
  return BRCNN_OK;
}

int brcnn_batched_nms(const float* boxes, const float* scores, const int64_t* idxs,
                      int32_t K, int32_t num_ids, float iou_threshold, int32_t offset,
                      int32_t max_num, int64_t* keep, float* dets, int32_t* num_keep,
                      void* workspace, size_t workspace_bytes,
                      brcnn_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  const int max_keep = (max_num > 0 && max_num < K) ? max_num : K;   // early stop of the sweeps
  if (K < 0 || !num_keep || (offset != 0 && offset != 1)) return BRCNN_ERR_ARG;
  if (K == 0) {
    cudaError_t e = cudaMemsetAsync(num_keep, 0, 4, stream);
    return e == cudaSuccess ? BRCNN_OK : (int)e;
  }
  if (!boxes || !scores || !keep || !workspace) return BRCNN_ERR_ARG;
  if (misaligned16(boxes) || misaligned16(workspace)) return BRCNN_ERR_ARG;
  if (K > 393216) return BRCNN_ERR_UNSUPPORTED;
  NmsWs w = nms_ws(K);
  if (workspace_bytes < w.total) return BRCNN_ERR_WORKSPACE;
  char* ws = (char*)workspace;
  float4* sboxes = (float4*)(ws + w.sorted_boxes);
  u64* skey = (u64*)(ws + w.sorted_key);
  int32_t* order = (int32_t*)(ws + w.order);
  int32_t* count = (int32_t*)(ws + w.count);
  float* maxc = (float*)(ws + w.maxc);
  u64* mask = (u64*)(ws + w.mask);
  int32_t* kept_pos = (int32_t*)(ws + w.kept_pos);
  int32_t* kept_count = (int32_t*)(ws + w.kept_count);
  if (idxs != nullptr) {
    unsigned int* maxbits = (unsigned int*)(ws + w.count);   // free until the fallback path
    cudaError_t em = cudaMemsetAsync(maxbits, 0, 4, stream);
    if (em != cudaSuccess) return (int)em;
    int mb = (4 * K + 256 * 16 - 1) / (256 * 16);
    if (mb > 512) mb = 512;
    nms_maxcoord_kernel<<<mb, 256, 0, stream>>>(boxes, K, maxbits);
    nms_maxcoord_finish_kernel<<<1, 1, 0, stream>>>(maxbits, maxc);
    g_launch_count_add(2);
    BRCNN_CUDA_CHECK_LAST();
  }
  // ---- clustered path: the ids are <= BRCNN_MAX_LEVELS sorted lists walked in global
  // score order by a cluster of 8 CTAs (rpn_nms.cuh); needs the caller's bound on the
  // id range (num_ids; 1 when idxs == NULL) and a kept list that fits shared memory
  {
    static const bool force_old = [] {
      const char* e = getenv("BRCNN_NMS_OP");
      return e && e[0] == 'o';
    }();
    const int L = (idxs == nullptr) ? 1 : num_ids;
    const RpnNmsImageSmem lay = rpn_nms_image_smem(L > 0 ? L : 1, max_keep, RNI_CLUSTER);
    // without an early stop the clustered walk pays one cluster round per 64 boxes of the WHOLE
    // input; several ids then run faster as independent per-id CTAs (the path below:
    // K = 15 150 / 5 ids 1.91 -> 1.51 ms, K = 4 693 / 5 ids 0.58 -> 0.42 ms)
    const bool prefer_perid = nms_prefer_per_id(idxs != nullptr, num_ids, max_keep, K);
    if (!force_old && !prefer_perid && L >= 1 && L <= BRCNN_MAX_LEVELS && lay.total <= 160 * 1024 &&
        lay.kp <= 65535) {
      int32_t* seg_start = (int32_t*)(ws + w.seg);
      int32_t* seg_count = seg_start + BRCNN_MAX_LEVELS;
      cudaError_t e = cudaMemsetAsync(seg_start, 0, 2 * BRCNN_MAX_LEVELS * 4, stream);
      if (e != cudaSuccess) return (int)e;
      const int rcs = launch_id_sort(boxes, scores, idxs, K, L, order, kept_pos, sboxes, skey,
                                     seg_start, seg_count, (int32_t*)(ws + w.count) + 1, stream);
      if (rcs) return rcs;
      if (lay.total > 32 * 1024) {
        e = ensure_dyn_smem((const void*)rpn_nms_image_kernel<RNI_CLUSTER>, lay.total);
        if (e != cudaSuccess) return (int)e;
      }
      cudaLaunchConfig_t cfg;
      memset(&cfg, 0, sizeof(cfg));
      cfg.gridDim = dim3((unsigned)RNI_CLUSTER);
      cfg.blockDim = dim3(RNI_THREADS);
      cfg.dynamicSmemBytes = (size_t)lay.total;
      cfg.stream = stream;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = (unsigned)RNI_CLUSTER;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      e = cudaLaunchKernelEx(&cfg, rpn_nms_image_kernel<RNI_CLUSTER>, (const float4*)sboxes,
                             (const u64*)skey, (const uint8_t*)nullptr,
                             (const int32_t*)seg_count, L, (int)K, iou_threshold,
                             idxs != nullptr ? (const float*)maxc : (const float*)nullptr,
                             (int)max_keep, (float*)nullptr, num_keep, lay, (long long*)nullptr,
                             (const int32_t*)seg_start, (float)offset, keep);
      if (e != cudaSuccess) return (int)e;
      g_launch_count_add(1);
      BRCNN_CUDA_CHECK_LAST();
      if (dets != nullptr) {
        nms_gather_dets_kernel<<<(K + 255) / 256, 256, 0, stream>>>(boxes, scores, keep,
                                                                    num_keep, dets);
        g_launch_count_add(1);
        BRCNN_CUDA_CHECK_LAST();
      }
      return BRCNN_OK;
    }
  }
  // ---- many ids (class-wise NMS at operator level): one fused-NMS CTA per id on the
  // id-sorted boxes, kept lists merged by an in-smem sort.  The kept list of a segment
  // lives in shared memory, so every segment must fit it: K <= 8192.
  {
    static const bool force_old2 = [] {
      const char* e = getenv("BRCNN_NMS_OP");
      return e && e[0] == 'o';
    }();
    // the kept list of a segment lives in shared memory up to 8192 boxes (160 KB); a longer one
    // (one id holding most of a huge input) spills to global memory inside the kernel
    int keep_pad = nms_keep_pad(K, K);
    if (keep_pad > 8192) keep_pad = 8192;
    int np2 = 1;
    while (np2 < K) np2 <<= 1;
    if (!force_old2 && idxs != nullptr && num_ids <= 1024 &&
        (num_ids > BRCNN_MAX_LEVELS || nms_prefer_per_id(true, num_ids, max_keep, K))) {
      int32_t* seg_start = (int32_t*)(ws + w.seg);
      int32_t* seg_count = seg_start + NMS_MAX_IDS;
      cudaError_t e = cudaMemsetAsync(seg_start, 0, 2 * NMS_MAX_IDS * 4, stream);
      if (e != cudaSuccess) return (int)e;
      const int rcs = launch_id_sort(boxes, scores, idxs, K, num_ids, order,
                                     (int32_t*)(ws + w.mask), sboxes, skey, seg_start, seg_count,
                                     (int32_t*)(ws + w.count) + 1, stream);
      if (rcs) return rcs;
      const size_t smem = (size_t)keep_pad * 20;
      if (smem > 48 * 1024) {
        e = ensure_dyn_smem((const void*)nms_fused_kernel, smem);
        if (e != cudaSuccess) return (int)e;
      }
      u64* kept_key = (u64*)(ws + w.mask);   // K u64, the bitmask area is unused on this path
      nms_fused_kernel<<<num_ids, NMS_FUSED_THREADS, smem, stream>>>(
          sboxes, nullptr, seg_count, K, iou_threshold, (float)offset, maxc, num_ids, skey,
          kept_pos, kept_key, kept_count_ids(ws, w), K, max_keep, keep_pad, seg_start);
      g_launch_count_add(1);
      BRCNN_CUDA_CHECK_LAST();
      const size_t sm = (size_t)np2 * 8;
      if (sm <= 160 * 1024) {
        // all kept keys fit one CTA's shared memory: in-smem bitonic sort
        if (sm > 32 * 1024) {
          e = ensure_dyn_smem((const void*)nms_merge_sort_kernel<OpMergeEpilogue>, sm);
          if (e != cudaSuccess) return (int)e;
        }
        OpMergeEpilogue ep{keep};
        nms_merge_sort_kernel<OpMergeEpilogue><<<1, 1024, sm, stream>>>(
            kept_key, kept_count_ids(ws, w), num_ids, K, max_keep, np2, num_keep, ep, seg_start);
      } else {
        // COCO-scale inputs (K = 20 480, 80 ids): rank counting over the whole GPU
        dim3 mgrid((unsigned)((K + 255) / 256), (unsigned)num_ids);
        nms_op_merge_rank_kernel<<<mgrid, 256, 0, stream>>>(kept_key, kept_count_ids(ws, w),
                                                           seg_start, num_ids, max_keep, keep,
                                                           num_keep);
      }
      g_launch_count_add(1);
      BRCNN_CUDA_CHECK_LAST();
      if (dets != nullptr) {
        nms_gather_dets_kernel<<<(K + 255) / 256, 256, 0, stream>>>(boxes, scores, keep,
                                                                    num_keep, dets);
        g_launch_count_add(1);
        BRCNN_CUDA_CHECK_LAST();
      }
      return BRCNN_OK;
    }
  }
  nms_rank_scatter_kernel<<<(K + 255) / 256, 256, 0, stream>>>(
      boxes, scores, idxs, K, maxc, sboxes, skey, order, count);
  g_launch_count_add(1);
  BRCNN_CUDA_CHECK_LAST();
  int rc = launch_nms_segments(sboxes, nullptr, count, 1, K, iou_threshold,
                               (float)offset, nullptr, 1, mask, nullptr, kept_pos,
                               nullptr, kept_count, K, max_keep, stream);
  if (rc) return rc;
  nms_finalize_kernel<<<(K + 255) / 256, 256, 0, stream>>>(
      boxes, scores, order, kept_pos, kept_count, keep, dets, num_keep);
  g_launch_count_add(1);
  BRCNN_CUDA_CHECK_LAST();
  return BRCNN_OK;
}

// ------------------------------- RoI --------------------------------------
int brcnn_map_roi_levels(const float* rois, int32_t R, float finest_scale,
                         int32_t num_levels, int64_t* target_lvls,
                         brcnn_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (R < 0 || num_levels <= 0) return BRCNN_ERR_ARG;
  if (R == 0) return BRCNN_OK;
  if (!rois || !target_lvls) return BRCNN_ERR_ARG;
  map_roi_levels_kernel<<<(R + 255) / 256, 256, 0, stream>>>(rois, R, finest_scale,
                                                             num_levels, target_lvls);
  g_launch_count_add(1);
  BRCNN_CUDA_CHECK_LAST();
  return BRCNN_OK;
}

int brcnn_bbox2roi_padded(const float* proposals, const int32_t* num_proposals,
                          int32_t batch, int32_t cap, float* rois, float* prior,
                          brcnn_stream_t stream_) {
  if (batch < 0 || cap < 0) return BRCNN_ERR_ARG;
  if (batch == 0 || cap == 0) return BRCNN_OK;
  if (!proposals || !num_proposals || !rois) return BRCNN_ERR_ARG;
  const long long n = (long long)batch * cap;
  if (n > 0x7fffffff) return BRCNN_ERR_UNSUPPORTED;
  bbox2roi_padded_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(
      proposals, num_proposals, batch, cap, rois, prior);
  g_launch_count_add(1);
  BRCNN_CUDA_CHECK_LAST();
  return BRCNN_OK;
}

static int roi_args_from(const brcnn_roi_params* p, RoiArgs* a) {
  if (!p || p->batch <= 0 || p->channels <= 0 || (p->channels & 3) ||
      p->num_levels <= 0 || p->num_levels > BRCNN_MAX_LEVELS || p->pooled_h <= 0 ||
      p->pooled_w <= 0 || (p->out_layout != 0 && p->out_layout != 1))
    return BRCNN_ERR_ARG;
  memset(a, 0, sizeof(*a));
  a->B = p->batch; a->C = p->channels; a->L = p->num_levels;
  a->PH = p->pooled_h; a->PW = p->pooled_w;
  a->sampling_ratio = p->sampling_ratio; a->aligned = p->aligned;
  a->finest_scale = p->finest_scale;
  for (int l = 0; l < p->num_levels; ++l) {
    if (p->feat_h[l] <= 0 || p->feat_w[l] <= 0) return BRCNN_ERR_ARG;
    a->H[l] = p->feat_h[l]; a->W[l] = p->feat_w[l]; a->scale[l] = p->spatial_scale[l];
  }
  return BRCNN_OK;
}

size_t brcnn_roi_extract_forward_workspace_bytes(const brcnn_roi_params* p) {
  (void)p;
  return 256;   // 2 counters per channel chunk (<= 32 chunks)
}

int brcnn_roi_extract_forward(const brcnn_roi_params* p,
                              const float* const* feats_nhwc_host,
                              const float* rois, int32_t R, float* out,
                              int32_t* roi_levels, void* workspace, size_t workspace_bytes,
                              brcnn_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (workspace != nullptr && (workspace_bytes < 256 || ((uintptr_t)workspace & 3)))
    return BRCNN_ERR_WORKSPACE;
  RoiArgs a;
  int rc = roi_args_from(p, &a);
  if (rc) return rc;
  if (R < 0) return BRCNN_ERR_ARG;
  if (R == 0) return BRCNN_OK;
  if (!feats_nhwc_host || !rois || !out) return BRCNN_ERR_ARG;
  for (int l = 0; l < a.L; ++l) {
    if (!feats_nhwc_host[l] || misaligned16(feats_nhwc_host[l])) return BRCNN_ERR_ARG;
    a.feat[l] = feats_nhwc_host[l];
  }
  const int nbins = a.PH * a.PW;
  if (a.PH > ROI_MAXP || a.PW > ROI_MAXP) return BRCNN_ERR_UNSUPPORTED;
  for (int l = 0; l < a.L; ++l) {
    if (a.H[l] > a.max_h) a.max_h = a.H[l];
    if (a.W[l] > a.max_w) a.max_w = a.W[l];
  }
  // ---- v3: persistent cross-RoI pipelined kernel, (R,ph,pw,C) output (roi_align_fwd3.cuh) ----
  if (p->out_layout == 1) {
    if (a.PH > RT_P || a.PW > RT_P || misaligned16(out)) return BRCNN_ERR_UNSUPPORTED;
    const int chunk = a.C < 4 * RT_SLAB_Q ? a.C : 4 * RT_SLAB_Q;
    Roi3Smem lay;
    lay.tab_floats = (R3_DESC + a.max_h * 8 + 8 * a.max_w + 31) & ~31;
    const size_t tables = (size_t)R3_TABS * lay.tab_floats * 4;
    const size_t budget = 110 * 1024;           // two CTAs per SM
    // a slot holds up to 16 footprint pixels of the channel chunk (wider rows: x-chunk passes)
    static const int slot_px = [] {
      const char* e = getenv("BRCNN_R3_SLOT_PX");   // tuning knob (developer)
      const int v = e ? atoi(e) : 0;
      return v >= 4 && v <= 64 ? v : 16;
    }();
    lay.slot_bytes = (slot_px * chunk * 4 + 127) & ~127;
    if (tables + (size_t)3 * 128 >= budget) return BRCNN_ERR_UNSUPPORTED;
    while ((size_t)3 * lay.slot_bytes + tables > budget && lay.slot_bytes > 128)
      lay.slot_bytes = ((lay.slot_bytes / 2) + 127) & ~127;
    if (lay.slot_bytes < chunk * 4 || (size_t)3 * lay.slot_bytes + tables > budget)
      return BRCNN_ERR_UNSUPPORTED;
    lay.ns = (int)((budget - tables) / lay.slot_bytes);
    if (lay.ns > R3_MAX_STAGES) lay.ns = R3_MAX_STAGES;
    lay.total = (int)((size_t)lay.ns * lay.slot_bytes + tables);
    a.chunk_c = chunk;
    static const bool split = [] {
      const char* e = getenv("BRCNN_R3_SPLIT");     // tuning knob (developer)
      return e ? atoi(e) != 0 : true;
    }();
    cudaError_t e = split
        ? ensure_dyn_smem((const void*)roi_align_fwd3_kernel<true>, lay.total, true)
        : ensure_dyn_smem((const void*)roi_align_fwd3_kernel<false>, lay.total, true);
    if (e != cudaSuccess) return (int)e;
    const int slots = 2 * sm_count();
    dim3 grid(R < slots ? R : slots, (a.C + chunk - 1) / chunk);
    unsigned int* sched = grid.y <= 32 ? (unsigned int*)workspace : nullptr;
#ifdef BRCNN_DEBUG_TIMING
    static unsigned long long* fdbg = [] {
      unsigned long long* pbuf = nullptr;
      cudaMalloc(&pbuf, 24 * 8);
      return pbuf;
    }();
    cudaMemsetAsync(fdbg, 0, 24 * 8, stream);
    if (split)
      roi_align_fwd3_kernel<true><<<grid, R3_THREADS, lay.total, stream>>>(
          a, rois, R, out, roi_levels, lay, sched, fdbg);
    else
      roi_align_fwd3_kernel<false><<<grid, RT_THREADS, lay.total, stream>>>(
          a, rois, R, out, roi_levels, lay, sched, fdbg);
    {
      unsigned long long h[24];
      cudaStreamSynchronize(stream);
      cudaMemcpy(h, fdbg, sizeof(h), cudaMemcpyDeviceToHost);
      if (h[0] && h[8]) {
        if (!h[16]) h[16] = 1;
        fprintf(stderr, "[fwd3] ctas=%llu rois/cta=%.1f | consumer0: total=%llu tab-wait=%llu "
                "row-wait=%llu stores=%llu | issuer: total=%llu tab-wait=%llu slot-wait=%llu "
                "publish=%llu | publisher: total=%llu tables=%llu buf-wait=%llu\n",
                h[0], (double)h[5] / h[0], h[1] / h[0], h[2] / h[0], h[3] / h[0], h[4] / h[0],
                h[9] / h[8], h[10] / h[8], h[11] / h[8], h[12] / h[8], h[17] / h[16],
                h[18] / h[16], h[19] / h[16]);
      }
    }
#else
    if (split)
      roi_align_fwd3_kernel<true><<<grid, R3_THREADS, lay.total, stream>>>(
          a, rois, R, out, roi_levels, lay, sched);
    else
      roi_align_fwd3_kernel<false><<<grid, RT_THREADS, lay.total, stream>>>(
          a, rois, R, out, roi_levels, lay, sched);
#endif
    g_launch_count_add(1);
    BRCNN_CUDA_CHECK_LAST();
    return BRCNN_OK;
  }
  // ---- v2: TMA row-streaming kernel (roi_align_tma.cuh) ----
  {
    static const bool force_v1 = [] {
      const char* e = getenv("BRCNN_ROI_FWD");
      return e && e[0] == 'v' && e[1] == '1';
    }();
    const int chunk = a.C < 4 * RT_SLAB_Q ? a.C : 4 * RT_SLAB_Q;
    const size_t tables = ((size_t)a.max_h * 8 + (size_t)8 * a.max_w) * 4;
    const size_t budget = 110 * 1024;  // two CTAs per SM
    size_t ring = 96 * 1024;
    if (tables + ring > budget) ring = tables < budget ? ((budget - tables) & ~(size_t)127) : 0;
    const size_t need = (size_t)chunk * nbins * 4 > (size_t)3 * 128 * ((chunk * 4 + 127) / 128)
                            ? (size_t)chunk * nbins * 4
                            : (size_t)3 * 128 * ((chunk * 4 + 127) / 128);
    if (!force_v1 && a.PH <= RT_P && a.PW <= RT_P && ring >= need && !misaligned16(out)) {
      a.chunk_c = chunk;
      const size_t smem = ring + tables;
      cudaError_t e = ensure_dyn_smem((const void*)roi_align_fwd_tma_kernel, smem, true);
      if (e != cudaSuccess) return (int)e;
      dim3 grid(R, (a.C + chunk - 1) / chunk);
      roi_align_fwd_tma_kernel<<<grid, RT_THREADS, smem, stream>>>(a, rois, R, out, roi_levels,
                                                                   (int)ring);
      g_launch_count_add(1);
      BRCNN_CUDA_CHECK_LAST();
      return BRCNN_OK;
    }
  }
  // channel slab per CTA: <= 256 channels (64 quads x 4 bin-row slots) and
  // <= ~56 KB of staging, multiple of 32
  int chunk = a.C < 256 ? a.C : 256;
  const int max_chunk = ((56 * 1024) / (nbins * 4)) & ~31;
  if (max_chunk < 32) return BRCNN_ERR_UNSUPPORTED;
  if (chunk > max_chunk) chunk = max_chunk;
  a.chunk_c = chunk;
  const int nchunks = (a.C + chunk - 1) / chunk;
  const size_t smem = ((size_t)chunk * nbins + (size_t)a.PH * a.max_h +
                       (size_t)a.PW * a.max_w) * 4;
  if (smem > 200 * 1024) return BRCNN_ERR_UNSUPPORTED;
  dim3 grid(R, nchunks);
  cudaError_t e;
  if (a.PW <= 7) {
    // 3 CTAs x ~58 KB per SM: ask for the large shared-memory carveout
    e = ensure_dyn_smem((const void*)roi_align_fwd_kernel<7>, smem, true);
    if (e != cudaSuccess) return (int)e;
    roi_align_fwd_kernel<7><<<grid, ROI_THREADS, smem, stream>>>(a, rois, R, out, roi_levels);
  } else {
    e = ensure_dyn_smem((const void*)roi_align_fwd_kernel<ROI_MAXP>, smem);
    if (e != cudaSuccess) return (int)e;
    roi_align_fwd_kernel<ROI_MAXP><<<grid, ROI_THREADS, smem, stream>>>(a, rois, R, out,
                                                                         roi_levels);
  }
  g_launch_count_add(1);
  BRCNN_CUDA_CHECK_LAST();
  return BRCNN_OK;
}

// 1 = v1 (pooled sizes > 7), 2 = v2 gather (whole-R key scan; BRCNN_ROI_BWD=v2, or buckets too
// large), 3 = per-tile lists + roi_bwd_gather5_kernel (the default)
static int roi_bwd_version(const RoiArgs& a, int R) {
  static const int forced = [] {
    const char* e = getenv("BRCNN_ROI_BWD");
    if (e && e[0] == 'v' && e[1] == '1') return 1;
    if (e && e[0] == 'v' && e[1] == '2') return 2;
    return 0;
  }();
  const bool small = a.PH <= B2_P && a.PW <= B2_P && (long long)a.B * a.L < 65535;
  if (forced == 1 || !small) return 1;
  // the (image, level) buckets of the tiled path hold R entries each
  const bool buckets_ok = (size_t)a.B * a.L * (size_t)(R > 0 ? R : 1) * 20 <= ((size_t)256 << 20);
  if (forced == 2 || !buckets_ok) return 2;
  return 3;
}

size_t brcnn_roi_extract_backward_workspace_bytes(const brcnn_roi_params* p,
                                                  int32_t R) {
  RoiArgs a;
  if (roi_args_from(p, &a) || R < 0) return 0;
  const int v = roi_bwd_version(a, R);
  if (v == 3) return roi_bwd3_ws(a, R, p->out_layout == 0).total;
  return v == 2 ? roi_bwd2_ws(a, R).total : roi_bwd_ws(a, R).total;
}

// grad_out (R,C,nbins) -> gt (R,nbins,C)
static int roi_bwd_transpose_grad(const float* grad_out, float* gt, int R, int C, int nbins,
                                  cudaStream_t stream) {
  const float* tin[1] = {grad_out};
  float* tout[1] = {gt};
  const int trows[1] = {C}, tcols[1] = {nbins};
  return transpose_launch_multi(tin, tout, 1, R, trows, tcols, stream);
}

static int roi_bwd2_launch(RoiArgs a, const float* grad_out, int out_layout, const float* rois,
                           int R, float* const* grad_feats, void* workspace,
                           size_t workspace_bytes, cudaStream_t stream) {
  const RoiBwd2Ws w = roi_bwd2_ws(a, R);
  if (!workspace || workspace_bytes < w.total) return BRCNN_ERR_WORKSPACE;
  if (misaligned16(workspace)) return BRCNN_ERR_ARG;
  RoiBwd2Args ba;
  memset(&ba, 0, sizeof(ba));
  for (int l = 0; l < a.L; ++l) {
    if (a.H[l] > a.max_h) a.max_h = a.H[l];
    if (a.W[l] > a.max_w) a.max_w = a.W[l];
  }
  ba.a = a;
  ba.TR = a.max_h + a.max_w;
  long long base = 0;
  for (int l = 0; l < a.L; ++l) {
    if (!grad_feats[l] || misaligned16(grad_feats[l])) return BRCNN_ERR_ARG;
    ba.grad[l] = grad_feats[l];
    ba.tiles_x[l] = (a.W[l] + B2_TS - 1) / B2_TS;
    ba.tiles_y[l] = (a.H[l] + B2_TS - 1) / B2_TS;
    ba.tile_base[l] = (int)base;
    base += (long long)ba.tiles_x[l] * ba.tiles_y[l] * a.B;
    if (base > 0x7fffffffLL) return BRCNN_ERR_UNSUPPORTED;
  }
  ba.tile_base[a.L] = (int)base;
  char* ws = (char*)workspace;
  RoiBwdRec* recs = (RoiBwdRec*)(ws + w.recs);
  unsigned short* keys = (unsigned short*)(ws + w.keys);
  float* tab = (float*)(ws + w.tab);
  const float* gt = grad_out;
  const int nbins = a.PH * a.PW;
  if (R > 0) {
    RoiBwdBuckets nob;
    memset(&nob, 0, sizeof(nob));
    roi_bwd_prep_kernel<<<R, B2_PREP_THREADS, 0, stream>>>(ba.a, rois, R, ba.TR, recs, keys, tab, nob);
    g_launch_count_add(1);
    BRCNN_CUDA_CHECK_LAST();
    if (out_layout == 0) {
      if (misaligned16(grad_out)) return BRCNN_ERR_ARG;
      const int rc = roi_bwd_transpose_grad(grad_out, (float*)(ws + w.gt), R, a.C, nbins, stream);
      if (rc) return rc;
      gt = (const float*)(ws + w.gt);
    }
  }
  dim3 grid((unsigned)base, (a.C + B2_CCH - 1) / B2_CCH);
  if (grid.y > 65535) return BRCNN_ERR_UNSUPPORTED;
  roi_bwd_gather2_kernel<<<grid, B2_THREADS, 0, stream>>>(ba, recs, keys, R, tab, gt);
  g_launch_count_add(1);
  BRCNN_CUDA_CHECK_LAST();
  return BRCNN_OK;
}

static int roi_bwd3_launch(RoiArgs a, const float* grad_out, int out_layout, const float* rois,
                           int R, float* const* grad_feats, void* workspace,
                           size_t workspace_bytes, cudaStream_t stream) {
  const RoiBwd3Ws w = roi_bwd3_ws(a, R, out_layout == 0);
  if (!workspace || workspace_bytes < w.total) return BRCNN_ERR_WORKSPACE;
  if (misaligned16(workspace) || (R > 0 && misaligned16(grad_out))) return BRCNN_ERR_ARG;
  RoiBwd3Args ba;
  memset(&ba, 0, sizeof(ba));
  for (int l = 0; l < a.L; ++l) {
    if (a.H[l] > a.max_h) a.max_h = a.H[l];
    if (a.W[l] > a.max_w) a.max_w = a.W[l];
  }
  ba.a = a;
  ba.TR = a.max_h + a.max_w;
  ba.bucket_cap = R > 0 ? R : 1;
  long long base = 0;
  for (int l = a.L - 1; l >= 0; --l) {      // coarse levels first
    if (!grad_feats[l] || misaligned16(grad_feats[l])) return BRCNN_ERR_ARG;
    ba.grad[l] = grad_feats[l];
    ba.tiles_x[l] = (a.W[l] + B3_TS - 1) / B3_TS;
    ba.tiles_y[l] = (a.H[l] + B3_TS - 1) / B3_TS;
    ba.tile_first[l] = (int)base;
    base += (long long)ba.tiles_x[l] * ba.tiles_y[l] * a.B;
    if (base > 0x7fffffffLL) return BRCNN_ERR_UNSUPPORTED;
  }
  char* ws = (char*)workspace;
  RoiBwdRec* recs = (RoiBwdRec*)(ws + w.recs);
  unsigned short* keys = (unsigned short*)(ws + w.keys);
  float* tab = (float*)(ws + w.tab);
  int32_t* bucket_cnt = (int32_t*)(ws + w.bucket_cnt);
  int32_t* tile_cnt = (int32_t*)(ws + w.tile_cnt);
  int32_t* bucket = (int32_t*)(ws + w.bucket);
  RoiBwdRec* bucket_rec = (RoiBwdRec*)(ws + w.bucket_rec);
  const float* gt = grad_out;
  const int nbins = a.PH * a.PW;
  if (R == 0) {              // nothing to gather: every level gets zeros
    for (int l = 0; l < a.L; ++l) {
      cudaError_t e0 = cudaMemsetAsync(ba.grad[l], 0, (size_t)a.B * a.H[l] * a.W[l] * a.C * 4, stream);
      if (e0 != cudaSuccess) return (int)e0;
    }
    return BRCNN_OK;
  }
  cudaError_t e = cudaMemsetAsync(bucket_cnt, 0, w.zero_bytes, stream);
  if (e != cudaSuccess) return (int)e;
  if (R > 0) {
    RoiBwdBuckets bk;
    memset(&bk, 0, sizeof(bk));
    bk.bucket = bucket; bk.bucket_rec = bucket_rec; bk.bucket_cnt = bucket_cnt;
    bk.tile_cnt = tile_cnt; bk.tile_side = B3_TS;
    bk.tile_r = (int32_t*)(ws + w.tile_r); bk.tile_rec = (RoiBwdRec*)(ws + w.tile_rec);
    bk.tile_cap = B4_TILE_CAP;
    for (int l = 0; l < a.L; ++l) {
      bk.tiles_x[l] = ba.tiles_x[l]; bk.tiles_y[l] = ba.tiles_y[l];
      bk.tile_first[l] = ba.tile_first[l];
    }
    roi_bwd_prep_kernel<<<R, B2_PREP_THREADS, 0, stream>>>(ba.a, rois, R, ba.TR, recs, keys, tab, bk);
    g_launch_count_add(1);
    BRCNN_CUDA_CHECK_LAST();
    if (out_layout == 0) {
      const int rc = roi_bwd_transpose_grad(grad_out, (float*)(ws + w.gt), R, a.C, nbins, stream);
      if (rc) return rc;
      gt = (const float*)(ws + w.gt);
    }
  }
  dim3 grid((unsigned)base, (a.C + B3_CS - 1) / B3_CS);
  if (grid.y > 65535) return BRCNN_ERR_UNSUPPORTED;
  // one CTA per (tile, slab), warp-private cp.async rings (roi_align_bwd5.cuh)
  static_assert(B4_TILE_CAP == B5_WIN, "tile list capacity");
  e = ensure_dyn_smem((const void*)roi_bwd_gather5_kernel, B5_RING_BYTES, true);
  if (e != cudaSuccess) return (int)e;
  roi_bwd_gather5_kernel<<<grid, B5_THREADS, B5_RING_BYTES, stream>>>(
      ba, (const int32_t*)(ws + w.tile_r), (const RoiBwdRec*)(ws + w.tile_rec), tile_cnt,
      bucket_rec, bucket, bucket_cnt, R, tab, gt);
  g_launch_count_add(1);
  BRCNN_CUDA_CHECK_LAST();
  return BRCNN_OK;
}

int brcnn_roi_extract_backward(const brcnn_roi_params* p, const float* grad_out,
                               const float* rois, int32_t R,
                               float* const* grad_feats_nhwc_host, void* workspace,
                               size_t workspace_bytes, brcnn_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  RoiArgs a;
  int rc = roi_args_from(p, &a);
  if (rc) return rc;
  if (R < 0 || !grad_feats_nhwc_host) return BRCNN_ERR_ARG;
  if (R > 0 && (!grad_out || !rois)) return BRCNN_ERR_ARG;
  const int v = roi_bwd_version(a, R);
  if (v == 3)
    return roi_bwd3_launch(a, grad_out, p->out_layout, rois, R, grad_feats_nhwc_host, workspace,
                           workspace_bytes, stream);
  if (v == 2)
    return roi_bwd2_launch(a, grad_out, p->out_layout, rois, R, grad_feats_nhwc_host, workspace,
                           workspace_bytes, stream);
  if (p->out_layout != 0) return BRCNN_ERR_UNSUPPORTED;
  return roi_bwd_launch(a, grad_out, rois, R, grad_feats_nhwc_host, workspace,
                        workspace_bytes, stream);
}

int brcnn_nchw_to_nhwc(const float* in, float* out, int32_t batch, int32_t channels,
                       int32_t hw, brcnn_stream_t stream) {
  const float* i1[1] = {in}; float* o1[1] = {out};
  const int r[1] = {channels}, c[1] = {hw};
  return transpose_launch_multi(i1, o1, 1, batch, r, c, (cudaStream_t)stream);
}
int brcnn_nhwc_to_nchw(const float* in, float* out, int32_t batch, int32_t channels,
                       int32_t hw, brcnn_stream_t stream) {
  const float* i1[1] = {in}; float* o1[1] = {out};
  const int r[1] = {hw}, c[1] = {channels};
  return transpose_launch_multi(i1, o1, 1, batch, r, c, (cudaStream_t)stream);
}
int brcnn_nchw_to_nhwc_multi(const float* const* in_host, float* const* out_host,
                             int32_t num_maps, int32_t batch, int32_t channels,
                             const int32_t* hw_host, brcnn_stream_t stream) {
  if (num_maps <= 0 || num_maps > BRCNN_MAX_LEVELS || !hw_host) return BRCNN_ERR_ARG;
  int r[BRCNN_MAX_LEVELS], c[BRCNN_MAX_LEVELS];
  for (int i = 0; i < num_maps; ++i) { r[i] = channels; c[i] = hw_host[i]; }
  return transpose_launch_multi(in_host, out_host, num_maps, batch, r, c, (cudaStream_t)stream);
}
int brcnn_nhwc_to_nchw_multi(const float* const* in_host, float* const* out_host,
                             int32_t num_maps, int32_t batch, int32_t channels,
                             const int32_t* hw_host, brcnn_stream_t stream) {
  if (num_maps <= 0 || num_maps > BRCNN_MAX_LEVELS || !hw_host) return BRCNN_ERR_ARG;
  int r[BRCNN_MAX_LEVELS], c[BRCNN_MAX_LEVELS];
  for (int i = 0; i < num_maps; ++i) { r[i] = hw_host[i]; c[i] = channels; }
  return transpose_launch_multi(in_host, out_host, num_maps, batch, r, c, (cudaStream_t)stream);
}

// ------------------------------- loss -------------------------------------
size_t brcnn_boost_loss_workspace_bytes(const brcnn_loss_params* p) {
  if (!p || p->num_rois < 0) return 0;
  int G, rpc;
  boost_loss_grid(p->num_rois, &G, &rpc);
  return align256((size_t)G * BL_REC * 4);
}

int brcnn_boost_loss(const brcnn_loss_params* p, const float* cls_score,
                     const int64_t* labels, const float* label_weights,
                     const float* prior, const float* bbox_pred,
                     const float* bbox_targets, const float* bbox_weights,
                     float* out_scalars, float* grad_cls_score,
                     float* grad_bbox_pred, void* workspace, size_t workspace_bytes,
                     brcnn_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!p || p->num_rois < 0 || p->num_classes <= 0 || !out_scalars) return BRCNN_ERR_ARG;
  if (p->num_rois > 0 && (!cls_score || !labels || !prior || !bbox_pred ||
                          !bbox_targets || !bbox_weights || !grad_cls_score))
    return BRCNN_ERR_ARG;
  if (!workspace || workspace_bytes < brcnn_boost_loss_workspace_bytes(p))
    return BRCNN_ERR_WORKSPACE;
  LossArgs a;
  a.N = p->num_rois; a.C = p->num_classes; a.agnostic = p->reg_class_agnostic;
  a.reg_norm_mean = p->reg_norm_mean; a.gamma = p->gamma; a.alpha = p->alpha;
  a.wcls = p->loss_cls_weight; a.wbbox = p->loss_bbox_weight;
  if (grad_bbox_pred != nullptr && a.N > 0) {
    const size_t nb = (size_t)a.N * (a.agnostic ? 4 : 4 * a.C) * 4;
    cudaError_t e = cudaMemsetAsync(grad_bbox_pred, 0, nb, stream);
    if (e != cudaSuccess) return (int)e;
  }
  int G, rpc;
  boost_loss_grid(a.N, &G, &rpc);
  float* partials = (float*)workspace;
  boost_loss_part1_kernel<<<G, BL_THREADS, 0, stream>>>(
      a, rpc, cls_score, labels, label_weights, prior, bbox_pred, bbox_targets, bbox_weights,
      partials, grad_cls_score);
  g_launch_count_add(1);
  BRCNN_CUDA_CHECK_LAST();
  boost_loss_part2_kernel<<<G, BL_THREADS, 0, stream>>>(
      a, rpc, G, partials, labels, label_weights, prior, bbox_pred, bbox_targets, bbox_weights,
      out_scalars, grad_cls_score, grad_bbox_pred);
  g_launch_count_add(1);
  BRCNN_CUDA_CHECK_LAST();
  return BRCNN_OK;
}

// --------------------------- RCNN train front-end -------------------------
int brcnn_rcnn_assign(const brcnn_assign_params* p, const float* proposals,
                      const int32_t* num_props, const float* gt_boxes, const int32_t* num_gt,
                      int32_t* gt_inds, int32_t* counts, brcnn_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!p || p->batch <= 0 || p->max_props <= 0 || p->max_gts < 0) return BRCNN_ERR_ARG;
  if (p->match_low_quality) return BRCNN_ERR_UNSUPPORTED;
  if (!proposals || !num_props || !num_gt || !gt_inds || !counts) return BRCNN_ERR_ARG;
  if (p->max_gts > 0 && (!gt_boxes || misaligned16(gt_boxes))) return BRCNN_ERR_ARG;
  if (p->batch > 65535 || p->max_gts > 2048) return BRCNN_ERR_UNSUPPORTED;
  cudaError_t e = cudaMemsetAsync(counts, 0, (size_t)p->batch * 2 * 4, stream);
  if (e != cudaSuccess) return (int)e;
  AssignArgs a;
  a.B = p->batch; a.M = p->max_props; a.Gmax = p->max_gts;
  a.pos_iou_thr = p->pos_iou_thr; a.neg_iou_thr = p->neg_iou_thr;
  const int slots = a.Gmax + a.M;
  dim3 grid((slots + 255) / 256, p->batch);
  const size_t smem = (size_t)(a.Gmax > 0 ? a.Gmax : 1) * 16;
  rcnn_assign_kernel<<<grid, 256, smem, stream>>>(a, proposals, num_props, gt_boxes, num_gt,
                                                  gt_inds, counts);
  g_launch_count_add(1);
  BRCNN_CUDA_CHECK_LAST();
  return BRCNN_OK;
}

int brcnn_rcnn_sample_targets(const brcnn_sample_params* p, const float* proposals,
                              const int32_t* num_props, const float* gt_boxes,
                              const int64_t* gt_labels, const int32_t* num_gt,
                              const int32_t* gt_inds, const int32_t* plan,
                              const int32_t* perm_pos, const int32_t* perm_neg,
                              float* rois, int64_t* labels, float* label_weights,
                              float* bbox_targets, float* bbox_weights, float* prior,
                              brcnn_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!p || p->batch <= 0 || p->max_props <= 0 || p->max_gts < 0 || p->max_sel <= 0 ||
      p->perm_cap <= 0)
    return BRCNN_ERR_ARG;
  if (!proposals || !num_props || !num_gt || !gt_inds || !plan || !perm_pos || !perm_neg ||
      !rois || !labels || !label_weights || !bbox_targets || !bbox_weights || !prior)
    return BRCNN_ERR_ARG;
  if (p->max_gts > 0 && (!gt_boxes || !gt_labels || misaligned16(gt_boxes))) return BRCNN_ERR_ARG;
  if (misaligned16(bbox_targets) || misaligned16(bbox_weights)) return BRCNN_ERR_ARG;
  SampleArgs a;
  a.B = p->batch; a.M = p->max_props; a.Gmax = p->max_gts; a.num_classes = p->num_classes;
  a.perm_cap = p->perm_cap; a.pos_weight = p->pos_weight;
  for (int i = 0; i < 4; ++i) { a.means[i] = p->means[i]; a.stds[i] = p->stds[i]; }
  const int sel_cap = next_pow2(p->max_sel);
  const size_t smem = (size_t)sel_cap * 8 + (size_t)(a.Gmax + a.M) * 4;
  if (smem > 200 * 1024) return BRCNN_ERR_UNSUPPORTED;
  if (smem > 48 * 1024) {
    cudaError_t e = ensure_dyn_smem((const void*)rcnn_sample_target_kernel, smem);
    if (e != cudaSuccess) return (int)e;
  }
  rcnn_sample_target_kernel<<<p->batch, ST_THREADS, smem, stream>>>(
      a, proposals, num_props, gt_boxes, gt_labels, num_gt, gt_inds, plan, perm_pos, perm_neg,
      sel_cap, rois, labels, label_weights, bbox_targets, bbox_weights, prior);
  g_launch_count_add(1);
  BRCNN_CUDA_CHECK_LAST();
  return BRCNN_OK;
}

// ------------------------------- RCNN post --------------------------------
int brcnn_rcnn_workspace_layout(const brcnn_rcnn_params* p, brcnn_rcnn_ws_layout* out) {
  RcnnWsInternal w;
  int rc = rcnn_ws(p, &w);
  if (rc) return rc;
  if (out) *out = w.pub;
  return BRCNN_OK;
}
size_t brcnn_rcnn_workspace_bytes(const brcnn_rcnn_params* p) {
  RcnnWsInternal w;
  if (rcnn_ws(p, &w)) return 0;
  return (size_t)w.pub.total_bytes;
}

int brcnn_rcnn_get_bboxes(const brcnn_rcnn_params* p, const float* rois,
                          const float* prior, const int32_t* num_rois,
                          const float* cls_score, const float* bbox_pred,
                          const float* img_hw, const float* scale_factor,
                          float* det_bboxes, int64_t* det_labels, int32_t* num_dets,
                          void* workspace, size_t workspace_bytes,
                          brcnn_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  RcnnWsInternal w;
  int rc = rcnn_ws(p, &w);
  if (rc) return rc;
  if (!rois || !num_rois || !cls_score || !bbox_pred || !img_hw || !det_bboxes ||
      !det_labels || !num_dets || !workspace)
    return BRCNN_ERR_ARG;
  if (p->prob && !prior) return BRCNN_ERR_ARG;
  if (p->rescale && (!scale_factor || misaligned16(scale_factor))) return BRCNN_ERR_ARG;
  if (misaligned16(workspace)) return BRCNN_ERR_ARG;
  if (workspace_bytes < (size_t)w.pub.total_bytes) return BRCNN_ERR_WORKSPACE;
  if (p->batch > 65535 || p->num_classes > 65535) return BRCNN_ERR_UNSUPPORTED;
  RcnnArgs a;
  a.B = p->batch; a.Rc = p->rois_per_img; a.C = p->num_classes;
  a.agnostic = p->reg_class_agnostic; a.prob = p->prob; a.rescale = p->rescale;
  for (int i = 0; i < 4; ++i) { a.means[i] = p->means[i]; a.stds[i] = p->stds[i]; }
  a.max_ratio = p->max_ratio; a.score_thr = p->score_thr;
  char* ws = (char*)workspace;
  float* scores = (float*)(ws + w.pub.scores);
  float4* bboxes = (float4*)(ws + w.pub.bboxes);
  int* img_maxc = (int*)(ws + w.pub.img_maxc);
  int32_t* seg_count = (int32_t*)(ws + w.pub.seg_count);
  u64* seg_key = (u64*)(ws + w.pub.seg_key);
  float4* seg_boxes = (float4*)(ws + w.seg_boxes);
  int32_t* kept_pos = (int32_t*)(ws + w.pub.kept_pos);
  int32_t* kept_count = (int32_t*)(ws + w.pub.kept_count);
  u64* kept_key = (u64*)(ws + w.kept_key);
  u64* mask = (u64*)(ws + w.mask);

  cudaError_t e = cudaMemsetAsync(img_maxc, 0, (size_t)a.B * 4, stream);
  if (e != cudaSuccess) return (int)e;
  const int rows = a.B * a.Rc;
  const size_t smem1 = (size_t)RCNN_FUSE_WARPS * (a.C + 1) * 4;
  if (smem1 > 48 * 1024) return BRCNN_ERR_UNSUPPORTED;
  rcnn_fuse_decode_kernel<<<(rows + RCNN_FUSE_WARPS - 1) / RCNN_FUSE_WARPS,
                            RCNN_FUSE_WARPS * 32, smem1, stream>>>(
      a, rois, prior, num_rois, cls_score, bbox_pred, img_hw, scale_factor, scores,
      bboxes, img_maxc);
  g_launch_count_add(1);
  BRCNN_CUDA_CHECK_LAST();

  const int rpow2 = next_pow2(a.Rc);
  dim3 grid2(a.C, a.B);
  rcnn_class_sort_kernel<<<grid2, 256, (size_t)rpow2 * 8, stream>>>(
      a, num_rois, scores, bboxes, rpow2, seg_key, seg_boxes, seg_count);
  g_launch_count_add(1);
  BRCNN_CUDA_CHECK_LAST();

  const int S = a.B * a.C;
  rc = launch_nms_segments(seg_boxes, nullptr, seg_count, S, a.Rc, p->iou_threshold,
                           0.f, (const float*)img_maxc, a.C, mask, seg_key, kept_pos,
                           kept_key, kept_count, w.keep_cap, p->max_per_img, stream);
  if (rc) return rc;

  RcnnMergeEpilogue ep{bboxes, det_bboxes, det_labels, a.Rc, a.C,
                       a.agnostic ? 1 : a.C, p->max_per_img};
  return launch_nms_merge(kept_pos, kept_key, kept_count, a.B, a.C, w.keep_cap,
                          p->max_per_img, num_dets, ep, stream);
}

}  // extern "C"
