// retrieval.cu -- DINOv2 CLS-token cosine similarity + the eval loop's slot-replacement top-k, on device.
//
// Replaces, for all R reference crops at once, the per-crop
//   score = F.cosine_similarity(ref_fea, fea, dim=1, eps=1e-8); if (score.item() > similarity_score).any(): ...
// of eval_linemod_json.py:94-101 (one `.item()` host sync per crop in the reference; none here).
// torch's cosine_similarity normalises each vector by max(|v|, eps) before the dot product.
#include <type_traits>

#include "common.cuh"

namespace pope {
namespace {

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

// 8 consecutive elements of a row as floats (one 16-byte load for bf16, two for fp32)
template <typename T> __device__ __forceinline__ void load8(const T* p, float (&f)[8]);
template <> __device__ __forceinline__ void load8<float>(const float* p, float (&f)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
template <> __device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16* p, float (&f)[8]) {
  const uint4 a = __ldg(reinterpret_cast<const uint4*>(p));
  const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) { f[2 * e] = __uint_as_float(w[e] << 16); f[2 * e + 1] = __uint_as_float(w[e] & 0xffff0000u); }
}

// One CTA per reference row (kCosThreads threads): HBM-bound for long rows (the patch-token shape streams 50-100 MB), so
// every thread keeps several 16-byte loads of the row in flight; the query is re-read from L2.  VEC: rows are 16-byte
// aligned and D % 8 == 0.
constexpr int kCosThreadsShort = 128, kCosThreadsLong = 1024;    // long rows (>= 64 KB): 256 rows must fill 148 SMs by themselves
template <typename T, bool VEC, int kCosThreads>
__global__ void __launch_bounds__(kCosThreads) cosine_kernel(const T* __restrict__ q, const T* __restrict__ refs, int R, int D,
                                                            float eps, float* __restrict__ scores) {
  __shared__ float red[3][kCosThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r = blockIdx.x;
  const T* x = refs + size_t(r) * D;
  float qq = 0.f, xx = 0.f, qx = 0.f;
  if (VEC) {
    float q1 = 0.f, x1 = 0.f, c1 = 0.f;                      // second accumulator set: two independent chains
    int d = threadIdx.x * 8;
    for (; d + kCosThreads * 8 < D; d += 2 * kCosThreads * 8) {
      float a[8], b[8], a2[8], b2[8];
      load8<T>(x + d, b); load8<T>(x + d + kCosThreads * 8, b2);
      load8<T>(q + d, a); load8<T>(q + d + kCosThreads * 8, a2);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        qq = fmaf(a[e], a[e], qq); xx = fmaf(b[e], b[e], xx); qx = fmaf(a[e], b[e], qx);
        q1 = fmaf(a2[e], a2[e], q1); x1 = fmaf(b2[e], b2[e], x1); c1 = fmaf(a2[e], b2[e], c1);
      }
    }
    for (; d < D; d += kCosThreads * 8) {
      float a[8], b[8];
      load8<T>(x + d, b); load8<T>(q + d, a);
#pragma unroll
      for (int e = 0; e < 8; ++e) { qq = fmaf(a[e], a[e], qq); xx = fmaf(b[e], b[e], xx); qx = fmaf(a[e], b[e], qx); }
    }
    qq += q1; xx += x1; qx += c1;
  } else {
    for (int d = threadIdx.x; d < D; d += kCosThreads) {
      const float a = to_f<T>(q[d]), b = to_f<T>(x[d]);
      qq = fmaf(a, a, qq); xx = fmaf(b, b, xx); qx = fmaf(a, b, qx);
    }
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    qq += __shfl_xor_sync(kFullMask, qq, o);
    xx += __shfl_xor_sync(kFullMask, xx, o);
    qx += __shfl_xor_sync(kFullMask, qx, o);
  }
  if (lane == 0) { red[0][warp] = qq; red[1][warp] = xx; red[2][warp] = qx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    qq = xx = qx = 0.f;
#pragma unroll
    for (int w = 0; w < kCosThreads / 32; ++w) { qq += red[0][w]; xx += red[1][w]; qx += red[2][w]; }
    scores[r] = qx / (fmaxf(sqrtf(qq), eps) * fmaxf(sqrtf(xx), eps));
  }
}

// The eval loop's slot update (eval_linemod_json.py:95-101) over nr scores in shared memory, continuing from the slots in
// registers: `if (score > slots).any(): slots[argmin(slots)] = score` -- i.e. a score enters iff it exceeds the smallest slot
// (first arg-min, like np.argmin).  Slots live in registers (fully unrolled over kTopkMaxK) and the smallest slot is cached:
// one shared-memory load and one compare per crop once the slots have filled (walking the slots in shared memory cost
// ~150 clk per crop: 20 us for 256 crops, the whole cost of the CLS-token case).
struct TopkSlots {
  float ss[16];
  int si[16];
  float smin;
  int lo;
};
__device__ __forceinline__ void topk_init(TopkSlots& T, int k) {
#pragma unroll
  for (int t = 0; t < 16; ++t) { T.ss[t] = t < k ? 0.f : INFINITY; T.si[t] = -1; }
  T.smin = 0.f;
  T.lo = 0;
}
__device__ __forceinline__ void topk_feed(TopkSlots& T, const float* sc, int nr, int r0) {
  auto feed = [&](float s, int r) {
    if (!(s > T.smin)) return;
#pragma unroll
    for (int t = 0; t < 16; ++t)
      if (t == T.lo) { T.ss[t] = s; T.si[t] = r0 + r; }
    T.smin = T.ss[0];
    T.lo = 0;
#pragma unroll
    for (int t = 1; t < 16; ++t)
      if (T.ss[t] < T.smin) { T.smin = T.ss[t]; T.lo = t; }     // (slots past k hold +inf)
  };
  int r = 0;
  for (; r + 4 <= nr; r += 4) {                                  // sc is 16-byte aligned: four crops per load
    const float4 v = *reinterpret_cast<const float4*>(sc + r);
    feed(v.x, r); feed(v.y, r + 1); feed(v.z, r + 2); feed(v.w, r + 3);
  }
  for (; r < nr; ++r) feed(sc[r], r);
}
__device__ __forceinline__ void topk_store(const TopkSlots& T, int k, float* slot_scores, int32_t* slot_idx) {
#pragma unroll
  for (int t = 0; t < 16; ++t)
    if (t < k) { slot_scores[t] = T.ss[t]; slot_idx[t] = T.si[t]; }
}

// Sequential by construction (slot order depends on arrival order); R is a few hundred.  The scores are staged in shared
// memory by the whole block first: one thread walking global memory pays a full L2 round trip per crop (measured 50 us
// for R = 256).
constexpr int kTopkThreads = 256, kTopkChunk = 2048, kTopkMaxK = 16;
__global__ void __launch_bounds__(kTopkThreads) running_topk_kernel(const float* __restrict__ scores, int R, int k,
                                                                   float* __restrict__ slot_scores,
                                                                   int32_t* __restrict__ slot_idx) {
  __shared__ __align__(16) float sc[kTopkChunk];
  TopkSlots T;
  topk_init(T, k);
  for (int r0 = 0; r0 < R; r0 += kTopkChunk) {
    const int nr = min(kTopkChunk, R - r0);
    __syncthreads();
    for (int r = threadIdx.x; r < nr; r += kTopkThreads) sc[r] = scores[r0 + r];
    __syncthreads();
    if (threadIdx.x == 0) topk_feed(T, sc, nr, r0);
  }
  if (threadIdx.x == 0) topk_store(T, k, slot_scores, slot_idx);
}

// The CLS-token case of the eval loop (R = a few hundred crops, D = 384: 0.4 MB) is bound by launch latency, not by memory:
// ONE CTA computes all scores (a warp per row, rows w, w + 32, ...; same per-lane accumulation and shuffle order for every
// row) and then runs the slot update over them -- one launch instead of two and no trip through global memory in between.
constexpr int kSmallThreads = 1024, kSmallMaxBytes = 2 << 20;
template <typename T, bool VEC>
__global__ void __launch_bounds__(kSmallThreads) cosine_topk_small_kernel(const T* __restrict__ q, const T* __restrict__ refs, int R,
                                                                          int D, float eps, int k, float* __restrict__ scores,
                                                                          float* __restrict__ slot_scores,
                                                                          int32_t* __restrict__ slot_idx) {
  __shared__ __align__(16) float sc[kTopkChunk];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float qq = 0.f;                                   // |q|^2: every warp computes it once, in the order it uses for the rows
  if (VEC) {
    for (int d = lane * 8; d < D; d += 256) {
      float a[8];
      load8<T>(q + d, a);
#pragma unroll
      for (int e = 0; e < 8; ++e) qq = fmaf(a[e], a[e], qq);
    }
  } else {
    for (int d = lane; d < D; d += 32) { const float a = to_f<T>(q[d]); qq = fmaf(a, a, qq); }
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) qq += __shfl_xor_sync(kFullMask, qq, o);
  // four rows of the warp at a time: their loads are in flight together (a warp has R / 32 rows; one at a time pays a full
  // memory round trip per row)
  constexpr int kRowsAtOnce = 4, kWarps = kSmallThreads / 32;
  for (int r0 = warp; r0 < R; r0 += kRowsAtOnce * kWarps) {
    float xx[kRowsAtOnce], qx[kRowsAtOnce];
#pragma unroll
    for (int u = 0; u < kRowsAtOnce; ++u) xx[u] = qx[u] = 0.f;
    if (VEC) {
      for (int d = lane * 8; d < D; d += 256) {
        float a[8], b[kRowsAtOnce][8];
#pragma unroll
        for (int u = 0; u < kRowsAtOnce; ++u) {
          const int r = min(r0 + u * kWarps, R - 1);            // (rows past the end repeat the last one; not stored)
          load8<T>(refs + size_t(r) * D + d, b[u]);
        }
        load8<T>(q + d, a);
#pragma unroll
        for (int u = 0; u < kRowsAtOnce; ++u)
#pragma unroll
          for (int e = 0; e < 8; ++e) { xx[u] = fmaf(b[u][e], b[u][e], xx[u]); qx[u] = fmaf(a[e], b[u][e], qx[u]); }
      }
    } else {
      for (int d = lane; d < D; d += 32) {
        const float a = to_f<T>(q[d]);
#pragma unroll
        for (int u = 0; u < kRowsAtOnce; ++u) {
          const float b = to_f<T>(refs[size_t(min(r0 + u * kWarps, R - 1)) * D + d]);
          xx[u] = fmaf(b, b, xx[u]); qx[u] = fmaf(a, b, qx[u]);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < kRowsAtOnce; ++u) {
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) { xx[u] += __shfl_xor_sync(kFullMask, xx[u], o); qx[u] += __shfl_xor_sync(kFullMask, qx[u], o); }
      const int r = r0 + u * kWarps;
      if (lane == 0 && r < R) {
        const float sres = qx[u] / (fmaxf(sqrtf(qq), eps) * fmaxf(sqrtf(xx[u]), eps));
        sc[r] = sres;
        scores[r] = sres;
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    TopkSlots T;
    topk_init(T, k);
    topk_feed(T, sc, R, 0);
    topk_store(T, k, slot_scores, slot_idx);
  }
}

// Match-list consumer of the eval loop (eval_linemod_json.py:118-119, :146): for every pair the number of matches with
// mconf > thr ("matching_score"), and for every group of `group` consecutive pairs (one query against its retrieved
// crops) the first arg-max of those scores (np.argmax).  One CTA per group; the pair's matches are the contiguous range
// [prefix(counts), +counts[n]) of the (pair, i)-sorted list written by pope_coarse_match.
__global__ void __launch_bounds__(256) match_score_kernel(const float* __restrict__ mconf, const int32_t* __restrict__ counts,
                                                         int n_pairs, int group, float thr, int32_t* __restrict__ scores,
                                                         int32_t* __restrict__ best) {
  __shared__ int red[8];
  const int g = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int p0 = g * group, p1 = min(p0 + group, n_pairs);
  auto block_sum = [&](int v) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    int t = lane < 8 ? red[lane] : 0;
#pragma unroll
    for (int o = 4; o >= 1; o >>= 1) t += __shfl_xor_sync(kFullMask, t, o);
    return __shfl_sync(kFullMask, t, 0);
  };
  int part = 0;
  for (int p = threadIdx.x; p < p0; p += blockDim.x) part += counts[p];
  int64_t off = block_sum(part);
  int best_score = -1, best_idx = 0;
  for (int p = p0; p < p1; ++p) {
    const int c = counts[p];
    int hits = 0;
    for (int k = threadIdx.x; k < c; k += blockDim.x) hits += mconf[off + k] > thr ? 1 : 0;
    hits = block_sum(hits);
    off += c;
    if (threadIdx.x == 0) scores[p] = hits;
    if (hits > best_score) { best_score = hits; best_idx = p - p0; }      // strict: the first maximum wins
  }
  if (threadIdx.x == 0) best[g] = best_idx;
}

// One 32-byte record per live match: (global pair index, i, j, mconf, x0, y0, x1, y1), floats as bit patterns -- the
// unit of the job-level gather of match lists.  The live count is read on the device.
__global__ void __launch_bounds__(256) pack_records_kernel(const int64_t* __restrict__ b_ids, const int64_t* __restrict__ i_ids,
                                                          const int64_t* __restrict__ j_ids, const float* __restrict__ mconf,
                                                          const float2* __restrict__ mk0, const float2* __restrict__ mk1,
                                                          const int32_t* __restrict__ m_dev, int64_t capacity,
                                                          int pair_offset, int4* __restrict__ rec,
                                                          const int64_t* __restrict__ base_dev) {
  const int64_t m = min(int64_t(*m_dev), capacity);
  if (base_dev) rec += 2 * *base_dev;                       // append behind the records already in the buffer
  for (int64_t k = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; k < m; k += int64_t(gridDim.x) * blockDim.x) {
    const float2 a = mk0[k], b = mk1[k];
    rec[2 * k] = make_int4(int(b_ids[k]) + pair_offset, int(i_ids[k]), int(j_ids[k]), __float_as_int(mconf[k]));
    rec[2 * k + 1] = make_int4(__float_as_int(a.x), __float_as_int(a.y), __float_as_int(b.x), __float_as_int(b.y));
  }
}

// The compact form of the record, 20 bytes (int32[5]): global pair index, i | j << 16, mconf, x1, y1 -- the keypoint of
// image 0 is implied by i (mkpts0_f = mkpts0_c = (i % w0c, i / w0c) * pixel_scale, fine_matching.py:66), so the gather moves
// 20 instead of 32 bytes per match.  Needs L, S <= 65536.
__global__ void __launch_bounds__(256) pack_records5_kernel(const int64_t* __restrict__ b_ids, const int64_t* __restrict__ i_ids,
                                                           const int64_t* __restrict__ j_ids, const float* __restrict__ mconf,
                                                           const float2* __restrict__ mk1, const int32_t* __restrict__ m_dev,
                                                           int64_t capacity, int pair_offset, int32_t* __restrict__ rec,
                                                           const int64_t* __restrict__ base_dev) {
  const int64_t m = min(int64_t(*m_dev), capacity);
  if (base_dev) rec += 5 * *base_dev;
  for (int64_t k = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; k < m; k += int64_t(gridDim.x) * blockDim.x) {
    const float2 b = mk1[k];
    int32_t* r = rec + 5 * k;
    r[0] = int(b_ids[k]) + pair_offset;
    r[1] = int(uint32_t(i_ids[k]) | (uint32_t(j_ids[k]) << 16));
    r[2] = __float_as_int(mconf[k]);
    r[3] = __float_as_int(b.x);
    r[4] = __float_as_int(b.y);
  }
}

}  // namespace
}  // namespace pope

using namespace pope;

extern "C" int pope_pack_records_compact(const int64_t* b_ids, const int64_t* i_ids, const int64_t* j_ids, const float* mconf,
                                         const float* mkpts1_f, const int32_t* m_dev, int64_t capacity, int pair_offset,
                                         int32_t* records, const int64_t* base_dev, void* stream) {
  if (!b_ids || !i_ids || !j_ids || !mconf || !mkpts1_f || !m_dev || !records || capacity < 0) return POPE_ERR_INVALID_ARG;
  if (reinterpret_cast<uintptr_t>(mkpts1_f) & 7u || reinterpret_cast<uintptr_t>(records) & 3u) return POPE_ERR_ALIGNMENT;
  if (capacity == 0) return POPE_OK;
  const int64_t want = (capacity + 255) / 256;
  const unsigned blocks = unsigned(want < 148 * 8 ? want : 148 * 8);
  pack_records5_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      b_ids, i_ids, j_ids, mconf, reinterpret_cast<const float2*>(mkpts1_f), m_dev, capacity, pair_offset, records, base_dev);
  return int(cudaGetLastError());
}

extern "C" int pope_pack_records(const int64_t* b_ids, const int64_t* i_ids, const int64_t* j_ids, const float* mconf,
                                 const float* mkpts0_f, const float* mkpts1_f, const int32_t* m_dev, int64_t capacity,
                                 int pair_offset, int32_t* records, const int64_t* base_dev, void* stream) {
  if (!b_ids || !i_ids || !j_ids || !mconf || !mkpts0_f || !mkpts1_f || !m_dev || !records || capacity < 0)
    return POPE_ERR_INVALID_ARG;
  if ((reinterpret_cast<uintptr_t>(mkpts0_f) | reinterpret_cast<uintptr_t>(mkpts1_f)) & 7u ||
      reinterpret_cast<uintptr_t>(records) & 15u)
    return POPE_ERR_ALIGNMENT;
  if (capacity == 0) return POPE_OK;
  const int64_t want = (capacity + 255) / 256;
  const unsigned blocks = unsigned(want < 148 * 8 ? want : 148 * 8);
  pack_records_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      b_ids, i_ids, j_ids, mconf, reinterpret_cast<const float2*>(mkpts0_f), reinterpret_cast<const float2*>(mkpts1_f), m_dev,
      capacity, pair_offset, reinterpret_cast<int4*>(records), base_dev);
  return int(cudaGetLastError());
}

extern "C" int pope_match_scores(const float* mconf, const int32_t* counts, int n_pairs, int group, float thr,
                                 int32_t* scores, int32_t* best, void* stream) {
  // mconf may be NULL when no pair has a match (an empty list has no storage); the kernel then never reads it
  if (!counts || !scores || !best || n_pairs <= 0 || group <= 0) return POPE_ERR_INVALID_ARG;
  const int groups = (n_pairs + group - 1) / group;
  match_score_kernel<<<groups, 256, 0, static_cast<cudaStream_t>(stream)>>>(mconf, counts, n_pairs, group, thr, scores, best);
  return int(cudaGetLastError());
}

extern "C" int pope_running_topk(const float* scores, int R, int k, float* slot_scores, int32_t* slot_idx, void* stream) {
  if (!scores || !slot_scores || !slot_idx || R <= 0 || k <= 0 || k > kTopkMaxK) return POPE_ERR_INVALID_ARG;
  running_topk_kernel<<<1, kTopkThreads, 0, static_cast<cudaStream_t>(stream)>>>(scores, R, k, slot_scores, slot_idx);
  return int(cudaGetLastError());
}

extern "C" int pope_cosine_topk(const void* q, const void* refs, int dtype, int R, int D, int k, float eps,
                                float* scores, float* slot_scores, int32_t* slot_idx, void* stream) {
  if (!q || !refs || !scores || !slot_scores || !slot_idx || R <= 0 || D <= 0 || k <= 0 || k > kTopkMaxK) return POPE_ERR_INVALID_ARG;
  if (dtype != POPE_F32 && dtype != POPE_BF16) return POPE_ERR_DTYPE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t esz = dtype == POPE_BF16 ? 2 : 4;
  const bool vec = D % 8 == 0 && ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(refs)) & 15u) == 0 &&
                   (size_t(D) * esz) % 16 == 0;
  if (R <= kTopkChunk && size_t(R) * D * esz <= size_t(kSmallMaxBytes)) {      // small problems: one launch, one CTA
    if (dtype == POPE_BF16) {
      const __nv_bfloat16* qq = static_cast<const __nv_bfloat16*>(q);
      const __nv_bfloat16* rr = static_cast<const __nv_bfloat16*>(refs);
      if (vec) cosine_topk_small_kernel<__nv_bfloat16, true><<<1, kSmallThreads, 0, st>>>(qq, rr, R, D, eps, k, scores, slot_scores, slot_idx);
      else cosine_topk_small_kernel<__nv_bfloat16, false><<<1, kSmallThreads, 0, st>>>(qq, rr, R, D, eps, k, scores, slot_scores, slot_idx);
    } else {
      const float* qq = static_cast<const float*>(q);
      const float* rr = static_cast<const float*>(refs);
      if (vec) cosine_topk_small_kernel<float, true><<<1, kSmallThreads, 0, st>>>(qq, rr, R, D, eps, k, scores, slot_scores, slot_idx);
      else cosine_topk_small_kernel<float, false><<<1, kSmallThreads, 0, st>>>(qq, rr, R, D, eps, k, scores, slot_scores, slot_idx);
    }
    return int(cudaGetLastError());
  }
  const bool long_rows = size_t(D) * esz >= (64u << 10);
  auto launch = [&](auto qq, auto rr) {
    using T = typename std::remove_cv<typename std::remove_pointer<decltype(qq)>::type>::type;
    if (long_rows) {
      if (vec) cosine_kernel<T, true, kCosThreadsLong><<<R, kCosThreadsLong, 0, st>>>(qq, rr, R, D, eps, scores);
      else cosine_kernel<T, false, kCosThreadsLong><<<R, kCosThreadsLong, 0, st>>>(qq, rr, R, D, eps, scores);
    } else {
      if (vec) cosine_kernel<T, true, kCosThreadsShort><<<R, kCosThreadsShort, 0, st>>>(qq, rr, R, D, eps, scores);
      else cosine_kernel<T, false, kCosThreadsShort><<<R, kCosThreadsShort, 0, st>>>(qq, rr, R, D, eps, scores);
    }
  };
  if (dtype == POPE_BF16) launch(static_cast<const __nv_bfloat16*>(q), static_cast<const __nv_bfloat16*>(refs));
  else launch(static_cast<const float*>(q), static_cast<const float*>(refs));
  running_topk_kernel<<<1, kTopkThreads, 0, st>>>(scores, R, k, slot_scores, slot_idx);
  return int(cudaGetLastError());
}
