// retrieval.cu -- DINOv2 CLS-token cosine similarity + the eval loop's slot-replacement top-k, on device.
//
// Replaces, for all R reference crops at once, the per-crop
//   score = F.cosine_similarity(ref_fea, fea, dim=1, eps=1e-8); if (score.item() > similarity_score).any(): ...
// of eval_linemod_json.py:94-101 (one `.item()` host sync per crop in the reference; none here).
// torch's cosine_similarity normalises each vector by max(|v|, eps) before the dot product.
#include "common.cuh"

namespace pope {
namespace {

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T>
__global__ void __launch_bounds__(256) cosine_kernel(const T* __restrict__ q, const T* __restrict__ refs, int R, int D,
                                                    float eps, float* __restrict__ scores) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= R) return;
  const T* x = refs + size_t(r) * D;
  float qq = 0.f, xx = 0.f, qx = 0.f;
  for (int d = lane; d < D; d += 32) {
    const float a = to_f<T>(q[d]), b = to_f<T>(x[d]);
    qq = fmaf(a, a, qq); xx = fmaf(b, b, xx); qx = fmaf(a, b, qx);
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    qq += __shfl_xor_sync(kFullMask, qq, o);
    xx += __shfl_xor_sync(kFullMask, xx, o);
    qx += __shfl_xor_sync(kFullMask, qx, o);
  }
  if (lane == 0) scores[r] = qx / (fmaxf(sqrtf(qq), eps) * fmaxf(sqrtf(xx), eps));
}

// Sequential by construction (slot order depends on arrival order); R is a few hundred.
__global__ void running_topk_kernel(const float* __restrict__ scores, int R, int k, float* __restrict__ slot_scores,
                                    int32_t* __restrict__ slot_idx) {
  if (threadIdx.x != 0) return;
  for (int t = 0; t < k; ++t) { slot_scores[t] = 0.f; slot_idx[t] = -1; }
  for (int r = 0; r < R; ++r) {
    const float s = scores[r];
    bool any = false;
    int lo = 0;
    for (int t = 0; t < k; ++t) {
      any |= s > slot_scores[t];
      if (slot_scores[t] < slot_scores[lo]) lo = t;     // first arg-min, like np.argmin
    }
    if (any) { slot_scores[lo] = s; slot_idx[lo] = r; }
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
                                                          int pair_offset, int4* __restrict__ rec) {
  const int64_t m = min(int64_t(*m_dev), capacity);
  for (int64_t k = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; k < m; k += int64_t(gridDim.x) * blockDim.x) {
    const float2 a = mk0[k], b = mk1[k];
    rec[2 * k] = make_int4(int(b_ids[k]) + pair_offset, int(i_ids[k]), int(j_ids[k]), __float_as_int(mconf[k]));
    rec[2 * k + 1] = make_int4(__float_as_int(a.x), __float_as_int(a.y), __float_as_int(b.x), __float_as_int(b.y));
  }
}

}  // namespace
}  // namespace pope

using namespace pope;

extern "C" int pope_pack_records(const int64_t* b_ids, const int64_t* i_ids, const int64_t* j_ids, const float* mconf,
                                 const float* mkpts0_f, const float* mkpts1_f, const int32_t* m_dev, int64_t capacity,
                                 int pair_offset, int32_t* records, void* stream) {
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
      capacity, pair_offset, reinterpret_cast<int4*>(records));
  return int(cudaGetLastError());
}

extern "C" int pope_match_scores(const float* mconf, const int32_t* counts, int n_pairs, int group, float thr,
                                 int32_t* scores, int32_t* best, void* stream) {
  if (!mconf || !counts || !scores || !best || n_pairs <= 0 || group <= 0) return POPE_ERR_INVALID_ARG;
  const int groups = (n_pairs + group - 1) / group;
  match_score_kernel<<<groups, 256, 0, static_cast<cudaStream_t>(stream)>>>(mconf, counts, n_pairs, group, thr, scores, best);
  return int(cudaGetLastError());
}

extern "C" int pope_cosine_topk(const void* q, const void* refs, int dtype, int R, int D, int k, float eps,
                                float* scores, float* slot_scores, int32_t* slot_idx, void* stream) {
  if (!q || !refs || !scores || !slot_scores || !slot_idx || R <= 0 || D <= 0 || k <= 0) return POPE_ERR_INVALID_ARG;
  if (dtype != POPE_F32 && dtype != POPE_BF16) return POPE_ERR_DTYPE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int warps = 8;
  const unsigned blocks = unsigned((R + warps - 1) / warps);
  if (dtype == POPE_BF16)
    cosine_kernel<__nv_bfloat16><<<blocks, warps * 32, 0, st>>>(static_cast<const __nv_bfloat16*>(q),
                                                               static_cast<const __nv_bfloat16*>(refs), R, D, eps, scores);
  else
    cosine_kernel<float><<<blocks, warps * 32, 0, st>>>(static_cast<const float*>(q), static_cast<const float*>(refs), R,
                                                       D, eps, scores);
  running_topk_kernel<<<1, 32, 0, st>>>(scores, R, k, slot_scores, slot_idx);
  return int(cudaGetLastError());
}
