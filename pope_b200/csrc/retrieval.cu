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

}  // namespace
}  // namespace pope

using namespace pope;

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
