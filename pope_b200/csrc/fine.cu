// fine.cu -- fine-level window gather and fine matching (HBM-bound, no tensor cores).
//
// pope_fine_gather replaces F.unfold(5x5, stride 4, pad 2) + advanced indexing of
//   src/matcher/loftr_module/fine_preprocess.py:40-47: only the M matched windows are read, straight from the
//   feature maps; the 25x unfold blow-up (2 x 61 MB per 480x640 pair) is never materialised.
// pope_fine_match replaces src/matcher/utils/fine_matching.py:43-57 and :62-74 (~10 small ATen kernels) with one
//   warp per match: 128-bit loads, warp-shuffle transpose-reduce of the 25 dot products, softmax + expectation in
//   registers.
#include "common.cuh"

namespace pope {
namespace {

__device__ __forceinline__ uint4 ld_stream16(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ uint2 ld_stream8(const void* p) {
  uint2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}

struct MapDesc {
  const char* base;
  int H, W, wc;              // map height/width (fine cells), coarse grid width
  int64_t sN, sC, sH, sW;    // element strides
};

// ---- channels-last gather: one warp per (match, image); a window pixel is one contiguous Cf-vector -----------
// VEC = 16-byte vectors per pixel (Cf * sizeof(T) / 16): 32 for fp32, 16 for bf16 at Cf = 128.
template <int VEC>
__global__ void __launch_bounds__(256) gather_cl_kernel(MapDesc m0, MapDesc m1, int esize, int stride, int W,
                                                       const int64_t* __restrict__ b_ids,
                                                       const int64_t* __restrict__ i_ids,
                                                       const int64_t* __restrict__ j_ids, int64_t M,
                                                       const int32_t* __restrict__ m_dev, char* __restrict__ win0,
                                                       char* __restrict__ win1) {
  const int lane = threadIdx.x & 31;
  const int64_t job = int64_t(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);   // 2 jobs per match
  const int64_t live = m_dev ? min(int64_t(*m_dev), M) : M;
  const int64_t m = job >> 1;
  if (m >= live) return;
  const bool second = job & 1;
  const MapDesc& md = second ? m1 : m0;
  const int cell = int(second ? j_ids[m] : i_ids[m]);
  const int64_t b = b_ids[m];
  const int cy = cell / md.wc, cx = cell - cy * md.wc;
  const int y0 = cy * stride - W / 2, x0 = cx * stride - W / 2;
  const int WW = W * W;
  constexpr int PIX_PER_IT = 32 / VEC;              // window pixels copied per warp iteration
  const int sub = lane / VEC, v = lane % VEC;
  char* out = (second ? win1 : win0) + size_t(m) * WW * VEC * 16;
  const char* img = md.base + size_t(b) * md.sN * esize;
  constexpr int UNROLL = 5;
  for (int p0 = 0; p0 < WW; p0 += PIX_PER_IT * UNROLL) {
    uint4 val[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const int p = p0 + u * PIX_PER_IT + sub;
      val[u] = make_uint4(0u, 0u, 0u, 0u);
      if (p < WW) {
        const int ky = p / W, kx = p - ky * W;
        const int y = y0 + ky, x = x0 + kx;
        if (y >= 0 && y < md.H && x >= 0 && x < md.W)
          val[u] = ld_stream16(img + (size_t(y) * md.sH + size_t(x) * md.sW) * esize + size_t(v) * 16);
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const int p = p0 + u * PIX_PER_IT + sub;
      if (p < WW) *reinterpret_cast<uint4*>(out + (size_t(p) * VEC + v) * 16) = val[u];
    }
  }
}

// ---- generic strided gather (plain NCHW maps): one thread per output element, coalesced on the write side --------
template <typename T>
__global__ void __launch_bounds__(256) gather_strided_kernel(MapDesc m0, MapDesc m1, int Cf, int stride, int W,
                                                            const int64_t* __restrict__ b_ids,
                                                            const int64_t* __restrict__ i_ids,
                                                            const int64_t* __restrict__ j_ids, int64_t M,
                                                            const int32_t* __restrict__ m_dev, T* __restrict__ win0,
                                                            T* __restrict__ win1) {
  const int64_t live = m_dev ? min(int64_t(*m_dev), M) : M;
  const int WW = W * W;
  const int64_t per_img = live * WW * Cf;
  for (int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; e < 2 * per_img;
       e += int64_t(gridDim.x) * blockDim.x) {
    const bool second = e >= per_img;
    const int64_t r = second ? e - per_img : e;
    const int c = int(r % Cf);
    const int p = int((r / Cf) % WW);
    const int64_t m = r / (int64_t(Cf) * WW);
    const MapDesc& md = second ? m1 : m0;
    const int64_t cell = second ? j_ids[m] : i_ids[m];
    const int cy = int(cell / md.wc), cx = int(cell - int64_t(cy) * md.wc);
    const int ky = p / W, kx = p - ky * W;
    const int y = cy * stride - W / 2 + ky, x = cx * stride - W / 2 + kx;
    T val = T(0.f);
    if (y >= 0 && y < md.H && x >= 0 && x < md.W)
      val = reinterpret_cast<const T*>(md.base)[b_ids[m] * md.sN + c * md.sC + y * md.sH + x * md.sW];
    (second ? win1 : win0)[r] = val;
  }
}

// ---- fine matching: one warp per match ---------------------------------------------------------------------------
template <typename T> struct Row4;   // each lane owns 4 consecutive channels of the 128
template <> struct Row4<float> {
  static __device__ __forceinline__ float4 load(const float* row, int lane) {
    uint4 u = ld_stream16(row + lane * 4);
    return make_float4(__uint_as_float(u.x), __uint_as_float(u.y), __uint_as_float(u.z), __uint_as_float(u.w));
  }
};
template <> struct Row4<__nv_bfloat16> {
  static __device__ __forceinline__ float4 load(const __nv_bfloat16* row, int lane) {
    uint2 u = ld_stream8(row + lane * 4);
    return make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u), __uint_as_float(u.y << 16),
                       __uint_as_float(u.y & 0xffff0000u));
  }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
  return v;
}

// Shared by both fine-match kernels: 25 per-lane partial dot products -> expectation / std / refined coordinate.
__device__ __forceinline__ void fine_match_finish(float (&p)[32], int lane, int64_t m, const float* __restrict__ mkpts1_c,
                                                  float inv_sqrt_c, float coord_scale, float* __restrict__ expec_f,
                                                  float* __restrict__ mkpts1_f) {
  constexpr int WW = 25, WIN = 5;
  // transpose-reduce: after the 5 steps lane r holds sum over lanes of p[r]   (31 shuffles instead of 25*5)
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool upper = lane & off;
#pragma unroll
    for (int k = 0; k < off; ++k) {
      const float send = upper ? p[k] : p[k + off];
      const float keep = upper ? p[k + off] : p[k];
      p[k] = keep + __shfl_xor_sync(kFullMask, send, off);
    }
  }
  const bool on = lane < WW;
  const float x = on ? p[0] * inv_sqrt_c : -INFINITY;
  float mx = x;
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(kFullMask, mx, o));
  const float e = on ? expf(x - mx) : 0.f;
  const float h = e / warp_sum(e);
  const float gx = -1.f + 0.5f * float(lane % WIN), gy = -1.f + 0.5f * float(lane / WIN);
  const float ex = warp_sum(h * gx), ey = warp_sum(h * gy);
  const float exx = warp_sum(h * gx * gx), eyy = warp_sum(h * gy * gy);
  if (lane == 0) {
    const float sd = sqrtf(fmaxf(exx - ex * ex, 1e-10f)) + sqrtf(fmaxf(eyy - ey * ey, 1e-10f));
    expec_f[3 * m + 0] = ex; expec_f[3 * m + 1] = ey; expec_f[3 * m + 2] = sd;
    mkpts1_f[2 * m + 0] = mkpts1_c[2 * m + 0] + ex * coord_scale;
    mkpts1_f[2 * m + 1] = mkpts1_c[2 * m + 1] + ey * coord_scale;
  }
}

template <typename T>
__global__ void __launch_bounds__(256) fine_match_kernel(const T* __restrict__ win0, const T* __restrict__ win1,
                                                        int64_t M, const int32_t* __restrict__ m_dev,
                                                        const float* __restrict__ mkpts1_c, float inv_sqrt_c,
                                                        float coord_scale, float* __restrict__ expec_f,
                                                        float* __restrict__ mkpts1_f) {
  constexpr int WW = 25, C = 128;
  const int lane = threadIdx.x & 31;
  const int64_t m = int64_t(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t live = m_dev ? min(int64_t(*m_dev), M) : M;
  if (m >= live) return;
  const T* w0 = win0 + size_t(m) * WW * C;
  const T* w1 = win1 + size_t(m) * WW * C;
  // all 26 row loads are issued before the first use (memory-level parallelism; the kernel is HBM-bound)
  const float4 ctr = Row4<T>::load(w0 + (WW / 2) * C, lane);
  float4 rows[WW];
#pragma unroll
  for (int r = 0; r < WW; ++r) rows[r] = Row4<T>::load(w1 + r * C, lane);
  float p[32];
#pragma unroll
  for (int r = 0; r < WW; ++r)
    p[r] = fmaf(ctr.x, rows[r].x, fmaf(ctr.y, rows[r].y, fmaf(ctr.z, rows[r].z, ctr.w * rows[r].w)));
#pragma unroll
  for (int r = WW; r < 32; ++r) p[r] = 0.f;
  fine_match_finish(p, lane, m, mkpts1_c, inv_sqrt_c, coord_scale, expec_f, mkpts1_f);
}

// Processing order for the fine stage: matches come out of the coarse stage sorted by (pair, i); their reference
// cells j are scattered over image 1, so consecutive warps would gather windows from random places of the fine map.
// One CTA per pair counting-sorts its matches by j (histogram over the S cells in shared memory, block scan, scatter):
// order[k] = index of the k-th match in (pair, j) order.  Neighbouring warps then read overlapping / adjacent windows.
__global__ void __launch_bounds__(1024) order_by_ref_kernel(const int32_t* __restrict__ counts, int n_pairs, int S,
                                                           const int64_t* __restrict__ j_ids, int32_t* __restrict__ order) {
  extern __shared__ int hist[];            // [S] counts, then exclusive offsets
  __shared__ int warp_tot[32];
  __shared__ int s_base;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    int base = 0;
    for (int p = 0; p < b; ++p) base += counts[p];
    s_base = base;
  }
  for (int j = tid; j < S; j += 1024) hist[j] = 0;
  __syncthreads();
  const int base = s_base, cnt = counts[b];
  for (int r = tid; r < cnt; r += 1024) atomicAdd(&hist[int(j_ids[base + r])], 1);
  __syncthreads();
  // exclusive scan of hist[0..S): each thread owns a contiguous slice
  const int per = (S + 1023) / 1024, lo = min(tid * per, S), hi = min(lo + per, S);
  int sum = 0;
  for (int j = lo; j < hi; ++j) sum += hist[j];
  int incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(kFullMask, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) warp_tot[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int v = warp_tot[lane], inc2 = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(kFullMask, inc2, o);
      if (lane >= o) inc2 += t;
    }
    warp_tot[lane] = inc2 - v;
  }
  __syncthreads();
  int run = warp_tot[warp] + incl - sum;
  for (int j = lo; j < hi; ++j) { const int c = hist[j]; hist[j] = run; run += c; }
  __syncthreads();
  for (int r = tid; r < cnt; r += 1024) {
    const int slot = atomicAdd(&hist[int(j_ids[base + r])], 1);
    order[base + slot] = base + r;
  }
}

// softmax over the 25 window positions + expectation / std / refined coordinate; lane holds position r (on = r < 25)
__device__ __forceinline__ void fine_match_tail(float sim, int r, bool on, int lane, int64_t m,
                                                const float* __restrict__ mkpts1_c, float inv_sqrt_c, float coord_scale,
                                                float* __restrict__ expec_f, float* __restrict__ mkpts1_f) {
  constexpr int WIN = 5;
  const float x = on ? sim * inv_sqrt_c : -INFINITY;
  float mx = x;
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(kFullMask, mx, o));
  const float e = on ? expf(x - mx) : 0.f;
  const float h = e / warp_sum(e);
  const float gx = -1.f + 0.5f * float(r % WIN), gy = -1.f + 0.5f * float(r / WIN);
  const float ex = warp_sum(h * gx), ey = warp_sum(h * gy);
  const float exx = warp_sum(h * gx * gx), eyy = warp_sum(h * gy * gy);
  if (lane == 0) {
    const float sd = sqrtf(fmaxf(exx - ex * ex, 1e-10f)) + sqrtf(fmaxf(eyy - ey * ey, 1e-10f));
    expec_f[3 * m + 0] = ex; expec_f[3 * m + 1] = ey; expec_f[3 * m + 2] = sd;
    mkpts1_f[2 * m + 0] = mkpts1_c[2 * m + 0] + ex * coord_scale;
    mkpts1_f[2 * m + 1] = mkpts1_c[2 * m + 1] + ey * coord_scale;
  }
}

template <typename T> struct Vec16;      // 16 bytes of channels -> floats
template <> struct Vec16<float> {
  static constexpr int N = 4;
  static __device__ __forceinline__ void unpack(const uint4& u, float (&f)[4]) {
    f[0] = __uint_as_float(u.x); f[1] = __uint_as_float(u.y); f[2] = __uint_as_float(u.z); f[3] = __uint_as_float(u.w);
  }
};
template <> struct Vec16<__nv_bfloat16> {
  static constexpr int N = 8;
  static __device__ __forceinline__ void unpack(const uint4& u, float (&f)[8]) {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) { f[2 * k] = __uint_as_float(w[k] << 16); f[2 * k + 1] = __uint_as_float(w[k] & 0xffff0000u); }
  }
};

// Fused window gather + fine match for the pipeline that has nothing between the two (no fine transformer): the
// centre pixel of window 0 and the 25 pixels of window 1 are read straight from the channels-last feature maps, the
// [M,25,128] windows are never written.  One warp per match; a pixel (128 channels) is LPP = 128*sizeof(T)/16 lanes
// x 16 bytes, so a warp covers 32/LPP pixels per load instruction (2 for bf16, 1 for fp32).  Loads are unconditional
// from clamped coordinates (out-of-map pixels contribute 0, like the reference's zero padding) so that all of them
// are in flight together.  Same arithmetic per channel as gather_cl_kernel + fine_match_kernel.
template <typename T>
__global__ void __launch_bounds__(256, sizeof(T) == 2 ? 4 : 2) fine_match_maps_kernel(MapDesc m0, MapDesc m1, int stride,
                                                             const int64_t* __restrict__ b_ids,
                                                             const int64_t* __restrict__ i_ids,
                                                             const int64_t* __restrict__ j_ids, int64_t M,
                                                             const int32_t* __restrict__ m_dev,
                                                             const int32_t* __restrict__ order,
                                                             const float* __restrict__ mkpts1_c, float inv_sqrt_c,
                                                             float coord_scale, float* __restrict__ expec_f,
                                                             float* __restrict__ mkpts1_f) {
  constexpr int WW = 25, WIN = 5, C = 128;
  constexpr int NCH = Vec16<T>::N;            // channels per lane
  constexpr int LPP = C / NCH;                // lanes per pixel: 32 (fp32) or 16 (bf16)
  constexpr int PPI = 32 / LPP;               // pixels per warp-wide load
  constexpr int IT = (WW + PPI - 1) / PPI;    // 25 or 13
  const int lane = threadIdx.x & 31, sub = lane / LPP, v = lane % LPP;
  const int64_t w = int64_t(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t live = m_dev ? min(int64_t(*m_dev), M) : M;
  if (w >= live) return;
  const int64_t m = order ? int64_t(order[w]) : w;      // optional processing order (sorted by reference cell)
  const int64_t b = b_ids[m];
  const int ci = int(i_ids[m]), cj = int(j_ids[m]);                                   // 32-bit: no 64-bit divisions
  const int y0c = (ci / m0.wc) * stride, x0c = (ci % m0.wc) * stride;                // centre pixel of window 0
  const int y1 = (cj / m1.wc) * stride - WIN / 2, x1 = (cj % m1.wc) * stride - WIN / 2;
  const T* img0 = reinterpret_cast<const T*>(m0.base) + b * m0.sN;
  const T* img1 = reinterpret_cast<const T*>(m1.base) + b * m1.sN;
  const uint4 craw = ld_stream16(img0 + int64_t(y0c) * m0.sH + int64_t(x0c) * m0.sW + v * NCH);
  uint4 raw[IT];
  bool ok[IT];
  if (y1 >= 0 && x1 >= 0 && y1 + WIN <= m1.H && x1 + WIN <= m1.W) {
    // window entirely inside the map (always the case when border_rm >= 1): no clamping, the pixel offset advances
    // incrementally (r -> r + PPI wraps to the next window row every 5 pixels)
    const T* p1 = img1 + int64_t(y1) * m1.sH + int64_t(x1) * m1.sW + v * NCH;
    int kx = sub;                                  // r = k * PPI + sub, (ky, kx) = divmod(r, 5); sub < 5
    int64_t off = int64_t(sub) * m1.sW;
    const int64_t wrap = m1.sH - int64_t(WIN) * m1.sW;
#pragma unroll
    for (int k = 0; k < IT; ++k) {
      ok[k] = k * PPI + sub < WW;
      raw[k] = ld_stream16(p1 + (ok[k] ? off : 0));
      kx += PPI; off += int64_t(PPI) * m1.sW;
      if (kx >= WIN) { kx -= WIN; off += wrap; }
    }
  } else {
#pragma unroll
    for (int k = 0; k < IT; ++k) {
      const int r = k * PPI + sub;
      const int y = y1 + r / WIN, x = x1 + r % WIN;
      ok[k] = r < WW && y >= 0 && y < m1.H && x >= 0 && x < m1.W;
      const int yc = min(max(y, 0), m1.H - 1), xc = min(max(x, 0), m1.W - 1);
      raw[k] = ld_stream16(img1 + int64_t(yc) * m1.sH + int64_t(xc) * m1.sW + v * NCH);
    }
  }
  float ctr[NCH];
  Vec16<T>::unpack(craw, ctr);
  float p[LPP];
#pragma unroll
  for (int k = 0; k < IT; ++k) {
    float f[NCH];
    Vec16<T>::unpack(raw[k], f);
    float d = 0.f;
#pragma unroll
    for (int c = NCH - 1; c >= 0; --c) d = fmaf(ctr[c], f[c], d);
    p[k] = ok[k] ? d : 0.f;
  }
#pragma unroll
  for (int k = IT; k < LPP; ++k) p[k] = 0.f;
  // transpose-reduce inside each group of LPP lanes: afterwards lane v of the group holds slot v = pixel v*PPI + sub
#pragma unroll
  for (int off = LPP / 2; off >= 1; off >>= 1) {
    const bool upper = v & off;
#pragma unroll
    for (int k = 0; k < off; ++k) {
      const float send = upper ? p[k] : p[k + off];
      const float keep = upper ? p[k + off] : p[k];
      p[k] = keep + __shfl_xor_sync(kFullMask, send, off);
    }
  }
  const int r = v * PPI + sub;
  fine_match_tail(p[0], r, r < WW, lane, m, mkpts1_c, inv_sqrt_c, coord_scale, expec_f, mkpts1_f);
}

}  // namespace
}  // namespace pope

using namespace pope;

extern "C" int pope_fine_gather(const void* feat_f0, const void* feat_f1, int dtype, int n_pairs, int Cf, int Hf0,
                                int Wf0, const int64_t strides0[4], int Hf1, int Wf1, const int64_t strides1[4],
                                int w0c, int w1c, int stride, int W, const int64_t* b_ids, const int64_t* i_ids,
                                const int64_t* j_ids, int64_t M, const int32_t* m_dev, void* win0, void* win1,
                                void* stream) {
  if (!feat_f0 || !feat_f1 || !strides0 || !strides1 || M < 0) return POPE_ERR_INVALID_ARG;
  if (dtype != POPE_F32 && dtype != POPE_BF16) return POPE_ERR_DTYPE;
  if (n_pairs <= 0 || Cf <= 0 || Hf0 <= 0 || Wf0 <= 0 || Hf1 <= 0 || Wf1 <= 0 || w0c <= 0 || w1c <= 0 || stride <= 0 ||
      W <= 0 || (W & 1) == 0)
    return POPE_ERR_INVALID_ARG;
  if (M == 0) return POPE_OK;
  if (!b_ids || !i_ids || !j_ids || !win0 || !win1) return POPE_ERR_INVALID_ARG;
  const int esize = dtype == POPE_BF16 ? 2 : 4;
  MapDesc m0{static_cast<const char*>(feat_f0), Hf0, Wf0, w0c, strides0[0], strides0[1], strides0[2], strides0[3]};
  MapDesc m1{static_cast<const char*>(feat_f1), Hf1, Wf1, w1c, strides1[0], strides1[1], strides1[2], strides1[3]};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int vec = Cf * esize / 16;
  auto vec_ok = [&](const MapDesc& m) {
    return m.sC == 1 && (m.sW * esize) % 16 == 0 && (m.sH * esize) % 16 == 0 && (m.sN * esize) % 16 == 0 &&
           (reinterpret_cast<uintptr_t>(m.base) & 15u) == 0;
  };
  const bool fast = (Cf * esize) % 16 == 0 && (vec == 32 || vec == 16) && vec_ok(m0) && vec_ok(m1) &&
                    ((reinterpret_cast<uintptr_t>(win0) | reinterpret_cast<uintptr_t>(win1)) & 15u) == 0;
  if (fast) {
    const int warps = 8;
    const unsigned blocks = unsigned((2 * M + warps - 1) / warps);
    if (vec == 32)
      gather_cl_kernel<32><<<blocks, warps * 32, 0, st>>>(m0, m1, esize, stride, W, b_ids, i_ids, j_ids, M, m_dev,
                                                         static_cast<char*>(win0), static_cast<char*>(win1));
    else
      gather_cl_kernel<16><<<blocks, warps * 32, 0, st>>>(m0, m1, esize, stride, W, b_ids, i_ids, j_ids, M, m_dev,
                                                         static_cast<char*>(win0), static_cast<char*>(win1));
  } else {
    const int64_t total = 2 * M * W * W * Cf;
    const int64_t want = (total + 255) / 256;
    const unsigned blocks = unsigned(want < 148 * 32 ? want : 148 * 32);
    if (dtype == POPE_BF16)
      gather_strided_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(m0, m1, Cf, stride, W, b_ids, i_ids, j_ids, M, m_dev,
                                                                  static_cast<__nv_bfloat16*>(win0),
                                                                  static_cast<__nv_bfloat16*>(win1));
    else
      gather_strided_kernel<float><<<blocks, 256, 0, st>>>(m0, m1, Cf, stride, W, b_ids, i_ids, j_ids, M, m_dev,
                                                          static_cast<float*>(win0), static_cast<float*>(win1));
  }
  return int(cudaGetLastError());
}

extern "C" int pope_fine_match(const void* win0, const void* win1, int dtype, int64_t M, const int32_t* m_dev, int WW,
                               int Cf, const float* mkpts1_c, float coord_scale, float* expec_f, float* mkpts1_f,
                               void* stream) {
  if (M < 0) return POPE_ERR_INVALID_ARG;
  if (dtype != POPE_F32 && dtype != POPE_BF16) return POPE_ERR_DTYPE;
  if (WW != 25 || Cf != 128) return POPE_ERR_SHAPE;
  if (M == 0) return POPE_OK;
  if (!win0 || !win1 || !mkpts1_c || !expec_f || !mkpts1_f) return POPE_ERR_INVALID_ARG;
  if ((reinterpret_cast<uintptr_t>(win0) | reinterpret_cast<uintptr_t>(win1)) & 15u) return POPE_ERR_ALIGNMENT;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int warps = 8;
  const unsigned blocks = unsigned((M + warps - 1) / warps);
  const float inv_sqrt_c = static_cast<float>(1.0 / sqrt(double(Cf)));
  if (dtype == POPE_BF16)
    fine_match_kernel<__nv_bfloat16><<<blocks, warps * 32, 0, st>>>(static_cast<const __nv_bfloat16*>(win0),
                                                                   static_cast<const __nv_bfloat16*>(win1), M, m_dev,
                                                                   mkpts1_c, inv_sqrt_c, coord_scale, expec_f, mkpts1_f);
  else
    fine_match_kernel<float><<<blocks, warps * 32, 0, st>>>(static_cast<const float*>(win0),
                                                           static_cast<const float*>(win1), M, m_dev, mkpts1_c,
                                                           inv_sqrt_c, coord_scale, expec_f, mkpts1_f);
  return int(cudaGetLastError());
}

extern "C" int pope_fine_match_maps(const void* feat_f0, const void* feat_f1, int dtype, int n_pairs, int Cf, int Hf0,
                                    int Wf0, const int64_t strides0[4], int Hf1, int Wf1, const int64_t strides1[4],
                                    int w0c, int w1c, int stride, int W, const int64_t* b_ids, const int64_t* i_ids,
                                    const int64_t* j_ids, int64_t M, const int32_t* m_dev, const int32_t* order,
                                    const float* mkpts1_c, float coord_scale, float* expec_f, float* mkpts1_f,
                                    void* stream) {
  if (!feat_f0 || !feat_f1 || !strides0 || !strides1 || M < 0) return POPE_ERR_INVALID_ARG;
  if (dtype != POPE_F32 && dtype != POPE_BF16) return POPE_ERR_DTYPE;
  if (n_pairs <= 0 || Hf0 <= 0 || Wf0 <= 0 || Hf1 <= 0 || Wf1 <= 0 || w0c <= 0 || w1c <= 0 || stride <= 0)
    return POPE_ERR_INVALID_ARG;
  if (W != 5 || Cf != 128) return POPE_ERR_SHAPE;
  if (strides0[1] != 1 || strides1[1] != 1) return POPE_ERR_SHAPE;        // channels-last maps only
  if (M == 0) return POPE_OK;
  if (!b_ids || !i_ids || !j_ids || !mkpts1_c || !expec_f || !mkpts1_f) return POPE_ERR_INVALID_ARG;
  const int esize = dtype == POPE_BF16 ? 2 : 4;
  for (int k = 0; k < 4; ++k)
    if (k != 1 && ((strides0[k] * esize) % 16 != 0 || (strides1[k] * esize) % 16 != 0)) return POPE_ERR_ALIGNMENT;
  if ((reinterpret_cast<uintptr_t>(feat_f0) | reinterpret_cast<uintptr_t>(feat_f1)) & 15u) return POPE_ERR_ALIGNMENT;
  MapDesc m0{static_cast<const char*>(feat_f0), Hf0, Wf0, w0c, strides0[0], strides0[1], strides0[2], strides0[3]};
  MapDesc m1{static_cast<const char*>(feat_f1), Hf1, Wf1, w1c, strides1[0], strides1[1], strides1[2], strides1[3]};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int warps = 8;
  const unsigned blocks = unsigned((M + warps - 1) / warps);
  const float inv_sqrt_c = static_cast<float>(1.0 / sqrt(double(Cf)));
  if (dtype == POPE_BF16)
    fine_match_maps_kernel<__nv_bfloat16><<<blocks, warps * 32, 0, st>>>(m0, m1, stride, b_ids, i_ids, j_ids, M, m_dev,
                                                                        order, mkpts1_c, inv_sqrt_c, coord_scale,
                                                                        expec_f, mkpts1_f);
  else
    fine_match_maps_kernel<float><<<blocks, warps * 32, 0, st>>>(m0, m1, stride, b_ids, i_ids, j_ids, M, m_dev, order,
                                                                mkpts1_c, inv_sqrt_c, coord_scale, expec_f, mkpts1_f);
  return int(cudaGetLastError());
}

extern "C" int pope_match_order_by_ref(const int32_t* counts, int n_pairs, int S, const int64_t* j_ids, int32_t* order,
                                       void* stream) {
  if (!counts || !j_ids || !order || n_pairs <= 0 || S <= 0) return POPE_ERR_INVALID_ARG;
  const size_t smem = size_t(S) * sizeof(int);
  if (smem > 200 * 1024) return POPE_ERR_SHAPE;
  cudaError_t e = cudaFuncSetAttribute(order_by_ref_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
  if (e != cudaSuccess) return int(e);
  order_by_ref_kernel<<<n_pairs, 1024, smem, static_cast<cudaStream_t>(stream)>>>(counts, n_pairs, S, j_ids, order);
  return int(cudaGetLastError());
}
