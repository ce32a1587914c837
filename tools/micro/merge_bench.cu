// Developer micro-benchmark: configurations of the column log-sum-exp merge (coarse_finalize.cu::colsum_reduce_kernel) on a
// synthetic [n, ngroups, S] array with equal shifts.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/micro/merge_bench tools/micro/merge_bench.cu && tools/micro/merge_bench
#include <cstdio>
#include <cuda_runtime.h>
#include <math.h>

__device__ __forceinline__ float scale_pow2(float x, float d) {
  const int b = __float_as_int(x);
  if (!(x > 0.f) || b >= 0x7f800000) return x;
  if (!(d > -280.f)) return 0.f;
  const int e = b + (int(d) << 23);
  return e >= 0x00800000 ? __int_as_float(e) : 0.f;
}
__device__ __forceinline__ void add_one(float (&acc)[4], float& mtop, const float4 t, float m) {
  const float q[4] = {t.x, t.y, t.z, t.w};
  if (m == mtop) {
#pragma unroll
    for (int v = 0; v < 4; ++v) acc[v] += q[v];
  } else if (m < mtop) {
#pragma unroll
    for (int v = 0; v < 4; ++v) acc[v] += scale_pow2(q[v], m - mtop);
  } else {
#pragma unroll
    for (int v = 0; v < 4; ++v) acc[v] = scale_pow2(acc[v], mtop - m) + q[v];
    mtop = m;
  }
}

// W warps per block (each: 32 lanes x 4 columns), warp w takes groups w, w + W, ...; B loads in flight; smem merge
template <int W, int B>
__global__ void __launch_bounds__(32 * W) merge_smem(const float* __restrict__ colpart, const float* __restrict__ cshift, int ngroups,
                                                     int S, int nblk, float* __restrict__ lse_c) {
  __shared__ float s_acc[W][128];
  __shared__ float s_m[W][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int j = (blockIdx.x * 32 + lane) * 4, n = blockIdx.y;
  float acc[4] = {0.f, 0.f, 0.f, 0.f}, mtop = -INFINITY;
  if (j < S) {
    const float* p = colpart + size_t(n) * ngroups * S + j;
    const float* sh = cshift + size_t(n) * ngroups * nblk + (j >> 5);
    for (int g0 = warp; g0 < ngroups; g0 += W * B) {
      float4 q[B]; float m[B];
#pragma unroll
      for (int b = 0; b < B; ++b) {
        const int g = g0 + b * W;
        m[b] = -INFINITY; q[b] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (g < ngroups) { q[b] = __ldcs(reinterpret_cast<const float4*>(p + size_t(g) * S)); m[b] = __ldg(sh + size_t(g) * nblk); }
      }
#pragma unroll
      for (int b = 0; b < B; ++b) add_one(acc, mtop, q[b], m[b]);
    }
  }
#pragma unroll
  for (int v = 0; v < 4; ++v) s_acc[warp][lane * 4 + v] = acc[v];
  s_m[warp][lane] = mtop;
  __syncthreads();
  if (warp != 0 || j >= S) return;
  float mall = -INFINITY;
#pragma unroll
  for (int w = 0; w < W; ++w) mall = fmaxf(mall, s_m[w][lane]);
#pragma unroll
  for (int v = 0; v < 4; ++v) {
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < W; ++w) { const float a = s_acc[w][lane * 4 + v], m = s_m[w][lane]; tot += (m == mall) ? a : scale_pow2(a, m - mall); }
    lse_c[size_t(n) * S + j + v] = mall + log2f(tot);
  }
}

// one warp = 8 column quads (128 B of a row) x 4 group phases, shuffle merge; B loads in flight
template <int B, int TPB>
__global__ void __launch_bounds__(TPB) merge_shfl(const float* __restrict__ colpart, const float* __restrict__ cshift, int ngroups, int S,
                                                  int nblk, float* __restrict__ lse_c) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int cq = lane & 7, ph = lane >> 3;
  const int j = ((blockIdx.x * (TPB / 32) + wib) * 8 + cq) * 4, n = blockIdx.y;
  float acc[4] = {0.f, 0.f, 0.f, 0.f}, mtop = -INFINITY;
  if (j < S) {
    const float* p = colpart + size_t(n) * ngroups * S + j;
    const float* sh = cshift + size_t(n) * ngroups * nblk + (j >> 5);
    for (int g0 = ph; g0 < ngroups; g0 += 4 * B) {
      float4 q[B]; float m[B];
#pragma unroll
      for (int b = 0; b < B; ++b) {
        const int g = g0 + b * 4;
        m[b] = -INFINITY; q[b] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (g < ngroups) { q[b] = __ldcs(reinterpret_cast<const float4*>(p + size_t(g) * S)); m[b] = __ldg(sh + size_t(g) * nblk); }
      }
#pragma unroll
      for (int b = 0; b < B; ++b) add_one(acc, mtop, q[b], m[b]);
    }
  }
  // merge the 4 phases (lanes cq, cq + 8, cq + 16, cq + 24): fixed butterfly order
#pragma unroll
  for (int o = 8; o <= 16; o <<= 1) {
    const float mo = __shfl_xor_sync(0xffffffffu, mtop, o);
    const float mm = fmaxf(mtop, mo);
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const float ao = __shfl_xor_sync(0xffffffffu, acc[v], o);
      const float a0 = (mtop == mm) ? acc[v] : scale_pow2(acc[v], mtop - mm), a1 = (mo == mm) ? ao : scale_pow2(ao, mo - mm);
      acc[v] = (lane & o) ? a1 + a0 : a0 + a1;     // same operand order in both partners
    }
    mtop = mm;
  }
  if (ph == 0 && j < S) {
#pragma unroll
    for (int v = 0; v < 4; ++v) lse_c[size_t(n) * S + j + v] = mtop + log2f(acc[v]);
  }
}


// NV float4 vectors per lane, adjacent in memory (a warp covers NV * 512 B of a row), W warps split the groups, B loads of each in flight
template <int W, int B, int NV>
__global__ void __launch_bounds__(32 * W) merge_wide(const float* __restrict__ colpart, const float* __restrict__ cshift, int ngroups,
                                                     int S, int nblk, float* __restrict__ lse_c) {
  __shared__ float s_acc[W][128 * NV];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int j0 = blockIdx.x * 128 * NV, n = blockIdx.y;
  float acc[NV][4];
#pragma unroll
  for (int k = 0; k < NV; ++k) { acc[k][0] = acc[k][1] = acc[k][2] = acc[k][3] = 0.f; }
  const float* p = colpart + size_t(n) * ngroups * S + j0 + lane * 4;
  for (int g0 = warp; g0 < ngroups; g0 += W * B) {
    float4 q[B][NV];
#pragma unroll
    for (int b = 0; b < B; ++b)
#pragma unroll
      for (int k = 0; k < NV; ++k) {
        const int g = g0 + b * W, j = j0 + lane * 4 + k * 128;
        q[b][k] = (g < ngroups && j < S) ? __ldcs(reinterpret_cast<const float4*>(p + size_t(g) * S + k * 128)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
    for (int b = 0; b < B; ++b)
#pragma unroll
      for (int k = 0; k < NV; ++k) { acc[k][0] += q[b][k].x; acc[k][1] += q[b][k].y; acc[k][2] += q[b][k].z; acc[k][3] += q[b][k].w; }
  }
#pragma unroll
  for (int k = 0; k < NV; ++k)
#pragma unroll
    for (int v = 0; v < 4; ++v) s_acc[warp][k * 128 + lane * 4 + v] = acc[k][v];
  __syncthreads();
  for (int c = threadIdx.x; c < 128 * NV; c += 32 * W) {
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < W; ++w) tot += s_acc[w][c];
    if (j0 + c < S) lse_c[size_t(n) * S + j0 + c] = log2f(tot);
  }
}

int main() {
  const int n = 64, L = 4800, S = 4800, ng = (L + 31) / 32, nblk = (S + 31) / 32;
  float *cp, *cs, *out, *flush;
  const size_t ncp = size_t(n) * ng * S;
  cudaMalloc(&cp, ncp * 4); cudaMalloc(&cs, size_t(n) * ng * nblk * 4); cudaMalloc(&out, size_t(n) * S * 4);
  cudaMalloc(&flush, 512u << 20);
  cudaMemset(cp, 0x3c, ncp * 4); cudaMemset(cs, 0, size_t(n) * ng * nblk * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto time = [&](const char* name, auto launch) {
    float best = 1e9f;
    for (int r = 0; r < 5; ++r) {
      cudaMemsetAsync(flush, r, 512u << 20);
      cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); best = fminf(best, ms);
    }
    printf("%-28s %7.1f us  (%.0f GB/s)  %s\n", name, best * 1e3f, ncp * 4 / (best * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
  };
  dim3 grid((S / 4 + 31) / 32, n);
  time("smem W=8  B=1", [&] { merge_smem<8, 1><<<grid, 256>>>(cp, cs, ng, S, nblk, out); });
  time("smem W=8  B=4", [&] { merge_smem<8, 4><<<grid, 256>>>(cp, cs, ng, S, nblk, out); });
  time("smem W=8  B=5", [&] { merge_smem<8, 5><<<grid, 256>>>(cp, cs, ng, S, nblk, out); });
  time("smem W=4  B=5", [&] { merge_smem<4, 5><<<grid, 128>>>(cp, cs, ng, S, nblk, out); });
  time("smem W=4  B=10", [&] { merge_smem<4, 10><<<grid, 128>>>(cp, cs, ng, S, nblk, out); });
  time("smem W=2  B=10", [&] { merge_smem<2, 10><<<grid, 64>>>(cp, cs, ng, S, nblk, out); });
  time("smem W=16 B=5", [&] { merge_smem<16, 5><<<grid, 512>>>(cp, cs, ng, S, nblk, out); });
  time("smem W=16 B=2", [&] { merge_smem<16, 2><<<grid, 512>>>(cp, cs, ng, S, nblk, out); });
  { dim3 g((S / 32 + 3) / 4, n); time("shfl B=5 128thr", [&] { merge_shfl<5, 128><<<g, 128>>>(cp, cs, ng, S, nblk, out); }); }
  { dim3 g((S / 32 + 3) / 4, n); time("shfl B=10 128thr", [&] { merge_shfl<10, 128><<<g, 128>>>(cp, cs, ng, S, nblk, out); }); }
  { dim3 g((S / 32 + 7) / 8, n); time("shfl B=5 256thr", [&] { merge_shfl<5, 256><<<g, 256>>>(cp, cs, ng, S, nblk, out); }); }
  { dim3 g((S / 32 + 1) / 2, n); time("shfl B=8 64thr", [&] { merge_shfl<8, 64><<<g, 64>>>(cp, cs, ng, S, nblk, out); }); }
  { dim3 g((S + 255) / 256, n); time("wide W=4 B=4 NV=2 (no shifts)", [&] { merge_wide<4, 4, 2><<<g, 128>>>(cp, cs, ng, S, nblk, out); }); }
  { dim3 g((S + 511) / 512, n); time("wide W=4 B=2 NV=4 (no shifts)", [&] { merge_wide<4, 2, 4><<<g, 128>>>(cp, cs, ng, S, nblk, out); }); }
  { dim3 g((S + 511) / 512, n); time("wide W=8 B=2 NV=4 (no shifts)", [&] { merge_wide<8, 2, 4><<<g, 256>>>(cp, cs, ng, S, nblk, out); }); }
  { dim3 g((S + 127) / 128, n); time("wide W=4 B=5 NV=1 (no shifts)", [&] { merge_wide<4, 5, 1><<<g, 128>>>(cp, cs, ng, S, nblk, out); }); }
  return 0;
}
