// Developer microtest: which (lane, column) does each register of tcgen05.ld.16x256b.x4 hold?
// Fills 32 lanes x 32 columns of TMEM with tcgen05.st.32x32b (lane = thread, value = lane*100 + col), reads them back with
// two .16x256b.x4 loads (lanes 0-15 and 16-31) and prints the map.  nvcc -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void k(float* out) {
  __shared__ uint32_t tptr;
  const int lane = threadIdx.x;
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&tptr)), "r"(32u) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t t = tptr;
  uint32_t v[32];
  for (int c = 0; c < 32; ++c) v[c] = __float_as_uint(float(lane * 100 + c));
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n\ttcgen05.wait::st.sync.aligned;"
      ::"r"(t), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]),
        "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]),
        "r"(v[30]), "r"(v[31]) : "memory");
  __syncwarp();
  for (int half = 0; half < 2; ++half) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(t + (uint32_t(half * 16) << 16)) : "memory");
    for (int i = 0; i < 16; ++i) out[(half * 32 + lane) * 16 + i] = __uint_as_float(r[i]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(t), "r"(32u) : "memory");
}

int main() {
  float* d;
  cudaMalloc(&d, 2 * 32 * 16 * 4);
  k<<<1, 32>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status %s\n", cudaGetErrorString(e));
  static float h[2 * 32 * 16];
  cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
  for (int half = 0; half < 2; ++half)
    for (int lane = 0; lane < 32; lane += (lane < 8 ? 1 : 8)) {
      printf("half %d lane %2d:", half, lane);
      for (int i = 0; i < 16; ++i) printf(" %5.0f", h[(half * 32 + lane) * 16 + i]);
      printf("\n");
    }
  // check the conjectured map: reg 4k+{0,1} -> (lane/4, 8k + 2(lane%4) + {0,1}), reg 4k+{2,3} -> row + 8
  int bad = 0;
  for (int half = 0; half < 2; ++half)
    for (int lane = 0; lane < 32; ++lane)
      for (int i = 0; i < 16; ++i) {
        const int kk = i >> 2, row = half * 16 + lane / 4 + ((i & 2) ? 8 : 0), col = 8 * kk + 2 * (lane % 4) + (i & 1);
        bad += h[(half * 32 + lane) * 16 + i] != float(row * 100 + col);
      }
  printf("conjecture mismatches: %d\n", bad);
  return 0;
}
