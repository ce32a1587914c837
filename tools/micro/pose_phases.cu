// Developer diagnostic: clock-stamps the phases of pm::five_point for one warp of random samples.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -fmad=false --expt-relaxed-constexpr -DPM_TRACE -I pope_b200/csrc \
//        -o tools/micro/pose_phases tools/micro/pose_phases.cu && tools/micro/pose_phases
#include <cstdio>
#include <cstdlib>
#include "pose_math.cuh"

__global__ void k(const double* pts, double* out, int* nm) {
    const int t = threadIdx.x;
    double x0[5], y0[5], x1[5], y1[5];
    for (int i = 0; i < 5; ++i) { x0[i] = pts[(t * 5 + i) * 4]; y0[i] = pts[(t * 5 + i) * 4 + 1]; x1[i] = pts[(t * 5 + i) * 4 + 2]; y1[i] = pts[(t * 5 + i) * 4 + 3]; }
    nm[t] = pm::five_point(x0, y0, x1, y1, reinterpret_cast<double (*)[9]>(out + t * 90));
}

int main() {
    const int T = 32;
    double h[T * 20];
    srand(1);
    // a consistent scene: random rotation about y plus translation, so that the solver finds real roots
    for (int i = 0; i < T * 5; ++i) {
        const double X = rand() / (double)RAND_MAX * 2 - 1, Y = rand() / (double)RAND_MAX * 2 - 1, Z = 3 + 3 * rand() / (double)RAND_MAX;
        const double c = 0.95, s = 0.3122;
        const double X1 = c * X + s * Z + 0.3, Y1 = Y + 0.1, Z1 = -s * X + c * Z + 0.2;
        h[i * 4] = X / Z; h[i * 4 + 1] = Y / Z; h[i * 4 + 2] = X1 / Z1; h[i * 4 + 3] = Y1 / Z1;
    }
    double *d, *o; int* n;
    cudaMalloc(&d, sizeof(h)); cudaMalloc(&o, T * 90 * 8); cudaMalloc(&n, T * 4);
    cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice);
    for (int rep = 0; rep < 3; ++rep) k<<<1, T>>>(d, o, n);
    cudaDeviceSynchronize();
    long long clk[8]; int hn[T];
    cudaMemcpyFromSymbol(clk, pm::pm_trace_clk, sizeof(clk));
    cudaMemcpy(hn, n, sizeof(hn), cudaMemcpyDeviceToHost);
    const char* names[] = {"null space", "constraints", "gauss-jordan", "B(z) + det poly", "real roots", "models"};
    for (int i = 0; i < 6; ++i) printf("%-16s %8lld clk\n", names[i], clk[i + 1] - clk[i]);
    printf("total            %8lld clk; models per sample:", clk[6] - clk[0]);
    for (int i = 0; i < T; ++i) printf(" %d", hn[i]);
    printf("\n%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
