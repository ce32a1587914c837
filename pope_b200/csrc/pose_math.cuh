// Scalar fp64 math of the batched pose solver (pose.cu): counter-based sampling, cyclic Jacobi eigen-solver, the five-point
// relative-pose solver (Nister 2004: null space -> ten cubic constraints -> Gauss-Jordan -> 3x3 polynomial matrix B(z) ->
// tenth-degree polynomial -> real roots), the Sampson inlier test, the essential-matrix decomposition and the DLT
// triangulation of the cheirality test.  Replaces what the reference gets from OpenCV 4.x behind
// src/utils/metrics.py:69-94 (cv2.findEssentialMat(..., method=cv2.RANSAC) + cv2.recoverPose).
//
// Everything here is written so that oracle/pose_oracle.py can restate it operation by operation: plain IEEE add / mul /
// div / sqrt in a fixed order, no fused multiply-add (pose.cu is compiled with -fmad=false), no libm calls other than
// sqrt / fabs.  The file also compiles as host C++ (tests/hostbuild) so that the CPU suite can check the very same source
// against the oracle without a GPU; the product only ever runs it on the device.
#pragma once
#include <stdint.h>
#include <math.h>

#ifdef __CUDACC__
#define PM_HD __host__ __device__ __forceinline__
#define PM_HDN __host__ __device__ __noinline__
#else
#define PM_HD inline
#define PM_HDN inline
#endif

// developer diagnostics (tools/micro/pose_phases.cu): clock stamps after each phase of the solver
#if defined(PM_TRACE) && defined(__CUDA_ARCH__)
#define PM_STAMP(k) pm::pm_trace_clk[k] = clock64()
#else
#define PM_STAMP(k)
#endif

namespace pm {

#if defined(PM_TRACE) && defined(__CUDACC__)
__device__ long long pm_trace_clk[8];
#endif

constexpr int kMaxModels = 10;    // a tenth-degree polynomial has at most ten real roots
constexpr int kGridCells = 128;   // sign-change cells on [-1, 1], for p(z) and for the reversed polynomial
constexpr int kBisect = 40;       // bisection steps per bracket: 2^-6 * 2^-40 wide at the end
constexpr int kMaxDraws = 64;     // hashed draws per minimal sample before the sample is given up

// ---- sampling ------------------------------------------------------------------------------------------------------------
PM_HD uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// Five distinct match indices of sample h of pair `pair`: hashed draws, duplicates rejected in draw order.
PM_HD bool draw5(uint64_t seed, uint64_t pair, uint64_t h, int m, int idx[5]) {
    const uint64_t base = ((seed * 0x100000001B3ull + pair) * 0x100000001B3ull + h) * (uint64_t)kMaxDraws;
    int got = 0;
    for (int d = 0; d < kMaxDraws && got < 5; ++d) {
        const int v = (int)(splitmix64(base + (uint64_t)d) % (uint64_t)m);
        bool dup = false;
        for (int j = 0; j < got; ++j) dup |= (idx[j] == v);
        if (!dup) idx[got++] = v;
    }
    return got == 5;
}

// ---- cyclic Jacobi for a symmetric N x N matrix: A -> diag(eigenvalues), V columns = eigenvectors ------------------------
template <int N, int SWEEPS>
PM_HD void jacobi_eig(double (&A)[N][N], double (&V)[N][N]) {
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) V[i][j] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < SWEEPS; ++sweep)
        for (int p = 0; p < N - 1; ++p)
            for (int q = p + 1; q < N; ++q) {
                const double apq = A[p][q];
                double t = 0.0, c = 1.0, s = 0.0;
                if (apq != 0.0) {
                    const double theta = (A[q][q] - A[p][p]) / (2.0 * apq);
                    t = 1.0 / (fabs(theta) + sqrt(theta * theta + 1.0));
                    if (theta < 0.0) t = -t;
                    c = 1.0 / sqrt(t * t + 1.0);
                    s = t * c;
                }
                A[p][p] = A[p][p] - t * apq;
                A[q][q] = A[q][q] + t * apq;
                A[p][q] = 0.0;
                A[q][p] = 0.0;
                for (int k = 0; k < N; ++k) {
                    if (k != p && k != q) {
                        const double akp = A[k][p], akq = A[k][q];
                        const double np_ = c * akp - s * akq, nq_ = s * akp + c * akq;
                        A[k][p] = np_; A[p][k] = np_;
                        A[k][q] = nq_; A[q][k] = nq_;
                    }
                }
                for (int k = 0; k < N; ++k) {
                    const double vkp = V[k][p], vkq = V[k][q];
                    V[k][p] = c * vkp - s * vkq;
                    V[k][q] = s * vkp + c * vkq;
                }
            }
}

// ---- polynomial bookkeeping of the five-point solver ---------------------------------------------------------------------
// Monomials of the ten cubic constraints in Nister's order:
//  0:x^3 1:y^3 2:x^2y 3:xy^2 4:x^2z 5:x^2 6:y^2z 7:y^2 8:xyz 9:xy | 10:xz^2 11:xz 12:x 13:yz^2 14:yz 15:y 16:z^3 17:z^2 18:z 19:1
// degree-1 polynomials are stored as [x, y, z, 1]; degree-2 as [x^2, y^2, xy, xz, x, yz, y, z^2, z, 1].
// The tables are only ever indexed by fully unrolled loop counters, so every index below folds to a literal.
PM_HD constexpr int mul11(int a, int b) {
    constexpr int8_t t[4][4] = {{0, 2, 3, 4}, {2, 1, 5, 6}, {3, 5, 7, 8}, {4, 6, 8, 9}};
    return t[a][b];
}
PM_HD constexpr int mul21(int a, int b) {
    constexpr int8_t t[10][4] = {{0, 2, 4, 5},   {3, 1, 6, 7},    {2, 3, 8, 9},    {4, 8, 10, 11},  {5, 9, 11, 12},
                                 {8, 6, 13, 14}, {9, 7, 14, 15}, {10, 13, 16, 17}, {11, 14, 17, 18}, {12, 15, 18, 19}};
    return t[a][b];
}

// r (degree 2) += sign * p * q, p and q of degree 1
PM_HD void acc11(double* r, const double* p, const double* q, double sign) {
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) r[mul11(a, b)] = r[mul11(a, b)] + sign * (p[a] * q[b]);
}
// r (degree 3, 20 monomials) += p (degree 2) * q (degree 1)
PM_HD void acc21(double* r, const double* p, const double* q) {
#pragma unroll
    for (int a = 0; a < 10; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) r[mul21(a, b)] = r[mul21(a, b)] + p[a] * q[b];
}

PM_HD double horner(const double* c, int deg, double z) {
    double v = c[deg];
    for (int k = deg - 1; k >= 0; --k) v = v * z + c[k];
    return v;
}

// out[0..da+db] (+)= sign * a * b for univariate polynomials (ascending coefficients); out must be initialised by the caller
PM_HD void conv_acc(double* out, const double* a, int da, const double* b, int db, double sign) {
    for (int i = 0; i <= da; ++i)
        for (int j = 0; j <= db; ++j) out[i + j] = out[i + j] + sign * (a[i] * b[j]);
}

// Real roots of a degree-10 polynomial (ascending coefficients c[0..10], any scale): sign changes of p on a uniform grid of
// [-1, 1] and of the reversed polynomial u^10 p(1/u) on the same grid (roots with |z| > 1), each bracket bisected kBisect
// times.  Returns the number of roots (<= 10), in grid order.  Two roots inside one cell are not separated.  The brackets
// are collected first and bisected afterwards so that the lanes of a warp bisect their k-th brackets in lockstep.
PM_HD int real_roots10(const double* c, double* roots) {
    int n = 0;
    int bracket[kMaxModels];
    double rev[11];
    for (int k = 0; k <= 10; ++k) rev[k] = c[10 - k];
    for (int part = 0; part < 2; ++part) {
        const double* p = part == 0 ? c : rev;
        bool slo = horner(p, 10, -1.0) > 0.0;
        for (int cell = 0; cell < kGridCells; cell += 4) {          // four independent Horner chains at a time
            double z[4], v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) { z[u] = (double)(cell + u + 1) * (2.0 / kGridCells) - 1.0; v[u] = p[10]; }
#pragma unroll
            for (int k = 9; k >= 0; --k) {
                const double ck = p[k];
#pragma unroll
                for (int u = 0; u < 4; ++u) v[u] = v[u] * z[u] + ck;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const bool shi = v[u] > 0.0;
                if (shi != slo && n < kMaxModels) bracket[n++] = part * kGridCells + cell + u;
                slo = shi;
            }
        }
    }
    for (int r = 0; r < n; r += 2) {                                 // two brackets at a time
        const double* p[2];
        double a[2], b[2];
        bool slo[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int br = bracket[r + u < n ? r + u : r], cell = br % kGridCells;
            p[u] = br / kGridCells == 0 ? c : rev;
            a[u] = (double)cell * (2.0 / kGridCells) - 1.0;
            b[u] = (double)(cell + 1) * (2.0 / kGridCells) - 1.0;
            slo[u] = horner(p[u], 10, a[u]) > 0.0;
        }
        for (int it = 0; it < kBisect; ++it) {
            double mid[2], v[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) { mid[u] = 0.5 * (a[u] + b[u]); v[u] = p[u][10]; }
#pragma unroll
            for (int k = 9; k >= 0; --k) {
#pragma unroll
                for (int u = 0; u < 2; ++u) v[u] = v[u] * mid[u] + p[u][k];
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                if ((v[u] > 0.0) == slo[u]) a[u] = mid[u]; else b[u] = mid[u];
            }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (r + u < n) {
                const double z = 0.5 * (a[u] + b[u]);          // never exactly 0: the bracket is still 2^-46 wide
                roots[r + u] = bracket[r + u] / kGridCells == 0 ? z : 1.0 / z;
            }
        }
    }
    return n;
}

// Five-point solver.  x0/y0, x1/y1: normalised coordinates of the five correspondences (image 0 / image 1), the models
// satisfy [x1 y1 1] E [x0 y0 1]^T = 0.  Writes up to ten unit-Frobenius-norm essential matrices (row-major); returns
// how many.
PM_HDN int five_point(const double* x0, const double* y0, const double* x1, const double* y1, double (*models)[9]) {
    PM_STAMP(0);
    // 1. null space of the 5 x 9 epipolar constraint matrix: orthonormalise the five rows (modified Gram-Schmidt, every
    //    projection done twice), then complete the basis four times with the unit vector e_j that has the largest residual
    double basis[9][9];
    for (int k = 0; k < 5; ++k) {
        const double row[9] = {x1[k] * x0[k], x1[k] * y0[k], x1[k], y1[k] * x0[k], y1[k] * y0[k], y1[k], x0[k], y0[k], 1.0};
        for (int i = 0; i < 9; ++i) basis[k][i] = row[i];
    }
#pragma unroll
    for (int nb = 0; nb < 9; ++nb) {
        if (nb >= 5) {
            int best = 0;
            double best_res = -1.0;
#pragma unroll
            for (int j = 0; j < 9; ++j) {
                double res = 1.0;
#pragma unroll
                for (int k = 0; k < nb; ++k) res = res - basis[k][j] * basis[k][j];
                if (res > best_res) { best_res = res; best = j; }
            }
            for (int i = 0; i < 9; ++i) basis[nb][i] = (i == best) ? 1.0 : 0.0;
        }
#pragma unroll
        for (int pass = 0; pass < 2; ++pass)
#pragma unroll
            for (int k = 0; k < nb; ++k) {
                double d = 0.0;
#pragma unroll
                for (int i = 0; i < 9; ++i) d = d + basis[k][i] * basis[nb][i];
#pragma unroll
                for (int i = 0; i < 9; ++i) basis[nb][i] = basis[nb][i] - d * basis[k][i];
            }
        double nrm = 0.0;
        for (int i = 0; i < 9; ++i) nrm = nrm + basis[nb][i] * basis[nb][i];
        nrm = sqrt(nrm);
        if (!(nrm > 0.0)) return 0;
        for (int i = 0; i < 9; ++i) basis[nb][i] = basis[nb][i] / nrm;
    }
    double Ec[9][4];   // E(x,y,z) entry e = Ec[e][0] x + Ec[e][1] y + Ec[e][2] z + Ec[e][3]
    for (int e = 0; e < 9; ++e)
        for (int s = 0; s < 4; ++s) Ec[e][s] = basis[5 + s][e];

    PM_STAMP(1);
    // 2. the ten cubic constraints: 2 E E^T E - trace(E E^T) E = 0 (nine, halved) and det E = 0
    double A[10][20];
    for (int r = 0; r < 10; ++r)
        for (int m = 0; m < 20; ++m) A[r][m] = 0.0;
    {
        double EEt[3][3][10];
        for (int i = 0; i < 3; ++i)
            for (int j = i; j < 3; ++j) {
                for (int m = 0; m < 10; ++m) EEt[i][j][m] = 0.0;
                for (int k = 0; k < 3; ++k) acc11(EEt[i][j], Ec[i * 3 + k], Ec[j * 3 + k], 1.0);
            }
        for (int m = 0; m < 10; ++m) {
            const double half_tr = 0.5 * ((EEt[0][0][m] + EEt[1][1][m]) + EEt[2][2][m]);
            EEt[0][0][m] = EEt[0][0][m] - half_tr;
            EEt[1][1][m] = EEt[1][1][m] - half_tr;
            EEt[2][2][m] = EEt[2][2][m] - half_tr;
            EEt[1][0][m] = EEt[0][1][m];
            EEt[2][0][m] = EEt[0][2][m];
            EEt[2][1][m] = EEt[1][2][m];
        }
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j)
                for (int k = 0; k < 3; ++k) acc21(A[i * 3 + j], EEt[i][k], Ec[k * 3 + j]);
        double m0[10], m1[10], m2[10];
        for (int m = 0; m < 10; ++m) { m0[m] = 0.0; m1[m] = 0.0; m2[m] = 0.0; }
        acc11(m0, Ec[4], Ec[8], 1.0); acc11(m0, Ec[5], Ec[7], -1.0);
        acc11(m1, Ec[5], Ec[6], 1.0); acc11(m1, Ec[3], Ec[8], -1.0);
        acc11(m2, Ec[3], Ec[7], 1.0); acc11(m2, Ec[4], Ec[6], -1.0);
        acc21(A[9], m0, Ec[0]);
        acc21(A[9], m1, Ec[1]);
        acc21(A[9], m2, Ec[2]);
    }

    PM_STAMP(2);
    // 3. Gauss-Jordan on the first ten columns (partial pivoting over rows).  Only rows 4..9 are read afterwards, so a
    //    pivot step updates the rows below the pivot and, of the rows above it, only those from 4 on; the pivot row is scaled
    //    by one reciprocal.  Fully unrolled: every index except the pivot row is a literal.
#pragma unroll
    for (int c = 0; c < 10; ++c) {
        int piv = c;
        double best = fabs(A[c][c]);
#pragma unroll
        for (int r = c + 1; r < 10; ++r) {
            const double v = fabs(A[r][c]);
            if (v > best) { best = v; piv = r; }
        }
#pragma unroll
        for (int r = c + 1; r < 10; ++r)
            if (r == piv) {
#pragma unroll
                for (int m = c; m < 20; ++m) { const double tmp = A[c][m]; A[c][m] = A[r][m]; A[r][m] = tmp; }
            }
        const double d = A[c][c];
        if (!(fabs(d) > 0.0)) return 0;
        const double inv = 1.0 / d;
#pragma unroll
        for (int m = c + 1; m < 20; ++m) A[c][m] = A[c][m] * inv;
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            if (r > c || (r >= 4 && r < c)) {
                const double f = A[r][c];
#pragma unroll
                for (int m = c + 1; m < 20; ++m) A[r][m] = A[r][m] - f * A[c][m];
            }
        }
    }

    PM_STAMP(3);
    // 4. B(z) [x y 1]^T = 0 with rows <k> = <e> - z<f>, <l> = <g> - z<h>, <m> = <i> - z<j>
    double bx[3][4], by[3][4], b1[3][5];
    for (int r = 0; r < 3; ++r) {
        const double* e = &A[4 + 2 * r][10];
        const double* f = &A[5 + 2 * r][10];
        bx[r][0] = e[2]; bx[r][1] = e[1] - f[2]; bx[r][2] = e[0] - f[1]; bx[r][3] = -f[0];
        by[r][0] = e[5]; by[r][1] = e[4] - f[5]; by[r][2] = e[3] - f[4]; by[r][3] = -f[3];
        b1[r][0] = e[9]; b1[r][1] = e[8] - f[9]; b1[r][2] = e[7] - f[8]; b1[r][3] = e[6] - f[7]; b1[r][4] = -f[6];
    }
    double p1[8], p2[8], p3[7], poly[11];
    for (int k = 0; k < 8; ++k) { p1[k] = 0.0; p2[k] = 0.0; }
    for (int k = 0; k < 7; ++k) p3[k] = 0.0;
    for (int k = 0; k < 11; ++k) poly[k] = 0.0;
    conv_acc(p1, by[1], 3, b1[2], 4, 1.0); conv_acc(p1, b1[1], 4, by[2], 3, -1.0);
    conv_acc(p2, b1[1], 4, bx[2], 3, 1.0); conv_acc(p2, bx[1], 3, b1[2], 4, -1.0);
    conv_acc(p3, bx[1], 3, by[2], 3, 1.0); conv_acc(p3, by[1], 3, bx[2], 3, -1.0);
    conv_acc(poly, bx[0], 3, p1, 7, 1.0);
    conv_acc(poly, by[0], 3, p2, 7, 1.0);
    conv_acc(poly, b1[0], 4, p3, 6, 1.0);
    double scale = 0.0;
    for (int k = 0; k < 11; ++k) scale = fabs(poly[k]) > scale ? fabs(poly[k]) : scale;
    if (!(scale > 0.0) || !(scale < 1e300)) return 0;
    for (int k = 0; k < 11; ++k) poly[k] = poly[k] / scale;

    PM_STAMP(4);
    // 5. real roots z -> (x, y) from the best-conditioned pair of rows of B(z) -> E
    double roots[kMaxModels];
    const int nroots = real_roots10(poly, roots);
    PM_STAMP(5);
    int n = 0;
    for (int i = 0; i < nroots; ++i) {
        const double z = roots[i];
        double rx[3], ry[3], r1[3];
        for (int r = 0; r < 3; ++r) {
            rx[r] = horner(bx[r], 3, z);
            ry[r] = horner(by[r], 3, z);
            r1[r] = horner(b1[r], 4, z);
        }
        double best0 = 0.0, best1 = 0.0, best2 = 0.0;
        for (int pr = 0; pr < 3; ++pr) {
            const int a = pr == 2 ? 1 : 0, b = pr == 0 ? 1 : 2;       // (k,l), (k,m), (l,m)
            const double v0 = ry[a] * r1[b] - r1[a] * ry[b];
            const double v1 = r1[a] * rx[b] - rx[a] * r1[b];
            const double v2 = rx[a] * ry[b] - ry[a] * rx[b];
            if (pr == 0 || fabs(v2) > fabs(best2)) { best0 = v0; best1 = v1; best2 = v2; }
        }
        const double x = best0 / best2, y = best1 / best2;
        double E[9], nrm = 0.0;
        for (int e = 0; e < 9; ++e) {
            E[e] = ((x * Ec[e][0] + y * Ec[e][1]) + z * Ec[e][2]) + Ec[e][3];
            nrm = nrm + E[e] * E[e];
        }
        nrm = sqrt(nrm);
        if (!(nrm > 0.0) || !(nrm < 1e300)) continue;
        for (int e = 0; e < 9; ++e) models[n][e] = E[e] / nrm;
        ++n;
    }
    PM_STAMP(6);
    return n;
}

// Sampson test of OpenCV's essential-matrix RANSAC: (x1^T E x0)^2 / (|E x0|_xy^2 + |E^T x1|_xy^2) <= thr^2, written
// without the division.
PM_HD bool sampson_inlier(const double* E, double x0, double y0, double x1, double y1, double thr2) {
    const double a0 = (E[0] * x0 + E[1] * y0) + E[2];
    const double a1 = (E[3] * x0 + E[4] * y0) + E[5];
    const double a2 = (E[6] * x0 + E[7] * y0) + E[8];
    const double b0 = (E[0] * x1 + E[3] * y1) + E[6];
    const double b1 = (E[1] * x1 + E[4] * y1) + E[7];
    const double d = (x1 * a0 + y1 * a1) + a2;
    const double den = ((a0 * a0 + a1 * a1) + b0 * b0) + b1 * b1;
    return d * d <= thr2 * den;
}

// E -> the two rotations and the translation direction of cv::decomposeEssentialMat (E = U diag(s,s,0) V^T with
// det U = det V = +1; R1 = U W V^T, R2 = U W^T V^T, t = U[:,2]).  The right singular vectors come from the Jacobi
// eigen-decomposition of E^T E.
PM_HD void decompose_essential(const double* E, double* R1, double* R2, double* t) {
    double S[3][3], V[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) S[i][j] = (E[0 * 3 + i] * E[0 * 3 + j] + E[1 * 3 + i] * E[1 * 3 + j]) + E[2 * 3 + i] * E[2 * 3 + j];
    jacobi_eig<3, 8>(S, V);
    int i0 = 0;                                            // two largest eigenvalues, first index wins ties
    for (int i = 1; i < 3; ++i) if (S[i][i] > S[i0][i0]) i0 = i;
    int i1 = i0 == 0 ? 1 : 0;
    for (int i = 0; i < 3; ++i) if (i != i0 && S[i][i] > S[i1][i1]) i1 = i;
    double v0[3], v1[3], v2[3], u0[3], u1[3], u2[3];
    for (int k = 0; k < 3; ++k) { v0[k] = V[k][i0]; v1[k] = V[k][i1]; }
    v2[0] = v0[1] * v1[2] - v0[2] * v1[1]; v2[1] = v0[2] * v1[0] - v0[0] * v1[2]; v2[2] = v0[0] * v1[1] - v0[1] * v1[0];
    for (int k = 0; k < 3; ++k) {
        u0[k] = (E[k * 3 + 0] * v0[0] + E[k * 3 + 1] * v0[1]) + E[k * 3 + 2] * v0[2];
        u1[k] = (E[k * 3 + 0] * v1[0] + E[k * 3 + 1] * v1[1]) + E[k * 3 + 2] * v1[2];
    }
    double n0 = sqrt((u0[0] * u0[0] + u0[1] * u0[1]) + u0[2] * u0[2]);
    for (int k = 0; k < 3; ++k) u0[k] = u0[k] / n0;
    const double dp = (u0[0] * u1[0] + u0[1] * u1[1]) + u0[2] * u1[2];
    for (int k = 0; k < 3; ++k) u1[k] = u1[k] - dp * u0[k];
    double n1 = sqrt((u1[0] * u1[0] + u1[1] * u1[1]) + u1[2] * u1[2]);
    for (int k = 0; k < 3; ++k) u1[k] = u1[k] / n1;
    u2[0] = u0[1] * u1[2] - u0[2] * u1[1]; u2[1] = u0[2] * u1[0] - u0[0] * u1[2]; u2[2] = u0[0] * u1[1] - u0[1] * u1[0];
    // U W V^T = -u1 v0^T + u0 v1^T + u2 v2^T ;  U W^T V^T = u1 v0^T - u0 v1^T + u2 v2^T
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            const double a = u0[i] * v1[j] - u1[i] * v0[j], b = u2[i] * v2[j];
            R1[i * 3 + j] = a + b;
            R2[i * 3 + j] = b - a;
        }
    for (int k = 0; k < 3; ++k) t[k] = u2[k];
}

// Cheirality test of cv::recoverPose for one correspondence and one candidate [R | t]: DLT triangulation against [I | 0]
// (null vector of the 4 x 4 system = eigenvector of A^T A with the smallest eigenvalue), positive and bounded depth in
// both cameras.
PM_HD bool cheirality(const double* R, const double* t, double x0, double y0, double x1, double y1, double dist) {
    double Am[4][4] = {{-1.0, 0.0, x0, 0.0}, {0.0, -1.0, y0, 0.0}, {0, 0, 0, 0}, {0, 0, 0, 0}};
    for (int j = 0; j < 3; ++j) {
        Am[2][j] = x1 * R[6 + j] - R[j];
        Am[3][j] = y1 * R[6 + j] - R[3 + j];
    }
    Am[2][3] = x1 * t[2] - t[0];
    Am[3][3] = y1 * t[2] - t[1];
    double S[4][4], V[4][4];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) S[i][j] = ((Am[0][i] * Am[0][j] + Am[1][i] * Am[1][j]) + Am[2][i] * Am[2][j]) + Am[3][i] * Am[3][j];
    jacobi_eig<4, 4>(S, V);   // four sweeps: the vote only needs the signs of the null vector
    int im = 0;
    for (int i = 1; i < 4; ++i) if (S[i][i] < S[im][im]) im = i;
    const double q0 = V[0][im], q1 = V[1][im], q2 = V[2][im], q3 = V[3][im];
    bool ok = q2 * q3 > 0.0;
    const double X = q0 / q3, Y = q1 / q3, Z = q2 / q3;
    ok = ok && (Z < dist);
    const double Z1 = ((R[6] * X + R[7] * Y) + R[8] * Z) + t[2];
    return ok && (Z1 > 0.0) && (Z1 < dist);
}

}  // namespace pm
