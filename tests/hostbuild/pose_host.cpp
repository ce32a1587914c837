// Host build of pope_b200/csrc/pose_math.cuh for the CPU test-suite (tests/test_pose.py): the same source the device runs,
// compiled by g++ without multiply-add contraction, checked bit for bit against oracle/pose_oracle.py.  Test infrastructure
// only; nothing in the product links it.
#include "pose_math.cuh"

extern "C" {
int pm_draw5(uint64_t seed, uint64_t pair, uint64_t h, int m, int* idx) { return pm::draw5(seed, pair, h, m, idx) ? 1 : 0; }
int pm_five_point(const double* x0, const double* y0, const double* x1, const double* y1, double* models) {
    return pm::five_point(x0, y0, x1, y1, reinterpret_cast<double (*)[9]>(models));
}
int pm_sampson(const double* E, double x0, double y0, double x1, double y1, double thr2) {
    return pm::sampson_inlier(E, x0, y0, x1, y1, thr2) ? 1 : 0;
}
void pm_decompose(const double* E, double* R1, double* R2, double* t) { pm::decompose_essential(E, R1, R2, t); }
int pm_cheirality(const double* R, const double* t, double x0, double y0, double x1, double y1, double dist) {
    return pm::cheirality(R, t, x0, y0, x1, y1, dist) ? 1 : 0;
}
}
