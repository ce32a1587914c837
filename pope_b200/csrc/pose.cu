// Batched relative-pose RANSAC for a whole batch of match lists (second half): the device-side
// replacement of the per-pair `estimate_pose` of the reference (src/utils/metrics.py:69-94: cv2.findEssentialMat with
// method=cv2.RANSAC, then cv2.recoverPose), consuming the match list where the hot path leaves it (device mkpts0_f /
// mkpts1_f / counts) instead of three D2H copies and one OpenCV call per pair.
//
// Scheme: the minimal samples of all pairs are solved and scored in parallel, in waves of growing size; a per-pair scan then
// consumes the scored models in sample order with exactly the bookkeeping of a sequential RANSAC loop (best count, adaptive
// iteration bound from `conf`), so the result equals the sequential algorithm run on the same samples, and pairs that have
// reached their bound skip the later waves.  All arithmetic is fp64 without fused multiply-add (this file is compiled
// with -fmad=false) so that oracle/pose_oracle.py reproduces it operation by operation.
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

#include <algorithm>

#include "pope_b200.h"
#include "pose_math.cuh"

namespace {

constexpr int kWaves = 5;
// 64 64 128 256 512: the solve kernel is bound by the latency of one thread's solver, so a first wave of 64 costs what 32 would
__host__ __device__ constexpr int wave_size(int w) { return w == 0 ? 64 : (64 << (w - 1)); }
constexpr int kMaxWave = 512;
constexpr int kScoreThreads = 128;
constexpr int kFinalThreads = 256;

struct PairState {
    int best_cnt;    // inliers of the best model so far
    int niters;      // current iteration bound (starts at max_iters)
    int done;        // the sequential loop has ended
    int used;        // samples consumed when it ended
};

struct PoseWs {
    int64_t* offsets;     // [n + 1]
    PairState* state;     // [n]
    double* thr2;         // [n]
    double* best_e;       // [n, 9]
    double* pts;          // [capacity, 4] normalised x0 y0 x1 y1
    double* models;       // [n, kMaxWave, 10, 9]
    int32_t* nmodels;     // [n, kMaxWave]
    int32_t* mcount;      // [n, kMaxWave, 10]
    uint8_t* flags;       // [capacity]
    double* cand;         // [n, 21]: R1, R2, t of the best model
    int32_t* good;        // [n, 4]: cheirality votes of the four candidates
    int32_t* work;        // [n * kMaxWave * 10]: (pair, sample, model) items of the current wave
    int32_t* work_n;      // [kWaves]: item count per wave
};

size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

size_t carve(PoseWs* w, void* base, int n, int64_t capacity) {
    size_t off = 0;
    auto take = [&](size_t bytes) {
        void* p = base ? (void*)((char*)base + off) : nullptr;
        off += align256(bytes);
        return p;
    };
    PoseWs tmp;
    tmp.offsets = (int64_t*)take(sizeof(int64_t) * (n + 1));
    tmp.state = (PairState*)take(sizeof(PairState) * n);
    tmp.thr2 = (double*)take(sizeof(double) * n);
    tmp.best_e = (double*)take(sizeof(double) * 9 * n);
    tmp.pts = (double*)take(sizeof(double) * 4 * capacity);
    tmp.models = (double*)take(sizeof(double) * 90 * (size_t)kMaxWave * n);
    tmp.nmodels = (int32_t*)take(sizeof(int32_t) * (size_t)kMaxWave * n);
    tmp.mcount = (int32_t*)take(sizeof(int32_t) * 10 * (size_t)kMaxWave * n);
    tmp.flags = (uint8_t*)take((size_t)capacity);
    tmp.cand = (double*)take(sizeof(double) * 21 * n);
    tmp.good = (int32_t*)take(sizeof(int32_t) * 4 * n);
    tmp.work = (int32_t*)take(sizeof(int32_t) * 10 * (size_t)kMaxWave * n);
    tmp.work_n = (int32_t*)take(sizeof(int32_t) * kWaves);
    if (w) *w = tmp;
    return off;
}

// Prefix of the per-pair match counts, per-pair threshold and initial loop state.  One block.
__global__ void pose_prepare_kernel(const int32_t* __restrict__ counts, const double* __restrict__ K0,
                                    const double* __restrict__ K1, int n, int64_t capacity, double thresh, int max_iters,
                                    PoseWs w) {
    __shared__ int64_t carry;
    __shared__ int64_t part[1024];
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += blockDim.x) {
        const int p = base + threadIdx.x;
        part[threadIdx.x] = p < n ? (int64_t)max(counts[p], 0) : 0;
        __syncthreads();
        for (int d = 1; d < (int)blockDim.x; d <<= 1) {          // inclusive Hillis-Steele scan
            const int64_t v = threadIdx.x >= (unsigned)d ? part[threadIdx.x - d] : 0;
            __syncthreads();
            part[threadIdx.x] += v;
            __syncthreads();
        }
        if (p < n) w.offsets[p + 1] = min(carry + part[threadIdx.x], capacity);   // lists never run past the arrays
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry += part[threadIdx.x];
        __syncthreads();
    }
    if (threadIdx.x == 0) w.offsets[0] = 0;
    if (threadIdx.x < kWaves) w.work_n[threadIdx.x] = 0;
    __syncthreads();
    for (int p = threadIdx.x; p < n; p += blockDim.x) {
        // thresh / np.mean([K0[0,0], K1[1,1], K0[0,0], K1[1,1]])   (metrics.py:77)
        const double f0 = K0[p * 9 + 0], f1 = K1[p * 9 + 4];
        const double thr = thresh / ((((f0 + f1) + f0) + f1) / 4.0);
        w.thr2[p] = thr * thr;
        PairState s;
        s.best_cnt = 0;
        s.niters = max_iters;
        s.done = w.offsets[p + 1] - w.offsets[p] < 5 ? 1 : 0;      // metrics.py:70-71
        s.used = 0;
        w.state[p] = s;
    }
}

// kpts = (kpts - K[[0,1],[2,2]]) / K[[0,1],[0,1]]   (metrics.py:73-74), fp32 pixels -> fp64 normalised coordinates
__global__ void pose_normalize_kernel(const float* __restrict__ mk0, const float* __restrict__ mk1,
                                      const double* __restrict__ K0, const double* __restrict__ K1, PoseWs w) {
    const int p = blockIdx.x;
    const int64_t lo = w.offsets[p], hi = w.offsets[p + 1];
    const double fx0 = K0[p * 9 + 0], fy0 = K0[p * 9 + 4], cx0 = K0[p * 9 + 2], cy0 = K0[p * 9 + 5];
    const double fx1 = K1[p * 9 + 0], fy1 = K1[p * 9 + 4], cx1 = K1[p * 9 + 2], cy1 = K1[p * 9 + 5];
    for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        const float2 a = reinterpret_cast<const float2*>(mk0)[i], b = reinterpret_cast<const float2*>(mk1)[i];
        double4 o;
        o.x = ((double)a.x - cx0) / fx0;
        o.y = ((double)a.y - cy0) / fy0;
        o.z = ((double)b.x - cx1) / fx1;
        o.w = ((double)b.y - cy1) / fy1;
        reinterpret_cast<double4*>(w.pts)[i] = o;
    }
}

// Waves that are SOLVED by one launch.  The solve kernel is bound by the latency of one thread's solver (~75 us for 64 samples
// per pair, ~20 us more per further 64: 255 registers, 4.5 KB of stack per thread), so for a high confidence, where the first
// wave rarely ends a pair, the first launch solves the samples of waves 0..2 (256 per pair) ahead of their scoring; scoring
// and the sequential scan still go wave by wave, and a wave's items of pairs that have finished meanwhile are skipped by the
// score kernel (the scan never looks at them), so the result is the one of the wave-by-wave schedule.  Measured (64 pairs x
// 2 500 matches): conf 0.99999 0.53 -> 0.51 ms (30 % outliers), 2.40 -> 2.25 ms (60 %); at conf 0.99 and 30 % outliers one wave
// is all a pair needs and solving ahead costs 0.42 -> 0.48 ms, hence the gate on the confidence.
constexpr int kAheadWaves = 3;
constexpr double kAheadConf = 0.9999;
__host__ __device__ constexpr int wave_start(int w) { return w == 0 ? 0 : (64 << (w - 1)); }     // 0 64 128 256 512
// where a wave's (pair, sample, model) items start in PoseWs::work: the waves of one solve launch need disjoint lists
__host__ __device__ inline size_t work_base(int wave, int first_wave, int n) {
    return (size_t)n * 10 * (size_t)(wave_start(wave) - wave_start(first_wave));
}

// One thread per (pair, sample of the waves [wave0, wave0 + nwaves)): hashed minimal sample -> five-point solver -> up to ten
// models, stored at slot = sample - first sample of the launch.
__global__ void pose_solve_kernel(int n, int start, int gsize, int wave0, int nwaves, uint64_t seed, PoseWs w) {
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= n * gsize) return;
    const int p = gid / gsize, s = gid % gsize, h = start + s;
    int wave = wave0;
    while (wave + 1 < wave0 + nwaves && h >= wave_start(wave + 1)) ++wave;
    const PairState st = w.state[p];
    int nm = 0;
    if (!st.done && h < st.niters) {
        const int64_t lo = w.offsets[p];
        const int m = (int)(w.offsets[p + 1] - lo);
        int idx[5];
        if (pm::draw5(seed, (uint64_t)p, (uint64_t)h, m, idx)) {
            double x0[5], y0[5], x1[5], y1[5];
            for (int k = 0; k < 5; ++k) {
                const double4 q = reinterpret_cast<const double4*>(w.pts)[lo + idx[k]];
                x0[k] = q.x; y0[k] = q.y; x1[k] = q.z; y1[k] = q.w;
            }
            double (*out)[9] = reinterpret_cast<double (*)[9]>(w.models + ((size_t)p * kMaxWave + s) * 90);
            nm = pm::five_point(x0, y0, x1, y1, out);
        }
    }
    w.nmodels[(size_t)p * kMaxWave + s] = nm;
    if (nm > 0) {
        const int at = atomicAdd(&w.work_n[wave], nm);
        int32_t* list = w.work + work_base(wave, wave0, n);
        for (int r = 0; r < nm; ++r) list[at + r] = (p * kMaxWave + s) * 10 + r;
    }
}

// Inlier count of each model over all matches of its pair: the warps walk the wave's list of (pair, sample, model) items
// (the list order is arbitrary, the counts are stored per item), the lanes stride over the matches.
__global__ void __launch_bounds__(kScoreThreads) pose_score_kernel(int wave, int first_wave, int n, PoseWs w) {
    const int items = w.work_n[wave];
    const int32_t* __restrict__ list = w.work + work_base(wave, first_wave, n);
    const int lane = threadIdx.x & 31, warps = (gridDim.x * kScoreThreads) >> 5;
    for (int item = (blockIdx.x * kScoreThreads + threadIdx.x) >> 5; item < items; item += warps) {
        const int id = list[item], p = id / (kMaxWave * 10);
        // solved ahead of the earlier waves' scans: the pair may have finished, or its bound may have dropped below this sample
        const PairState st = w.state[p];
        if (st.done || wave_start(first_wave) + (id / 10) % kMaxWave >= st.niters) continue;
        double E[9];
#pragma unroll
        for (int e = 0; e < 9; ++e) E[e] = w.models[(size_t)id * 9 + e];
        const int64_t lo = w.offsets[p], hi = w.offsets[p + 1];
        const double thr2 = w.thr2[p];
        // The scan only acts on a count that exceeds the best count of the earlier waves (and 4), so a model that can no
        // longer get there is dropped: every 256 matches the warp compares its count plus the matches still ahead.
        const int need = max(st.best_cnt, 4);
        int cnt = 0;
        bool dropped = false;
        if (need <= 4) {                              // first wave: nothing to compare with yet
#pragma unroll 2
            for (int64_t i = lo + lane; i < hi; i += 32) {
                const double4 q = reinterpret_cast<const double4*>(w.pts)[i];
                cnt += pm::sampson_inlier(E, q.x, q.y, q.z, q.w, thr2) ? 1 : 0;
            }
        } else {
            int it = 0;
            for (int64_t base = lo; base < hi; base += 32) {
                const int64_t i = base + lane;
                if (i < hi) {
                    const double4 q = reinterpret_cast<const double4*>(w.pts)[i];
                    cnt += pm::sampson_inlier(E, q.x, q.y, q.z, q.w, thr2) ? 1 : 0;
                }
                if ((++it & 7) == 0) {
                    const int64_t ahead = hi - (base + 32);
                    if (__reduce_add_sync(0xffffffffu, cnt) + ahead <= need) { dropped = true; break; }
                }
            }
        }
        cnt = dropped ? 0 : __reduce_add_sync(0xffffffffu, cnt);
        if (lane == 0) w.mcount[id] = cnt;
    }
}

// cv::RANSACUpdateNumIters of OpenCV 4.x (calib3d/src/ptsetreg.cpp), model size 5.
__device__ int update_num_iters(double p, double ep, int max_iters) {
    p = fmax(p, 0.0); p = fmin(p, 1.0);
    ep = fmax(ep, 0.0); ep = fmin(ep, 1.0);
    double num = fmax(1.0 - p, DBL_MIN);
    const double wv = 1.0 - ep;
    double denom = 1.0 - (((wv * wv) * wv) * wv) * wv;
    if (denom < DBL_MIN) return 0;
    num = log(num);
    denom = log(denom);
    return (denom >= 0.0 || -num >= max_iters * (-denom)) ? max_iters : (int)rint(num / denom);
}

// One warp per pair: consume this wave's scored models in sample order like the sequential loop would.  Each lane fetches
// one sample's best model (the first one with the sample's largest count: within a sample only that one can end up as the
// running best, and the iteration bound after the sample depends on it alone); the warp then walks the 32 samples in order.
// slot0 = slot of the wave's first sample in the models / counts arrays (the solve launch's numbering).
__global__ void pose_scan_kernel(int n, int start, int wsize, int slot0, double conf, int last_wave, PoseWs w) {
    const int p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (p >= n) return;
    PairState st = w.state[p];
    if (st.done) return;
    const int m = (int)(w.offsets[p + 1] - w.offsets[p]);
    for (int s0 = 0; s0 < wsize && !st.done; s0 += 32) {
        const int s = s0 + lane;
        int bc = 0, br = 0;
        if (s < wsize) {
            const int nm = w.nmodels[(size_t)p * kMaxWave + slot0 + s];
            for (int r = 0; r < nm; ++r) {
                const int cnt = w.mcount[((size_t)p * kMaxWave + slot0 + s) * 10 + r];
                if (cnt > bc) { bc = cnt; br = r; }
            }
        }
        for (int l = 0; l < 32 && s0 + l < wsize; ++l) {
            if (start + s0 + l >= st.niters) { st.done = 1; st.used = st.niters; break; }
            const int cnt = __shfl_sync(0xffffffffu, bc, l), r = __shfl_sync(0xffffffffu, br, l);
            if (cnt > max(st.best_cnt, 4)) {
                st.best_cnt = cnt;
                const double* src = w.models + (((size_t)p * kMaxWave + slot0 + s0 + l) * 10 + r) * 9;
                if (lane < 9) w.best_e[p * 9 + lane] = src[lane];
                st.niters = update_num_iters(conf, (double)(m - cnt) / (double)m, st.niters);
            }
        }
    }
    if (!st.done && (start + wsize >= st.niters || last_wave)) {
        st.done = 1;
        st.used = min(st.niters, start + wsize);
    }
    __syncwarp();
    if (lane == 0) w.state[p] = st;
}

// One thread per pair: best model -> the two rotations and the translation direction; clears the votes.
__global__ void pose_decompose_kernel(int n, PoseWs w) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    for (int c = 0; c < 4; ++c) w.good[p * 4 + c] = 0;
    if (w.state[p].best_cnt == 0) return;
    double E[9], R1[9], R2[9], t[3];
    for (int e = 0; e < 9; ++e) E[e] = w.best_e[p * 9 + e];
    pm::decompose_essential(E, R1, R2, t);
    for (int e = 0; e < 9; ++e) { w.cand[p * 21 + e] = R1[e]; w.cand[p * 21 + 9 + e] = R2[e]; }
    for (int k = 0; k < 3; ++k) w.cand[p * 21 + 18 + k] = t[k];
}

// One thread per match: inlier test of the pair's best model and the cheirality test of the four candidate poses.
__global__ void __launch_bounds__(128) pose_cheirality_kernel(int n, PoseWs w) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = w.offsets[n];
    int p = -1;
    uint8_t f = 0;
    if (i < total) {
        int lo = 0, hi = n;                      // largest p with offsets[p] <= i
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (w.offsets[mid] <= i) lo = mid; else hi = mid;
        }
        p = lo;
        if (w.state[p].best_cnt > 0) {
            const double4 q = reinterpret_cast<const double4*>(w.pts)[i];
            const double* E = w.best_e + p * 9;
            if (pm::sampson_inlier(E, q.x, q.y, q.z, q.w, w.thr2[p])) {
                const double* cd = w.cand + p * 21;
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {          // [R1|t] [R2|t] [R1|-t] [R2|-t], the order of cv::recoverPose
                    double tt[3];
                    for (int k = 0; k < 3; ++k) tt[k] = (c & 2) ? -cd[18 + k] : cd[18 + k];
                    if (pm::cheirality(cd + (c & 1) * 9, tt, q.x, q.y, q.z, q.w, 1e9)) f |= (uint8_t)(1u << c);
                }
            }
        }
        w.flags[i] = f;
    }
    const int p0 = __shfl_sync(0xffffffffu, p, 0);
    if (__all_sync(0xffffffffu, p == p0)) {
        if (p0 < 0) return;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int tot = __reduce_add_sync(0xffffffffu, (f >> c) & 1);
            if ((threadIdx.x & 31) == 0 && tot) atomicAdd(&w.good[p0 * 4 + c], tot);
        }
    } else if (p >= 0) {
        for (int c = 0; c < 4; ++c)
            if ((f >> c) & 1) atomicAdd(&w.good[p * 4 + c], 1);
    }
}

// One block per pair: the vote, the outputs, the final mask.
__global__ void __launch_bounds__(kFinalThreads) pose_select_kernel(PoseWs w, double* __restrict__ R_out,
                                                                    double* __restrict__ t_out, double* __restrict__ E_out,
                                                                    uint8_t* __restrict__ inliers,
                                                                    int32_t* __restrict__ n_inliers,
                                                                    int32_t* __restrict__ status, int32_t* __restrict__ iters) {
    const int p = blockIdx.x;
    const int64_t lo = w.offsets[p], hi = w.offsets[p + 1];
    const PairState st = w.state[p];
    if (st.best_cnt == 0) {       // fewer than five matches, or no model with at least five inliers: `return None`
        for (int64_t i = lo + threadIdx.x; i < hi; i += kFinalThreads) inliers[i] = 0;
        if (threadIdx.x < 9) { R_out[p * 9 + threadIdx.x] = 0.0; E_out[p * 9 + threadIdx.x] = 0.0; }
        if (threadIdx.x < 3) t_out[p * 3 + threadIdx.x] = 0.0;
        if (threadIdx.x == 0) { n_inliers[p] = 0; status[p] = 0; iters[p] = st.used; }
        return;
    }
    int c = 0;
    for (int k = 1; k < 4; ++k) if (w.good[p * 4 + k] > w.good[p * 4 + c]) c = k;   // first maximum = recoverPose's if-chain
    if (threadIdx.x == 0) {
        n_inliers[p] = w.good[p * 4 + c];
        status[p] = w.good[p * 4 + c] > 0 ? 1 : 0;        // metrics.py:91 `if n > best_num_inliers`
        iters[p] = st.used;
    }
    if (threadIdx.x < 9) {
        R_out[p * 9 + threadIdx.x] = w.cand[p * 21 + (c & 1) * 9 + threadIdx.x];
        E_out[p * 9 + threadIdx.x] = w.best_e[p * 9 + threadIdx.x];
    }
    if (threadIdx.x < 3) t_out[p * 3 + threadIdx.x] = (c & 2) ? -w.cand[p * 21 + 18 + threadIdx.x] : w.cand[p * 21 + 18 + threadIdx.x];
    for (int64_t i = lo + threadIdx.x; i < hi; i += kFinalThreads) inliers[i] = (w.flags[i] >> c) & 1;
}

}  // namespace

extern "C" size_t pope_pose_workspace_bytes(int n_pairs, int64_t capacity) {
    if (n_pairs <= 0 || capacity < 0) return 0;
    return carve(nullptr, nullptr, n_pairs, capacity);
}

extern "C" int pope_estimate_pose_batch(const float* mkpts0, const float* mkpts1, const int32_t* counts, int n_pairs,
                                        int64_t capacity, const double* K0, const double* K1, double thresh, double conf,
                                        int max_iters, uint64_t seed, double* R, double* t, double* E, uint8_t* inliers,
                                        int32_t* n_inliers, int32_t* status, int32_t* iters, void* workspace,
                                        size_t workspace_bytes, void* stream) {
    if (n_pairs == 0) return POPE_OK;
    if (n_pairs < 0 || n_pairs > 65535 || capacity < 0 || max_iters < 1 || max_iters > POPE_POSE_MAX_ITERS || !(thresh > 0.0) || !counts || !K0 || !K1 || !R || !t || !E ||
        !n_inliers || !status || !iters || (capacity > 0 && (!mkpts0 || !mkpts1 || !inliers)))
        return POPE_ERR_INVALID_ARG;
    PoseWs w;
    if (!workspace || workspace_bytes < carve(&w, workspace, n_pairs, capacity)) return POPE_ERR_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    pose_prepare_kernel<<<1, 1024, 0, st>>>(counts, K0, K1, n_pairs, capacity, thresh, max_iters, w);
    pose_normalize_kernel<<<n_pairs, 256, 0, st>>>(mkpts0, mkpts1, K0, K1, w);
    int start = 0, first_wave = 0;          // first_wave: first wave of the solve launch the current wave belongs to
    const int ahead = conf >= kAheadConf ? kAheadWaves : 1;
    for (int wv = 0; wv < kWaves && start < max_iters; ++wv) {
        const int ws = wave_size(wv);
        const bool last = (wv == kWaves - 1) || (start + ws >= max_iters);
        if (wv == 0 || wv >= ahead) {
            const int nw = wv == 0 ? ahead : 1;
            const int gsize = wave_start(wv + nw) - wave_start(wv);
            first_wave = wv;
            pose_solve_kernel<<<(n_pairs * gsize + 31) / 32, 32, 0, st>>>(n_pairs, start, gsize, wv, nw, seed, w);
        }
        // about four models per sample: one block of four warps per sample, capped at a few waves of the machine
        pose_score_kernel<<<(int)std::min<int64_t>((int64_t)n_pairs * ws, 148 * 16), kScoreThreads, 0, st>>>(wv, first_wave, n_pairs, w);
        pose_scan_kernel<<<(n_pairs + 3) / 4, 128, 0, st>>>(n_pairs, start, ws, start - wave_start(first_wave), conf, last ? 1 : 0, w);
        start += ws;
    }
    pose_decompose_kernel<<<(n_pairs + 63) / 64, 64, 0, st>>>(n_pairs, w);
    if (capacity > 0) pose_cheirality_kernel<<<(unsigned)((capacity + 127) / 128), 128, 0, st>>>(n_pairs, w);
    pose_select_kernel<<<n_pairs, kFinalThreads, 0, st>>>(w, R, t, E, inliers, n_inliers, status, iters);
    return (int)cudaGetLastError();
}
