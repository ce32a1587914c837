// coarse_finalize.cu -- mutual-nearest test, border removal and ordered compaction of the coarse matches.
//
// Replaces src/matcher/utils/coarse_matching.py:176-196 and :239-259 (reference tree): after the sweeps every row
// and column holds its best above-threshold candidate (value = log2 conf, index).  Because conf <= min(p_row,
// p_col), any cell above the threshold beats every cell below it, so "== row max and == column max of the full
// matrix" is "the row's best candidate is also (by value) its column's best candidate".  Border cells take part
// in the maxima and are only removed afterwards, exactly like the reference (mask_border clears the threshold
// mask only, :176-184, while :187-189 use the full conf matrix).
// Output order is torch.where order: sorted by (pair, i), at most one match per row (:192-195).
#include "common.cuh"

namespace pope {
namespace {

constexpr int FT = 1024;

struct Grid2 { int h, w, b; };
__device__ __forceinline__ bool interior(int idx, Grid2 g) {
  int y = idx / g.w, x = idx - y * g.w;
  return y >= g.b && y < g.h - g.b && x >= g.b && x < g.w - g.b;
}

// is row i of pair n a match?  returns j and log2-conf
__device__ __forceinline__ bool row_match(const u64* __restrict__ rowbest, const u64* __restrict__ colbest, int L,
                                          int S, int i, Grid2 g0, Grid2 g1, int& j, float& t2) {
  const u64 rb = rowbest[i];
  if (rb == 0ull) return false;
  j = best_index(rb);
  if (static_cast<unsigned>(j) >= static_cast<unsigned>(S)) return false;
  if (best_key(colbest[j]) != best_key(rb)) return false;   // some other row holds a larger value in column j
  if (!interior(i, g0) || !interior(j, g1)) return false;
  t2 = key_to_float(best_key(rb));
  return true;
}

__device__ __forceinline__ int block_sum(int v, int* smem) {
  // returns the block-wide sum to every thread (FT threads)
  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
  __syncthreads();
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  int t = (lane < FT / 32) ? smem[lane] : 0;
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) t += __shfl_xor_sync(kFullMask, t, o);
  return t;
}

// One CTA per pair: count the pair's matches, publish the count, wait for the counts of all earlier pairs, then emit in
// (pair, i) order.  ready[0..n) must be 0 on entry and ready[n] (the ticket counter) too; the kernel leaves them set (the
// caller's memset / the column-merge kernel clears them).
// phase 0: count, look back, emit in one launch.  The pair a CTA works on is a TICKET drawn when the CTA starts, not its
// blockIdx: a CTA only ever waits for pairs whose CTAs drew their tickets earlier, i.e. are running or done, whatever
// order the hardware dispatches the grid in and whatever else occupies the SMs.  Batches of more than 2 x SMs pairs keep the
// two-launch form (phase 1 = count only, phase 2 = emit only).
// Rows are matched once: a thread keeps the outcome of its (up to kKeep x FT rows) in registers between counting and
// emitting (L <= kKeep * FT; longer rows are matched again).
constexpr int kKeep = 5;
__global__ void __launch_bounds__(FT) count_emit_kernel(const u64* __restrict__ rowbest, const u64* __restrict__ colbest,
                                                       const float* __restrict__ lse_r, const float* __restrict__ lse_c,
                                                       int L, int S, Grid2 g0, Grid2 g1, float pixel_scale,
                                                       int32_t* __restrict__ counts, int n_pairs, int* __restrict__ ready,
                                                       int64_t* __restrict__ b_ids, int64_t* __restrict__ i_ids,
                                                       int64_t* __restrict__ j_ids, float* __restrict__ mconf,
                                                       float* __restrict__ mk0, float* __restrict__ mk1, int phase,
                                                       int64_t capacity) {
  __shared__ int smem[32];
  __shared__ int warp_off[FT / 32];
  __shared__ int s_pair;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (phase == 0) {
    if (threadIdx.x == 0) s_pair = atomicAdd(ready + n_pairs, 1);
    __syncthreads();
  }
  const int n = phase == 0 ? s_pair : int(blockIdx.x);
  rowbest += size_t(n) * L; colbest += size_t(n) * S;
  const bool keep = L <= kKeep * FT;
  int kj[kKeep];
  float kt[kKeep];
  uint32_t khit = 0;
  int c = 0, bad = 0;
  if (phase != 2) {
    if (keep) {
#pragma unroll
      for (int k = 0; k < kKeep; ++k) {
        const int i = k * FT + threadIdx.x;
        kj[k] = 0; kt[k] = 0.f;
        if (i < L) {
          if (row_match(rowbest, colbest, L, S, i, g0, g1, kj[k], kt[k])) { khit |= 1u << k; ++c; }
          bad |= !isfinite(lse_r[size_t(n) * L + i]);
        }
      }
    } else {
      for (int i = threadIdx.x; i < L; i += FT) {
        int j; float t2;
        c += row_match(rowbest, colbest, L, S, i, g0, g1, j, t2) ? 1 : 0;
        bad |= !isfinite(lse_r[size_t(n) * L + i]);
      }
    }
    for (int j = threadIdx.x; j < S; j += FT) bad |= !isfinite(lse_c[size_t(n) * S + j]);
    c = block_sum(c, smem);
    bad = block_sum(bad, smem);
    if (threadIdx.x == 0) {
      counts[n] = c;
      if (bad) atomicOr(reinterpret_cast<unsigned*>(counts + n_pairs + 1), POPE_FLAG_NONFINITE_LSE);
      __threadfence();
      atomicExch(ready + n, 1);
    }
    if (phase == 1) return;
  } else {
    c = counts[n];
  }
  // exclusive prefix of the per-pair counts = where this pair's matches start
  int part = 0;
  for (int p = threadIdx.x; p < n; p += FT) {
    if (phase == 0) {
      const long long t0 = clock64();
      while (atomicAdd(ready + p, 0) == 0) {
        __nanosleep(64);
        if (clock64() - t0 > 4000000000ll) __trap();   // a pair with an earlier ticket always publishes: unreachable
      }
      __threadfence();
    }
    part += *reinterpret_cast<volatile int32_t*>(counts + p);
  }
  int base = block_sum(part, smem);
  if (n == n_pairs - 1 && threadIdx.x == 0) {
    // one match per row, so n * L always suffices; the documented minimum n * min(L, S) can fall short only when L > S and
    // several rows hold bit-identical confidences in one column (the reference emits them all, too)
    counts[n_pairs] = int32_t(min(int64_t(base + c), capacity));
    if (base + c > capacity) atomicOr(reinterpret_cast<unsigned*>(counts + n_pairs + 1), POPE_FLAG_CAPACITY);
  }
  for (int i0 = 0, k = 0; i0 < L; i0 += FT, ++k) {
    const int i = i0 + threadIdx.x;
    int j = 0; float t2 = 0.f;
    bool hit;
    if (keep && phase == 0) {
      hit = false;
#pragma unroll
      for (int q = 0; q < kKeep; ++q)
        if (q == k) { hit = (khit >> q) & 1u; j = kj[q]; t2 = kt[q]; }
    } else {
      hit = (i < L) && row_match(rowbest, colbest, L, S, i, g0, g1, j, t2);
    }
    const unsigned ballot = __ballot_sync(kFullMask, hit);
    __syncthreads();
    if (lane == 0) warp_off[warp] = __popc(ballot);
    __syncthreads();
    if (warp == 0) {   // exclusive scan of the 32 warp totals
      int v = warp_off[lane], incl = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(kFullMask, incl, o);
        if (lane >= o) incl += t;
      }
      warp_off[lane] = incl - v;
      if (lane == 31) smem[0] = incl;
    }
    __syncthreads();
    const int64_t pos = base + warp_off[warp] + __popc(ballot & ((1u << lane) - 1u));
    if (hit && pos < capacity) {
      b_ids[pos] = n; i_ids[pos] = i; j_ids[pos] = j;
      mconf[pos] = exp2f(t2);
      mk0[2 * pos + 0] = float(i % g0.w) * pixel_scale; mk0[2 * pos + 1] = float(i / g0.w) * pixel_scale;
      mk1[2 * pos + 0] = float(j % g1.w) * pixel_scale; mk1[2 * pos + 1] = float(j / g1.w) * pixel_scale;
    }
    base += smem[0];
  }
}

// Two-sweep path, between the sweeps: cell (i, j) can only have conf > thr if p_row(i, j) > thr, i.e. if its raw
// accumulator exceeds (lse_r[i] + log2 thr) / scale.  One warp per aligned group of 32 rows writes that bound (margin on
// the safe side, +inf for non-finite lse) and the group's minimum.
__device__ __forceinline__ bool gate_closed(const int32_t* gate) {
  return gate && !(uint32_t(*gate) & POPE_FLAG_ROBUST_PATH);
}

__global__ void __launch_bounds__(256) gated_clear_kernel(uint4* __restrict__ ptr, size_t n16, const int32_t* __restrict__ gate) {
  if (gate_closed(gate)) return;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n16; i += size_t(gridDim.x) * blockDim.x)
    ptr[i] = make_uint4(0u, 0u, 0u, 0u);
}

__global__ void __launch_bounds__(256) cand_bounds_kernel(const float* __restrict__ lse_r, int n_pairs, int L, float scale,
                                                         float log2_thr, float* __restrict__ cbound,
                                                         float* __restrict__ cminb, const int32_t* __restrict__ gate) {
  if (gate_closed(gate)) return;
  const int lane = threadIdx.x & 31;
  const int nchunks = (L + 31) / 32;
  const size_t g = size_t(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (g >= size_t(n_pairs) * nchunks) return;
  const int n = int(g / nchunks), i = int(g - size_t(n) * nchunks) * 32 + lane;
  float bound = INFINITY;
  if (i < L) {
    const float inv_s = 1.f / scale;
    const float b = (lse_r[size_t(n) * L + i] + log2_thr) * inv_s;
    if (isfinite(b)) bound = b - (1e-5f * fabsf(b) + 0.005f * inv_s);
  }
  cbound[g * 32 + lane] = bound;        // rows padded to a multiple of 32 with +inf
  float mb = bound;
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) mb = fminf(mb, __shfl_xor_sync(kFullMask, mb, o));
  if (lane == 0) cminb[g] = mb;
}

// One thread per row i: evaluate the (at most kCandSlots) cells with p_row > thr found by the column sweep.
__global__ void __launch_bounds__(256) cand_eval_kernel(const int* __restrict__ cand_cnt, const u64* __restrict__ cand,
                                                       const float* __restrict__ lse_r, const float* __restrict__ lse_c,
                                                       int n_pairs, int L, int S, float scale, float log2_thr,
                                                       u64* __restrict__ rowbest, u64* __restrict__ colbest,
                                                       const int32_t* __restrict__ gate) {
  if (gate_closed(gate)) return;
  const size_t r = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r >= size_t(n_pairs) * L) return;
  const int n = int(r / L), i = int(r - size_t(n) * L);
  const int c = min(cand_cnt[r], kCandSlots);
  if (c == 0) return;
  const float lr = lse_r[r];
  u64 best = 0;
  for (int k = 0; k < c; ++k) {
    const u64 rec = cand[r * kCandSlots + k];
    const int j = int(uint32_t(rec));
    const float x = __uint_as_float(uint32_t(rec >> 32)) * scale;
    const float t2 = (x - lr) + (x - lse_c[size_t(n) * S + j]);
    if (t2 > log2_thr) {
      const u64 mine = pack_best(t2, j);
      best = mine > best ? mine : best;
      atomicMax(colbest + size_t(n) * S + j, pack_best(t2, i));
    }
  }
  rowbest[r] = best;
}

// tcgen05 paths: one thread per row i evaluates the cells its four epilogue threads listed during the row sweep (a
// superset of the cells with p_row > thr: the test there ran against the RUNNING row sum, which only grows).  Same
// arithmetic as cand_eval_kernel; the lists hold similarities in log2 units on every path (single sweep, its gated
// online-softmax redo of flagged pairs, two-sweep).  mode 2 (fp32 split path): nothing at all once POPE_FLAG_ROBUST_PATH is set -- the
// fp32-FMA fallback evaluates its own lists.
// Every row's rowbest is written (0 = no candidate), so the caller need not clear it.
__global__ void __launch_bounds__(256) cand_eval_lists_kernel(const int* __restrict__ cand_cnt, const u64* __restrict__ cand,
                                                             const float* __restrict__ lse_r,
                                                             const float* __restrict__ lse_c, int n_pairs, int L, int S,
                                                             float scale, float log2_thr, u64* __restrict__ rowbest,
                                                             u64* __restrict__ colbest, const int32_t* __restrict__ flags,
                                                             int mode) {
  if (mode == 2 && (uint32_t(*flags) & POPE_FLAG_ROBUST_PATH)) return;
  const size_t r = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r >= size_t(n_pairs) * L) return;
  const uint2 cw = reinterpret_cast<const uint2*>(cand_cnt)[r];       // four uint16: nibble counts of the row's 16 sub-lists
  u64 best = 0;
  if ((cw.x | cw.y) != 0u) {
    const int n = int(r / L), i = int(r - size_t(n) * L);
    const float lr = lse_r[r];
    u64 todo = u64(cw.x) | (u64(cw.y) << 32);
    while (todo) {                                          // the non-empty sub-lists only (usually one or two per row)
      const int sub = (__ffsll((long long)todo) - 1) >> 2;
      const int c = min(int((todo >> (4 * sub)) & 0xfull), kLaneSlots);
      todo &= ~(0xfull << (4 * sub));
      for (int k = 0; k < c; ++k) {
        const u64 rec = cand[(r * (kListGroups * kListStride / kLaneSlots) + sub) * kLaneSlots + k];
        const int j = int(uint32_t(rec));
        const float x = __uint_as_float(uint32_t(rec >> 32));   // the similarity in log2 units
        if (!(x - lr > log2_thr - 0.01f)) continue;         // conf <= p_row: stale entries of the running-bound test go here
        const float t2 = (x - lr) + (x - lse_c[size_t(n) * S + j]);
        if (t2 > log2_thr) {
          const u64 mine = pack_best(t2, j);
          best = mine > best ? mine : best;
          atomicMax(colbest + size_t(n) * S + j, pack_best(t2, i));
        }
      }
    }
  }
  rowbest[r] = best;
}

// single-sweep path: column log-sum-exp from the per-32-row partial sums written by the sweep.  Partial (group g, column j)
// is a sum of 2^(x - m) with m = cshift[g][j / 32], an integer, so moving it to another domain is an exact scaling; the
// column's total is taken in the domain of the largest shift that occurs (usually all shifts of a column are equal and no
// exponential is evaluated).  A total below 2^-90 of that domain may have lost terms to underflow: the pair is flagged for
// the online-softmax kernels.  Also clears the column's best-candidate record and the pair's "count published" word (the
// single-sweep launch sequence needs no memset of the scratch).  HBM-bound: n * ceil(L/32) * S * 4 bytes are read once; a
// thread owns VEC adjacent columns (inside one 32-column block) for every eighth row group.
constexpr int kMergeWarps = 4, kMergeBatch = 5;    // tools/micro/merge_bench.cu: 41 us for 184 MB (8 warps: 45, 16 warps: 76)
template <int VEC>
__global__ void __launch_bounds__(32 * kMergeWarps) colsum_reduce_kernel(const float* __restrict__ colpart,
                                                                        const float* __restrict__ cshift, int ngroups, int S,
                                                                        int nblk, float* __restrict__ lse_c,
                                                                        u64* __restrict__ colbest, int* __restrict__ ready,
                                                                        int32_t* __restrict__ flags, int* __restrict__ pairflag,
                                                                        int n_base, int n_total, int wait_from) {
  // block = kMergeWarps warps x (32 lanes x VEC adjacent columns): warp w takes the row groups g = w, w + kMergeWarps, ...,
  // kMergeBatch loads in flight per thread; the partial results of a column meet in shared memory and are merged in warp
  // order (fixed order: deterministic)
  __shared__ float s_acc[kMergeWarps][32 * VEC];
  __shared__ float s_m[kMergeWarps][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int j = (blockIdx.x * 32 + lane) * VEC, n = n_base + blockIdx.y;
  // launched beside the tail of a split sweep (programmatic stream serialisation): the pairs of the head are complete, the
  // blocks of the tail's pairs -- the last ones of the grid -- wait here until that sweep has finished and its stores are visible
  if (n >= wait_from) asm volatile("griddepcontrol.wait;" ::: "memory");
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    ready[n] = 0;
    if (n == 0) ready[n_total] = 0;                   // the compaction kernel's ticket counter
  }
  float acc[VEC], mtop = -INFINITY;
#pragma unroll
  for (int v = 0; v < VEC; ++v) acc[v] = 0.f;
  if (j < S) {
    const float* p = colpart + size_t(n) * ngroups * S + j;
    const float* sh = cshift + size_t(n) * ngroups * nblk + (j >> 5);
    for (int g0 = warp; g0 < ngroups; g0 += kMergeWarps * kMergeBatch) {
      float q[kMergeBatch][VEC], m[kMergeBatch];
#pragma unroll
      for (int b = 0; b < kMergeBatch; ++b) {
        const int g = g0 + b * kMergeWarps;
        m[b] = -INFINITY;                                  // past the last group: (0, -inf) is neutral below
#pragma unroll
        for (int v = 0; v < VEC; ++v) q[b][v] = 0.f;
        if (g < ngroups) {
          if (VEC == 4) {
            const float4 t = __ldcs(reinterpret_cast<const float4*>(p + size_t(g) * S));
            q[b][0] = t.x; q[b][1 % VEC] = t.y; q[b][2 % VEC] = t.z; q[b][3 % VEC] = t.w;
          } else {
            q[b][0] = __ldcs(p + size_t(g) * S);
          }
          m[b] = __ldg(sh + size_t(g) * nblk);
        }
      }
#pragma unroll
      for (int b = 0; b < kMergeBatch; ++b) {
        if (m[b] == mtop) {
#pragma unroll
          for (int v = 0; v < VEC; ++v) acc[v] += q[b][v];
        } else if (m[b] < mtop) {
#pragma unroll
          for (int v = 0; v < VEC; ++v) acc[v] += scale_pow2(q[b][v], m[b] - mtop);
        } else {                                          // also the first group (mtop = -inf, zero sums)
#pragma unroll
          for (int v = 0; v < VEC; ++v) acc[v] = scale_pow2(acc[v], mtop - m[b]) + q[b][v];
          mtop = m[b];
        }
      }
    }
  }
#pragma unroll
  for (int v = 0; v < VEC; ++v) s_acc[warp][lane * VEC + v] = acc[v];
  s_m[warp][lane] = mtop;
  __syncthreads();
  if (warp != 0 || j >= S) return;
  float mall = -INFINITY;
#pragma unroll
  for (int w = 0; w < kMergeWarps; ++w) mall = fmaxf(mall, s_m[w][lane]);
  bool bad = false;
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < kMergeWarps; ++w) {
      const float a = s_acc[w][lane * VEC + v], m = s_m[w][lane];
      tot += (m == mall) ? a : scale_pow2(a, m - mall);          // a warp without groups holds (0, -inf)
    }
    lse_c[size_t(n) * S + j + v] = mall + log2f(tot);
    colbest[size_t(n) * S + j + v] = 0ull;
    bad |= !(tot > kSumLo && tot < kSumHi);
  }
  if (bad) {
    atomicOr(reinterpret_cast<unsigned*>(flags), POPE_FLAG_ROBUST_PATH);
    *reinterpret_cast<volatile int*>(pairflag + n) = 1;
  }
}

}  // namespace

cudaError_t cand_bounds_run(const CoarseProblem& p, const CoarseScratch& w, cudaStream_t st, const int32_t* gate) {
  const size_t groups = size_t(p.n) * ((p.L + 31) / 32);
  cand_bounds_kernel<<<unsigned((groups + 7) / 8), 256, 0, st>>>(w.lse_r, p.n, p.L, p.scale_log2, p.log2_thr, w.cbound, w.cminb,
                                                                gate);
  return cudaGetLastError();
}

cudaError_t cand_eval_run(const CoarseProblem& p, const CoarseScratch& w, cudaStream_t st, const int32_t* gate) {
  const size_t rows = size_t(p.n) * p.L;
  cand_eval_kernel<<<unsigned((rows + 255) / 256), 256, 0, st>>>(w.cand_cnt, w.cand, w.lse_r, w.lse_c, p.n, p.L, p.S,
                                                                p.scale_log2, p.log2_thr, w.rowbest, w.colbest, gate);
  return cudaGetLastError();
}

cudaError_t gated_clear_run(void* ptr, size_t bytes, const int32_t* gate, cudaStream_t st) {
  gated_clear_kernel<<<148 * 2, 256, 0, st>>>(static_cast<uint4*>(ptr), bytes / 16, gate);
  return cudaGetLastError();
}

cudaError_t cand_eval_lists_run(const CoarseProblem& p, const CoarseScratch& w, const int32_t* flags, int mode,
                                cudaStream_t st) {
  const size_t rows = size_t(p.n) * p.L;
  cand_eval_lists_kernel<<<unsigned((rows + 255) / 256), 256, 0, st>>>(w.cand_cnt, w.cand, w.lse_r, w.lse_c, p.n, p.L, p.S,
                                                                      p.scale_log2, p.log2_thr, w.rowbest, w.colbest, flags, mode);
  return cudaGetLastError();
}

cudaError_t colsum_reduce_run(const CoarseProblem& p, const CoarseScratch& w, int32_t* flags, cudaStream_t st, int n_base,
                              int n_count, int wait_from) {
  const bool overlap_previous = wait_from >= 0;
  if (wait_from < 0) wait_from = 0x7fffffff;
  const int ngroups = (p.L + 31) / 32, nblk = (p.S + 31) / 32;
  const bool vec4 = p.S % 4 == 0;   // rows of the partial-sum array are then 16-byte aligned (the array itself is 256-byte aligned)
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = vec4 ? dim3((p.S / 4 + 31) / 32, n_count) : dim3((p.S + 31) / 32, n_count);
  cfg.blockDim = dim3(32 * kMergeWarps);
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = overlap_previous ? 1 : 0;
  const float* colpart = w.colpart;
  const float* cshift = w.cshift;
  if (vec4)
    return cudaLaunchKernelEx(&cfg, colsum_reduce_kernel<4>, colpart, cshift, ngroups, p.S, nblk, w.lse_c, w.colbest, w.ready, flags,
                              w.pairflag, n_base, p.n, wait_from);
  return cudaLaunchKernelEx(&cfg, colsum_reduce_kernel<1>, colpart, cshift, ngroups, p.S, nblk, w.lse_c, w.colbest, w.ready, flags,
                            w.pairflag, n_base, p.n, wait_from);
}

cudaError_t coarse_finalize_run(const CoarseProblem& p, const CoarseScratch& w, int64_t* b_ids, int64_t* i_ids,
                                int64_t* j_ids, float* mconf, float* mk0, float* mk1, int32_t* counts, int64_t capacity,
                                cudaStream_t st) {
  Grid2 g0{p.h0c, p.w0c, p.border}, g1{p.h1c, p.w1c, p.border};
  int dev = 0, sms = 0;
  cudaError_t e;
  if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
  if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
  const bool one_launch = p.n <= 2 * sms;            // two 1024-thread CTAs per SM: every CTA of the grid is resident
  for (int phase = one_launch ? 0 : 1; phase <= (one_launch ? 0 : 2); ++phase)
    count_emit_kernel<<<p.n, FT, 0, st>>>(w.rowbest, w.colbest, w.lse_r, w.lse_c, p.L, p.S, g0, g1, p.pixel_scale, counts, p.n,
                                          w.ready, b_ids, i_ids, j_ids, mconf, mk0, mk1, phase, capacity);
  return cudaGetLastError();
}

}  // namespace pope
