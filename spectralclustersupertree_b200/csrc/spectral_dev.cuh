// Device code shared by the per-node spectral path (spectral.cu) and the batched medium-node path (medium.cu):
// reductions, the streamed row dot product of the Laplacian matvec, the projected tridiagonal eigenproblem, one
// Lanczos step tail, the Ritz vector, the true residual and the exact 1-D 2-means -- each as a __device__ function
// executed by one CTA, so that "one CTA per launch" (spectral.cu) and "one CTA per node of a batch" (medium.cu)
// run the same arithmetic.  See spectral.cu for what is computed and why.
#pragma once

#include "common.cuh"

#include <cmath>

namespace scs {
namespace specdev {

constexpr int kMaxBasis = 256;        // Lanczos vectors per (re)start
constexpr int kMaxRestarts = 8;
constexpr double kResidualTol = 1e-12;  // |beta_j s_j| of the Fiedler Ritz pair (|N| <= 1)
constexpr double kBreakdown = 1e-11;  // beta_j below this: the Krylov space is invariant
constexpr double kGapTie = 1e-7;      // lambda_3 - lambda_2 below this: eigenvector is ill-defined
constexpr double kMarginTie = 1e-9;   // a vertex this close (relative) to the 2-means boundary

constexpr int kMvThreads = 256;
constexpr int kVecThreads = 256;
constexpr int kOneCta = 1024;

// ---- small helpers --------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
    return v;
}

// Sum over the block in a fixed order; result valid in every thread.  `scratch` holds 33 doubles.
__device__ __forceinline__ double block_sum(double v, double *scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();  // scratch may still be read from a previous call
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    if (warp == 0) {
        double t = lane < nwarp ? scratch[lane] : 0.0;
        t = warp_sum(t);
        if (lane == 0) scratch[32] = t;
    }
    __syncthreads();
    return scratch[32];
}

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// ---- the operator: y = isd .* (W z) ----------------------------------------------------------
// Streams W exactly once with 16-byte loads that bypass L1 allocation; z (8 m bytes) stays in
// L1/L2.  Rows are 8 m bytes apart, so odd rows of an odd-m matrix start on an 8-byte boundary:
// each row is split into an optional 1-element head, an aligned double2 body and a tail.
__device__ __forceinline__ double2 load_stream(const double2 *p) {
    double2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}

// Partial dot product of one row with z over the calling group of `nthr` threads (`t` = rank).
__device__ __forceinline__ double row_dot_partial(const double *__restrict__ row, const double *__restrict__ z,
                                                  int m, int t, int nthr) {
    const int head = static_cast<int>((reinterpret_cast<uintptr_t>(row) >> 3) & 1u);
    const int nvec = (m - head) >> 1;
    const double2 *body = reinterpret_cast<const double2 *>(row + head);
    const double *zb = z + head;
    double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
    int v = t;
    // four independent 16-byte loads in flight per thread per iteration
    for (; v + 3 * nthr < nvec; v += 4 * nthr) {
        const double2 a = load_stream(body + v);
        const double2 b = load_stream(body + v + nthr);
        const double2 c = load_stream(body + v + 2 * nthr);
        const double2 d = load_stream(body + v + 3 * nthr);
        acc0 = fma(a.x, zb[2 * v], acc0);
        acc0 = fma(a.y, zb[2 * v + 1], acc0);
        acc1 = fma(b.x, zb[2 * (v + nthr)], acc1);
        acc1 = fma(b.y, zb[2 * (v + nthr) + 1], acc1);
        acc2 = fma(c.x, zb[2 * (v + 2 * nthr)], acc2);
        acc2 = fma(c.y, zb[2 * (v + 2 * nthr) + 1], acc2);
        acc3 = fma(d.x, zb[2 * (v + 3 * nthr)], acc3);
        acc3 = fma(d.y, zb[2 * (v + 3 * nthr) + 1], acc3);
    }
    // what is left of the row (fewer than 4 nthr vectors): again all loads at once, clamped to the last vector
    // and masked, not one dependent load after the other
    if (v < nvec) {
        const int last = nvec - 1;
        const int vb = v + nthr, vc = v + 2 * nthr;
        const double mb = vb < nvec ? 1.0 : 0.0, mc = vc < nvec ? 1.0 : 0.0;
        const int ib = min(vb, last), ic = min(vc, last);
        const double2 a = load_stream(body + v);
        const double2 b = load_stream(body + ib);
        const double2 c = load_stream(body + ic);
        acc0 = fma(a.x, zb[2 * v], acc0);
        acc0 = fma(a.y, zb[2 * v + 1], acc0);
        acc1 = fma(b.x * mb, zb[2 * ib], acc1);
        acc1 = fma(b.y * mb, zb[2 * ib + 1], acc1);
        acc2 = fma(c.x * mc, zb[2 * ic], acc2);
        acc2 = fma(c.y * mc, zb[2 * ic + 1], acc2);
    }
    double acc = (acc0 + acc1) + (acc2 + acc3);
    if (t == 0) {
        if (head) acc = fma(row[0], z[0], acc);
        if ((m - head) & 1) acc = fma(row[m - 1], z[m - 1], acc);
    }
    return acc;
}

// ---- projected problem ------------------------------------------------------------------------
// T = tridiag(alpha[1..j]; beta[1..j-1]).  Finds its two largest eigenvalues by Sturm-count
// multi-section (blockDim.x probes per round), the eigenvector s of the largest by a twisted
// factorisation, and the Lanczos residual estimate beta[j] * |s_j|.
// If state[0] latched a breakdown step jb <= j, only the leading jb x jb block is solved.
// out: [0] theta1, [1] theta2 (NaN if the block is 1 x 1), [2] residual estimate, [3] beta[j_eff],
//      [4] j_eff
__device__ __forceinline__ int sturm_count(const double *a, const double *b2, int j, double x, double pivmin) {
    int cnt = 0;
    double q = a[1] - x;
    if (fabs(q) < pivmin) q = -pivmin;
    cnt += q < 0.0;
    for (int i = 2; i <= j; ++i) {
        q = (a[i] - x) - b2[i - 1] / q;
        if (fabs(q) < pivmin) q = -pivmin;
        cnt += q < 0.0;
    }
    return cnt;  // number of eigenvalues below x
}

// All threads of the CTA take part.  sm: 5 (jrun + 2) doubles, counts: blockDim.x ints, pair: 4 doubles.
// (No __restrict__ here: the fused tail writes alpha / beta / state in the same launch that reads them.)
__device__ inline void tridiag_solve(int jrun, const double *alpha, const double *beta, const int32_t *state, double *coef,
                              double *out, double *sm, int *counts, double *pair) {
    const int latched = state[0];
    const int j = latched > 0 && latched < jrun ? latched : jrun;
    double *a = sm;                 // [j + 2], 1-based
    double *b = a + (j + 2);        // [j + 2]
    double *b2 = b + (j + 2);       // squares
    double *dplus = b2 + (j + 2);
    double *dminus = dplus + (j + 2);
    double *bounds = pair, *theta = pair + 2;
    const int tid = threadIdx.x, P = blockDim.x;
    __syncthreads();  // sm / counts may still be in use by the caller
    for (int i = tid + 1; i <= j; i += P) {
        a[i] = alpha[i];
        b[i] = beta[i];
        b2[i] = beta[i] * beta[i];
    }
    __syncthreads();
    if (tid == 0) {
        double lo = a[1], hi = a[1], bmax = 0.0;
        for (int i = 1; i <= j; ++i) {
            const double left = i > 1 ? fabs(b[i - 1]) : 0.0;
            const double right = i < j ? fabs(b[i]) : 0.0;
            lo = fmin(lo, a[i] - left - right);
            hi = fmax(hi, a[i] + left + right);
            bmax = fmax(bmax, right);
        }
        const double span = fmax(hi - lo, 1e-300);
        bounds[0] = lo - 1e-12 * span - 1e-300;
        bounds[1] = hi + 1e-12 * span + 1e-300;
    }
    __syncthreads();
    const double pivmin = 1e-290;
    const double glo = bounds[0], ghi = bounds[1];
    const int wanted = j >= 2 ? 2 : 1;
    for (int which = 0; which < wanted; ++which) {
        const int k = j - which;  // k-th smallest eigenvalue
        double lo = glo, hi = which == 0 ? ghi : theta[0];
        if (which == 1) hi = hi + fabs(hi) * 4.5e-16 + 1e-300;
        for (int round = 0; round < 9; ++round) {
            const double step = (hi - lo) / (P + 1);
            const double x = lo + step * (tid + 1);
            counts[tid] = sturm_count(a, b2, j, x, pivmin);
            __syncthreads();
            // the first probe with count >= k bounds the eigenvalue from above; counts are monotone,
            // so exactly one thread sees the step (or the last thread sees none)
            {
                const bool here = counts[tid] >= k;
                const bool before = tid > 0 && counts[tid - 1] >= k;
                if (here && !before) {
                    bounds[0] = tid == 0 ? lo : lo + step * tid;
                    bounds[1] = lo + step * (tid + 1);
                } else if (tid == P - 1 && !here) {
                    bounds[0] = lo + step * P;
                    bounds[1] = hi;
                }
            }
            __syncthreads();
            lo = bounds[0];
            hi = bounds[1];
            __syncthreads();
            if (!(hi > lo) || (hi - lo) <= 2.3e-16 * fmax(fabs(lo), fabs(hi))) break;
        }
        if (tid == 0) theta[which] = 0.5 * (lo + hi);
        __syncthreads();
    }
    if (tid == 0) {
        const double th = theta[0];
        // twisted factorisation of T - th I
        double d = a[1] - th;
        if (fabs(d) < pivmin) d = pivmin;
        dplus[1] = d;
        for (int i = 2; i <= j; ++i) {
            d = (a[i] - th) - b2[i - 1] / dplus[i - 1];
            if (fabs(d) < pivmin) d = pivmin;
            dplus[i] = d;
        }
        d = a[j] - th;
        if (fabs(d) < pivmin) d = pivmin;
        dminus[j] = d;
        for (int i = j - 1; i >= 1; --i) {
            d = (a[i] - th) - b2[i] / dminus[i + 1];
            if (fabs(d) < pivmin) d = pivmin;
            dminus[i] = d;
        }
        int kbest = 1;
        double gbest = INFINITY;
        for (int i = 1; i <= j; ++i) {
            const double g = fabs(dplus[i] + dminus[i] - (a[i] - th));
            if (g < gbest) { gbest = g; kbest = i; }
        }
        // solve outward from the twist index (coef reuses no shared memory: written straight out)
        // going down needs dplus[i] for i < kbest, going up needs dminus[i] for i > kbest;
        // the unnormalised vector is kept in b2 (no longer needed)
        double prev = 1.0;
        double norm2 = 1.0;
        b2[kbest] = 1.0;
        for (int i = kbest - 1; i >= 1; --i) {
            prev = -(b[i] / dplus[i]) * prev;
            b2[i] = prev;
            norm2 += prev * prev;
        }
        prev = 1.0;
        for (int i = kbest + 1; i <= j; ++i) {
            prev = -(b[i - 1] / dminus[i]) * prev;
            b2[i] = prev;
            norm2 += prev * prev;
        }
        const double inv = 1.0 / sqrt(norm2);
        for (int i = 1; i <= j; ++i) coef[i] = b2[i] * inv;
        for (int i = j + 1; i <= jrun; ++i) coef[i] = 0.0;  // vectors after a breakdown are noise
        coef[0] = 0.0;
        out[4] = static_cast<double>(j);
        out[0] = th;
        out[1] = j >= 2 ? theta[1] : nan("");
        out[2] = fabs(b[j] * b2[j] * inv);
        out[3] = b[j];
    }
    __syncthreads();
}


__device__ __forceinline__ void prepare_scaling_body(int m, const double *__restrict__ degree, double *__restrict__ isd, double *__restrict__ q0,
                int32_t *__restrict__ flags) {
    __shared__ double scratch[33];
    double vol = 0.0;
    bool bad = false;
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
        const double d = degree[i];
        if (!(d >= 0.0) || isinf(d)) bad = true;
        vol += d > 0.0 ? d : 0.0;
    }
    vol = block_sum(vol, scratch);
    if (bad) flags[0] = 1;
    const double inv_norm = vol > 0.0 ? 1.0 / sqrt(vol) : 0.0;
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
        const double d = degree[i];
        const double s = d > 0.0 ? sqrt(d) : 0.0;
        isd[i] = d > 0.0 ? 1.0 / s : 1.0;
        q0[i] = s * inv_norm;
    }
}

__device__ __forceinline__ void lanczos_tail_body(int m, int j, int nb, int hj, int check, const double *basis, const double *__restrict__ w_in,
             const double *__restrict__ isd, double *alpha, double *beta, double *next, double *__restrict__ z,
             int32_t *state, double *coef, double *ritz) {
    extern __shared__ double dyn[];
    double *w = dyn;                      // [m]
    double *h = w + m;                    // [nb] coefficients of one pass
    double *tri = h + (kMaxBasis + 8);    // 5 (j + 2) doubles for the projected problem
    __shared__ double scratch[33];
    __shared__ int counts[kOneCta];
    __shared__ double pair[4];
    if (state[1]) return;  // converged earlier in this chunk
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < m; i += kOneCta) w[i] = w_in[i];
    __syncthreads();
    double hkeep = 0.0;  // coefficient along v_j summed over both passes (thread 0 of warp owning it)
    for (int pass = 0; pass < 2; ++pass) {
        // h[k] = basis[k] . w : one warp per basis vector
        for (int k = warp; k < nb; k += kOneCta / 32) {
            const double *b = basis + static_cast<size_t>(k) * m;
            double acc = 0.0;
            for (int i = lane; i < m; i += 32) acc = fma(b[i], w[i], acc);
            acc = warp_sum(acc);
            if (lane == 0) h[k] = acc;
        }
        __syncthreads();
        if (tid == 0 && j > 0) hkeep += h[hj];
        // w -= sum_k h[k] basis[k]
        for (int i = tid; i < m; i += kOneCta) {
            double acc = w[i];
            for (int k = 0; k < nb; ++k) acc = fma(-h[k], basis[static_cast<size_t>(k) * m + i], acc);
            w[i] = acc;
        }
        __syncthreads();
    }
    double ss = 0.0;
    for (int i = tid; i < m; i += kOneCta) ss = fma(w[i], w[i], ss);
    ss = block_sum(ss, scratch);
    const double bnorm = sqrt(ss);
    if (tid == 0) {
        beta[j] = bnorm;
        if (j > 0) {
            alpha[j] = hkeep;
            if (bnorm <= kBreakdown && state[0] == 0) state[0] = j;
        } else {
            state[0] = 0;
        }
    }
    const double inv = bnorm > 0.0 ? 1.0 / bnorm : 0.0;
    for (int i = tid; i < m; i += kOneCta) {
        const double v = w[i] * inv;
        next[i] = v;
        z[i] = isd[i] * v;
    }
    __syncthreads();  // alpha / beta / state written by thread 0 are read below
    if (check && j > 0) {
        tridiag_solve(j, alpha, beta, state, coef, ritz, tri, counts, pair);
        if (tid == 0) {
            ritz[5] = static_cast<double>(j);
            if (ritz[3] <= kBreakdown || ritz[2] <= kResidualTol) {
                state[1] = 1;
                state[2] = j;
            }
        }
    }
}

__device__ __forceinline__ void ritz_vector_body(int m, int j, const double *__restrict__ basis, const double *__restrict__ coef,
            const double *__restrict__ isd, double *__restrict__ y, double *__restrict__ z) {
    __shared__ double scratch[33];
    extern __shared__ double cs[];
    for (int k = threadIdx.x; k <= j; k += blockDim.x) cs[k] = coef[k];
    __syncthreads();
    double ss = 0.0;
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
        double acc = 0.0;
        for (int k = 1; k <= j; ++k) acc = fma(cs[k], basis[static_cast<size_t>(k) * m + i], acc);
        y[i] = acc;
        ss = fma(acc, acc, ss);
    }
    ss = block_sum(ss, scratch);
    const double inv = ss > 0.0 ? 1.0 / sqrt(ss) : 0.0;
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
        const double v = y[i] * inv;
        y[i] = v;
        z[i] = isd[i] * v;
    }
}

__device__ __forceinline__ void true_residual_body(int m, const double *__restrict__ y, const double *__restrict__ Ny, const double *__restrict__ ritz,
              double *__restrict__ out) {
    __shared__ double scratch[33];
    const double th = ritz[0];
    double rr = 0.0, rq = 0.0;
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
        const double r = Ny[i] - th * y[i];
        rr = fma(r, r, rr);
        rq = fma(y[i], Ny[i], rq);
    }
    rr = block_sum(rr, scratch);
    rq = block_sum(rq, scratch);
    if (threadIdx.x == 0) {
        out[0] = sqrt(rr);
        out[1] = rq;
    }
}

// kStage: 0 everything in this CTA; 1 only the sign, u and the unsorted keys (sorted_g), for a sort by many CTAs
// (spectral.cu: sort_keys); 2 the keys in sorted_g are already sorted: everything after the sort.
template <bool kShared, int kStage = 0>
__device__ __forceinline__ void two_means_1d_body(int m, int P, const double *__restrict__ y, const double *__restrict__ isd, double *__restrict__ u,
             double *__restrict__ sorted_g, int32_t *__restrict__ side, double *__restrict__ result) {
    extern __shared__ double sort_s[];
    __shared__ double scratch[33];
    __shared__ double best_score[32];
    __shared__ int best_index[32];
    __shared__ double bcast[4];
    __shared__ int ibcast;
    double *keys = kShared ? sort_s : sorted_g;
    const int tid = threadIdx.x, nthr = blockDim.x;
    double total = 0.0;
    if (kStage != 2) {
    // sign: entry of largest magnitude (smallest index on ties) must be positive
    double amax = -1.0;
    int imax = 0x7fffffff;
    for (int i = tid; i < m; i += nthr) {
        const double v = fabs(y[i] * isd[i]);
        if (v > amax) { amax = v; imax = i; }
    }
    for (int off = 16; off > 0; off >>= 1) {
        const double oa = __shfl_down_sync(0xffffffffu, amax, off);
        const int oi = __shfl_down_sync(0xffffffffu, imax, off);
        if (oa > amax || (oa == amax && oi < imax)) { amax = oa; imax = oi; }
    }
    if ((tid & 31) == 0) { best_score[tid >> 5] = amax; best_index[tid >> 5] = imax; }
    __syncthreads();
    if (tid == 0) {
        double ba = best_score[0];
        int bi = best_index[0];
        for (int w = 1; w < (nthr >> 5); ++w)
            if (best_score[w] > ba || (best_score[w] == ba && best_index[w] < bi)) { ba = best_score[w]; bi = best_index[w]; }
        ibcast = bi;
    }
    __syncthreads();
    const int arg = ibcast < m ? ibcast : 0;
    const double sign = y[arg] * isd[arg] < 0.0 ? -1.0 : 1.0;

    for (int i = tid; i < P; i += nthr) {
        double v = INFINITY;
        if (i < m) {
            v = sign * (y[i] * isd[i]);
            u[i] = v;
            total += v;
        }
        keys[i] = v;
    }
    if (kStage == 1) return;
    } else {
        // the same partial sums in the same order as the stage that wrote u
        for (int i = tid; i < m; i += nthr) total += u[i];
    }
    total = block_sum(total, scratch);
    const double mean = total / m;
    __syncthreads();

    // bitonic sort, ascending
    for (int k = 2; kStage == 0 && k <= P; k <<= 1) {
        for (int jj = k >> 1; jj > 0; jj >>= 1) {
            for (int i = tid; i < P; i += nthr) {
                const int partner = i ^ jj;
                if (partner > i) {
                    const double x0 = kShared ? keys[i] : __ldcg(keys + i);
                    const double x1 = kShared ? keys[partner] : __ldcg(keys + partner);
                    const bool up = (i & k) == 0;
                    if ((x0 > x1) == up) {
                        if (kShared) { keys[i] = x1; keys[partner] = x0; }
                        else { __stcg(keys + i, x1); __stcg(keys + partner, x0); }
                    }
                }
            }
            __syncthreads();
        }
    }

    // prefix sums of the centred sorted values: thread t owns a contiguous segment
    const int seg = (m + nthr - 1) / nthr;
    const int lo = min(tid * seg, m), hi = min(lo + seg, m);
    double local = 0.0;
    for (int i = lo; i < hi; ++i) local += (kShared ? keys[i] : __ldcg(keys + i)) - mean;
    // exclusive scan of the per-thread sums (warp scan, then warps in order)
    double incl = local;
    for (int off = 1; off < 32; off <<= 1) {
        const double o = __shfl_up_sync(0xffffffffu, incl, off);
        if ((tid & 31) >= off) incl += o;
    }
    __syncthreads();
    if ((tid & 31) == 31) scratch[tid >> 5] = incl;
    __syncthreads();
    double before = 0.0;
    for (int w = 0; w < (tid >> 5); ++w) before += scratch[w];
    double all = 0.0;
    for (int w = 0; w < (nthr >> 5); ++w) all += scratch[w];
    double prefix = before + incl - local;  // sum of the elements before `lo`
    // split after i elements (i = 1..m-1) maximises P_i^2 / i + (S - P_i)^2 / (m - i)
    double bscore = -1.0;
    int bidx = 0x7fffffff;
    for (int i = lo; i < hi; ++i) {
        prefix += (kShared ? keys[i] : __ldcg(keys + i)) - mean;
        const int cnt = i + 1;
        if (cnt < m) {
            const double rest = all - prefix;
            const double score = prefix * prefix / cnt + rest * rest / (m - cnt);
            if (score > bscore) { bscore = score; bidx = cnt; }
        }
    }
    for (int off = 16; off > 0; off >>= 1) {
        const double os = __shfl_down_sync(0xffffffffu, bscore, off);
        const int oi = __shfl_down_sync(0xffffffffu, bidx, off);
        if (os > bscore || (os == bscore && oi < bidx)) { bscore = os; bidx = oi; }
    }
    __syncthreads();
    if ((tid & 31) == 0) { best_score[tid >> 5] = bscore; best_index[tid >> 5] = bidx; }
    __syncthreads();
    if (tid == 0) {
        double bs = best_score[0];
        int bi = best_index[0];
        for (int w = 1; w < (nthr >> 5); ++w)
            if (best_score[w] > bs || (best_score[w] == bs && best_index[w] < bi)) { bs = best_score[w]; bi = best_index[w]; }
        if (bi >= m) bi = 1;  // degenerate (m == 1 never reaches here)
        ibcast = bi;
        bcast[1] = bs;
    }
    __syncthreads();
    // Lloyd-stable splits (both neighbours of the cut on their own side of the centroid midpoint):
    // k-means as a local search can stop at any of them, so more than one means the reference's
    // answer depends on its random initialisation.  Count them and score the best runner-up.
    {
        const int chosen = ibcast;
        double p2 = before + incl - local;
        double stable = 0.0, runner = -1.0;
        for (int i = lo; i < hi; ++i) {
            const double key = kShared ? keys[i] : __ldcg(keys + i);
            p2 += key - mean;
            const int cnt = i + 1;
            if (cnt < m) {
                const double rest = all - p2;
                const double mid = mean + 0.5 * (p2 / cnt + rest / (m - cnt));
                const double next = kShared ? keys[i + 1] : __ldcg(keys + i + 1);
                if (key < mid && mid < next) {
                    stable += 1.0;
                    if (cnt != chosen) runner = fmax(runner, p2 * p2 / cnt + rest * rest / (m - cnt));
                }
            }
        }
        stable = block_sum(stable, scratch);
        for (int off = 16; off > 0; off >>= 1) runner = fmax(runner, __shfl_down_sync(0xffffffffu, runner, off));
        __syncthreads();
        if ((tid & 31) == 0) best_score[tid >> 5] = runner;
        __syncthreads();
        if (tid == 0) {
            double r = best_score[0];
            for (int w = 1; w < (nthr >> 5); ++w) r = fmax(r, best_score[w]);
            result[4] = stable;
            result[5] = (r >= 0.0 && bcast[1] > 0.0) ? r / bcast[1] : 0.0;
        }
    }
    if (tid == 0) {
        const int bi = ibcast;
        // centroids of the two parts
        double lower = 0.0;
        for (int i = 0; i < bi; ++i) lower += (kShared ? keys[i] : __ldcg(keys + i)) - mean;
        const double c0 = mean + lower / bi;
        const double c1 = mean + (all - lower) / (m - bi);
        const double mid = 0.5 * (c0 + c1);
        const double below = kShared ? keys[bi - 1] : __ldcg(keys + bi - 1);
        const double above = kShared ? keys[bi] : __ldcg(keys + bi);
        const double first = kShared ? keys[0] : __ldcg(keys);
        const double last = kShared ? keys[m - 1] : __ldcg(keys + m - 1);
        const double range = fmax(last - first, 1e-300);
        result[0] = fmin(fabs(below - mid), fabs(above - mid)) / range;
        result[1] = c0;
        result[2] = c1;
        result[3] = static_cast<double>(m - bi);
        bcast[0] = above;
    }
    __syncthreads();
    const double threshold = bcast[0];
    for (int i = tid; i < m; i += nthr) side[i] = u[i] >= threshold ? 1 : 0;
}

}  // namespace specdev
}  // namespace scs
