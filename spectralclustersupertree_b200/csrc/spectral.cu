// Spectral bipartition of the (contracted) proper cluster graph on the GPU.
//
// Replaces spectral_cluster_graph of the reference (/root/reference/src/sc_supertree/scs.py:210-258),
// i.e. sklearn.cluster.SpectralClustering(2, affinity="precomputed", assign_labels="kmeans"):
//   L = I - D^-1/2 W D^-1/2   (scipy csgraph.laplacian(normed=True): degree 0 is scaled by 1)
//   x_0, x_1 = eigenvectors of the two smallest eigenvalues (sklearn: shift-invert ARPACK after a
//              dense LU of L); embedding rows u_k = x_k / sqrt(d); sign flip; 2-means.
//
// Here nothing is factorised.  With N = D^-1/2 W D^-1/2 the smallest eigenpair of L is known in
// closed form (eigenvalue 0, x_0 = sqrt(d) / |sqrt(d)|), so the Fiedler pair is the LARGEST
// eigenpair of N on the orthogonal complement of x_0: Lanczos with full (twice-applied classical
// Gram-Schmidt) re-orthogonalisation against x_0 and the whole basis.  The operator is the fused
// matvec  y = isd .* (W (isd .* v))  -- W is read once per application, straight from HBM, in
// fp64 (the parity contract needs a 1e-3 tree-weight difference and 1e-6 eigenvalues to survive).
// The projected tridiagonal problem is solved on the device (Sturm multi-section + twisted
// factorisation), so the host only reads back a residual estimate to decide when to stop.
// u_0 is constant on a connected graph, so sklearn's 2-means on (u_0, u_1) is 1-D 2-means on the
// Fiedler coordinate; it is solved exactly (sort, prefix sums, best split) instead of by ten
// random Lloyd restarts, which makes the partition deterministic.

#include "common.cuh"
#include "shard.cuh"

#include <algorithm>
#include <chrono>
#include <cmath>

namespace scs {

namespace {

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

// ---- degree and scaling ---------------------------------------------------------------------
// one warp per row: d[a] = sum_b W[a][b]   (scipy _laplacian.py:543: column sums of a symmetric matrix)
__global__ void row_sums(int m, const double *__restrict__ W, double *__restrict__ degree) {
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= m) return;
    const double *w = W + static_cast<size_t>(row) * m;
    double acc = 0.0;
    for (int j = lane; j < m; j += 32) acc += w[j];
    acc = warp_sum(acc);
    if (lane == 0) degree[row] = acc;
}

// isd = 1/sqrt(d) (1 where d == 0, scipy _laplacian.py:544-545); q0 = sqrt(d)/|sqrt(d)|; one CTA.
// flags[0] is raised if a degree is negative or not finite (sklearn would produce NaNs there).
__global__ void __launch_bounds__(kOneCta)
prepare_scaling(int m, const double *__restrict__ degree, double *__restrict__ isd, double *__restrict__ q0,
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

__global__ void scale_vector(int m, const double *__restrict__ isd, const double *__restrict__ x,
                             double *__restrict__ z) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) z[i] = isd[i] * x[i];
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

// A CTA of kMvThreads threads works on kMvThreads / kGroup rows, kGroup threads each:
//   kGroup = 32   a warp per row: short rows (m < 2048);
//   kGroup = 256  a CTA per row: long rows (tens of KB), many waves.
// 64 and 128 threads per row exist for tuning (SCS_MATVEC_GROUP); measured under ncu on the C4 nodes they are
// within noise of 256 at every size from 2362 to 8765 columns (tools/matvec_bench.py, profiles/README.md): a
// 45 MB matrix takes 13 us whatever the shape -- launch ramp and tail, not the row split, separate it from the
// 7 us of the HBM roofline.
// The partial sums of a row are combined in a fixed order (lanes by shuffle tree, then warps in order), the
// same in the single-GPU and in the row-sharded kernel, so both produce the same bits.
__host__ __device__ constexpr int mv_group_of(int m) { return m < 2048 ? 32 : kMvThreads; }

// Row `row` of W (stored at `Wrow`) times z, by the kGroup threads of the calling group; the result is valid
// in the group's first thread.  `part` holds kMvThreads / 32 doubles.
template <int kGroup>
__device__ __forceinline__ double group_row_dot(const double *__restrict__ Wrow, const double *__restrict__ z, int m,
                                                double *part) {
    const int t = threadIdx.x % kGroup, warp = threadIdx.x >> 5;
    double acc = row_dot_partial(Wrow, z, m, t, kGroup);
    acc = warp_sum(acc);
    if (kGroup == 32) return acc;
    if ((threadIdx.x & 31) == 0) part[warp] = acc;
    __syncthreads();
    double s = 0.0;
    if (t == 0) {
#pragma unroll
        for (int w = 0; w < kGroup / 32; ++w) s += part[warp + w];
    }
    return s;
}

template <int kGroup>
__global__ void __launch_bounds__(kMvThreads)
matvec_rows(int m, const double *__restrict__ W, const double *__restrict__ isd, const double *__restrict__ z,
            double *__restrict__ y, const int32_t *__restrict__ done) {
    __shared__ double part[kMvThreads / 32];
    if (done && *done) return;
    constexpr int kRows = kMvThreads / kGroup;
    const int row = blockIdx.x * kRows + threadIdx.x / kGroup;
    const bool live = row < m;
    // rows past the end redo the last row (every thread must reach the barrier in group_row_dot)
    const double s = group_row_dot<kGroup>(W + static_cast<size_t>(live ? row : m - 1) * m, z, m, part);
    if (live && threadIdx.x % kGroup == 0) y[row] = isd[row] * s;
}

int launch_matvec_single(scs_ctx *ctx, int m, const double *W, const double *isd, const double *z, double *y,
                         const int32_t *done);

// threads per row; SCS_MATVEC_GROUP (32 / 64 / 128 / 256) overrides the choice for rows of >= 2048 columns (tuning)
int matvec_group(int m) {
    static const int forced = [] {
        const char *env = std::getenv("SCS_MATVEC_GROUP");
        const int v = env ? std::atoi(env) : 0;
        return (v == 32 || v == 64 || v == 128 || v == 256) ? v : 0;
    }();
    return forced && m >= 2048 ? forced : mv_group_of(m);
}

// Row-sharded operator fused with its all-gather (one process per GPU, rows [row0, row0 + gridDim.x) of
// W on this rank): every CTA computes one entry of y and stores it straight into the `vec` buffer of
// EVERY rank's exchange window over NVLink; the last CTA to finish raises this rank's flag in every
// window and then waits until every rank has raised its flag here, so when the kernel ends the whole
// vector is in this rank's window.  Every entry is computed by the same code whatever rank owns the row,
// so the assembled vector is bit-identical to the single-GPU matvec's.
// kWarpRows mirrors the single-GPU choice (one warp per row below 2048 columns, one CTA per row above), so
// that the summation order of every entry is the same in both paths.
template <int kGroup>
__global__ void __launch_bounds__(kMvThreads)
matvec_rows_allgather(int m, int row0, int nrows, const double *__restrict__ W, const double *__restrict__ isd,
                      const double *__restrict__ z, const PeerTable peers, size_t vec_offset,
                      unsigned long long epoch, unsigned long long timeout_ns, const int32_t *__restrict__ done) {
    __shared__ double part[kMvThreads / 32];
    __shared__ unsigned int last;
    if (done && *done) return;  // the same on every rank: nobody signals, nobody waits
    constexpr int kRows = kMvThreads / kGroup;
    const int local = blockIdx.x * kRows + threadIdx.x / kGroup;
    const bool live = local < nrows;
    double s = group_row_dot<kGroup>(W + static_cast<size_t>(live ? local : nrows - 1) * m, z, m, part);
    // the group's first warp stores the entry into every rank's window
    s = __shfl_sync(0xffffffffu, s, 0);
    if (live && threadIdx.x % kGroup < 32) {
        const int lane = threadIdx.x & 31;
        const double out = isd[row0 + local] * s;
        if (lane < peers.world) reinterpret_cast<double *>(peers.window[lane] + vec_offset)[row0 + local] = out;
        __threadfence_system();
    }
    __syncthreads();  // every store of this CTA is fenced before its ticket is drawn
    if (threadIdx.x == 0) {
        ShardHeader *mine = reinterpret_cast<ShardHeader *>(peers.window[peers.rank]);
        last = atomicInc(&mine->ticket, gridDim.x - 1) == gridDim.x - 1;
    }
    __syncthreads();
    if (last && threadIdx.x < 32) peer_signal_and_wait(peers, epoch, timeout_ns);
}

int launch_matvec(scs_ctx *ctx, int m, const double *W, const double *isd, const double *z, double **y,
                  const int32_t *done = nullptr, RowBlock rows = RowBlock()) {
    if (rows.sharded()) {
        // the result lands in the window (double-buffered); *y is redirected there
        const ShardState &sh = ctx->shard;
        const ShardMatvecTicket ticket = shard_next_matvec(ctx);
        PeerTable peers;
        for (int r = 0; r < kMaxPeers; ++r) peers.window[r] = sh.peer[r];
        peers.rank = sh.rank;
        peers.world = sh.world;
        const int nrows = rows.row1 - rows.row0;
        *y = reinterpret_cast<double *>(sh.window + ticket.vec_offset);
        if (nrows <= 0) return fail(ctx, SCS_ERR_INVALID, "sharded matvec: a rank owns no rows");
        if (m >= kProfileMinSize)
            profile_begin(ctx, PROFILE_MATVEC, 8.0 * nrows * m + 8.0 * m + 8.0 * nrows * (1 + sh.world), 2.0 * nrows * m);
#define SCS_MV_AG(G)                                                                                          \
    matvec_rows_allgather<G><<<ceil_div(nrows, kMvThreads / G), kMvThreads, 0, ctx->stream>>>(                \
        m, rows.row0, nrows, W, isd, z, peers, ticket.vec_offset, ticket.epoch, ticket.timeout_ns, done)
        switch (matvec_group(m)) {
        case 32: SCS_MV_AG(32); break;
        case 64: SCS_MV_AG(64); break;
        case 128: SCS_MV_AG(128); break;
        default: SCS_MV_AG(256); break;
        }
#undef SCS_MV_AG
        if (m >= kProfileMinSize) profile_end(ctx);
        SCS_LAUNCHED(ctx, "matvec_rows_allgather");
        return SCS_OK;
    }
    return launch_matvec_single(ctx, m, W, isd, z, *y, done);
}

int launch_matvec_single(scs_ctx *ctx, int m, const double *W, const double *isd, const double *z, double *y,
                         const int32_t *done) {
    // algorithmic bytes: W once + read z, isd, write y
    if (m >= kProfileMinSize) profile_begin(ctx, PROFILE_MATVEC, 8.0 * m * m + 24.0 * m, 2.0 * m * m);
#define SCS_MV(G) matvec_rows<G><<<ceil_div(m, kMvThreads / G), kMvThreads, 0, ctx->stream>>>(m, W, isd, z, y, done)
    switch (matvec_group(m)) {
    case 32: SCS_MV(32); break;
    case 64: SCS_MV(64); break;
    case 128: SCS_MV(128); break;
    default: SCS_MV(256); break;
    }
#undef SCS_MV
    if (m >= kProfileMinSize) profile_end(ctx);
    SCS_LAUNCHED(ctx, "matvec_rows");
    return SCS_OK;
}

// ---- Lanczos vector kernels ------------------------------------------------------------------
// start vector: uniform(-1, 1) from a counter-based generator (sklearn: random_state.uniform(-1, 1, n),
// _arpack.py:31-33)
__global__ void random_start(int m, uint64_t seed, double *__restrict__ w) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const uint64_t r = splitmix64(seed * 0xD1342543DE82EF95ull + static_cast<uint64_t>(i) + 1ull);
    w[i] = 2.0 * (static_cast<double>(r >> 11) * (1.0 / 9007199254740992.0)) - 1.0;
}

// h[k] = basis[k] . w   for k in [0, nb); one CTA per basis vector, fixed summation order
__global__ void __launch_bounds__(kVecThreads)
multi_dot(int m, const double *__restrict__ basis, const double *__restrict__ w, double *__restrict__ h) {
    __shared__ double scratch[33];
    const double *b = basis + static_cast<size_t>(blockIdx.x) * m;
    double acc = 0.0;
    for (int i = threadIdx.x; i < m; i += kVecThreads) acc = fma(b[i], w[i], acc);
    acc = block_sum(acc, scratch);
    if (threadIdx.x == 0) h[blockIdx.x] = acc;
}

// w -= sum_k h[k] basis[k]
__global__ void __launch_bounds__(kVecThreads)
multi_axpy(int m, int nb, const double *__restrict__ basis, const double *__restrict__ h, double *__restrict__ w) {
    extern __shared__ double hs[];
    for (int k = threadIdx.x; k < nb; k += kVecThreads) hs[k] = h[k];
    __syncthreads();
    const int i = blockIdx.x * kVecThreads + threadIdx.x;
    if (i >= m) return;
    double acc = w[i];
    for (int k = 0; k < nb; ++k) acc = fma(-hs[k], basis[static_cast<size_t>(k) * m + i], acc);
    w[i] = acc;
}

// beta = |w|; next = w / beta; z = isd .* next; records alpha_j = h1[hj] + h2[hj] (hj = position of
// v_j in the basis) and beta_j.  j == 0 is the normalisation of the start vector (nothing recorded).
// The first j whose beta_j is negligible (the Krylov space is invariant: later vectors are noise)
// is latched in state[0]; the projected problem is then solved on the leading block only.  One CTA.
__global__ void __launch_bounds__(kOneCta)
normalize_step(int m, int j, int hj, const double *__restrict__ w, const double *__restrict__ isd,
               const double *__restrict__ h1, const double *__restrict__ h2, double *__restrict__ alpha,
               double *__restrict__ beta, double *__restrict__ next, double *__restrict__ z,
               int32_t *__restrict__ state) {
    __shared__ double scratch[33];
    double ss = 0.0;
    for (int i = threadIdx.x; i < m; i += blockDim.x) ss = fma(w[i], w[i], ss);
    ss = block_sum(ss, scratch);
    const double b = sqrt(ss);
    if (threadIdx.x == 0) {
        beta[j] = b;
        if (j > 0) {
            alpha[j] = h1[hj] + h2[hj];
            if (b <= kBreakdown && state[0] == 0) state[0] = j;
        } else {
            state[0] = 0;
        }
    }
    const double inv = b > 0.0 ? 1.0 / b : 0.0;
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
        const double v = w[i] * inv;
        next[i] = v;
        z[i] = isd[i] * v;
    }
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
__device__ void tridiag_solve(int jrun, const double *alpha, const double *beta, const int32_t *state, double *coef,
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

__global__ void __launch_bounds__(256)
tridiag_ritz(int jrun, const double *__restrict__ alpha, const double *__restrict__ beta,
             const int32_t *__restrict__ state, double *__restrict__ coef, double *__restrict__ out) {
    extern __shared__ double sm[];
    __shared__ int counts[256];
    __shared__ double pair[4];
    tridiag_solve(jrun, alpha, beta, state, coef, out, sm, counts, pair);
}

// ---- fused Lanczos step tail (m <= kFusedMax) -----------------------------------------------------
// Everything of a Lanczos step except the operator application, in one single-CTA launch: both
// Gram-Schmidt passes against the nb leading basis vectors, the norm, the next basis vector and its
// scaled copy, and -- at check steps -- the projected eigenproblem, whose verdict is latched in
// state[1] so that the launches already queued behind it (matvec and tail both test the flag) fall
// through.  The host therefore synchronises once per chunk of steps instead of once per check.
// state: [0] breakdown step, [1] done, [2] step at which done was raised.
constexpr int kFusedMax = 4096;

__global__ void __launch_bounds__(kOneCta)
lanczos_tail(int m, int j, int nb, int hj, int check, const double *basis, const double *__restrict__ w_in,
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

// y = sum_{k=1..j} coef[k] v_k, normalised (`basis` points at v_0, the row before v_1);
// z = isd .* y.  One CTA.
__global__ void __launch_bounds__(kOneCta)
ritz_vector(int m, int j, const double *__restrict__ basis, const double *__restrict__ coef,
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

// out[0] = | Ny - theta y |, out[1] = y . Ny   (Ny given).  One CTA.
__global__ void __launch_bounds__(kOneCta)
true_residual(int m, const double *__restrict__ y, const double *__restrict__ Ny, const double *__restrict__ ritz,
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

// ---- embedding, sign flip, exact 1-D 2-means --------------------------------------------------
// u = y .* isd is the Fiedler coordinate sklearn clusters (_spectral_embedding.py:378-381); its sign is
// fixed so that the entry of largest magnitude is positive (extmath.py:1283-1286).  The 2-means optimum
// in one dimension is a threshold on the sorted values: sort, prefix sums, best split.
// result: [0] margin, [1] lower centroid, [2] upper centroid, [3] size of the upper part,
//         [4] number of Lloyd-stable splits, [5] best other stable split's score / the optimum's
template <bool kShared>
__global__ void __launch_bounds__(kOneCta)
two_means_1d(int m, int P, const double *__restrict__ y, const double *__restrict__ isd, double *__restrict__ u,
             double *__restrict__ sorted_g, int32_t *__restrict__ side, double *__restrict__ result) {
    extern __shared__ double sort_s[];
    __shared__ double scratch[33];
    __shared__ double best_score[32];
    __shared__ int best_index[32];
    __shared__ double bcast[4];
    __shared__ int ibcast;
    double *keys = kShared ? sort_s : sorted_g;
    const int tid = threadIdx.x, nthr = blockDim.x;

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

    double total = 0.0;
    for (int i = tid; i < P; i += nthr) {
        double v = INFINITY;
        if (i < m) {
            v = sign * (y[i] * isd[i]);
            u[i] = v;
            total += v;
        }
        keys[i] = v;
    }
    total = block_sum(total, scratch);
    const double mean = total / m;
    __syncthreads();

    // bitonic sort, ascending
    for (int k = 2; k <= P; k <<= 1) {
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

__global__ void trivial_pair(int32_t *side) {
    side[0] = 0;
    side[1] = 1;
}

int next_pow2(int x) {
    int p = 1;
    while (p < x) p <<= 1;
    return p;
}

}  // namespace

int normalized_matvec(scs_ctx *ctx, int m, const double *W, const double *isd, const double *x, double *y) {
    if (m <= 0 || !W || !isd || !x || !y) return fail(ctx, SCS_ERR_INVALID, "normalized_matvec: bad argument");
    double *z;
    int rc;
    if ((rc = reserve_as(ctx, SLOT_UVEC, static_cast<size_t>(m), &z))) return rc;
    scale_vector<<<ceil_div(m, 256), 256, 0, ctx->stream>>>(m, isd, x, z);
    SCS_LAUNCHED(ctx, "scale_vector");
    return launch_matvec_single(ctx, m, W, isd, z, y, nullptr);
}

namespace {

struct LanczosOutcome {
    double theta1 = 0.0, theta2 = 0.0, estimate = 0.0;
    int steps = 0;        // size of the projected problem that produced the pair
    bool converged = false;
    bool invariant = false;  // stopped on a breakdown: the Krylov space is an invariant subspace
    int matvecs = 0;
    int restarts = 0;
};

struct LanczosBuffers {
    double *isd, *basis, *w, *z, *coef, *h1, *h2, *alpha, *beta, *ritz;
    int32_t *state;
    double *pin;
};

// Largest eigenpair of N = D^-1/2 W D^-1/2 on the orthogonal complement of the first `ndefl` rows of
// `basis`.  Lanczos vectors go to rows ndefl, ndefl+1, ...; the Ritz vector to `y` (and isd .* y to z).
int lanczos_largest(scs_ctx *ctx, int m, const double *W, const LanczosBuffers &b, int ndefl, uint64_t seed,
                    double *y, LanczosOutcome *out, RowBlock rows) {
    const int dim = m - ndefl;  // dimension of the deflated space
    const int jmax = dim < kMaxBasis ? dim : kMaxBasis;
    const int vec_blocks = ceil_div(m, kVecThreads);
    double *v0 = b.basis + static_cast<size_t>(ndefl - 1) * m;  // v_k lives at v0 + k m
    double *w = b.w;  // the vector being orthogonalised: b.w, or the window buffer a sharded matvec filled
    auto orthogonalise = [&](int nb, double *h) -> int {
        multi_dot<<<nb, kVecThreads, 0, ctx->stream>>>(m, b.basis, w, h);
        SCS_LAUNCHED(ctx, "multi_dot");
        multi_axpy<<<vec_blocks, kVecThreads, nb * sizeof(double), ctx->stream>>>(m, nb, b.basis, h, w);
        SCS_LAUNCHED(ctx, "multi_axpy");
        return SCS_OK;
    };
    int rc;
    int jdone = 0;
    const bool fused = m <= kFusedMax;
    if (fused && !ctx->tail_configured) {
        SCS_CUDA(ctx, cudaFuncSetAttribute(lanczos_tail, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           static_cast<int>((kFusedMax + 6 * (kMaxBasis + 8)) * sizeof(double))));
        ctx->tail_configured = true;
    }
    const size_t tail_smem = (static_cast<size_t>(m) + 6 * (kMaxBasis + 8)) * sizeof(double);
    for (int attempt = 0; attempt <= kMaxRestarts && !out->converged; ++attempt) {
        w = b.w;
        if (attempt == 0) {
            random_start<<<vec_blocks, kVecThreads, 0, ctx->stream>>>(m, seed, b.w);
            SCS_LAUNCHED(ctx, "random_start");
        } else {
            // explicit restart from the current Ritz vector
            SCS_CUDA(ctx, cudaMemcpyAsync(b.w, y, sizeof(double) * m, cudaMemcpyDeviceToDevice, ctx->stream));
            out->restarts = attempt;
        }
        if (fused) {
            SCS_CUDA(ctx, cudaMemsetAsync(b.state, 0, 4 * sizeof(int32_t), ctx->stream));
            lanczos_tail<<<1, kOneCta, tail_smem, ctx->stream>>>(m, 0, ndefl, 0, 0, b.basis, b.w, b.isd, b.alpha, b.beta,
                                                                v0 + m, b.z, b.state, b.coef, b.ritz);
            SCS_LAUNCHED(ctx, "lanczos_tail");
            // steps are queued a chunk at a time; the tail of a check step latches convergence on the
            // device and everything queued behind it falls through
            int j = 1;
            int attempt_matvecs = 0;
            bool done = false;
            while (j <= jmax && !done) {
                const int chunk_end = std::min(jmax, j == 1 ? 16 : j + 3);  // chunks end on check steps
                for (; j <= chunk_end; ++j) {
                    if ((rc = launch_matvec(ctx, m, W, b.isd, b.z, &w, b.state + 1, rows))) return rc;
                    const int check = j == jmax || (j % 4) == 0;
                    lanczos_tail<<<1, kOneCta, tail_smem, ctx->stream>>>(
                        m, j, ndefl + j, ndefl - 1 + j, check, b.basis, w, b.isd, b.alpha, b.beta,
                        v0 + static_cast<size_t>(j + 1) * m, b.z, b.state, b.coef, b.ritz);
                    SCS_LAUNCHED(ctx, "lanczos_tail");
                }
                SCS_CUDA(ctx, cudaMemcpyAsync(b.pin, b.ritz, 6 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
                SCS_CUDA(ctx, cudaMemcpyAsync(b.pin + 6, b.state, 4 * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
                SCS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
                const int32_t *st = reinterpret_cast<const int32_t *>(b.pin + 6);
                out->theta1 = b.pin[0];
                out->theta2 = b.pin[1];
                out->estimate = b.pin[2];
                out->steps = static_cast<int>(b.pin[4]);
                jdone = static_cast<int>(b.pin[5]);  // the step of the last check that ran
                if (st[1]) {
                    done = true;
                    out->converged = true;
                    out->invariant = b.pin[3] <= kBreakdown;
                    attempt_matvecs = st[2];  // steps actually applied before the flag went up
                } else {
                    attempt_matvecs = chunk_end;
                }
            }
            out->matvecs += attempt_matvecs;
        } else {
            if ((rc = orthogonalise(ndefl, b.h1))) return rc;
            if ((rc = orthogonalise(ndefl, b.h2))) return rc;
            normalize_step<<<1, kOneCta, 0, ctx->stream>>>(m, 0, 0, w, b.isd, b.h1, b.h2, b.alpha, b.beta, v0 + m, b.z,
                                                           b.state);
            SCS_LAUNCHED(ctx, "normalize_step");
            for (int j = 1; j <= jmax; ++j) {
                if ((rc = launch_matvec(ctx, m, W, b.isd, b.z, &w, nullptr, rows))) return rc;
                out->matvecs += 1;
                if ((rc = orthogonalise(ndefl + j, b.h1))) return rc;
                if ((rc = orthogonalise(ndefl + j, b.h2))) return rc;
                normalize_step<<<1, kOneCta, 0, ctx->stream>>>(m, j, ndefl - 1 + j, w, b.isd, b.h1, b.h2, b.alpha,
                                                               b.beta, v0 + static_cast<size_t>(j + 1) * m, b.z, b.state);
                SCS_LAUNCHED(ctx, "normalize_step");
                jdone = j;
                const bool check = j == jmax || (j % 4) == 0;
                if (!check) continue;
                tridiag_ritz<<<1, 256, 5 * (j + 2) * sizeof(double), ctx->stream>>>(j, b.alpha, b.beta, b.state, b.coef,
                                                                                   b.ritz);
                SCS_LAUNCHED(ctx, "tridiag_ritz");
                SCS_CUDA(ctx, cudaMemcpyAsync(b.pin, b.ritz, 5 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
                SCS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
                out->theta1 = b.pin[0];
                out->theta2 = b.pin[1];
                out->estimate = b.pin[2];
                out->steps = static_cast<int>(b.pin[4]);
                if (b.pin[3] <= kBreakdown) {
                    out->converged = true;
                    out->invariant = true;
                    break;
                }
                if (out->estimate <= kResidualTol) {
                    out->converged = true;
                    break;
                }
            }
        }
        // Ritz vector of the current basis (also the restart vector); coefficients past a breakdown are 0
        ritz_vector<<<1, kOneCta, (jdone + 1) * sizeof(double), ctx->stream>>>(m, jdone, v0, b.coef, b.isd, y, b.z);
        SCS_LAUNCHED(ctx, "ritz_vector");
    }
    return SCS_OK;
}

}  // namespace

int spectral_bipartition(scs_ctx *ctx, int m, const double *W, const double *degree, uint64_t seed, int32_t *side,
                         scs_node_stats *stats, RowBlock rows) {
    if (rows.sharded() && !degree) return fail(ctx, SCS_ERR_INVALID, "sharded spectral step needs the degrees");
    if (m < 2 || !W || !side || !stats) return fail(ctx, SCS_ERR_INVALID, "spectral_bipartition: bad argument");
    int rc;
    if (m == 2) {
        // sklearn falls back to a dense eigh (arpack.py:1691-1707); two vertices always separate
        trivial_pair<<<1, 1, 0, ctx->stream>>>(side);
        SCS_LAUNCHED(ctx, "trivial_pair");
        stats->solver = 1;
        stats->eig[1] = 2.0;  // L = [[1,-1],[-1,1]] whatever the weight
        stats->residual = 0.0;
        stats->margin = 0.5;
        return SCS_OK;
    }
    const int jcap = m - 1 < kMaxBasis ? m - 1 : kMaxBasis;
    LanczosBuffers b;
    double *tri, *embed, *sorted, *deg_own, *work;
    int32_t *flags;
    if ((rc = reserve_as(ctx, SLOT_ISD, static_cast<size_t>(m), &b.isd))) return rc;
    if ((rc = reserve_as(ctx, SLOT_BASIS, static_cast<size_t>(jcap + 3) * m, &b.basis))) return rc;
    if ((rc = reserve_as(ctx, SLOT_WORK, 3 * static_cast<size_t>(m), &work))) return rc;
    if ((rc = reserve_as(ctx, SLOT_UVEC, static_cast<size_t>(m), &b.z))) return rc;
    if ((rc = reserve_as(ctx, SLOT_COEF, 3 * static_cast<size_t>(kMaxBasis + 8), &b.coef))) return rc;
    if ((rc = reserve_as(ctx, SLOT_TRIDIAG, 2 * static_cast<size_t>(kMaxBasis + 8), &tri))) return rc;
    if ((rc = reserve_as(ctx, SLOT_RITZ, 32, &b.ritz))) return rc;
    if ((rc = reserve_as(ctx, SLOT_EMBED, static_cast<size_t>(m), &embed))) return rc;
    const int P = next_pow2(m);
    if ((rc = reserve_as(ctx, SLOT_SORTED, static_cast<size_t>(P), &sorted))) return rc;
    if ((rc = reserve_as(ctx, SLOT_SPEC_SCALARS, 16, &flags))) return rc;
    b.h1 = b.coef + (kMaxBasis + 8);
    b.h2 = b.h1 + (kMaxBasis + 8);
    b.alpha = tri;
    b.beta = tri + (kMaxBasis + 8);
    b.w = work;
    b.state = flags + 4;
    double *yvec = work + m, *y2 = work + 2 * static_cast<size_t>(m);
    void *pin_v;
    if ((rc = reserve_pinned(ctx, 256, &pin_v))) return rc;
    b.pin = static_cast<double *>(pin_v);
    double *pin = b.pin;

    SCS_CUDA(ctx, cudaMemsetAsync(flags, 0, 16 * sizeof(int32_t), ctx->stream));
    if (!degree) {
        if ((rc = reserve_as(ctx, SLOT_DEGREE_C, static_cast<size_t>(m), &deg_own))) return rc;
        row_sums<<<ceil_div(static_cast<int64_t>(m) * 32, 256), 256, 0, ctx->stream>>>(m, W, deg_own);
        SCS_LAUNCHED(ctx, "row_sums");
        degree = deg_own;
    }
    prepare_scaling<<<1, kOneCta, 0, ctx->stream>>>(m, degree, b.isd, b.basis, flags);
    SCS_LAUNCHED(ctx, "prepare_scaling");

    stats->solver = 3;
    LanczosOutcome first;
    const auto t_lanczos = std::chrono::steady_clock::now();
    if ((rc = lanczos_largest(ctx, m, W, b, 1, seed, yvec, &first, rows))) return rc;
    ctx->stage_seconds[5] += std::chrono::duration<double>(std::chrono::steady_clock::now() - t_lanczos).count();
    stats->matvecs = first.matvecs;
    stats->restarts = first.restarts;

    // true residual of the accepted pair: one more operator application (z = isd .* y is current)
    double *Ny = b.w;
    if ((rc = launch_matvec(ctx, m, W, b.isd, b.z, &Ny, nullptr, rows))) return rc;
    stats->matvecs += 1;
    true_residual<<<1, kOneCta, 0, ctx->stream>>>(m, yvec, Ny, b.ritz, b.ritz + 8);
    SCS_LAUNCHED(ctx, "true_residual");
    SCS_CUDA(ctx, cudaMemcpyAsync(pin + 8, b.ritz + 8, 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));

    // The second Ritz value of the same Krylov space only bounds the next eigenvalue from below.  If the
    // space closed early (fewer than m - 1 steps), an eigenvalue of N is repeated -- possibly the Fiedler
    // one.  Settle it with a second run deflated by the Fiedler vector as well.
    double theta_next = first.theta2;
    if (first.invariant && first.steps < m - 1 && m >= 3) {
        SCS_CUDA(ctx, cudaMemcpyAsync(b.basis + m, yvec, sizeof(double) * m, cudaMemcpyDeviceToDevice, ctx->stream));
        LanczosOutcome second;
        b.ritz += 16;  // keep the first pair's scalars
        rc = lanczos_largest(ctx, m, W, b, 2, seed + 0x5bd1e995u, y2, &second, rows);
        b.ritz -= 16;
        if (rc) return rc;
        stats->matvecs += second.matvecs;
        theta_next = second.theta1;
    }

    if (P <= 8192) {
        auto kernel = two_means_1d<true>;
        const size_t smem = static_cast<size_t>(P) * sizeof(double);
        if (!ctx->kmeans_configured) {
            SCS_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 8));
            ctx->kmeans_configured = true;
        }
        kernel<<<1, kOneCta, smem, ctx->stream>>>(m, P, yvec, b.isd, embed, sorted, side, b.ritz + 12);
    } else {
        two_means_1d<false><<<1, kOneCta, 0, ctx->stream>>>(m, P, yvec, b.isd, embed, sorted, side, b.ritz + 12);
    }
    SCS_LAUNCHED(ctx, "two_means_1d");
    SCS_CUDA(ctx, cudaMemcpyAsync(pin + 12, b.ritz + 12, 6 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    SCS_CUDA(ctx, cudaMemcpyAsync(pin + 20, flags, 4 * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    SCS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    stats->eig[0] = 0.0;
    stats->eig[1] = 1.0 - first.theta1;
    stats->eig[2] = 1.0 - theta_next;
    stats->residual = pin[8];
    stats->margin = pin[12];
    const int32_t bad_degree = reinterpret_cast<const int32_t *>(pin + 20)[0];
    stats->kmeans_stable_splits = static_cast<int32_t>(pin[16]);
    stats->kmeans_runner_up = pin[17];
    stats->tie_flag = 0;
    if (!std::isnan(theta_next) && (first.theta1 - theta_next) < kGapTie) stats->tie_flag |= 1;
    if (!(stats->margin >= kMarginTie)) stats->tie_flag |= 2;
    if (!first.converged) stats->tie_flag |= 4;  // accepted at the restart limit: see stats->residual
    if (bad_degree) stats->tie_flag |= 8;        // negative / non-finite degree: sklearn's result is NaN-driven
    if (stats->kmeans_stable_splits > 1) stats->tie_flag |= 16;  // k-means has several local optima
    return SCS_OK;
}

}  // namespace scs
