// One-CTA path for small recursion nodes (n <= 64 vertices).
//
// Three quarters of the recursion nodes of a supertree job have fewer than 32 taxa; on them the
// staged path (components, contraction, Lanczos, 2-means: dozens of launches and several host
// round trips) is pure latency.  Here everything after the graph build -- the reference's
// _get_graph_components (/root/reference/src/sc_supertree/scs.py:458-492),
// _contract_proper_cluster_graph (:261-387) and spectral_cluster_graph (:210-258) -- runs in one
// launch of one CTA with the graph in shared memory, and the host reads one small record back.
//
//  * components: transitive closure of the 64-bit adjacency rows by repeated squaring;
//  * contraction: the same closure on the max-graph rows, then a max-merge of the weights;
//  * spectral: N = D^-1/2 W D^-1/2 is formed explicitly and diagonalised by a parallel cyclic
//    Jacobi iteration (round-robin pair ordering).  The known eigenpair (1, sqrt(d)) is moved out
//    of the way with M = N - 2 q0 q0^T, so the Fiedler pair is the largest eigenpair of M and the
//    next eigenvalue is exact as well (exact repeated-eigenvalue detection);
//  * sign flip and exact 1-D 2-means as in spectral.cu.

#include "common.cuh"

namespace scs {

namespace {

constexpr int kN = kSmallNode;  // 64
constexpr int kThreads = 256;
constexpr double kGapTie = 1e-7;
constexpr double kMarginTie = 1e-9;

struct SmallShared {
    double A[kN][kN + 1];   // W, then Wc, then M (padded: column sweeps hit distinct banks)
    double V[kN][kN + 1];   // Jacobi eigenvectors
    double B[kN][kN + 1];   // the (contracted) weights the split was computed from: residual check
    unsigned short cnt[kN][kN];  // co-occurrence counts (batched path: graph built in shared memory)
    int occ[kN];
    int tour_taxon[kN], tour_depth[kN];
    double tour_val[kN];
    double cs[kN / 2][2];
    double deg[kN], isd[kN], q0[kN], u[kN], sorted[kN];
    unsigned long long adj[kN], mx[kN], reach[kN];
    int label[kN], group[kN], member_of[kN], side[kN];
    double red[kThreads / 32 + 1];
    int ired[4];
};

__device__ __forceinline__ double block_sum_small(double v, double *red) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < kThreads / 32; ++w) s += red[w];
        red[kThreads / 32] = s;
    }
    __syncthreads();
    return red[kThreads / 32];
}

// reach[v] = bit set of the vertices connected to v (v included); label[v] = smallest of them
__device__ void closure(int n, const unsigned long long *rows, unsigned long long *reach, int *label) {
    const int v = threadIdx.x;
    if (v < n) reach[v] = rows[v] | (1ull << v);
    __syncthreads();
    for (int round = 0; round < 6; ++round) {  // path lengths double every round: 2^6 >= 64
        unsigned long long acc = 0;
        if (v < n) {
            unsigned long long bits = reach[v];
            acc = bits;
            while (bits) {
                const int b = __ffsll(static_cast<long long>(bits)) - 1;
                bits &= bits - 1;
                acc |= reach[b];
            }
        }
        __syncthreads();
        if (v < n) reach[v] = acc;
        __syncthreads();
    }
    if (v < n) label[v] = __ffsll(static_cast<long long>(reach[v])) - 1;
    __syncthreads();
}

// Everything after the graph build for one node whose weights are in S.A (n x n) and whose adjacency /
// max-graph rows are in S.adj / S.mx.  part, out, group_out, Wc_out are global pointers (the last two
// may be null).
__device__ void small_finish(SmallShared &S, int n, int contract_edges, int32_t *__restrict__ part,
                             scs_node_stats *__restrict__ out, int32_t *__restrict__ group_out,
                             double *__restrict__ Wc_out) {
    const int tid = threadIdx.x;
    const double nan_v = __longlong_as_double(0x7ff8000000000000ll);

    // ---- components (scs.py:458-492) -----------------------------------------------------------
    closure(n, S.adj, S.reach, S.label);
    if (tid == 0) {
        int count = 0;
        for (int v = 0; v < n; ++v) {
            S.member_of[v] = count;  // rank of v among the representatives, valid where label[v] == v
            count += S.label[v] == v;
        }
        S.ired[0] = count;
    }
    __syncthreads();
    const int ncomp = S.ired[0];
    if (tid == 0) {
        out->n_components = ncomp;
        out->contracted_size = n;
        out->spectral_ran = 0;
        out->solver = 0;
        out->matvecs = 0;
        out->restarts = 0;
        out->tie_flag = 0;
        out->kmeans_stable_splits = 0;
        out->eig[0] = 0.0;
        out->eig[1] = nan_v;
        out->eig[2] = nan_v;
        out->residual = nan_v;
        out->margin = nan_v;
        out->kmeans_runner_up = 0.0;
    }
    if (ncomp != 1) {
        if (tid < n) part[tid] = S.member_of[S.label[tid]];
        return;
    }

    // ---- contraction (scs.py:261-387) ----------------------------------------------------------
    int m = n;
    if (contract_edges) {
        closure(n, S.mx, S.reach, S.label);
        if (tid == 0) {
            int count = 0;
            for (int v = 0; v < n; ++v) {
                S.member_of[v] = count;
                count += S.label[v] == v;
            }
            S.ired[1] = count;
        }
        __syncthreads();
        m = S.ired[1];
        if (tid < n) S.group[tid] = S.member_of[S.label[tid]];
        __syncthreads();
        if (m != n) {
            // Wc[A][B] = max over existing edges (u in A, v in B) of W[u][v]; staged in V as ordered keys
            for (int e = tid; e < m * m; e += kThreads) S.V[e / m][e % m] = __longlong_as_double(0ll);
            __syncthreads();
            for (int e = tid; e < n * n; e += kThreads) {
                const int u = e / n, v = e % n;
                if (!((S.adj[u] >> v) & 1ull)) continue;
                const int A = S.group[u], B = S.group[v];
                if (A == B) continue;  // edges inside a merged vertex vanish (scs.py:352-354)
                atomicMax(reinterpret_cast<unsigned long long *>(&S.V[A][B]), order_key(S.A[u][v]));
            }
            __syncthreads();
            for (int e = tid; e < m * m; e += kThreads) {
                const unsigned long long k = static_cast<unsigned long long>(__double_as_longlong(S.V[e / m][e % m]));
                S.V[e / m][e % m] = k ? order_value(k) : 0.0;
            }
            __syncthreads();
            for (int e = tid; e < m * m; e += kThreads) {
                S.A[e / m][e % m] = S.V[e / m][e % m];
                if (Wc_out) Wc_out[e] = S.V[e / m][e % m];
            }
            __syncthreads();
        }
        if (group_out && tid < n) group_out[tid] = S.group[tid];
    } else if (tid < n) {
        S.group[tid] = tid;
    }
    __syncthreads();
    if (tid == 0) {
        out->contracted_size = m;
        out->spectral_ran = 1;
    }
    if (m < 2) {
        if (tid == 0) out->solver = -1;  // nothing to split: the host raises SCS_ERR_TOO_SMALL
        return;
    }
    if (m == 2) {
        if (tid < n) part[tid] = S.group[tid];
        if (tid == 0) {
            out->solver = 1;
            out->eig[1] = 2.0;
            out->residual = 0.0;
            out->margin = 0.5;
        }
        return;
    }

    for (int e = tid; e < m * m; e += kThreads) S.B[e / m][e % m] = S.A[e / m][e % m];
    // ---- normalised affinity, trivial pair shifted away ------------------------------------------
    if (tid < m) {
        double d = 0.0;
        for (int j = 0; j < m; ++j) d += S.A[tid][j];
        S.deg[tid] = d;
    }
    __syncthreads();
    if (tid == 0) {
        double vol = 0.0;
        bool bad = false;
        for (int i = 0; i < m; ++i) {
            const double d = S.deg[i];
            if (!(d >= 0.0) || isinf(d)) bad = true;
            vol += d > 0.0 ? d : 0.0;
        }
        const double inv = vol > 0.0 ? 1.0 / sqrt(vol) : 0.0;
        for (int i = 0; i < m; ++i) {
            const double d = S.deg[i];
            const double s = d > 0.0 ? sqrt(d) : 0.0;
            S.isd[i] = d > 0.0 ? 1.0 / s : 1.0;
            S.q0[i] = s * inv;
        }
        S.ired[2] = bad;
    }
    __syncthreads();
    for (int e = tid; e < m * m; e += kThreads) {
        const int i = e / m, j = e % m;
        const double w = i == j ? 0.0 : S.A[i][j];  // scipy zeroes the diagonal (_laplacian.py:536-539)
        S.V[i][j] = S.isd[i] * w * S.isd[j] - 2.0 * S.q0[i] * S.q0[j];
    }
    __syncthreads();
    for (int e = tid; e < m * m; e += kThreads) {
        const int i = e / m, j = e % m;
        S.A[i][j] = 0.5 * (S.V[i][j] + S.V[j][i]);  // exactly symmetric
    }
    __syncthreads();
    for (int e = tid; e < m * m; e += kThreads) S.V[e / m][e % m] = (e / m == e % m) ? 1.0 : 0.0;
    __syncthreads();

    // ---- parallel cyclic Jacobi ------------------------------------------------------------------
    const int me = (m + 1) & ~1;  // even number of players; the extra one is a bye
    const int half = me >> 1;
    if (tid == 0) S.ired[1] = 0;  // sweeps done
    for (int sweep = 0; sweep < 30; ++sweep) {
        double off = 0.0, diag = 0.0;
        for (int e = tid; e < m * m; e += kThreads) {
            const int i = e / m, j = e % m;
            const double x = S.A[i][j];
            if (i == j) diag += x * x; else off += x * x;
        }
        off = block_sum_small(off, S.red);
        diag = block_sum_small(diag, S.red);
        // off-diagonal Frobenius norm below 1e-14 of the whole: eigenvalues are then converged to
        // rounding (their error is quadratic in it) and eigenvectors to ~1e-14 / gap
        if (off <= 1e-28 * (diag + off) || off < 1e-300) break;
        if (tid == 0) S.ired[1] = sweep + 1;
        for (int step = 0; step < me - 1; ++step) {
            if (tid < half) {
                int p, q;
                if (tid == 0) { p = me - 1; q = step; }
                else { p = (step + tid) % (me - 1); q = (step - tid + (me - 1)) % (me - 1); }
                if (p > q) { const int t = p; p = q; q = t; }
                double c = 1.0, s = 0.0;
                if (q < m) {
                    const double apq = S.A[p][q];
                    if (fabs(apq) > 1e-300) {
                        const double tau = (S.A[q][q] - S.A[p][p]) / (2.0 * apq);
                        const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
                        c = 1.0 / sqrt(1.0 + t * t);
                        s = t * c;
                    }
                }
                S.cs[tid][0] = c;
                S.cs[tid][1] = s;
            }
            __syncthreads();
            // columns: A <- A J, V <- V J
            for (int e = tid; e < m * half; e += kThreads) {
                const int r = e / half, k = e % half;
                int p, q;
                if (k == 0) { p = me - 1; q = step; }
                else { p = (step + k) % (me - 1); q = (step - k + (me - 1)) % (me - 1); }
                if (p > q) { const int t = p; p = q; q = t; }
                if (q >= m) continue;
                const double c = S.cs[k][0], s = S.cs[k][1];
                const double arp = S.A[r][p], arq = S.A[r][q];
                S.A[r][p] = c * arp - s * arq;
                S.A[r][q] = s * arp + c * arq;
                const double vrp = S.V[r][p], vrq = S.V[r][q];
                S.V[r][p] = c * vrp - s * vrq;
                S.V[r][q] = s * vrp + c * vrq;
            }
            __syncthreads();
            // rows: A <- J^T A
            for (int e = tid; e < m * half; e += kThreads) {
                const int r = e / half, k = e % half;
                int p, q;
                if (k == 0) { p = me - 1; q = step; }
                else { p = (step + k) % (me - 1); q = (step - k + (me - 1)) % (me - 1); }
                if (p > q) { const int t = p; p = q; q = t; }
                if (q >= m) continue;
                const double c = S.cs[k][0], s = S.cs[k][1];
                const double apr = S.A[p][r], aqr = S.A[q][r];
                S.A[p][r] = c * apr - s * aqr;
                S.A[q][r] = s * apr + c * aqr;
            }
            __syncthreads();
        }
    }

    // ---- Fiedler pair: largest eigenvalue of M (smallest index on exact ties) ------------------------
    if (tid == 0) {
        int best = 0;
        for (int i = 1; i < m; ++i)
            if (S.A[i][i] > S.A[best][best]) best = i;
        double second = -INFINITY;
        for (int i = 0; i < m; ++i)
            if (i != best && S.A[i][i] > second) second = S.A[i][i];
        S.ired[3] = best;
        S.red[0] = S.A[best][best];
        S.red[1] = second;
    }
    __syncthreads();
    const int col = S.ired[3];
    const double theta1 = S.red[0], theta2 = S.red[1];
    __syncthreads();

    // ---- embedding, sign flip, exact 1-D 2-means (see spectral.cu) ----------------------------------
    if (tid < m) S.u[tid] = S.V[tid][col] * S.isd[tid];
    __syncthreads();
    if (tid == 0) {
        int arg = 0;
        for (int i = 1; i < m; ++i)
            if (fabs(S.u[i]) > fabs(S.u[arg])) arg = i;
        const double sign = S.u[arg] < 0.0 ? -1.0 : 1.0;
        double mean = 0.0;
        for (int i = 0; i < m; ++i) {
            S.u[i] *= sign;
            mean += S.u[i];
        }
        S.red[2] = mean / m;
    }
    __syncthreads();
    if (tid < m) {  // rank sort (ties by index)
        const double x = S.u[tid];
        int rank = 0;
        for (int j = 0; j < m; ++j) rank += (S.u[j] < x) || (S.u[j] == x && j < tid);
        S.sorted[rank] = x;
    }
    __syncthreads();
    if (tid == 0) {
        const double mean = S.red[2];
        double all = 0.0;
        for (int i = 0; i < m; ++i) all += S.sorted[i] - mean;
        double prefix = 0.0, bscore = -1.0;
        int bi = 1;
        for (int i = 0; i + 1 < m; ++i) {
            prefix += S.sorted[i] - mean;
            const int cnt = i + 1;
            const double rest = all - prefix;
            const double score = prefix * prefix / cnt + rest * rest / (m - cnt);
            if (score > bscore) { bscore = score; bi = cnt; }
        }
        int stable = 0;
        double runner = -1.0, lower = 0.0;
        prefix = 0.0;
        for (int i = 0; i + 1 < m; ++i) {
            prefix += S.sorted[i] - mean;
            const int cnt = i + 1;
            if (cnt == bi) lower = prefix;
            const double rest = all - prefix;
            const double mid = mean + 0.5 * (prefix / cnt + rest / (m - cnt));
            if (S.sorted[i] < mid && mid < S.sorted[i + 1]) {
                stable += 1;
                if (cnt != bi) runner = fmax(runner, prefix * prefix / cnt + rest * rest / (m - cnt));
            }
        }
        const double c0 = mean + lower / bi;
        const double c1 = mean + (all - lower) / (m - bi);
        const double mid = 0.5 * (c0 + c1);
        const double range = fmax(S.sorted[m - 1] - S.sorted[0], 1e-300);
        const double margin = fmin(fabs(S.sorted[bi - 1] - mid), fabs(S.sorted[bi] - mid)) / range;
        S.red[3] = S.sorted[bi];
        out->solver = 2;
        out->matvecs = S.ired[1];  // Jacobi sweeps (the dense solver applies no matvec)
        out->eig[1] = 1.0 - theta1;
        out->eig[2] = 1.0 - theta2;
        out->margin = margin;
        out->kmeans_stable_splits = stable;
        out->kmeans_runner_up = (runner >= 0.0 && bscore > 0.0) ? runner / bscore : 0.0;
        int flag = 0;
        if (theta1 - theta2 < kGapTie) flag |= 1;
        if (!(margin >= kMarginTie)) flag |= 2;
        if (S.ired[2]) flag |= 8;
        if (stable > 1) flag |= 16;
        out->tie_flag = flag;
    }
    __syncthreads();
    const double threshold = S.red[3];
    if (tid < m) S.side[tid] = S.u[tid] >= threshold ? 1 : 0;
    __syncthreads();
    if (tid < n) part[tid] = S.side[S.group[tid]];

    // residual of the accepted pair against the operator it was computed from: |N y - theta y|
    if (tid < m) {
        double acc = 0.0;
        for (int j = 0; j < m; ++j)
            if (j != tid) acc += S.isd[tid] * S.B[tid][j] * S.isd[j] * S.V[j][col];
        acc -= theta1 * S.V[tid][col];
        S.deg[tid] = acc * acc;
    }
    __syncthreads();
    if (tid == 0) {
        double rr = 0.0;
        for (int i = 0; i < m; ++i) rr += S.deg[i];
        out->residual = sqrt(rr);
    }
}

// ---- per-node entry: graph already built in global memory (csrc/pcg.cu) ------------------------------
__global__ void __launch_bounds__(kThreads)
small_node_kernel(int n, int words, int contract_edges, const double *__restrict__ W,
                  const uint32_t *__restrict__ adj_bits, const uint32_t *__restrict__ max_bits,
                  int32_t *__restrict__ part, scs_node_stats *__restrict__ out,
                  int32_t *__restrict__ group_out, double *__restrict__ Wc_out) {
    extern __shared__ __align__(16) unsigned char raw[];
    SmallShared &S = *reinterpret_cast<SmallShared *>(raw);
    const int tid = threadIdx.x;
    for (int e = tid; e < n * n; e += kThreads) S.A[e / n][e % n] = W[e];
    if (tid < n) {
        unsigned long long a = adj_bits[static_cast<size_t>(tid) * words];
        unsigned long long m = max_bits ? max_bits[static_cast<size_t>(tid) * words] : 0ull;
        if (words > 1) {
            a |= static_cast<unsigned long long>(adj_bits[static_cast<size_t>(tid) * words + 1]) << 32;
            if (max_bits) m |= static_cast<unsigned long long>(max_bits[static_cast<size_t>(tid) * words + 1]) << 32;
        }
        S.adj[tid] = a;
        S.mx[tid] = m;
    }
    __syncthreads();
    small_finish(S, n, contract_edges, part, out, group_out, Wc_out);
}

// Cycle counters of the batched kernel, summed over its CTAs: [0] graph build, [1] everything after it (profiling aid:
// scs_debug_small_cycles).
__device__ unsigned long long g_small_cycles[2];

// ---- batched entry: one CTA per node builds the node's graph from its leaf tours in shared memory
//      (the reference's _proper_cluster_graph_edges, scs.py:495-663) and finishes it ------------------
// Trees are taken in input order with a barrier between them; inside a tree thread i owns the pairs
// (i, j > i) and walks j upwards with a running minimum of the consecutive-leaf LCA depths, so every
// W entry receives its terms in tree order with separately rounded multiply and add -- the same sums,
// bit for bit, as pcg_rows_kernel and the reference.
__global__ void __launch_bounds__(kThreads)
small_batch_kernel(const scs_small_node *__restrict__ nodes, const int64_t *__restrict__ leaf_offsets,
                   const int32_t *__restrict__ leaf_taxon, const int32_t *__restrict__ adj_depth,
                   const double *__restrict__ adj_val, const int32_t *__restrict__ root_depth,
                   const double *__restrict__ tree_weight, int contract_edges, int32_t *__restrict__ part,
                   scs_node_stats *__restrict__ stats, int32_t *__restrict__ bad, int absolute_offsets) {
    extern __shared__ __align__(16) unsigned char raw[];
    SmallShared &S = *reinterpret_cast<SmallShared *>(raw);
    const int tid = threadIdx.x;
    const scs_small_node node = nodes[blockIdx.x];
    const int n = node.n;
    const long long clock_start = clock64();
    if (node.num_trees > kSmallMaxTrees || n < 1 || n > kN) {
        // the co-occurrence counts are 16-bit here (the staged path switches to 32-bit counts at 65 536 trees,
        // pcg.cu `narrow`); the host entry points route such nodes to the staged path -- refuse loudly if one
        // arrives anyway
        if (tid == 0) {
            *bad = 1;
            stats[blockIdx.x].solver = -1;
        }
        return;
    }
    for (int e = tid; e < n * n; e += kThreads) {
        S.A[e / n][e % n] = 0.0;
        S.cnt[e / n][e % n] = 0;
    }
    if (tid < n) S.occ[tid] = 0;
    __syncthreads();
    // T + 1 offsets per node, relative to its first leaf -- or the forest's absolute offsets
    const int64_t *offs = leaf_offsets + node.tree_base + (absolute_offsets ? 0 : blockIdx.x);
    const int64_t leaf_base = absolute_offsets ? 0 : node.leaf_base;
    for (int t = 0; t < node.num_trees; ++t) {
        const int64_t tb = leaf_base + offs[t];
        const int k = static_cast<int>(offs[t + 1] - offs[t]);
        const double w = tree_weight[node.tree_base + t];
        const int rd = root_depth[node.tree_base + t];
        if (k > n) {  // more leaves than vertices: malformed input
            if (tid == 0) *bad = 1;
            continue;
        }
        if (tid < k) {
            const int a = leaf_taxon[tb + tid];
            S.tour_taxon[tid] = a;
            S.tour_depth[tid] = adj_depth[tb + tid];
            S.tour_val[tid] = adj_val[tb + tid];
            if (a < 0 || a >= n) *bad = 1;
        }
        __syncthreads();
        // four threads share a leaf i: each walks the whole staircase (the running minimum needs every j) but adds
        // only every fourth pair; a pair is stored once, at [smaller vertex][larger vertex], and mirrored after the
        // last tree -- every entry still receives its terms in tree order, one rounded multiply and one rounded add
        // each, so the sums are the reference's bit for bit
        const int i = tid & (kN - 1), q = tid >> 6;
        if (i < k) {
            const int a = S.tour_taxon[i];
            if (a >= 0 && a < n) {
                if (q == 0) S.occ[a] += 1;
                int best_depth = 0x7fffffff, best_idx = i;
                for (int j = i + 1; j < k; ++j) {
                    const int d = S.tour_depth[j - 1];
                    if (d < best_depth) { best_depth = d; best_idx = j - 1; }  // leftmost shallowest entry
                    if (best_depth == rd) break;  // the root separates everything further right too
                    if (((j - i - 1) & 3) != q) continue;
                    const int b = S.tour_taxon[j];
                    if (b < 0 || b >= n) continue;
                    const int lo = min(a, b), hi = max(a, b);
                    S.A[lo][hi] = __dadd_rn(S.A[lo][hi], __dmul_rn(S.tour_val[best_idx], w));
                    S.cnt[lo][hi] += 1;
                }
            }
        }
        __syncthreads();
    }
    for (int e = tid; e < n * n; e += kThreads) {  // mirror the upper triangle
        const int r = e / n, c = e % n;
        if (r > c) {
            S.A[r][c] = S.A[c][r];
            S.cnt[r][c] = S.cnt[c][r];
        }
    }
    __syncthreads();
    if (tid < n) {
        unsigned long long a = 0, m = 0;
        const int occ_a = S.occ[tid];
        for (int b = 0; b < n; ++b) {
            const int c = S.cnt[tid][b];
            if (c > 0) {
                a |= 1ull << b;
                if (c == max(occ_a, S.occ[b])) m |= 1ull << b;
            }
        }
        S.adj[tid] = a;
        S.mx[tid] = contract_edges ? m : 0ull;
    }
    __syncthreads();
    const long long clock_built = clock64();
    small_finish(S, n, contract_edges, part + node.vertex_base, stats + blockIdx.x, nullptr, nullptr);
    if (tid == 0) {
        atomicAdd(&g_small_cycles[0], static_cast<unsigned long long>(clock_built - clock_start));
        atomicAdd(&g_small_cycles[1], static_cast<unsigned long long>(clock64() - clock_built));
    }
}

}  // namespace

int small_node(scs_ctx *ctx, int n, int contract_edges, const double *W, const uint32_t *adj_bits,
               const uint32_t *max_bits, int32_t *part, scs_node_stats *out_dev, int32_t *group_out,
               double *Wc_out) {
    if (n < 1 || n > kSmallNode) return fail(ctx, SCS_ERR_INVALID, "small_node: size out of range");
    const size_t smem = sizeof(SmallShared);
    if (!ctx->small_configured) {
        SCS_CUDA(ctx, cudaFuncSetAttribute(small_node_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           static_cast<int>(smem)));
        ctx->small_configured = true;
    }
    small_node_kernel<<<1, kThreads, smem, ctx->stream>>>(n, scs_bit_words(n), contract_edges, W, adj_bits,
                                                         contract_edges ? max_bits : nullptr, part, out_dev, group_out,
                                                         Wc_out);
    SCS_LAUNCHED(ctx, "small_node_kernel");
    return SCS_OK;
}

int small_cycles(scs_ctx *ctx, unsigned long long *out2, int reset) {
    SCS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    SCS_CUDA(ctx, cudaMemcpyFromSymbol(out2, g_small_cycles, 2 * sizeof(unsigned long long)));
    if (reset) {
        const unsigned long long zero[2] = {0, 0};
        SCS_CUDA(ctx, cudaMemcpyToSymbol(g_small_cycles, zero, sizeof(zero)));
    }
    return SCS_OK;
}

int small_batch(scs_ctx *ctx, int num_nodes, const scs_small_node *nodes_dev, const int64_t *leaf_offsets,
                const int32_t *leaf_taxon, const int32_t *adj_depth, const double *adj_val, const int32_t *root_depth,
                const double *tree_weight, int contract_edges, int32_t *part_dev, scs_node_stats *stats_dev,
                int32_t *bad_dev, int absolute_offsets) {
    if (num_nodes <= 0) return SCS_OK;
    const size_t smem = sizeof(SmallShared);
    if (!ctx->batch_configured) {
        SCS_CUDA(ctx, cudaFuncSetAttribute(small_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           static_cast<int>(smem)));
        ctx->batch_configured = true;
    }
    small_batch_kernel<<<num_nodes, kThreads, smem, ctx->stream>>>(nodes_dev, leaf_offsets, leaf_taxon, adj_depth,
                                                                  adj_val, root_depth, tree_weight, contract_edges,
                                                                  part_dev, stats_dev, bad_dev, absolute_offsets);
    SCS_LAUNCHED(ctx, "small_batch_kernel");
    return SCS_OK;
}

}  // namespace scs
