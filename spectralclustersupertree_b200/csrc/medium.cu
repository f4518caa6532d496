// Batched path for medium-sized recursion nodes (kSmallNode < n <= kMediumMax).
//
// On the 10 000-taxon workload 295 recursion nodes have between 65 and 4096 taxa.  Taken one by one
// (context.cu: node_split) each of them is ~73 launches and ~5 host round trips, and the GPU idles between them:
// latency, not throughput, was 90 % of the job.  Here ALL such nodes of a wave of the breadth-first recursion go
// through every stage together -- each stage is ONE launch over the rows / trees / nodes of the whole batch:
//
//   graph build        pcg_build_batch (pcg.cu): the row kernel with one CTA per row of the batch's global row space
//                      (replaces _proper_cluster_graph_edges + _dfs_pcg_weights,
//                      /root/reference/src/sc_supertree/scs.py:495-663, for every node of the batch)
//   components         lock-free union-find over the global rows, adjacency bits and max-graph bits (scs.py:458-492)
//   decisions          per node on the device: disconnected / contract to m / trivial pair / Lanczos -- no round trip
//   contraction        max-merge, one CTA per contracted row of the batch (scs.py:261-387)
//   spectral split     Lanczos in lock-step: per step one matvec launch over all rows of all unfinished nodes and one
//                      tail launch with a CTA per node (both Gram-Schmidt passes, norm, projected eigenproblem,
//                      convergence latched per node on the device); the host only reads a counter of unfinished
//                      nodes every few steps, pipelined behind the next chunk of launches (scs.py:210-258)
//   2-means, labels    one CTA per node
//
// The arithmetic per node is the arithmetic of the per-node path (same device functions: spectral_dev.cuh, the same
// row kernel), so W, components and contraction are bit-identical to it; the Lanczos iterates differ from it only
// in the summation order of the matvec for m >= 2048 (a warp per row here).
// Nodes that need what a lock-step batch cannot give -- an explicit restart after kMaxBasis steps, or the second,
// deflated run that settles a repeated Fiedler eigenvalue -- are flagged and re-run through the per-node path.

#include "common.cuh"
#include "spectral_dev.cuh"
#include "uf.cuh"

#include <algorithm>
#include <vector>

namespace scs {

namespace {

using namespace specdev;

enum MedMode : int32_t { MED_DISCONNECTED = 0, MED_TOO_SMALL = 1, MED_PAIR = 2, MED_LANCZOS = 3 };
enum MedOutcome : int32_t { MED_RUNNING = 0, MED_CONVERGED = 1, MED_INVARIANT = 2, MED_OUT_OF_STEPS = 3 };

struct MedState {
    int32_t ncomp[2];    // components of the adjacency graph / of the max-graph (counted by flatten)
    int32_t giant[2];    // sampled giant component root (global row) of either graph
    int32_t m;           // contracted size
    int32_t mode;        // MedMode
    int32_t jmax;        // Lanczos steps this node may take: min(m - 1, jcap)
    int32_t outcome;     // MedOutcome
    int32_t lanczos[4];  // [0] breakdown step, [1] done, [2] step at which done was raised
    int32_t bad_degree;  // a degree is negative or not finite
    int32_t pad[3];
};

constexpr int kScalarStride = 3 * (kMaxBasis + 8) + 32;  // doubles per node: alpha | beta | coef | ritz[32]

struct MedBuffers {
    const MedNode *nodes;
    MedState *state;
    const int32_t *row_node;
    double *W, *Wc;
    uint32_t *adj_bits, *max_bits;
    int32_t *occ;
    double *degree, *degree_c;
    int32_t *uf_parent, *label[2];
    int32_t *group, *gptr, *members;
    double *isd, *z, *y, *yvec, *wstart, *embed;
    double *basis, *scalars;
    int32_t *side;
    int32_t *remaining;  // Lanczos nodes still running
};

__device__ __forceinline__ double *node_alpha(const MedBuffers &mb, int b) { return mb.scalars + static_cast<size_t>(b) * kScalarStride; }
__device__ __forceinline__ double *node_beta(const MedBuffers &mb, int b) { return node_alpha(mb, b) + (kMaxBasis + 8); }
__device__ __forceinline__ double *node_coef(const MedBuffers &mb, int b) { return node_alpha(mb, b) + 2 * (kMaxBasis + 8); }
__device__ __forceinline__ double *node_ritz(const MedBuffers &mb, int b) { return node_alpha(mb, b) + 3 * (kMaxBasis + 8); }

// ---- maps ----------------------------------------------------------------------------------------------------
__global__ void med_fill_maps(int B, int R, int T, const MedNode *__restrict__ nodes, int32_t *__restrict__ row_node,
                              int32_t *__restrict__ tree_node) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < R) {
        int lo = 0, hi = B;  // last node with row_base <= i
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (nodes[mid].row_base <= i) lo = mid; else hi = mid;
        }
        row_node[i] = lo;
    }
    if (i < T) {
        int lo = 0, hi = B;  // last node with tree_begin <= i (of a run of equal tree_begin values the last owns tree i)
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (nodes[mid].tree_begin <= i) lo = mid; else hi = mid;
        }
        // trees between the nodes' ranges belong to sub-problems that are not part of this batch
        tree_node[i] = (i >= nodes[lo].tree_begin && i < nodes[lo].tree_end) ? lo : -1;
    }
}

__global__ void med_clear_state(int B, MedState *state, int32_t *remaining) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b == 0) *remaining = 0;
    if (b >= B) return;
    MedState s;
    memset(&s, 0, sizeof(s));
    state[b] = s;
}

// ---- components (scs.py:458-492) over the global row space ---------------------------------------------------
__global__ void med_uf_init(int R, int32_t *parent) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < R) parent[i] = i;
}

constexpr int kSample = 2;
// one warp per row: hook the vertex to its first kSample neighbours
__global__ void med_uf_hook_sample(int R, const MedNode *__restrict__ nodes, const int32_t *__restrict__ row_node,
                                   const uint32_t *__restrict__ bits, int32_t *parent) {
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= R) return;
    const MedNode &nd = nodes[row_node[r]];
    const int base = nd.row_base, n = nd.n, words = nd.words;
    const int a = r - base;
    const uint32_t *row = bits + nd.bit_off + static_cast<size_t>(a) * words;
    int found = 0;
    for (int j0 = 0; j0 < words && found < kSample; j0 += 32) {
        const int j = j0 + lane;
        uint32_t wbits = j < words ? row[j] : 0u;
        if (j == (a >> 5)) wbits &= ~(1u << (a & 31));
        unsigned active = __ballot_sync(0xffffffffu, wbits != 0u);
        while (active && found < kSample) {
            const int src = __ffs(active) - 1;
            uint32_t wsrc = __shfl_sync(0xffffffffu, wbits, src);
            while (wsrc && found < kSample) {
                const int b = ((j0 + src) << 5) + __ffs(wsrc) - 1;
                wsrc &= wsrc - 1;
                if (b < n) {
                    if (lane == 0) uf_union(parent, r, base + b);
                    ++found;
                }
            }
            active &= active - 1;
        }
    }
}

__global__ void med_uf_snapshot(int R, int32_t *parent, int32_t *snapshot) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R) return;
    int r = i;
    while (parent[r] != r) r = parent[r];
    snapshot[i] = r;
}

// the most frequent root among up to 1024 evenly spaced vertices of the node; one CTA of 1024 threads per node
__global__ void __launch_bounds__(1024)
med_uf_pick_giant(const MedNode *__restrict__ nodes, const int32_t *__restrict__ snapshot, MedState *state, int which) {
    __shared__ int32_t roots[1024];
    __shared__ int best_count[32];
    __shared__ int32_t best_root[32];
    const MedNode &nd = nodes[blockIdx.x];
    const int n = nd.n, tid = threadIdx.x;
    const int samples = n < 1024 ? n : 1024;
    const int32_t mine = tid < samples ? snapshot[nd.row_base + static_cast<int64_t>(tid) * n / samples] : -1;
    roots[tid] = mine;
    __syncthreads();
    int count = 0;
    if (mine >= 0)
        for (int i = 0; i < samples; ++i) count += roots[i] == mine;
    int32_t root = mine;
    for (int off = 16; off > 0; off >>= 1) {
        const int oc = __shfl_down_sync(0xffffffffu, count, off);
        const int32_t orr = __shfl_down_sync(0xffffffffu, root, off);
        if (oc > count || (oc == count && orr >= 0 && (root < 0 || orr < root))) { count = oc; root = orr; }
    }
    if ((tid & 31) == 0) { best_count[tid >> 5] = count; best_root[tid >> 5] = root; }
    __syncthreads();
    if (tid == 0) {
        int bc = best_count[0];
        int32_t br = best_root[0];
        for (int w = 1; w < 32; ++w)
            if (best_count[w] > bc || (best_count[w] == bc && best_root[w] >= 0 && (br < 0 || best_root[w] < br))) {
                bc = best_count[w];
                br = best_root[w];
            }
        state[blockIdx.x].giant[which] = br;
    }
}

// one warp per row outside its node's giant component; every neighbour is hooked
__global__ void med_uf_hook_rest(int R, const MedNode *__restrict__ nodes, const int32_t *__restrict__ row_node,
                                 const uint32_t *__restrict__ bits, const int32_t *__restrict__ snapshot,
                                 const MedState *__restrict__ state, int which, int32_t *parent) {
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= R) return;
    const int b_node = row_node[r];
    if (snapshot[r] == state[b_node].giant[which]) return;
    const MedNode &nd = nodes[b_node];
    const int base = nd.row_base, n = nd.n, words = nd.words;
    const int a = r - base;
    const uint32_t *row = bits + nd.bit_off + static_cast<size_t>(a) * words;
    for (int j = lane; j < words; j += 32) {
        uint32_t wbits = row[j];
        const int first = j << 5;
        if (j == (a >> 5)) wbits &= ~(1u << (a & 31));
        while (wbits) {
            const int b = first + __ffs(wbits) - 1;
            wbits &= wbits - 1;
            if (b < n) uf_union(parent, r, base + b);
        }
    }
}

__global__ void med_uf_flatten(int R, const int32_t *__restrict__ row_node, int32_t *parent, int32_t *label,
                               MedState *state, int which) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= R) return;
    int r = i;
    while (parent[r] != r) r = parent[r];
    label[i] = r;
    if (r == i) atomicAdd(&state[row_node[i]].ncomp[which], 1);
}

// ---- per-node decisions (scs.py:124-134) -----------------------------------------------------------------------
__global__ void med_decide(int B, int contract_edges, const MedNode *__restrict__ nodes, MedState *state,
                           int32_t *remaining) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    MedState &s = state[b];
    const int n = nodes[b].n;
    const int m = contract_edges ? s.ncomp[1] : n;
    s.m = m;
    s.outcome = MED_RUNNING;
    s.lanczos[0] = s.lanczos[1] = s.lanczos[2] = s.lanczos[3] = 0;
    if (s.ncomp[0] != 1) s.mode = MED_DISCONNECTED;
    else if (m < 2) s.mode = MED_TOO_SMALL;
    else if (m == 2) s.mode = MED_PAIR;
    else {
        s.mode = MED_LANCZOS;
        s.jmax = min(m - 1, nodes[b].jcap);
        atomicAdd(remaining, 1);
    }
}

// ---- ranks of representatives: component numbers of a disconnected node, contraction groups of a connected one ----
// One CTA of 1024 threads per node, thread t owns vertices [4t, 4t + 4).  rank_s: n ints of shared memory.
__device__ void rank_representatives(int n, int base, const int32_t *__restrict__ label, int32_t *rank_s, int32_t *warp_part,
                                     int *total_out) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int32_t v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int i = 4 * tid + k;
        v[k] = i < n ? (label[base + i] == base + i) : 0;
    }
    const int32_t mine = v[0] + v[1] + v[2] + v[3];
    int32_t inc = mine;
    for (int off = 1; off < 32; off <<= 1) {
        const int32_t o = __shfl_up_sync(0xffffffffu, inc, off);
        if (lane >= off) inc += o;
    }
    __syncthreads();  // warp_part may still be read from a previous call
    if (lane == 31) warp_part[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int32_t w = warp_part[lane];
        for (int off = 1; off < 32; off <<= 1) {
            const int32_t o = __shfl_up_sync(0xffffffffu, w, off);
            if (lane >= off) w += o;
        }
        warp_part[lane] = w;
    }
    __syncthreads();
    int32_t run = (warp ? warp_part[warp - 1] : 0) + inc - mine;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int i = 4 * tid + k;
        if (i < n) rank_s[i] = run;
        run += v[k];
    }
    if (tid == 0) *total_out = warp_part[31];
    __syncthreads();
}

__global__ void __launch_bounds__(1024)
med_ranks(int contract_edges, MedBuffers mb, int32_t *__restrict__ part) {
    extern __shared__ int32_t rank_s[];  // [max_n] ranks, then [max_n] group sizes / cursors
    __shared__ int32_t warp_part[32];
    __shared__ int total;
    const int b = blockIdx.x;
    const MedNode &nd = mb.nodes[b];
    const MedState &st = mb.state[b];
    const int n = nd.n, base = nd.row_base, tid = threadIdx.x;
    if (st.mode == MED_DISCONNECTED) {
        // number the components by smallest member (scs.py:139 iterates them in arbitrary order)
        rank_representatives(n, base, mb.label[0], rank_s, warp_part, &total);
        for (int v = tid; v < n; v += 1024) part[nd.part_off + v] = rank_s[mb.label[0][base + v] - base];
        return;
    }
    if (!contract_edges || st.m == n) {
        for (int v = tid; v < n; v += 1024) mb.group[base + v] = v;
        return;
    }
    rank_representatives(n, base, mb.label[1], rank_s, warp_part, &total);
    int32_t *count_s = rank_s + n;
    for (int v = tid; v < n; v += 1024) count_s[v] = 0;
    __syncthreads();
    for (int v = tid; v < n; v += 1024) {
        const int g = rank_s[mb.label[1][base + v] - base];
        mb.group[base + v] = g;
        atomicAdd(&count_s[g], 1);
    }
    __syncthreads();
    // offsets of the groups' member lists: exclusive scan of the m sizes by one warp (m <= n <= kMediumMax)
    const int m = st.m;
    int32_t *gptr = mb.gptr + base + b;  // m + 1 entries per node
    if (tid < 32) {
        int32_t carry = 0;
        for (int g0 = 0; g0 < m; g0 += 32) {
            const int g = g0 + tid;
            const int32_t c = g < m ? count_s[g] : 0;
            int32_t inc = c;
            for (int off = 1; off < 32; off <<= 1) {
                const int32_t o = __shfl_up_sync(0xffffffffu, inc, off);
                if (tid >= off) inc += o;
            }
            if (g < m) gptr[g] = carry + inc - c;
            carry += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (tid == 0) gptr[m] = carry;
    }
    __syncthreads();
    for (int v = tid; v < n; v += 1024) count_s[v] = 0;  // now cursors
    __syncthreads();
    for (int v = tid; v < n; v += 1024) {
        const int g = mb.group[base + v];
        mb.members[base + gptr[g] + atomicAdd(&count_s[g], 1)] = v;  // order inside a group is irrelevant: max-merge
    }
}

// ---- contraction (scs.py:261-387): one CTA per contracted row of the batch ---------------------------------------
__global__ void __launch_bounds__(256)
med_contract_rows(MedBuffers mb) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long *best = reinterpret_cast<unsigned long long *>(smem_raw);
    __shared__ double warp_part[8];
    const int r = blockIdx.x;
    const int b = mb.row_node[r];
    const MedState &st = mb.state[b];
    const MedNode &nd = mb.nodes[b];
    const int n = nd.n, m = st.m, base = nd.row_base;
    const int A = r - base;
    if (st.mode != MED_LANCZOS || m == n || A >= m) return;
    const int tid = threadIdx.x;
    for (int c = tid; c < m; c += 256) best[c] = 0ull;
    __syncthreads();
    const int32_t *gptr = mb.gptr + base + b;
    const int32_t *group = mb.group + base;
    const double *W = mb.W + nd.w_off;
    const uint32_t *bits = mb.adj_bits + nd.bit_off;
    for (int mi = gptr[A]; mi < gptr[A + 1]; ++mi) {
        const int u = mb.members[base + mi];
        const uint32_t *bits_u = bits + static_cast<size_t>(u) * nd.words;
        const double *W_u = W + static_cast<size_t>(u) * n;
        for (int v = tid; v < n; v += 256) {
            if ((bits_u[v >> 5] >> (v & 31)) & 1u) {
                const int Bv = group[v];
                if (Bv != A) atomicMax(&best[Bv], order_key(W_u[v]));  // edges inside a merged vertex vanish
            }
        }
    }
    __syncthreads();
    double partial = 0.0;
    double *out = mb.Wc + nd.w_off + static_cast<size_t>(A) * m;
    for (int c = tid; c < m; c += 256) {
        const unsigned long long k = best[c];
        const double x = k ? order_value(k) : 0.0;
        out[c] = x;
        partial += x;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) partial += __shfl_down_sync(0xffffffffu, partial, off);
    if ((tid & 31) == 0) warp_part[tid >> 5] = partial;
    __syncthreads();
    if (tid == 0) {
        double s = 0.0;
        for (int wi = 0; wi < 8; ++wi) s += warp_part[wi];
        mb.degree_c[base + A] = s;
    }
}

// ---- Lanczos in lock-step ------------------------------------------------------------------------------------------
__device__ __forceinline__ const double *node_matrix(const MedBuffers &mb, const MedNode &nd, const MedState &st) {
    return (st.m == nd.n ? mb.W : mb.Wc) + nd.w_off;
}

// scaling, start vector and step 0 (the start vector orthogonalised against q0 and normalised); one CTA per node
__global__ void __launch_bounds__(kOneCta)
med_prepare(MedBuffers mb) {
    const int b = blockIdx.x;
    const MedNode &nd = mb.nodes[b];
    MedState &st = mb.state[b];
    if (st.mode != MED_LANCZOS) return;
    const int m = st.m, base = nd.row_base;
    const double *degree = (m == nd.n ? mb.degree : mb.degree_c) + base;
    double *basis = mb.basis + nd.basis_off;
    prepare_scaling_body(m, degree, mb.isd + base, basis, &st.bad_degree);
    double *w = mb.wstart + base;
    for (int i = threadIdx.x; i < m; i += kOneCta) {
        const uint64_t r = splitmix64(nd.seed * 0xD1342543DE82EF95ull + static_cast<uint64_t>(i) + 1ull);
        w[i] = 2.0 * (static_cast<double>(r >> 11) * (1.0 / 9007199254740992.0)) - 1.0;
    }
    __syncthreads();
    lanczos_tail_body(m, 0, 1, 0, 0, basis, w, mb.isd + base, node_alpha(mb, b), node_beta(mb, b), basis + m, mb.z + base,
                      st.lanczos, node_coef(mb, b), node_ritz(mb, b));
}

// y = isd .* (W z) for every row of every node that is still iterating (final = 1: of every Lanczos node); a warp per row
__global__ void __launch_bounds__(kMvThreads)
med_matvec(int R, int final, MedBuffers mb) {
    const int r = (blockIdx.x * kMvThreads + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= R) return;
    const int b = mb.row_node[r];
    const MedState &st = mb.state[b];
    if (st.mode != MED_LANCZOS || (!final && st.lanczos[1])) return;
    const MedNode &nd = mb.nodes[b];
    const int m = st.m, base = nd.row_base;
    const int A = r - base;
    if (A >= m) return;
    const double *Wm = node_matrix(mb, nd, st);
    double acc = row_dot_partial(Wm + static_cast<size_t>(A) * m, mb.z + base, m, lane, 32);
    acc = warp_sum(acc);
    if (lane == 0) mb.y[r] = mb.isd[r] * acc;
}

// step j of every node that is still iterating: both Gram-Schmidt passes, norm, next vector, and at check steps the
// projected eigenproblem; a node that converges (or runs out of steps) leaves the count of running nodes
__global__ void __launch_bounds__(kOneCta)
med_tail(int j, MedBuffers mb) {
    const int b = blockIdx.x;
    const MedNode &nd = mb.nodes[b];
    MedState &st = mb.state[b];
    if (st.mode != MED_LANCZOS || st.lanczos[1] || j > st.jmax) return;
    const int m = st.m, base = nd.row_base;
    const int check = (j == st.jmax) || (j % 4) == 0;
    double *basis = mb.basis + nd.basis_off;
    double *ritz = node_ritz(mb, b);
    lanczos_tail_body(m, j, 1 + j, j, check, basis, mb.y + base, mb.isd + base, node_alpha(mb, b), node_beta(mb, b),
                      basis + static_cast<size_t>(j + 1) * m, mb.z + base, st.lanczos, node_coef(mb, b), ritz);
    if (threadIdx.x == 0) {
        if (st.lanczos[1]) {
            st.outcome = ritz[3] <= kBreakdown ? MED_INVARIANT : MED_CONVERGED;
            atomicSub(mb.remaining, 1);
        } else if (j == st.jmax) {
            st.outcome = MED_OUT_OF_STEPS;
            st.lanczos[1] = 1;
            st.lanczos[2] = j;
            atomicSub(mb.remaining, 1);
        }
    }
}

// Ritz vector of the last projected problem solved -> yvec, z = isd .* yvec
__global__ void __launch_bounds__(kOneCta)
med_ritz_vector(MedBuffers mb) {
    const int b = blockIdx.x;
    const MedNode &nd = mb.nodes[b];
    const MedState &st = mb.state[b];
    if (st.mode != MED_LANCZOS) return;
    const int base = nd.row_base;
    const int jdone = static_cast<int>(node_ritz(mb, b)[5]);
    ritz_vector_body(st.m, jdone, mb.basis + nd.basis_off, node_coef(mb, b), mb.isd + base, mb.yvec + base, mb.z + base);
}

// true residual, 2-means, labels and the node's record; one CTA per node
__global__ void __launch_bounds__(kOneCta)
med_finish(MedBuffers mb, int32_t *__restrict__ part, scs_node_stats *__restrict__ stats) {
    const int b = blockIdx.x;
    const MedNode &nd = mb.nodes[b];
    const MedState &st = mb.state[b];
    const int n = nd.n, m = st.m, base = nd.row_base, tid = threadIdx.x;
    const double nan_v = __longlong_as_double(0x7ff8000000000000ll);
    scs_node_stats out;
    out.n_components = st.ncomp[0];
    out.contracted_size = st.mode == MED_DISCONNECTED ? n : m;
    out.spectral_ran = st.mode != MED_DISCONNECTED;
    out.solver = 0;
    out.matvecs = 0;
    out.restarts = 0;
    out.tie_flag = 0;
    out.kmeans_stable_splits = 0;
    out.eig[0] = 0.0;
    out.eig[1] = out.eig[2] = nan_v;
    out.residual = nan_v;
    out.margin = nan_v;
    out.kmeans_runner_up = 0.0;
    if (st.mode == MED_DISCONNECTED) {  // labels were written by med_ranks
        if (tid == 0) stats[b] = out;
        return;
    }
    if (st.mode == MED_TOO_SMALL) {
        out.solver = -1;  // the host raises SCS_ERR_TOO_SMALL (sklearn: ensure_min_samples = 2, _spectral.py:699)
        if (tid == 0) stats[b] = out;
        return;
    }
    const int32_t *group = mb.group + base;
    if (st.mode == MED_PAIR) {
        // sklearn falls back to a dense eigh (arpack.py:1691-1707); two vertices always separate
        for (int v = tid; v < n; v += kOneCta) part[nd.part_off + v] = group[v];
        out.solver = 1;
        out.eig[1] = 2.0;
        out.residual = 0.0;
        out.margin = 0.5;
        if (tid == 0) stats[b] = out;
        return;
    }
    double *ritz = node_ritz(mb, b);
    true_residual_body(m, mb.yvec + base, mb.y + base, ritz, ritz + 8);
    const int P = 1 << (32 - __clz(max(m - 1, 1)));  // next power of two >= m (m >= 3)
    two_means_1d_body<true>(m, P, mb.yvec + base, mb.isd + base, mb.embed + base, nullptr, mb.side + base, ritz + 12);
    __syncthreads();
    for (int v = tid; v < n; v += kOneCta) part[nd.part_off + v] = mb.side[base + group[v]];
    if (tid == 0) {
        const double theta1 = ritz[0], theta2 = ritz[1];
        const int steps = static_cast<int>(ritz[4]);
        out.solver = 3;
        out.matvecs = st.lanczos[2] + 1;
        out.eig[1] = 1.0 - theta1;
        out.eig[2] = 1.0 - theta2;
        out.residual = ritz[8];
        out.margin = ritz[12];
        out.kmeans_stable_splits = static_cast<int32_t>(ritz[16]);
        out.kmeans_runner_up = ritz[17];
        int flag = 0;
        if (!isnan(theta2) && (theta1 - theta2) < kGapTie) flag |= 1;
        if (!(out.margin >= kMarginTie)) flag |= 2;
        if (st.bad_degree) flag |= 8;
        if (out.kmeans_stable_splits > 1) flag |= 16;
        out.tie_flag = flag;
        // what a lock-step batch cannot do: an explicit restart after the basis is full, and the second, deflated run
        // that settles a repeated Fiedler eigenvalue when the Krylov space closed early -- the host re-runs the node
        if (st.outcome == MED_OUT_OF_STEPS || (st.outcome == MED_INVARIANT && steps < m - 1)) out.solver = -3;
        stats[b] = out;
    }
}

template <typename T>
int carve(unsigned char *&cursor, size_t count, T **out) {
    const size_t bytes = (count * sizeof(T) + 255) & ~static_cast<size_t>(255);
    *out = reinterpret_cast<T *>(cursor);
    cursor += bytes;
    return 0;
}

}  // namespace

int medium_batch(scs_ctx *ctx, int B, const int32_t *node_n, const int32_t *tree_begin, const int32_t *tree_end,
                 const int64_t *part_off, const uint64_t *seeds, int T, int64_t L, const int64_t *leaf_offsets, const int32_t *leaf_taxon,
                 const int32_t *adj_depth, const double *adj_val, const int32_t *root_depth, const double *tree_weight,
                 int contract_edges, int32_t *part_dev, scs_node_stats *stats_host, uint8_t *needs_rerun) {
    if (B <= 0) return SCS_OK;
    if (!node_n || !tree_begin || !tree_end || !part_off || !seeds || !part_dev || !stats_host || !needs_rerun || T < 0 || L < 0)
        return fail(ctx, SCS_ERR_INVALID, "medium batch: bad argument");
    // ---- node table -------------------------------------------------------------------------------------------------
    std::vector<MedNode> nodes(static_cast<size_t>(B));
    int64_t R = 0, w_total = 0, bit_total = 0, basis_total = 0, blocks = 0;
    int max_n = 0, max_trees = 0;
    for (int b = 0; b < B; ++b) {
        const int n = node_n[b];
        if (n < 1 || n > kMediumMax) return fail(ctx, SCS_ERR_INVALID, "medium batch: node size out of range");
        if (tree_begin[b] < 0 || tree_end[b] < tree_begin[b] || tree_end[b] > T || (b > 0 && tree_begin[b] < tree_end[b - 1]))
            return fail(ctx, SCS_ERR_INVALID, "medium batch: tree ranges out of order");
        MedNode &nd = nodes[b];
        nd.n = n;
        nd.words = scs_bit_words(n);
        nd.row_base = static_cast<int32_t>(R);
        nd.tree_begin = tree_begin[b];
        nd.tree_end = tree_end[b];
        nd.jcap = std::min(n - 1, kMaxBasis);
        nd.blk_base = static_cast<int32_t>(blocks);
        nd.pad0 = 0;
        blocks += (n + 31) / 32;
        nd.w_off = w_total;
        nd.bit_off = bit_total;
        nd.basis_off = basis_total;
        nd.part_off = part_off[b];
        nd.seed = seeds[b];
        R += n;
        w_total += static_cast<int64_t>(n) * n;
        bit_total += static_cast<int64_t>(n) * nd.words;
        basis_total += static_cast<int64_t>(nd.jcap + 3) * n;
        max_n = std::max(max_n, n);
        max_trees = std::max(max_trees, nd.tree_end - nd.tree_begin);
    }
    if (R >= (1ll << 31)) return fail(ctx, SCS_ERR_INVALID, "medium batch: too many rows");

    // ---- workspace ---------------------------------------------------------------------------------------------------
    MedBuffers mb;
    int rc;
    MedNode *nodes_dev;
    int32_t *row_node, *tree_node;
    if ((rc = reserve_as(ctx, SLOT_MED_NODES, static_cast<size_t>(B), &nodes_dev))) return rc;
    if ((rc = reserve_as(ctx, SLOT_MED_ROW_NODE, static_cast<size_t>(R), &row_node))) return rc;
    if ((rc = reserve_as(ctx, SLOT_MED_TREE_NODE, static_cast<size_t>(T) + 1, &tree_node))) return rc;
    if ((rc = reserve_as(ctx, SLOT_W, static_cast<size_t>(w_total), &mb.W))) return rc;
    if ((rc = reserve_as(ctx, SLOT_WC, static_cast<size_t>(w_total), &mb.Wc))) return rc;
    if ((rc = reserve_as(ctx, SLOT_ADJ_BITS, static_cast<size_t>(bit_total), &mb.adj_bits))) return rc;
    if ((rc = reserve_as(ctx, SLOT_MAX_BITS, static_cast<size_t>(bit_total), &mb.max_bits))) return rc;
    if ((rc = reserve_as(ctx, SLOT_BASIS, static_cast<size_t>(basis_total), &mb.basis))) return rc;
    {
        // everything that is one value per row (or per node), carved out of two slots
        const size_t rows = static_cast<size_t>(R);
        const size_t vec_bytes = 9 * (rows * sizeof(double) + 256) + static_cast<size_t>(B) * kScalarStride * sizeof(double) + 256;
        const size_t int_bytes = 9 * (rows * sizeof(int32_t) + 256) + (rows + B + 1) * sizeof(int32_t) + 256 +
                                 static_cast<size_t>(B) * sizeof(MedState) + 512;
        unsigned char *vec, *ints;
        if ((rc = reserve_as(ctx, SLOT_MED_VEC, vec_bytes, &vec))) return rc;
        if ((rc = reserve_as(ctx, SLOT_MED_STATE, int_bytes, &ints))) return rc;
        carve(vec, rows, &mb.degree);
        carve(vec, rows, &mb.degree_c);
        carve(vec, rows, &mb.isd);
        carve(vec, rows, &mb.z);
        carve(vec, rows, &mb.y);
        carve(vec, rows, &mb.yvec);
        carve(vec, rows, &mb.wstart);
        carve(vec, rows, &mb.embed);
        carve(vec, static_cast<size_t>(B) * kScalarStride, &mb.scalars);
        carve(ints, rows, &mb.occ);
        carve(ints, rows, &mb.uf_parent);
        carve(ints, rows, &mb.label[0]);
        carve(ints, rows, &mb.label[1]);
        carve(ints, rows, &mb.group);
        carve(ints, rows, &mb.members);
        carve(ints, rows, &mb.side);
        carve(ints, rows + B + 1, &mb.gptr);
        carve(ints, static_cast<size_t>(B), &mb.state);
        carve(ints, 64, &mb.remaining);
    }
    mb.nodes = nodes_dev;
    mb.row_node = row_node;
    int32_t *bad_dev = mb.remaining + 8;
    scs_node_stats *stats_dev;
    if ((rc = reserve_as(ctx, SLOT_NODE_STATS, static_cast<size_t>(B), &stats_dev))) return rc;

    // the node table goes through pinned memory so that the copy is asynchronous
    void *pin_v;
    const size_t table_bytes = sizeof(MedNode) * static_cast<size_t>(B);
    if ((rc = reserve_pinned(ctx, table_bytes + 1024, &pin_v))) return rc;
    unsigned char *pin = static_cast<unsigned char *>(pin_v);
    std::memcpy(pin + 1024, nodes.data(), table_bytes);
    SCS_CUDA(ctx, cudaMemcpyAsync(nodes_dev, pin + 1024, table_bytes, cudaMemcpyHostToDevice, ctx->stream));
    ctx->h2d_bytes += static_cast<int64_t>(table_bytes);
    SCS_CUDA(ctx, cudaMemsetAsync(bad_dev, 0, sizeof(int32_t), ctx->stream));
    const int Ri = static_cast<int>(R);
    med_fill_maps<<<ceil_div(std::max<int64_t>(R, T), 256), 256, 0, ctx->stream>>>(B, Ri, T, nodes_dev, row_node, tree_node);
    SCS_LAUNCHED(ctx, "med_fill_maps");
    med_clear_state<<<ceil_div(B, 256), 256, 0, ctx->stream>>>(B, mb.state, mb.remaining);
    SCS_LAUNCHED(ctx, "med_clear_state");

    // ---- graph build ---------------------------------------------------------------------------------------------------
    if ((rc = pcg_build_batch(ctx, B, static_cast<int>(blocks), Ri, T, L, max_n, max_trees, nodes_dev, tree_node, row_node, leaf_offsets, leaf_taxon,
                              adj_depth, adj_val, root_depth, tree_weight, mb.W, mb.occ, mb.adj_bits,
                              contract_edges ? mb.max_bits : nullptr, mb.degree, bad_dev)))
        return rc;

    // ---- components of the graph and of the max-graph -------------------------------------------------------------------
    const int row_blocks = ceil_div(R * 32, 256), vec_blocks = ceil_div(R, 256);
    for (int which = 0; which < (contract_edges ? 2 : 1); ++which) {
        const uint32_t *bits = which == 0 ? mb.adj_bits : mb.max_bits;
        med_uf_init<<<vec_blocks, 256, 0, ctx->stream>>>(Ri, mb.uf_parent);
        SCS_LAUNCHED(ctx, "med_uf_init");
        med_uf_hook_sample<<<row_blocks, 256, 0, ctx->stream>>>(Ri, nodes_dev, row_node, bits, mb.uf_parent);
        SCS_LAUNCHED(ctx, "med_uf_hook_sample");
        med_uf_snapshot<<<vec_blocks, 256, 0, ctx->stream>>>(Ri, mb.uf_parent, mb.label[which]);
        SCS_LAUNCHED(ctx, "med_uf_snapshot");
        med_uf_pick_giant<<<B, 1024, 0, ctx->stream>>>(nodes_dev, mb.label[which], mb.state, which);
        SCS_LAUNCHED(ctx, "med_uf_pick_giant");
        med_uf_hook_rest<<<row_blocks, 256, 0, ctx->stream>>>(Ri, nodes_dev, row_node, bits, mb.label[which], mb.state,
                                                             which, mb.uf_parent);
        SCS_LAUNCHED(ctx, "med_uf_hook_rest");
        med_uf_flatten<<<vec_blocks, 256, 0, ctx->stream>>>(Ri, row_node, mb.uf_parent, mb.label[which], mb.state, which);
        SCS_LAUNCHED(ctx, "med_uf_flatten");
    }
    med_decide<<<ceil_div(B, 256), 256, 0, ctx->stream>>>(B, contract_edges, nodes_dev, mb.state, mb.remaining);
    SCS_LAUNCHED(ctx, "med_decide");

    // ---- labels of disconnected nodes, contraction of connected ones -----------------------------------------------------
    if (!ctx->medium_configured) {
        const int tail_smem = static_cast<int>((kMediumMax + 6 * (kMaxBasis + 8)) * sizeof(double));
        SCS_CUDA(ctx, cudaFuncSetAttribute(med_prepare, cudaFuncAttributeMaxDynamicSharedMemorySize, tail_smem));
        SCS_CUDA(ctx, cudaFuncSetAttribute(med_tail, cudaFuncAttributeMaxDynamicSharedMemorySize, tail_smem));
        SCS_CUDA(ctx, cudaFuncSetAttribute(med_finish, cudaFuncAttributeMaxDynamicSharedMemorySize, kMediumMax * 8));
        SCS_CUDA(ctx, cudaFuncSetAttribute(med_ritz_vector, cudaFuncAttributeMaxDynamicSharedMemorySize, (kMaxBasis + 8) * 8));
        SCS_CUDA(ctx, cudaFuncSetAttribute(med_ranks, cudaFuncAttributeMaxDynamicSharedMemorySize, kMediumMax * 8));
        SCS_CUDA(ctx, cudaFuncSetAttribute(med_contract_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, kMediumMax * 8));
        ctx->medium_configured = true;
    }
    med_ranks<<<B, 1024, static_cast<size_t>(max_n) * 8, ctx->stream>>>(contract_edges, mb, part_dev);
    SCS_LAUNCHED(ctx, "med_ranks");
    if (contract_edges) {
        med_contract_rows<<<Ri, 256, static_cast<size_t>(max_n) * 8, ctx->stream>>>(mb);
        SCS_LAUNCHED(ctx, "med_contract_rows");
    }

    // ---- Lanczos in lock-step ----------------------------------------------------------------------------------------------
    const size_t tail_smem = (static_cast<size_t>(max_n) + 6 * (kMaxBasis + 8)) * sizeof(double);
    med_prepare<<<B, kOneCta, tail_smem, ctx->stream>>>(mb);
    SCS_LAUNCHED(ctx, "med_prepare");
    const int steps_max = std::min(std::max(max_n - 1, 1), kMaxBasis);
    int32_t *pin_remaining = reinterpret_cast<int32_t *>(pin);  // [0..15]: one slot per chunk in flight
    cudaEvent_t probe[2] = {nullptr, nullptr};
    for (cudaEvent_t &e : probe) SCS_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    // chunks of steps end on check steps (multiples of 4); the count of running nodes after chunk k is read while chunk
    // k + 1 is already queued, so the GPU never waits for the host
    int j = 1, chunk = 0, pending = -1;
    bool finished = false;
    while (j <= steps_max && !finished) {
        const int chunk_end = std::min(steps_max, j == 1 ? 16 : j + 3);
        for (; j <= chunk_end; ++j) {
            med_matvec<<<row_blocks, kMvThreads, 0, ctx->stream>>>(Ri, 0, mb);
            SCS_LAUNCHED(ctx, "med_matvec");
            med_tail<<<B, kOneCta, tail_smem, ctx->stream>>>(j, mb);
            SCS_LAUNCHED(ctx, "med_tail");
        }
        SCS_CUDA(ctx, cudaMemcpyAsync(pin_remaining + (chunk & 1), mb.remaining, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
        SCS_CUDA(ctx, cudaEventRecord(probe[chunk & 1], ctx->stream));
        if (pending >= 0) {
            SCS_CUDA(ctx, cudaEventSynchronize(probe[pending & 1]));
            if (pin_remaining[pending & 1] == 0) finished = true;
        }
        pending = chunk++;
    }
    for (cudaEvent_t &e : probe) cudaEventDestroy(e);

    // ---- Ritz vectors, true residuals, 2-means, labels ----------------------------------------------------------------------
    med_ritz_vector<<<B, kOneCta, (kMaxBasis + 8) * sizeof(double), ctx->stream>>>(mb);
    SCS_LAUNCHED(ctx, "med_ritz_vector");
    med_matvec<<<row_blocks, kMvThreads, 0, ctx->stream>>>(Ri, 1, mb);
    SCS_LAUNCHED(ctx, "med_matvec");
    int P = 1;
    while (P < max_n) P <<= 1;
    med_finish<<<B, kOneCta, static_cast<size_t>(P) * sizeof(double), ctx->stream>>>(mb, part_dev, stats_dev);
    SCS_LAUNCHED(ctx, "med_finish");

    SCS_CUDA(ctx, cudaMemcpyAsync(stats_host, stats_dev, sizeof(scs_node_stats) * B, cudaMemcpyDeviceToHost, ctx->stream));
    SCS_CUDA(ctx, cudaMemcpyAsync(pin + 512, bad_dev, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    SCS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->d2h_bytes += static_cast<int64_t>(sizeof(scs_node_stats)) * B;
    if (*reinterpret_cast<int32_t *>(pin + 512) != 0) return fail(ctx, SCS_ERR_INPUT, "leaf tour: taxon id out of range");
    for (int b = 0; b < B; ++b) {
        needs_rerun[b] = stats_host[b].solver == -3;
        if (stats_host[b].solver == -1) {
            stats_host[b].solver = 0;
            return fail(ctx, SCS_ERR_TOO_SMALL, "spectral step on a graph contracted to one vertex");
        }
    }
    return SCS_OK;
}

}  // namespace scs
