// Proper-cluster-graph build on the GPU.
//
// Replaces _proper_cluster_graph_edges + _dfs_pcg_weights of the reference
// (/root/reference/src/sc_supertree/scs.py:495-583, 586-663).
//
// Layout.  Source trees arrive as leaf tours (flatten.py / forest.cpp): leaves in depth-first order
// and, for consecutive leaves i, i+1, the depth and weighting value of their LCA.  The LCA of the
// leaves at tour positions p < q is the shallowest entry of adj[p..q-1]: a range-minimum query.
// pcg_sparse_table builds, once per tree, the classic doubling table over (depth, position) keys
// (level j holds the minimum of 2^j consecutive entries), so every leaf pair's LCA is two table
// reads and a min -- no pointer chasing, no per-row rescan of the tour.
//
// Kernel shape.  One CTA owns one row `a` of W (one taxon) and keeps the row's accumulators in
// shared memory (8 B weight + 2/4 B count per column).  It visits the trees containing `a` in
// input order (their headers are staged 64 at a time, so the dependent loads that locate a tree
// overlap); for each, every thread takes other leaves q of the tree, looks up LCA(a, q) and, unless
// it is the root (scs.py:570-579: pairs separated by the root are not proper clusters), adds
// fl(val(LCA) * w_t) to column taxon(q) (scs.py:644-658).  Within a tree every leaf is a distinct
// column, so threads never collide; one barrier separates consecutive trees, so every W entry is
// summed in tree input order with separately rounded multiply and add (scs.py:655-657) --
// bit-identical to the reference, no atomics on W, and W[a][b] == W[b][a] bit for bit.  The finished
// row is written once, coalesced, together with its adjacency bits (C > 0, scs.py:651-652),
// max-graph bits (C == max(occ_a, occ_b), scs.py:302-305) and its row sum (the degree the spectral
// step needs), so the co-occurrence matrix never has to be written to HBM unless the caller asks.
//
// Rows wider than shared memory (n > ~22k columns) are split into column chunks (gridDim.y);
// each chunk CTA walks the same trees and keeps only its columns.

#include "common.cuh"

namespace scs {

namespace {

constexpr int kRowThreads = 512;  // 2 CTAs/SM at n = 10^4 (100 KB of row accumulators each): 32 warps/SM
constexpr int kWarps = kRowThreads / 32;
constexpr int kHeaderBatch = 64;  // (row, tree) incidences whose headers are staged together
constexpr unsigned long long kNoKey = ~0ull;

// ---- index: leaf -> tree, occurrences, taxon -> leaves ------------------------------------
__global__ void pcg_index_leaves(int n, int T, int64_t L, const int64_t *__restrict__ leaf_offsets,
                                 const int32_t *__restrict__ leaf_taxon, int32_t *__restrict__ leaf_tree,
                                 int32_t *__restrict__ occ, int32_t *__restrict__ bad) {
    int64_t g = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (g >= L) return;
    int lo = 0, hi = T;  // largest t with leaf_offsets[t] <= g
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (leaf_offsets[mid] <= g) lo = mid; else hi = mid;
    }
    leaf_tree[g] = lo;
    int a = leaf_taxon[g];
    if (a < 0 || a >= n) { *bad = 1; return; }
    atomicAdd(&occ[a], 1);
}

// out[i] = sum of in[0..i), out[n] = total.  One block.
__global__ void exclusive_scan_i32(int n, const int32_t *__restrict__ in, int32_t *__restrict__ out) {
    __shared__ int32_t warp_sum[32];
    __shared__ int32_t carry_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < n; base += blockDim.x) {
        int i = base + threadIdx.x;
        int32_t v = i < n ? in[i] : 0;
        int32_t inc = v;
        for (int off = 1; off < 32; off <<= 1) {
            int32_t o = __shfl_up_sync(0xffffffffu, inc, off);
            if (lane >= off) inc += o;
        }
        if (lane == 31) warp_sum[warp] = inc;
        __syncthreads();
        int32_t before = carry_s;
        for (int w = 0; w < warp; ++w) before += warp_sum[w];
        if (i < n) out[i] = before + inc - v;
        __syncthreads();
        if (threadIdx.x == 0) {
            int32_t tot = carry_s;
            for (int w = 0; w < nwarp; ++w) tot += warp_sum[w];
            carry_s = tot;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) out[n] = carry_s;
}

__global__ void pcg_fill_inverse(int n, int64_t L, const int32_t *__restrict__ leaf_taxon,
                                 const int32_t *__restrict__ row_ptr, int32_t *__restrict__ cursor,
                                 int32_t *__restrict__ inv) {
    int64_t g = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (g >= L) return;
    int a = leaf_taxon[g];
    if (a < 0 || a >= n) return;  // flagged by pcg_index_leaves
    int slot = atomicAdd(&cursor[a], 1);
    inv[row_ptr[a] + slot] = static_cast<int32_t>(g);
}

// The atomics above leave each taxon's leaf list in arbitrary order; leaves are numbered tree
// by tree, so sorting a list ascending restores tree input order.  Rank sort, one CTA per row.
__global__ void pcg_sort_inverse(int row0, const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ inv,
                                 int32_t *__restrict__ inv_sorted) {
    __shared__ int32_t stage[1024];
    const int a = row0 + blockIdx.x;
    const int base = row_ptr[a];
    const int cnt = row_ptr[a + 1] - base;
    for (int i0 = 0; i0 < cnt; i0 += blockDim.x) {
        const int i = i0 + threadIdx.x;
        const int32_t mine = i < cnt ? inv[base + i] : 0;
        int rank = 0;
        for (int j0 = 0; j0 < cnt; j0 += 1024) {
            const int len = min(1024, cnt - j0);
            __syncthreads();
            for (int j = threadIdx.x; j < len; j += blockDim.x) stage[j] = inv[base + j0 + j];
            __syncthreads();
            if (i < cnt)
                for (int j = 0; j < len; ++j) rank += stage[j] < mine;
        }
        if (i < cnt) inv_sorted[base + rank] = mine;
    }
}

// ---- range-minimum table over the consecutive-leaf LCAs --------------------------------------------
// st[j * L + g] = the shallowest adj entry among [g, g + 2^j) of the same tree, as {key, value} with
// key = depth << 32 | position in tree (ties: leftmost; equal depth in a range means the same node) and
// value = the weighting value of that LCA, so a lookup needs no second, dependent read.
// One CTA per tree builds all levels of its tree.
struct __align__(16) LcaEntry {
    unsigned long long key;
    double val;
};

__global__ void __launch_bounds__(256)
pcg_sparse_table(int64_t L, int levels, const int64_t *__restrict__ leaf_offsets,
                 const int32_t *__restrict__ adj_depth, const double *__restrict__ adj_val, LcaEntry *st) {
    const int t = blockIdx.x;
    const int64_t tb = leaf_offsets[t];
    const int k = static_cast<int>(leaf_offsets[t + 1] - tb);
    const int entries = k - 1;  // adj entries of this tree (the last leaf has none)
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        LcaEntry e;
        e.key = i < entries
                    ? (static_cast<unsigned long long>(static_cast<uint32_t>(adj_depth[tb + i])) << 32) | static_cast<uint32_t>(i)
                    : kNoKey;
        e.val = i < entries ? adj_val[tb + i] : 0.0;
        st[tb + i] = e;
    }
    for (int j = 1; j < levels; ++j) {
        const int span = 1 << j;
        if (span > entries) break;
        __syncthreads();
        const LcaEntry *prev = st + static_cast<size_t>(j - 1) * L + tb;
        LcaEntry *cur = st + static_cast<size_t>(j) * L + tb;
        for (int i = threadIdx.x; i + span <= entries; i += blockDim.x) {
            const LcaEntry x = prev[i], y = prev[i + (span >> 1)];
            cur[i] = x.key < y.key ? x : y;
        }
    }
}

// ---- the row kernel -----------------------------------------------------------------------
struct TreeHeader {
    int64_t base;  // first leaf of the tree in the tour arrays
    double weight;
    int leaves, position, root_depth, pad;
};

constexpr int kPerThread = 2;                          // visits per thread per segment
constexpr int kSegment = kRowThreads * kPerThread;     // visits per segment

struct Pending {
    int col[kPerThread];     // column of the chunk, or -1
    double term[kPerThread];  // fl(val(LCA) * w_t)
};

// Everything a segment needs from global memory, into registers.  v enumerates the other leaves of the
// tree: v < p is leaf v, v >= p is leaf v + 1.
__device__ __forceinline__ LcaEntry load_entry(const LcaEntry *p) {
    const ulonglong2 raw = *reinterpret_cast<const ulonglong2 *>(p);  // one 16-byte load
    LcaEntry e;
    e.key = raw.x;
    e.val = __longlong_as_double(static_cast<long long>(raw.y));
    return e;
}

__device__ __forceinline__ void load_segment(const TreeHeader &h, int seg, int64_t L,
                                             const LcaEntry *__restrict__ st,
                                             const int32_t *__restrict__ leaf_taxon, int col0, int ncols, int tid,
                                             Pending &out) {
    const int p = h.position;
    const LcaEntry *st_t = st + h.base;
    const int32_t *taxon_t = leaf_taxon + h.base;
    LcaEntry x[kPerThread], y[kPerThread];
    int col[kPerThread];
#pragma unroll
    for (int r = 0; r < kPerThread; ++r) {
        const int v = seg * kSegment + r * kRowThreads + tid;
        col[r] = -1;
        x[r].key = y[r].key = kNoKey;
        x[r].val = y[r].val = 0.0;
        if (v < h.leaves - 1) {
            const int q = v < p ? v : v + 1;
            const int lo = v < p ? q : p;  // adj entries [lo, lo + len) lie between the two leaves
            const int len = v < p ? p - q : q - p;
            const int j = 31 - __clz(len);
            const LcaEntry *level = st_t + static_cast<size_t>(j) * L;
            x[r] = load_entry(level + lo);
            y[r] = load_entry(level + lo + len - (1 << j));
            col[r] = taxon_t[q] - col0;
        }
    }
#pragma unroll
    for (int r = 0; r < kPerThread; ++r) {
        const LcaEntry e = x[r].key < y[r].key ? x[r] : y[r];
        const bool proper = col[r] != -1 && static_cast<int>(e.key >> 32) != h.root_depth &&
                            static_cast<unsigned>(col[r]) < static_cast<unsigned>(ncols);
        out.col[r] = proper ? col[r] : -1;
        out.term[r] = __dmul_rn(e.val, h.weight);
    }
}

// next (tree, segment) of the batch
__device__ __forceinline__ void advance(const TreeHeader *headers, int batch, int &e, int &seg) {
    if (e >= batch) return;
    ++seg;
    if (seg * kSegment >= headers[e].leaves - 1) {
        seg = 0;
        do {
            ++e;
        } while (e < batch && headers[e].leaves < 2);
    }
}

template <typename CountT, bool kWriteC>
__global__ void __launch_bounds__(kRowThreads)
pcg_rows_kernel(int n, int row0, int words_per_row, int cols_per_chunk, int64_t L,
                const int64_t *__restrict__ leaf_offsets, const int32_t *__restrict__ leaf_taxon,
                const LcaEntry *__restrict__ st, const int32_t *__restrict__ root_depth,
                const double *__restrict__ tree_weight,
                const int32_t *__restrict__ leaf_tree, const int32_t *__restrict__ row_ptr,
                const int32_t *__restrict__ inv_sorted, const int32_t *__restrict__ occ,
                double *__restrict__ W, int32_t *__restrict__ C, uint32_t *__restrict__ adj_bits,
                uint32_t *__restrict__ max_bits, double *__restrict__ degree_part, int32_t *__restrict__ bad) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *accW = reinterpret_cast<double *>(smem_raw);
    CountT *accC = reinterpret_cast<CountT *>(accW + cols_per_chunk);
    __shared__ TreeHeader headers[kHeaderBatch];
    __shared__ double warp_sum[kWarps];

    const int a = row0 + blockIdx.x;  // W / C point at the row block: its first row is row0
    const int col0 = blockIdx.y * cols_per_chunk;
    const int ncols = min(cols_per_chunk, n - col0);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (int c = tid; c < cols_per_chunk; c += kRowThreads) {
        accW[c] = 0.0;
        accC[c] = 0;
    }

    const int ebase = row_ptr[a];
    const int cnt = row_ptr[a + 1] - ebase;
    for (int e0 = 0; e0 < cnt; e0 += kHeaderBatch) {
        const int batch = min(kHeaderBatch, cnt - e0);
        __syncthreads();  // the previous batch's headers (and the initialisation) are done with
        if (tid < batch) {
            const int g = inv_sorted[ebase + e0 + tid];
            const int t = leaf_tree[g];
            const int64_t tb = leaf_offsets[t];
            TreeHeader h;
            h.base = tb;
            h.weight = tree_weight[t];
            h.leaves = static_cast<int>(leaf_offsets[t + 1] - tb);
            if (h.leaves > n) {  // more leaves than taxa: a taxon is repeated, the table would be too shallow
                *bad = 1;
                h.leaves = 0;
            }
            h.position = static_cast<int>(g - tb);
            h.root_depth = root_depth[t];
            h.pad = 0;
            headers[tid] = h;
        }
        __syncthreads();
        // Software pipeline over segments of kSegment visits: the loads of the next segment (table,
        // taxon, value: three dependent L2 round trips) are issued before the shared-memory updates of
        // the current one, and each thread keeps kPerThread independent chains in flight.  Only the
        // updates are ordered: a barrier whenever the tree changes (segments of one tree touch
        // distinct columns).
        int next_e = 0, next_seg = 0;
        Pending cur, nxt;
        load_segment(headers[0], 0, L, st, leaf_taxon, col0, ncols, tid, cur);
        int cur_e = 0;
        advance(headers, batch, next_e, next_seg);
        while (true) {
            const bool more = next_e < batch;
            if (more) load_segment(headers[next_e], next_seg, L, st, leaf_taxon, col0, ncols, tid, nxt);
#pragma unroll
            for (int r = 0; r < kPerThread; ++r) {
                const int c = cur.col[r];
                if (c >= 0) {
                    accW[c] = __dadd_rn(accW[c], cur.term[r]);
                    accC[c] = static_cast<CountT>(accC[c] + 1);
                }
            }
            if (!more) break;
            if (next_e != cur_e) __syncthreads();  // tree order: the next tree may touch the same columns
            cur = nxt;
            cur_e = next_e;
            advance(headers, batch, next_e, next_seg);
        }
    }
    __syncthreads();

    // ---- write the finished row --------------------------------------------------------------
    const int occ_a = occ[a];
    double *Wrow = W + static_cast<size_t>(blockIdx.x) * n + col0;
    double partial = 0.0;
    for (int c = tid; c < ncols; c += kRowThreads) {
        const double x = accW[c];
        Wrow[c] = x;
        partial += x;
        if (kWriteC) C[static_cast<size_t>(blockIdx.x) * n + col0 + c] = static_cast<int32_t>(accC[c]);
    }
    const int word0 = col0 >> 5;
    const int nwords = (ncols + 31) >> 5;
    for (int j = warp; j < nwords; j += kWarps) {
        const int c = (j << 5) + lane;
        bool edge = false, top = false;
        if (c < ncols) {
            const int cc = static_cast<int>(accC[c]);
            edge = cc > 0;
            if (edge && max_bits != nullptr) top = cc == max(occ_a, occ[col0 + c]);
        }
        const uint32_t eb = __ballot_sync(0xffffffffu, edge);
        const uint32_t tb2 = __ballot_sync(0xffffffffu, top);
        if (lane == 0) {
            adj_bits[static_cast<size_t>(a) * words_per_row + word0 + j] = eb;
            if (max_bits != nullptr) max_bits[static_cast<size_t>(a) * words_per_row + word0 + j] = tb2;
        }
    }
    // row sum in a fixed order: strided per-thread sums, shuffle tree, then warps in order
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) partial += __shfl_down_sync(0xffffffffu, partial, off);
    if (lane == 0) warp_sum[warp] = partial;
    __syncthreads();
    if (tid == 0) {
        double s = 0.0;
        for (int wi = 0; wi < kWarps; ++wi) s += warp_sum[wi];
        degree_part[static_cast<size_t>(blockIdx.y) * n + a] = s;
    }
}

__global__ void pcg_sum_degree_parts(int n, int row0, int row1, int nchunks, const double *__restrict__ part,
                                     double *__restrict__ degree) {
    int a = row0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= row1) return;
    double s = 0.0;
    for (int c = 0; c < nchunks; ++c) s += part[static_cast<size_t>(c) * n + a];
    degree[a] = s;
}

template <typename CountT, bool kWriteC>
int launch_rows(scs_ctx *ctx, int n, int row0, int nrows, int words, int cols_per_chunk, int nchunks, size_t smem, int64_t L,
                const int64_t *leaf_offsets, const int32_t *leaf_taxon, const LcaEntry *st,
                const int32_t *root_depth, const double *tree_weight,
                const int32_t *leaf_tree, const int32_t *row_ptr, const int32_t *inv_sorted,
                const int32_t *occ, double *W, int32_t *C, uint32_t *adj_bits, uint32_t *max_bits,
                double *degree_part, int32_t *bad) {
    auto kernel = pcg_rows_kernel<CountT, kWriteC>;
    // always the same (maximal) opt-in size: contexts on other host threads launch this kernel concurrently
    bool &configured = ctx->rows_configured[(sizeof(CountT) == 2 ? 0 : 2) + (kWriteC ? 1 : 0)];
    if (!configured) {
        const size_t optin = ctx->smem_optin > 8192 ? ctx->smem_optin - 4096 : 44 * 1024;
        SCS_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(optin)));
        configured = true;
    }
    dim3 grid(nrows, nchunks);
    if (n >= kProfileMinSize) {
        // algorithmic bytes: W + both bit matrices + occ/degree written once, C if asked
        const double out_bytes = 8.0 * nrows * n + (C ? 4.0 * nrows * n : 0.0) + 8.0 * nrows * words + 12.0 * nrows;
        profile_begin(ctx, PROFILE_PCG_ROWS, out_bytes, ctx->pending_units);
    }
    kernel<<<grid, kRowThreads, smem, ctx->stream>>>(n, row0, words, cols_per_chunk, L, leaf_offsets, leaf_taxon, st,
                                                     root_depth, tree_weight, leaf_tree, row_ptr,
                                                     inv_sorted, occ, W, C, adj_bits, max_bits, degree_part, bad);
    if (n >= kProfileMinSize) profile_end(ctx);
    SCS_LAUNCHED(ctx, "pcg_rows_kernel");
    return SCS_OK;
}

}  // namespace

int exclusive_scan(scs_ctx *ctx, int n, const int32_t *in, int32_t *out) {
    exclusive_scan_i32<<<1, 1024, 0, ctx->stream>>>(n, in, out);
    SCS_LAUNCHED(ctx, "exclusive_scan_i32");
    return SCS_OK;
}

int pcg_build(scs_ctx *ctx, int n, int T, int64_t L, const int64_t *leaf_offsets,
              const int32_t *leaf_taxon, const int32_t *adj_depth, const double *adj_val,
              const int32_t *root_depth, const double *tree_weight, double *W, int32_t *C,
              int32_t *occ, uint32_t *adj_bits, uint32_t *max_bits, double *degree, RowBlock rows) {
    const int row0 = rows.sharded() ? rows.row0 : 0;
    const int row1 = rows.sharded() ? rows.row1 : n;
    const int nrows = row1 - row0;
    if (row0 < 0 || row1 > n || nrows < 0) return fail(ctx, SCS_ERR_INVALID, "pcg_build: bad row block");
    if (n <= 0 || T < 0 || L < 0 || L >= (1ll << 31) || !W || !occ || !adj_bits)
        return fail(ctx, SCS_ERR_INVALID, "pcg_build: bad argument");
    if (T > 0 && (!leaf_offsets || !root_depth || !tree_weight))
        return fail(ctx, SCS_ERR_INVALID, "pcg_build: null tour array");
    if (L > 0 && (!leaf_taxon || !adj_depth || !adj_val))
        return fail(ctx, SCS_ERR_INVALID, "pcg_build: null tour array");
    const int words = scs_bit_words(n);

    int32_t *leaf_tree, *row_ptr, *cursor, *inv, *inv_sorted, *scalars;
    double *degree_part;
    int rc;
    if ((rc = reserve_as(ctx, SLOT_LEAF_TREE, static_cast<size_t>(L) + 1, &leaf_tree))) return rc;
    if ((rc = reserve_as(ctx, SLOT_ROW_PTR, static_cast<size_t>(n) + 1, &row_ptr))) return rc;
    if ((rc = reserve_as(ctx, SLOT_CURSOR, static_cast<size_t>(n), &cursor))) return rc;
    if ((rc = reserve_as(ctx, SLOT_INV, static_cast<size_t>(L) + 1, &inv))) return rc;
    if ((rc = reserve_as(ctx, SLOT_INV_SORTED, static_cast<size_t>(L) + 1, &inv_sorted))) return rc;
    if ((rc = reserve_as(ctx, SLOT_SCALARS, 64, &scalars))) return rc;

    SCS_CUDA(ctx, cudaMemsetAsync(occ, 0, sizeof(int32_t) * n, ctx->stream));
    SCS_CUDA(ctx, cudaMemsetAsync(cursor, 0, sizeof(int32_t) * n, ctx->stream));
    SCS_CUDA(ctx, cudaMemsetAsync(scalars, 0, sizeof(int32_t) * 64, ctx->stream));
    if (L > 0) {
        pcg_index_leaves<<<ceil_div(L, 256), 256, 0, ctx->stream>>>(n, T, L, leaf_offsets, leaf_taxon, leaf_tree, occ,
                                                                   scalars);
        SCS_LAUNCHED(ctx, "pcg_index_leaves");
    }
    if ((rc = exclusive_scan(ctx, n, occ, row_ptr))) return rc;
    if (L > 0) {
        pcg_fill_inverse<<<ceil_div(L, 256), 256, 0, ctx->stream>>>(n, L, leaf_taxon, row_ptr, cursor, inv);
        SCS_LAUNCHED(ctx, "pcg_fill_inverse");
        if (nrows > 0) {
            pcg_sort_inverse<<<nrows, 128, 0, ctx->stream>>>(row0, row_ptr, inv, inv_sorted);
            SCS_LAUNCHED(ctx, "pcg_sort_inverse");
        }
    }
    // a tree has at most n leaves (distinct taxa), so ceil(log2(n)) doubling levels always suffice
    int levels = 1;
    while ((1 << levels) < n) ++levels;
    LcaEntry *st;
    if ((rc = reserve_as(ctx, SLOT_SPARSE, static_cast<size_t>(levels) * static_cast<size_t>(L > 0 ? L : 1), &st)))
        return rc;
    if (L > 0 && T > 0) {
        pcg_sparse_table<<<T, 256, 0, ctx->stream>>>(L, levels, leaf_offsets, adj_depth, adj_val, st);
        SCS_LAUNCHED(ctx, "pcg_sparse_table");
    }

    // column chunking: the whole row if it fits in shared memory, else equal chunks of 32-multiples
    const bool narrow = T < 65536;
    const size_t per_col = sizeof(double) + (narrow ? sizeof(uint16_t) : sizeof(int32_t));
    const size_t budget = ctx->smem_optin > 8192 ? ctx->smem_optin - 4096 : 44 * 1024;
    const int max_cols = static_cast<int>((budget / per_col) / 32 * 32);
    const int padded = words * 32;
    int nchunks = ceil_div(padded, max_cols);
    int cols_per_chunk = ceil_div(ceil_div(padded, nchunks), 32) * 32;
    nchunks = ceil_div(n, cols_per_chunk);
    const size_t smem = static_cast<size_t>(cols_per_chunk) * per_col + 16;
    if ((rc = reserve_as(ctx, SLOT_DEGREE_PART, static_cast<size_t>(n) * nchunks, &degree_part))) return rc;

#define SCS_ROWS(CT, WC)                                                                                      \
    launch_rows<CT, WC>(ctx, n, row0, nrows, words, cols_per_chunk, nchunks, smem, L, leaf_offsets, leaf_taxon, st, \
                        root_depth, tree_weight, leaf_tree, row_ptr, inv_sorted, occ, W, C, adj_bits, max_bits, \
                        degree_part, scalars)
    if (nrows == 0) return SCS_OK;
    if (narrow) rc = C ? SCS_ROWS(uint16_t, true) : SCS_ROWS(uint16_t, false);
    else rc = C ? SCS_ROWS(int32_t, true) : SCS_ROWS(int32_t, false);
#undef SCS_ROWS
    if (rc) return rc;
    if (degree) {
        pcg_sum_degree_parts<<<ceil_div(nrows, 256), 256, 0, ctx->stream>>>(n, row0, row1, nchunks, degree_part, degree);
        SCS_LAUNCHED(ctx, "pcg_sum_degree_parts");
    }
    return SCS_OK;
}

}  // namespace scs
