// Edge contraction of the proper cluster graph.
//
// Replaces _contract_proper_cluster_graph of the reference
// (/root/reference/src/sc_supertree/scs.py:261-387): taxa that sit in the same proper cluster in
// every tree containing either of them (co-occurrence == max of the two occurrence counts,
// scs.py:302-305 -- the max-graph bits the row kernel emitted) are merged; the merged vertices
// are the components of that max-graph (scs.py:316); parallel edges between two merged vertices
// keep the MAXIMUM weight (scs.py:382-387) and edges inside a merged vertex vanish
// (scs.py:352-354).  max is order-free, so shared-memory atomics keep this exact.

#include "common.cuh"

namespace scs {

namespace {

constexpr int kThreads = 256;
__global__ void mark_representatives(int n, const int32_t *__restrict__ label, int32_t *__restrict__ is_rep) {
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v < n) is_rep[v] = label[v] == v;
}

__global__ void assign_groups(int n, const int32_t *__restrict__ label, const int32_t *__restrict__ rep_rank,
                              int32_t *__restrict__ group, int32_t *__restrict__ group_size) {
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    int gidx = rep_rank[label[v]];
    group[v] = gidx;
    atomicAdd(&group_size[gidx], 1);
}

__global__ void fill_members(int n, const int32_t *__restrict__ group, const int32_t *__restrict__ group_ptr,
                             int32_t *__restrict__ cursor, int32_t *__restrict__ members) {
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    int gidx = group[v];
    members[group_ptr[gidx] + atomicAdd(&cursor[gidx], 1)] = v;
}

// Where the rows of W live: one block when the node is on one GPU, else rank r's window holds rows
// [r * rows_per_rank, (r + 1) * rows_per_rank) and the others are read over NVLink.
struct WRows {
    const double *block[kMaxPeers];
    int rows_per_rank;
};

// one CTA per contracted vertex A (and per column chunk of the contracted matrix); the CTA of
// blockIdx.x produces contracted row A = A0 + blockIdx.x into row blockIdx.x of Wc
__global__ void __launch_bounds__(kThreads)
contract_rows(int n, int m, int A0, int words, int cols_per_chunk, const WRows W,
              const uint32_t *__restrict__ adj_bits, const int32_t *__restrict__ group,
              const int32_t *__restrict__ group_ptr, const int32_t *__restrict__ members,
              double *__restrict__ Wc, double *__restrict__ degree_part) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long *best = reinterpret_cast<unsigned long long *>(smem_raw);
    __shared__ double warp_sum[kThreads / 32];
    const int A = A0 + blockIdx.x;
    const int c0 = blockIdx.y * cols_per_chunk;
    const int ncols = min(cols_per_chunk, m - c0);
    const int tid = threadIdx.x;
    for (int c = tid; c < ncols; c += kThreads) best[c] = 0ull;
    __syncthreads();
    const int mb = group_ptr[A], me = group_ptr[A + 1];
    for (int mi = mb; mi < me; ++mi) {
        const int u = members[mi];
        const uint32_t *bits_u = adj_bits + static_cast<size_t>(u) * words;
        const int owner = u / W.rows_per_rank;
        const double *W_u = W.block[owner] + static_cast<size_t>(u - owner * W.rows_per_rank) * n;
        for (int v = tid; v < n; v += kThreads) {
            if ((bits_u[v >> 5] >> (v & 31)) & 1u) {
                const int B = group[v];
                const int c = B - c0;
                if (B != A && static_cast<unsigned>(c) < static_cast<unsigned>(ncols))
                    atomicMax(&best[c], order_key(W_u[v]));
            }
        }
    }
    __syncthreads();
    double partial = 0.0;
    double *out = Wc + static_cast<size_t>(blockIdx.x) * m + c0;
    for (int c = tid; c < ncols; c += kThreads) {
        const unsigned long long k = best[c];
        const double x = k ? order_value(k) : 0.0;
        out[c] = x;
        partial += x;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) partial += __shfl_down_sync(0xffffffffu, partial, off);
    if ((tid & 31) == 0) warp_sum[tid >> 5] = partial;
    __syncthreads();
    if (tid == 0) {
        double s = 0.0;
        for (int wi = 0; wi < kThreads / 32; ++wi) s += warp_sum[wi];
        degree_part[static_cast<size_t>(blockIdx.y) * m + A] = s;
    }
}

__global__ void sum_parts(int m, int row0, int row1, int nchunks, const double *__restrict__ part,
                          double *__restrict__ out) {
    int a = row0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= row1) return;
    double s = 0.0;
    for (int c = 0; c < nchunks; ++c) s += part[static_cast<size_t>(c) * m + a];
    out[a] = s;
}

}  // namespace

// The merge itself, given the max-graph component labels (label[v] = smallest member) and their number m.
int contract_with_labels(scs_ctx *ctx, int n, const double *W, const uint32_t *adj_bits, const int32_t *label, int m,
                         int32_t *group, double *Wc, double *degree_c, RowBlock rows) {
    const int words = scs_bit_words(n);
    int32_t *flags, *rank, *gptr, *cursor, *members;
    int rc;
    // scratch: flags[n] | rank[n+1] | gptr[n+1] | cursor[n]
    if ((rc = reserve_as(ctx, SLOT_GROUP_PTR, 4 * static_cast<size_t>(n) + 8, &flags))) return rc;
    rank = flags + n;
    gptr = rank + n + 1;
    cursor = gptr + n + 1;
    if ((rc = reserve_as(ctx, SLOT_GROUP_MEMBERS, static_cast<size_t>(n), &members))) return rc;
    const int blocks = ceil_div(n, 256);
    mark_representatives<<<blocks, 256, 0, ctx->stream>>>(n, label, flags);
    SCS_LAUNCHED(ctx, "mark_representatives");
    if ((rc = exclusive_scan(ctx, n, flags, rank))) return rc;
    // group sizes go to `cursor` (m <= n entries), then become offsets in gptr
    SCS_CUDA(ctx, cudaMemsetAsync(cursor, 0, sizeof(int32_t) * n, ctx->stream));
    assign_groups<<<blocks, 256, 0, ctx->stream>>>(n, label, rank, group, cursor);
    SCS_LAUNCHED(ctx, "assign_groups");
    if (m == n) return SCS_OK;  // nothing merged: the caller keeps using W
    if ((rc = exclusive_scan(ctx, m, cursor, gptr))) return rc;
    SCS_CUDA(ctx, cudaMemsetAsync(cursor, 0, sizeof(int32_t) * n, ctx->stream));
    fill_members<<<blocks, 256, 0, ctx->stream>>>(n, group, gptr, cursor, members);
    SCS_LAUNCHED(ctx, "fill_members");

    const size_t budget = ctx->smem_optin > 8192 ? ctx->smem_optin - 4096 : 44 * 1024;
    const int max_cols = static_cast<int>(budget / sizeof(unsigned long long));
    int nchunks = ceil_div(m, max_cols);
    const int cols_per_chunk = ceil_div(m, nchunks);
    nchunks = ceil_div(m, cols_per_chunk);
    const size_t smem = static_cast<size_t>(cols_per_chunk) * sizeof(unsigned long long);
    double *part;
    if ((rc = reserve_as(ctx, SLOT_DEGREE_PART, static_cast<size_t>(m) * nchunks, &part))) return rc;
    // always the same (maximal) opt-in size: contexts on other host threads launch this kernel concurrently
    if (!ctx->contract_configured) {
        SCS_CUDA(ctx, cudaFuncSetAttribute(contract_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(budget)));
        ctx->contract_configured = true;
    }
    WRows where;
    int A0 = 0, A1 = m;
    if (rows.sharded()) {
        // W (the argument) is ignored: row u is in the window of rank u / rows_per_rank(n)
        const ShardState &sh = ctx->shard;
        for (int r = 0; r < sh.world; ++r) where.block[r] = reinterpret_cast<const double *>(sh.peer[r] + sh.layout.W);
        where.rows_per_rank = shard_rows_per_rank(n, sh.world);
        A0 = rows.row0;
        A1 = rows.row1;
    } else {
        where.block[0] = W;
        where.rows_per_rank = n;
    }
    if (A1 <= A0) return SCS_OK;
    contract_rows<<<dim3(A1 - A0, nchunks), kThreads, smem, ctx->stream>>>(n, m, A0, words, cols_per_chunk, where,
                                                                           adj_bits, group, gptr, members, Wc, part);
    SCS_LAUNCHED(ctx, "contract_rows");
    if (degree_c) {
        sum_parts<<<ceil_div(A1 - A0, 256), 256, 0, ctx->stream>>>(m, A0, A1, nchunks, part, degree_c);
        SCS_LAUNCHED(ctx, "sum_parts");
    }
    return SCS_OK;
}

int contract(scs_ctx *ctx, int n, const double *W, const uint32_t *adj_bits, const uint32_t *max_bits,
             int32_t *group, int32_t *m_host, double *Wc, double *degree_c) {
    if (n <= 0 || !W || !adj_bits || !max_bits || !group || !m_host || !Wc)
        return fail(ctx, SCS_ERR_INVALID, "contract: bad argument");
    int32_t *label;
    int rc;
    if ((rc = reserve_as(ctx, SLOT_LABEL2, static_cast<size_t>(n), &label))) return rc;
    int m = 0;
    if ((rc = components(ctx, n, max_bits, label, &m))) return rc;  // synchronises: m is known
    *m_host = m;
    return contract_with_labels(ctx, n, W, adj_bits, label, m, group, Wc, degree_c);
}

}  // namespace scs
