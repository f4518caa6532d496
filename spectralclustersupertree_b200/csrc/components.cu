// Connected components of a graph given as a bit adjacency matrix.
//
// Replaces _get_graph_components of the reference (/root/reference/src/sc_supertree/scs.py:458-492).
// The reference explores the `edges` sets, i.e. "the pair co-occurred in some tree", which is the
// adjacency bit matrix the row kernel emits (C > 0), not W > 0.
//
// Lock-free union-find: a warp takes a row and hooks the row's vertex to every neighbour with
// a larger id.  Roots are always hooked larger-under-smaller, so the final root of a component
// is its smallest vertex id whatever the interleaving -- labels are deterministic.

#include "common.cuh"

namespace scs {

namespace {

__device__ __forceinline__ int uf_find(volatile int32_t *parent, int x) {
    int r = x;
    while (true) {
        int up = parent[r];
        if (up == r) break;
        r = up;
    }
    // path halving towards the root we found (benign races: parents only ever decrease)
    while (true) {
        int up = parent[x];
        if (up <= r) break;
        parent[x] = r;
        x = up;
    }
    return r;
}

__device__ __forceinline__ void uf_union(int32_t *parent, int a, int b) {
    while (true) {
        int ra = uf_find(parent, a);
        int rb = uf_find(parent, b);
        if (ra == rb) return;
        if (ra > rb) { int t = ra; ra = rb; rb = t; }
        // hook the larger root under the smaller one
        int old = atomicCAS(&parent[rb], rb, ra);
        if (old == rb) return;
        a = ra;
        b = rb;
    }
}

__global__ void uf_init(int n, int32_t *parent) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) parent[i] = i;
}

// one warp per row; lanes take words of the row
__global__ void uf_hook_rows(int n, int words, const uint32_t *__restrict__ bits, int32_t *parent) {
    const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp_global >= n) return;
    const int a = warp_global;
    const uint32_t *row = bits + static_cast<size_t>(a) * words;
    for (int j = lane; j < words; j += 32) {
        uint32_t wbits = row[j];
        // only neighbours with a larger id: the matrix is symmetric
        const int first = j << 5;
        if (first + 31 <= a) continue;
        if (first <= a) wbits &= (a - first == 31) ? 0u : (~0u << (a - first + 1));
        while (wbits) {
            const int b = first + __ffs(wbits) - 1;
            wbits &= wbits - 1;
            if (b < n) uf_union(parent, a, b);
        }
    }
}

__global__ void uf_flatten(int n, int32_t *parent, int32_t *label, int32_t *n_roots) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int r = i;
    while (parent[r] != r) r = parent[r];
    label[i] = r;
    if (r == i) atomicAdd(n_roots, 1);
}

}  // namespace

// Enqueue only: label[v] = smallest vertex of v's component, *count_dev = number of components.
int components_async(scs_ctx *ctx, int n, const uint32_t *bits, int32_t *label, int32_t *count_dev) {
    if (n <= 0 || !bits || !label || !count_dev) return fail(ctx, SCS_ERR_INVALID, "components: bad argument");
    const int words = scs_bit_words(n);
    int32_t *parent;
    int rc;
    // the union-find forest is built in `label` itself, then flattened in place
    if ((rc = reserve_as(ctx, SLOT_UF_PARENT, static_cast<size_t>(n), &parent))) return rc;
    SCS_CUDA(ctx, cudaMemsetAsync(count_dev, 0, sizeof(int32_t), ctx->stream));
    uf_init<<<ceil_div(n, 256), 256, 0, ctx->stream>>>(n, parent);
    SCS_LAUNCHED(ctx, "uf_init");
    uf_hook_rows<<<ceil_div(static_cast<int64_t>(n) * 32, 256), 256, 0, ctx->stream>>>(n, words, bits, parent);
    SCS_LAUNCHED(ctx, "uf_hook_rows");
    uf_flatten<<<ceil_div(n, 256), 256, 0, ctx->stream>>>(n, parent, label, count_dev);
    SCS_LAUNCHED(ctx, "uf_flatten");
    return SCS_OK;
}

int components(scs_ctx *ctx, int n, const uint32_t *bits, int32_t *label, int32_t *n_components_host) {
    int32_t *scalars;
    int rc;
    if ((rc = reserve_as(ctx, SLOT_SCALARS, 64, &scalars))) return rc;
    int32_t *n_roots = scalars + 8;
    if ((rc = components_async(ctx, n, bits, label, n_roots))) return rc;
    if (n_components_host) {
        void *pin;
        if ((rc = reserve_pinned(ctx, 64, &pin))) return rc;
        SCS_CUDA(ctx, cudaMemcpyAsync(pin, n_roots, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
        SCS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        *n_components_host = *static_cast<int32_t *>(pin);
    }
    return SCS_OK;
}

}  // namespace scs
