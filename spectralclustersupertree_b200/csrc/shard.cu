// One recursion node spread over the GPUs of one NVLink/NVSwitch box (one process per GPU).
//
// rows of W are independent both in the graph build (every row is its own sum over the
// trees, scs.py:569) and in the Laplacian matvec, so a large node is ROW-SHARDED: rank r builds rows
// [r * ceil(n / G), (r + 1) * ceil(n / G)) of W from the (replicated) leaf tours -- no reduction over trees,
// so the fixed tree-order summation and with it bit-exactness survive, unlike a tree-sharded all-reduce --
// and keeps only that block (n^2 / G doubles).  Every row computes only the cyclic half window of columns after
// its own, so each pair of the node is visited once over all ranks; the other half of a row is fetched, as the
// mirror image of what the owners of those rows computed, out of their W blocks (pcg.cu: RowBlock::half_window,
// pcg_fetch_transposed).  What the other ranks need is pushed into their exchange
// windows over NVLink (shard.cuh): the adjacency / max-graph bit rows and degrees after the build (the
// components and contraction groups are then computed redundantly, they are cheap and deterministic),
// and one slice of the iterate per Lanczos step, written by the matvec kernel itself
// (matvec_rows_allgather in spectral.cu: compute + all-gather + barrier in one launch).  Contraction
// reads the member rows of W it needs straight out of the owners' windows.  The Lanczos vector work is
// replicated: it is L2-resident and identical on every rank, so no further exchange is needed and all
// ranks end with the same partition, bit for bit the single-GPU one.
//
// Exchange window of one rank (byte offsets in ShardLayout, sized for the largest node n_max):
//   header (flags, error word) | vec[2][n_max] | degree[n_max] | degree_c[n_max] |
//   adj_bits[n_max][words] | max_bits[n_max][words] | W block [ceil(n_max / G)][n_max]
// The window is plain cudaMalloc memory exported with cudaIpcGetMemHandle; peers in other processes map
// it with cudaIpcOpenMemHandle, peers in the same process (tests) use the pointer directly.

#include "shard.cuh"

#include <cmath>

namespace scs {

namespace {

constexpr size_t kHeaderBytes = 4096;

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

ShardLayout make_layout(int n_max, int world) {
    ShardLayout lay;
    const size_t n = static_cast<size_t>(n_max);
    const size_t words = static_cast<size_t>(scs_bit_words(n_max));
    size_t at = kHeaderBytes;
    auto take = [&](size_t bytes) {
        const size_t here = at;
        at = align_up(at + bytes, 256);
        return here;
    };
    lay.vec[0] = take(8 * n);
    lay.vec[1] = take(8 * n);
    lay.degree = take(8 * n);
    lay.degree_c = take(8 * n);
    lay.adj_bits = take(4 * n * words);
    lay.max_bits = take(4 * n * words);
    lay.W = take(8 * static_cast<size_t>(shard_rows_per_rank(n_max, world)) * n);
    lay.total = at;
    return lay;
}

PeerTable peer_table(const ShardState &sh) {
    PeerTable peers;
    for (int r = 0; r < kMaxPeers; ++r) peers.window[r] = sh.peer[r];
    peers.rank = sh.rank;
    peers.world = sh.world;
    return peers;
}

unsigned long long timeout_ns(const ShardState &sh) {
    return static_cast<unsigned long long>(sh.timeout_s * 1e9);
}

__global__ void shard_barrier_kernel(const PeerTable peers, unsigned long long epoch, unsigned long long timeout) {
    peer_signal_and_wait(peers, epoch, timeout);
}

__global__ void relabel_by_rank(int n, const int32_t *__restrict__ label, const int32_t *__restrict__ rank,
                                int32_t *__restrict__ part) {
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v < n) part[v] = rank[label[v]];
}

__global__ void flag_roots(int n, const int32_t *__restrict__ label, int32_t *__restrict__ flag) {
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v < n) flag[v] = label[v] == v;
}

__global__ void sides_of_groups(int n, const int32_t *__restrict__ group, const int32_t *__restrict__ side,
                                int32_t *__restrict__ part) {
    int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v < n) part[v] = side[group ? group[v] : v];
}

void release_window(scs_ctx *ctx) {
    ShardState &sh = ctx->shard;
    for (int r = 0; r < kMaxPeers; ++r) {
        if (sh.opened[r] && sh.peer[r]) cudaIpcCloseMemHandle(sh.peer[r]);
        sh.peer[r] = nullptr;
        sh.opened[r] = false;
    }
    if (sh.window) cudaFree(sh.window);
    cudaGetLastError();
    const double timeout = sh.timeout_s;
    const int min_n = sh.min_n;
    sh = ShardState();
    sh.timeout_s = timeout;
    sh.min_n = min_n;
}

}  // namespace

bool shard_applies(const scs_ctx *ctx, int n) {
    const ShardState &sh = ctx->shard;
    return sh.connected && sh.engaged && sh.world > 1 && n >= sh.min_n && n <= sh.n_max;
}

int shard_barrier(scs_ctx *ctx) {
    ShardState &sh = ctx->shard;
    sh.epoch += 1;
    shard_barrier_kernel<<<1, 32, 0, ctx->stream>>>(peer_table(sh), sh.epoch, timeout_ns(sh));
    SCS_LAUNCHED(ctx, "shard_barrier_kernel");
    return SCS_OK;
}

int shard_push(scs_ctx *ctx, size_t offset, size_t bytes) {
    ShardState &sh = ctx->shard;
    if (bytes == 0) return SCS_OK;
    for (int step = 1; step < sh.world; ++step) {
        const int p = (sh.rank + step) % sh.world;  // staggered so that the ranks do not all hit one peer
        SCS_CUDA(ctx, cudaMemcpyAsync(sh.peer[p] + offset, sh.window + offset, bytes, cudaMemcpyDeviceToDevice,
                                      ctx->stream));
    }
    return SCS_OK;
}

ShardMatvecTicket shard_next_matvec(scs_ctx *ctx) {
    ShardState &sh = ctx->shard;
    sh.epoch += 1;
    sh.parity ^= 1;
    ShardMatvecTicket t;
    t.epoch = sh.epoch;
    t.vec_offset = sh.layout.vec[sh.parity];
    t.timeout_ns = timeout_ns(sh);
    return t;
}

// The block scs.py:110-134 for one node, cooperatively: every rank calls this with the same arguments.
int node_split_sharded(scs_ctx *ctx, int n, int T, int64_t L, const int64_t *leaf_offsets, const int32_t *leaf_taxon,
                       const int32_t *adj_depth, const double *adj_val, const int32_t *root_depth,
                       const double *tree_weight, int contract_edges, uint64_t seed, int32_t *part_dev,
                       int32_t *part_host, scs_node_stats *stats) {
    ShardState &sh = ctx->shard;
    const ShardLayout &lay = sh.layout;
    const int words = scs_bit_words(n);
    RowBlock rows = shard_block(n, sh.rank, sh.world);
    const int nrows = rows.row1 - rows.row0;
    double *W = reinterpret_cast<double *>(sh.window + lay.W);
    double *degree = reinterpret_cast<double *>(sh.window + lay.degree);
    double *degree_c = reinterpret_cast<double *>(sh.window + lay.degree_c);
    uint32_t *adj_bits = reinterpret_cast<uint32_t *>(sh.window + lay.adj_bits);
    uint32_t *max_bits = reinterpret_cast<uint32_t *>(sh.window + lay.max_bits);
    ShardHeader *header = reinterpret_cast<ShardHeader *>(sh.window);
    int32_t *occ, *label, *label2 = nullptr, *group, *side, *scalars;
    double *Wc;
    int rc;
    if ((rc = reserve_as(ctx, SLOT_OCC, static_cast<size_t>(n), &occ))) return rc;
    if ((rc = reserve_as(ctx, SLOT_LABEL, static_cast<size_t>(n), &label))) return rc;
    if ((rc = reserve_as(ctx, SLOT_SCALARS, 64, &scalars))) return rc;
    void *pin_v;
    if ((rc = reserve_pinned(ctx, 512, &pin_v))) return rc;
    unsigned char *pin = static_cast<unsigned char *>(pin_v);
    ctx->last_n = n;
    ctx->last_m = 0;
    sh.nodes += 1;
    auto fetch_part = [&]() -> int {
        if (part_host) {
            SCS_CUDA(ctx, cudaMemcpyAsync(part_host, part_dev, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, ctx->stream));
            ctx->d2h_bytes += static_cast<int64_t>(sizeof(int32_t)) * n;
        }
        SCS_CUDA(ctx, cudaMemcpyAsync(pin + 384, &header->error, sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->stream));
        SCS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (*reinterpret_cast<unsigned int *>(pin + 384) != 0)
            return fail(ctx, SCS_ERR_PEER, "sharded node: a wait for a peer GPU timed out");
        return SCS_OK;
    };

    // nobody may still be reading this rank's window (W rows, bits) from the previous node
    if ((rc = shard_barrier(ctx))) return rc;
    // Every pair of the node once over all ranks: a row computes the cyclic half window of columns after its own
    // (pcg.cu, RowBlock::half_window); the other half of a row is the mirror image of what the owners of those rows
    // computed, fetched from their blocks once everybody is that far.  (scs_ctx_set_full_rows: every row in full.)
    rows.half_window = !ctx->full_rows;
    if ((rc = pcg_build(ctx, n, T, L, leaf_offsets, leaf_taxon, adj_depth, adj_val, root_depth, tree_weight, W, nullptr,
                        occ, adj_bits, contract_edges ? max_bits : nullptr, degree, rows)))
        return rc;
    const size_t row_bits = static_cast<size_t>(words) * sizeof(uint32_t);
    if ((rc = shard_push(ctx, lay.adj_bits + rows.row0 * row_bits, nrows * row_bits))) return rc;
    if (contract_edges && (rc = shard_push(ctx, lay.max_bits + rows.row0 * row_bits, nrows * row_bits))) return rc;
    if (rows.half_window) {
        if ((rc = shard_barrier(ctx))) return rc;  // every rank's half rows and bit rows are in place
        const double *peer_W[kMaxPeers] = {};
        for (int r = 0; r < sh.world; ++r) peer_W[r] = reinterpret_cast<const double *>(sh.peer[r] + lay.W);
        if ((rc = pcg_fetch_transposed(ctx, n, rows, shard_rows_per_rank(n, sh.world), sh.world, peer_W, W))) return rc;
        if ((rc = pcg_symmetrize_bits(ctx, n, adj_bits, contract_edges ? max_bits : nullptr))) return rc;
        if ((rc = pcg_degree_block(ctx, n, T, rows, W, degree))) return rc;
    }
    if ((rc = shard_push(ctx, lay.degree + sizeof(double) * rows.row0, sizeof(double) * nrows))) return rc;
    if ((rc = shard_barrier(ctx))) return rc;

    // replicated: components of the graph and of the max-graph, one round trip for the counts
    if ((rc = components_async(ctx, n, adj_bits, label, scalars + 8))) return rc;
    if (contract_edges) {
        if ((rc = reserve_as(ctx, SLOT_LABEL2, static_cast<size_t>(n), &label2))) return rc;
        if ((rc = components_async(ctx, n, max_bits, label2, scalars + 9))) return rc;
    }
    SCS_CUDA(ctx, cudaMemcpyAsync(pin + 256, scalars, 16 * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    SCS_CUDA(ctx, cudaMemcpyAsync(pin + 384, &header->error, sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->stream));
    SCS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (*reinterpret_cast<unsigned int *>(pin + 384) != 0)
        return fail(ctx, SCS_ERR_PEER, "sharded node: a wait for a peer GPU timed out");
    const int32_t *host_scalars = reinterpret_cast<const int32_t *>(pin + 256);
    if (host_scalars[0] != 0) return fail(ctx, SCS_ERR_INPUT, "leaf tour: taxon id out of range");
    const int ncomp = host_scalars[8];
    stats->n_components = ncomp;
    stats->contracted_size = n;
    const int blocks = ceil_div(n, 256);

    if (ncomp != 1) {
        int32_t *flag, *rank;
        if ((rc = reserve_as(ctx, SLOT_GROUP_PTR, 4 * static_cast<size_t>(n) + 8, &flag))) return rc;
        rank = flag + n;
        flag_roots<<<blocks, 256, 0, ctx->stream>>>(n, label, flag);
        SCS_LAUNCHED(ctx, "flag_roots");
        if ((rc = exclusive_scan(ctx, n, flag, rank))) return rc;
        relabel_by_rank<<<blocks, 256, 0, ctx->stream>>>(n, label, rank, part_dev);
        SCS_LAUNCHED(ctx, "relabel_by_rank");
        return fetch_part();
    }

    int m = n;
    const double *Wm = W;
    const double *deg_m = degree;
    const int32_t *group_m = nullptr;
    RowBlock rows_m = rows;
    if (contract_edges) {
        m = host_scalars[9];
        if ((rc = reserve_as(ctx, SLOT_GROUP, static_cast<size_t>(n), &group))) return rc;
        if (m != n) {
            // a contracted graph too small to share out is finished on every rank by itself
            const bool share = m >= sh.min_n / 2 && m >= 2 * sh.world;
            rows_m = share ? shard_block(m, sh.rank, sh.world) : RowBlock{0, m};
            const size_t wc_rows = static_cast<size_t>(rows_m.row1 - rows_m.row0);
            if ((rc = reserve_as(ctx, SLOT_WC, wc_rows * m, &Wc))) return rc;
            double *deg_out = degree_c;
            if (!share && (rc = reserve_as(ctx, SLOT_DEGREE_C, static_cast<size_t>(m), &deg_out))) return rc;
            if ((rc = contract_with_labels(ctx, n, nullptr, adj_bits, label2, m, group, Wc, deg_out, rows_m))) return rc;
            if (share) {
                if ((rc = shard_push(ctx, lay.degree_c + sizeof(double) * rows_m.row0,
                                     sizeof(double) * (rows_m.row1 - rows_m.row0))))
                    return rc;
                if ((rc = shard_barrier(ctx))) return rc;
            } else {
                rows_m = RowBlock();
            }
            Wm = Wc;
            deg_m = deg_out;
            group_m = group;
        } else if ((rc = contract_with_labels(ctx, n, nullptr, adj_bits, label2, m, group, nullptr, nullptr, rows))) {
            return rc;
        }
    }
    ctx->last_m = m;
    stats->contracted_size = m;
    stats->spectral_ran = 1;
    if ((rc = reserve_as(ctx, SLOT_SIDE, static_cast<size_t>(n), &side))) return rc;
    if (m < 2) return fail(ctx, SCS_ERR_TOO_SMALL, "spectral step on a graph contracted to one vertex");
    if ((rc = spectral_bipartition(ctx, m, Wm, deg_m, seed, side, stats, rows_m))) return rc;
    sides_of_groups<<<blocks, 256, 0, ctx->stream>>>(n, group_m, side, part_dev);
    SCS_LAUNCHED(ctx, "sides_of_groups");
    return fetch_part();
}

}  // namespace scs

using namespace scs;

extern "C" {

int scs_shard_create(scs_ctx *ctx, int rank, int world, int n_max, unsigned char *handle_out) {
    if (!ctx || world < 1 || world > kMaxPeers || rank < 0 || rank >= world || n_max < 2)
        return ctx ? fail(ctx, SCS_ERR_INVALID, "shard_create: bad argument") : SCS_ERR_INVALID;
    static_assert(sizeof(cudaIpcMemHandle_t) == SCS_IPC_HANDLE_BYTES, "handle size");
    cudaSetDevice(ctx->device);
    if (ctx->shard.window) {
        SCS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        release_window(ctx);
    }
    ShardState &sh = ctx->shard;
    sh.rank = rank;
    sh.world = world;
    sh.n_max = n_max;
    sh.layout = make_layout(n_max, world);
    void *window = nullptr;
    SCS_CUDA(ctx, cudaMalloc(&window, sh.layout.total));
    sh.window = static_cast<unsigned char *>(window);
    SCS_CUDA(ctx, cudaMemset(sh.window, 0, kHeaderBytes));
    SCS_CUDA(ctx, cudaDeviceSynchronize());
    if (handle_out) {
        cudaIpcMemHandle_t handle;
        SCS_CUDA(ctx, cudaIpcGetMemHandle(&handle, sh.window));
        std::memcpy(handle_out, &handle, sizeof(handle));
    }
    return SCS_OK;
}

int scs_shard_window(scs_ctx *ctx, void **window_dev, size_t *bytes) {
    if (!ctx || !ctx->shard.window) return SCS_ERR_INVALID;
    if (window_dev) *window_dev = ctx->shard.window;
    if (bytes) *bytes = ctx->shard.layout.total;
    return SCS_OK;
}

int scs_shard_connect_ipc(scs_ctx *ctx, const unsigned char *handles) {
    if (!ctx || !handles || !ctx->shard.window) return SCS_ERR_INVALID;
    cudaSetDevice(ctx->device);
    ShardState &sh = ctx->shard;
    for (int r = 0; r < sh.world; ++r) {
        if (r == sh.rank) {
            sh.peer[r] = sh.window;
            continue;
        }
        cudaIpcMemHandle_t handle;
        std::memcpy(&handle, handles + static_cast<size_t>(r) * SCS_IPC_HANDLE_BYTES, sizeof(handle));
        void *mapped = nullptr;
        SCS_CUDA(ctx, cudaIpcOpenMemHandle(&mapped, handle, cudaIpcMemLazyEnablePeerAccess));
        sh.peer[r] = static_cast<unsigned char *>(mapped);
        sh.opened[r] = true;
    }
    sh.connected = true;
    return SCS_OK;
}

int scs_shard_connect_ptrs(scs_ctx *ctx, void *const *windows) {
    if (!ctx || !windows || !ctx->shard.window) return SCS_ERR_INVALID;
    cudaSetDevice(ctx->device);
    ShardState &sh = ctx->shard;
    for (int r = 0; r < sh.world; ++r) {
        if (!windows[r]) return fail(ctx, SCS_ERR_INVALID, "shard_connect: null window");
        if (r != sh.rank) {
            cudaPointerAttributes attr;
            SCS_CUDA(ctx, cudaPointerGetAttributes(&attr, windows[r]));
            if (attr.device != ctx->device) {
                const cudaError_t err = cudaDeviceEnablePeerAccess(attr.device, 0);
                if (err != cudaSuccess && err != cudaErrorPeerAccessAlreadyEnabled)
                    return fail(ctx, SCS_ERR_CUDA, "cudaDeviceEnablePeerAccess", err);
                cudaGetLastError();
            }
        }
        sh.peer[r] = r == sh.rank ? sh.window : static_cast<unsigned char *>(windows[r]);
    }
    sh.connected = true;
    return SCS_OK;
}

int scs_shard_engage(scs_ctx *ctx, int on) {
    if (!ctx) return SCS_ERR_INVALID;
    if (on && !ctx->shard.connected) return fail(ctx, SCS_ERR_INVALID, "shard_engage: peers are not connected");
    ctx->shard.engaged = on != 0;
    return SCS_OK;
}

int scs_shard_configure(scs_ctx *ctx, int min_n, double timeout_seconds) {
    if (!ctx) return SCS_ERR_INVALID;
    if (min_n > 0) ctx->shard.min_n = min_n < 2 * kMaxPeers ? 2 * kMaxPeers : min_n;
    if (timeout_seconds > 0.0) ctx->shard.timeout_s = timeout_seconds;
    return SCS_OK;
}

int64_t scs_shard_nodes(const scs_ctx *ctx) { return ctx ? ctx->shard.nodes : 0; }

int scs_shard_barrier(scs_ctx *ctx) {
    if (!ctx || !ctx->shard.connected) return SCS_ERR_INVALID;
    cudaSetDevice(ctx->device);
    int rc = shard_barrier(ctx);
    if (rc) return rc;
    unsigned int error = 0;
    SCS_CUDA(ctx, cudaMemcpyAsync(&error, &reinterpret_cast<ShardHeader *>(ctx->shard.window)->error, sizeof(error),
                                  cudaMemcpyDeviceToHost, ctx->stream));
    SCS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (error) return fail(ctx, SCS_ERR_PEER, "shard barrier: a wait for a peer GPU timed out");
    return SCS_OK;
}

int scs_shard_destroy(scs_ctx *ctx) {
    if (!ctx) return SCS_ERR_INVALID;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    release_window(ctx);
    return SCS_OK;
}

}  // extern "C"
