// Device-resident source trees: upload, leaf tours, restriction to the children of a wave (see devforest.cuh).
//
// Restriction.  The reference restricts every source tree to the taxa of a component with
// PhyloNode.get_sub_tree(names, ignore_missing=True, as_rooted=True) (/root/reference/src/sc_supertree/scs.py:444-453);
// forest.cpp does the same on flat arrays on the host, one linear pass per (tree, child).  Here a WARP owns one
// (tree, child) pair and the passes are prefix sums over the tree's pre-order node array:
//   kept[k]     tip whose taxon belongs to the child;           P = exclusive prefix sum of kept
//   cnt[k]      kept tips below k = P[k + size[k]] - P[k]       (pre-order: a subtree is a contiguous run)
//   unary[k]    some child of k carries all of cnt[k]           (set by that child: cnt[child] == cnt[k])
//   retained[k] cnt[k] > 0 and (tip or not unary[k]);           R = exclusive prefix sum of retained = new index
// A retained node then walks up its chain of merged (unary) ancestors, folding their lengths bottom-up in the
// reference's operand order  length(node) + length(child)  -- the same additions in the same order as forest.cpp and
// tree.py, so branch-length weights stay bit-identical at every recursion level.  The first retained node in
// pre-order is the new root and loses its length.  Which (tree, child) pairs exist (>= 2 kept tips, scs.py:447-448)
// comes from a histogram of the tips over the children, scanned into a dense pair list ordered by child, then tree:
// the new forest is laid out by child with its trees in source order (the graph build sums W in tree input order).

#include "devforest.cuh"

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdlib>
#include <vector>

#include <omp.h>

#include "forest.hpp"
#include "shard.cuh"

namespace scs {

int grow(scs_ctx *ctx, GrowBuf &buf, size_t bytes) {
    if (bytes == 0) bytes = 16;
    if (buf.bytes >= bytes) return SCS_OK;
    const size_t want = bytes + bytes / 4 + 256;
    if (buf.ptr) {
        SCS_CUDA(ctx, cudaFreeAsync(buf.ptr, ctx->stream));
        buf.ptr = nullptr;
        buf.bytes = 0;
    }
    const cudaError_t err = cudaMallocAsync(&buf.ptr, want, ctx->stream);
    if (err != cudaSuccess) {
        buf.ptr = nullptr;
        return fail(ctx, SCS_ERR_CUDA, "cudaMallocAsync (device forest)", err);
    }
    buf.bytes = want;
    return SCS_OK;
}

void release(scs_ctx *ctx, GrowBuf &buf) {
    if (buf.ptr) cudaFreeAsync(buf.ptr, ctx->stream);
    buf.ptr = nullptr;
    buf.bytes = 0;
}

void DevForest::free_all(scs_ctx *ctx) {
    for (GrowBuf *b : {&tree_off, &leaf_off, &parent, &size, &taxon, &length, &support, &weight, &tree_job}) release(ctx, *b);
    trees = nodes = leaves = 0;
}

void DevTours::free_all(scs_ctx *ctx) {
    for (GrowBuf *b : {&leaf_taxon, &adj_depth, &root_depth, &adj_val, &depth_s, &val_s}) release(ctx, *b);
}

namespace {

constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ int warp_excl_scan(int v, int lane, int &total) {
    int inc = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int o = __shfl_up_sync(kFull, inc, off);
        if (lane >= off) inc += o;
    }
    total = __shfl_sync(kFull, inc, 31);
    return inc - v;
}

// ---- tours (scs.py:555-567, 628: the value handed top-down; forest.cpp:scs_forest_tours) -----------------------------
// One warp per tree.  Pass A: depth and weighting value of every internal node, top-down, 32 pre-order nodes at a
// time; a node whose parent sits in the same chunk waits for it (parents come first in pre-order, so the smallest
// waiting node is always ready: at most 32 rounds, a few in practice).  Every value is  value(parent) + own  in that
// operand order: the sums of forest.cpp, bit for bit.  Pass B: every tip writes its tour entry.
__global__ void __launch_bounds__(256)
df_tours(int64_t T, const int64_t *__restrict__ tree_off, const int64_t *__restrict__ leaf_off,
         const int32_t *__restrict__ parent, const int32_t *__restrict__ taxon, const double *__restrict__ length,
         const double *__restrict__ support, int weighting, const int32_t *__restrict__ taxon_vertex,
         int32_t *depth_s, double *val_s, int32_t *__restrict__ leaf_taxon, int32_t *__restrict__ adj_depth,
         double *__restrict__ adj_val, int32_t *__restrict__ root_depth, int32_t *__restrict__ flags) {
    const int64_t t = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (t >= T) return;
    const int64_t base = tree_off[t];
    const int count = static_cast<int>(tree_off[t + 1] - base);
    if (lane == 0) root_depth[t] = 0;
    if (count <= 1) return;  // a lone tip has no sides (scs.py:570)
    const int32_t *par = parent + base;
    const int32_t *tax = taxon + base;
    for (int c0 = 0; c0 < count; c0 += 32) {
        const int k = c0 + lane;
        const bool valid = k < count;
        const bool internal = valid && tax[k] < 0;
        const int p = internal && k > 0 ? par[k] : 0;
        double own = 0.0;
        if (internal && k > 0) {
            if (weighting == 1) {
                const double len = length[base + k];
                own = isnan(len) ? 1.0 : len;
            } else if (weighting == 3) {
                own = support[base + k];
            }
        }
        double v = 0.0;  // the value handed to the root's children is 0 (scs.py:577)
        int d = 0;
        bool pending = internal && k > 0;
        auto value_of = [&](double above) {
            switch (weighting) {
            case 0: return 1.0;
            case 1: return above + own;
            case 2: return above + 1.0;
            default: return own;
            }
        };
        if (pending && p < c0) {
            v = value_of(val_s[base + p]);
            d = depth_s[base + p] + 1;
            pending = false;
        }
        unsigned waiting = __ballot_sync(kFull, pending);
        while (waiting) {
            const int src = pending ? p - c0 : 0;
            const double above = __shfl_sync(kFull, v, src);
            const int dabove = __shfl_sync(kFull, d, src);
            if (pending && !((waiting >> src) & 1u)) {
                v = value_of(above);
                d = dabove + 1;
                pending = false;
            }
            waiting = __ballot_sync(kFull, pending);
        }
        if (internal) {
            val_s[base + k] = v;
            depth_s[base + k] = d;
        }
        __syncwarp();
    }
    const int64_t o = leaf_off[t];
    int tip_base = 0;
    for (int c0 = 0; c0 < count; c0 += 32) {
        const int k = c0 + lane;
        const bool tip = k < count && tax[k] >= 0;
        const unsigned tips = __ballot_sync(kFull, tip);
        if (tip) {
            const int64_t pos = o + tip_base + __popc(tips & ((1u << lane) - 1u));
            leaf_taxon[pos] = taxon_vertex[tax[k]];
            if (k + 1 < count) {
                const int lca = par[k + 1];  // the next pre-order node hangs off the LCA with the next tip
                const double x = val_s[base + lca];
                adj_depth[pos] = depth_s[base + lca];
                adj_val[pos] = x;
                if (lca != 0 && isnan(x)) flags[1] = 1;  // bootstrap weighting without a support value
            } else {
                adj_depth[pos] = -1;
                adj_val[pos] = 0.0;
            }
        }
        tip_base += __popc(tips);
    }
}

// ---- restriction ---------------------------------------------------------------------------------------------------
struct RestrictPlan {
    const int32_t *job_tree_begin;  // [J + 1]
    const int32_t *job_parts;       // [J]
    const int32_t *job_part_base;   // [J]
    const int32_t *job_cell_base;   // [J + 1]: cells of job j = parts x trees, cell = base + part * trees + tree
    const int32_t *part_newjob;     // [parts]
    const int32_t *owner;           // [taxa]
    int J;
};

// tips of every tree counted per part of the tree's job; one warp per tree
__global__ void __launch_bounds__(256)
df_count(int64_t T, const int64_t *__restrict__ tree_off, const int32_t *__restrict__ taxon,
         const int32_t *__restrict__ tree_job, RestrictPlan plan, int32_t *__restrict__ hist) {
    const int64_t t = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (t >= T) return;
    const int j = tree_job[t];
    const int parts = plan.job_parts[j];
    if (parts == 0) return;
    const int trees = plan.job_tree_begin[j + 1] - plan.job_tree_begin[j];
    const int part0 = plan.job_part_base[j];
    const int64_t cell0 = static_cast<int64_t>(plan.job_cell_base[j]) + (t - plan.job_tree_begin[j]);
    const int64_t base = tree_off[t];
    const int count = static_cast<int>(tree_off[t + 1] - base);
    if (count < 3) return;  // fewer than two tips
    for (int k = lane; k < count; k += 32) {
        const int x = taxon[base + k];
        if (x < 0) continue;
        const int part = plan.owner[x] - part0;
        if (part >= 0 && part < parts) atomicAdd(&hist[cell0 + static_cast<int64_t>(part) * trees], 1);
    }
}

__global__ void df_flag_cells(int cells, const int32_t *__restrict__ hist, int32_t *__restrict__ flag) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < cells) flag[c] = hist[c] >= 2;  // scs.py:447-448: a tree with fewer than two tips of the part is dropped
}

// one entry per (tree, part) pair that yields a restricted tree, in cell order = by new job, trees in source order
__global__ void df_make_pairs(int cells, const int32_t *__restrict__ hist, const int32_t *__restrict__ pair_index,
                              const int64_t *__restrict__ tree_off, RestrictPlan plan, int32_t *__restrict__ pair_tree,
                              int32_t *__restrict__ pair_part, int32_t *__restrict__ pair_tips,
                              int32_t *__restrict__ pair_need) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cells || hist[c] < 2) return;
    int lo = 0, hi = plan.J;  // last job with cell_base <= c (jobs without cells share a base with their successor)
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (plan.job_cell_base[mid] <= c) lo = mid; else hi = mid;
    }
    const int j = lo;
    const int trees = plan.job_tree_begin[j + 1] - plan.job_tree_begin[j];
    const int within = c - plan.job_cell_base[j];
    const int part = within / trees, tl = within - part * trees;
    const int t = plan.job_tree_begin[j] + tl;
    const int p = pair_index[c];
    pair_tree[p] = t;
    pair_part[p] = plan.job_part_base[j] + part;
    pair_tips[p] = hist[c];
    pair_need[p] = static_cast<int32_t>(tree_off[t + 1] - tree_off[t]) + 1;
}

// kept / retained prefix sums of one (tree, part) pair; one warp per pair
__global__ void __launch_bounds__(256)
df_mark(const int32_t *__restrict__ pair_count, const int32_t *__restrict__ pair_tree, const int32_t *__restrict__ pair_part,
        const int32_t *__restrict__ scratch_off, const int64_t *__restrict__ tree_off, const int32_t *__restrict__ parent,
        const int32_t *__restrict__ size, const int32_t *__restrict__ taxon, const int32_t *__restrict__ owner,
        int32_t *P_all, int32_t *R_all, uint8_t *U_all, int32_t *__restrict__ pair_nodes) {
    const int p = static_cast<int>((blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (p >= *pair_count) return;
    const int t = pair_tree[p], part = pair_part[p];
    const int64_t base = tree_off[t];
    const int count = static_cast<int>(tree_off[t + 1] - base);
    const int32_t *par = parent + base;
    const int32_t *sz = size + base;
    const int32_t *tax = taxon + base;
    int32_t *P = P_all + scratch_off[p];
    int32_t *R = R_all + scratch_off[p];
    uint8_t *U = U_all + scratch_off[p];
    int carry = 0;
    for (int c0 = 0; c0 < count; c0 += 32) {
        const int k = c0 + lane;
        const bool valid = k < count;
        int kept = 0;
        if (valid) {
            const int x = tax[k];
            kept = x >= 0 && owner[x] == part;
            U[k] = 0;
        }
        int total;
        const int excl = warp_excl_scan(kept, lane, total);
        if (valid) P[k] = carry + excl;
        carry += total;
    }
    if (lane == 0) P[count] = carry;
    __syncwarp();
    for (int k = 1 + lane; k < count; k += 32) {
        const int cnt = P[k + sz[k]] - P[k];
        if (cnt > 0) {
            const int up = par[k];
            if (cnt == P[up + sz[up]] - P[up]) U[up] = 1;  // the parent has all its kept tips below this child
        }
    }
    __syncwarp();
    carry = 0;
    for (int c0 = 0; c0 < count; c0 += 32) {
        const int k = c0 + lane;
        const bool valid = k < count;
        int retained = 0;
        if (valid) {
            const int cnt = P[k + sz[k]] - P[k];
            retained = cnt > 0 && (tax[k] >= 0 || !U[k]);
        }
        int total;
        const int excl = warp_excl_scan(retained, lane, total);
        if (valid) R[k] = carry + excl;
        carry += total;
    }
    if (lane == 0) {
        R[count] = carry;
        pair_nodes[p] = carry;
    }
}

__global__ void df_widen(int n, const int32_t *__restrict__ in, int64_t *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= n) out[i] = in[i];
}

// the restricted tree of one pair written into the new forest; one warp per pair
__global__ void __launch_bounds__(256)
df_write(const int32_t *__restrict__ pair_count, const int32_t *__restrict__ pair_tree, const int32_t *__restrict__ pair_part,
         const int32_t *__restrict__ pair_tips, const int32_t *__restrict__ scratch_off, const int32_t *__restrict__ node_off32,
         const int32_t *__restrict__ part_newjob, const int64_t *__restrict__ tree_off, const int32_t *__restrict__ parent,
         const int32_t *__restrict__ size, const int32_t *__restrict__ taxon, const double *__restrict__ length,
         const double *__restrict__ support, const double *__restrict__ weight, const int32_t *__restrict__ R_all,
         int32_t *__restrict__ o_parent, int32_t *__restrict__ o_size, int32_t *__restrict__ o_taxon,
         double *__restrict__ o_length, double *__restrict__ o_support, double *__restrict__ o_weight,
         int32_t *__restrict__ o_tree_job, uint8_t *__restrict__ present, int32_t *__restrict__ job_trees,
         unsigned long long *__restrict__ job_visits) {
    const int p = static_cast<int>((blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (p >= *pair_count) return;
    const int t = pair_tree[p];
    const int64_t base = tree_off[t];
    const int count = static_cast<int>(tree_off[t + 1] - base);
    const int32_t *par = parent + base;
    const int32_t *sz = size + base;
    const int32_t *tax = taxon + base;
    const int32_t *R = R_all + scratch_off[p];
    const int64_t ob = node_off32[p];
    const double nan_v = __longlong_as_double(0x7ff8000000000000ll);
    for (int k = lane; k < count; k += 32) {
        const int q = R[k];
        if (R[k + 1] == q) continue;  // not retained
        int up = -1;
        double acc = nan_v;
        if (q > 0) {
            // merged unary ancestors, bottom-up: length(node) + length(child); NaN (missing) propagates like None
            up = par[k];
            if (length) acc = length[base + k];
            while (R[up + 1] == R[up]) {
                if (length) acc = __dadd_rn(length[base + up], acc);
                up = par[up];
            }
            up = R[up];
        }
        // q == 0: the first retained node in pre-order is the new root; its own length is dropped
        o_parent[ob + q] = up;
        o_size[ob + q] = R[k + sz[k]] - q;
        o_taxon[ob + q] = tax[k];
        if (o_length) o_length[ob + q] = acc;
        if (o_support) o_support[ob + q] = support[base + k];
        if (tax[k] >= 0) present[tax[k]] = 1;  // racing writers all store 1
    }
    if (lane == 0) {
        const int job = part_newjob[pair_part[p]];
        o_weight[p] = weight[t];
        o_tree_job[p] = job;
        if (job >= 0) {  // always: taxa of parts without a job have owner -1 and never make a pair
            atomicAdd(&job_trees[job], 1);
            const unsigned long long k = static_cast<unsigned long long>(pair_tips[p]);
            atomicAdd(&job_visits[job], k * (k - 1ull));
        }
    }
}

__global__ void df_job_info(int jobs, const int32_t *__restrict__ job_trees, const int32_t *__restrict__ job_tree_begin,
                            const int32_t *__restrict__ node_off32, const int32_t *__restrict__ leaf_off32,
                            const unsigned long long *__restrict__ job_visits, DevJobInfo *__restrict__ info) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= jobs) return;
    DevJobInfo out;
    out.trees = job_trees[j];
    out.tree_begin = job_tree_begin[j];
    out.leaf_begin = leaf_off32[out.tree_begin];
    out.node_begin = node_off32[out.tree_begin];
    out.first_tree_nodes = out.trees > 0 ? node_off32[out.tree_begin + 1] - node_off32[out.tree_begin] : 0;
    out.pair_visits = static_cast<int64_t>(job_visits[j]);
    info[j] = out;
}

struct FetchItem {
    int64_t src, dst, count;
};

__global__ void df_fetch(int items, const FetchItem *__restrict__ item, const int32_t *__restrict__ parent,
                         const int32_t *__restrict__ taxon, int32_t *__restrict__ o_parent, int32_t *__restrict__ o_taxon) {
    const int i = blockIdx.x;
    if (i >= items) return;
    const FetchItem it = item[i];
    for (int64_t k = threadIdx.x; k < it.count; k += blockDim.x) {
        o_parent[it.dst + k] = parent[it.src + k];
        o_taxon[it.dst + k] = taxon[it.src + k];
    }
}

template <typename T>
T *carve(unsigned char *&cursor, size_t count) {
    T *out = reinterpret_cast<T *>(cursor);
    cursor += (count * sizeof(T) + 255) & ~static_cast<size_t>(255);
    return out;
}

}  // namespace

int devforest_upload(scs_ctx *ctx, const scs_forest *host, int weighting, DevForest *out, bool cooperative) {
    if (!host || !out) return fail(ctx, SCS_ERR_INVALID, "devforest_upload: bad argument");
    const int64_t T = host->num_trees(), M = host->node_offsets.back();
    if (M >= (1ll << 31)) return fail(ctx, SCS_ERR_INVALID, "device forest: more than 2^31 tree nodes");
    out->trees = T;
    out->nodes = M;
    out->leaves = host->leaf_offsets.back();
    out->has_length = weighting == 1;
    out->has_support = weighting == 3;
    int rc;
    const size_t nT = static_cast<size_t>(T), nM = static_cast<size_t>(M);
    if ((rc = grow(ctx, out->tree_off, (nT + 1) * sizeof(int64_t)))) return rc;
    if ((rc = grow(ctx, out->leaf_off, (nT + 1) * sizeof(int64_t)))) return rc;
    if ((rc = grow(ctx, out->parent, nM * sizeof(int32_t)))) return rc;
    if ((rc = grow(ctx, out->size, nM * sizeof(int32_t)))) return rc;
    if ((rc = grow(ctx, out->taxon, nM * sizeof(int32_t)))) return rc;
    if ((rc = grow(ctx, out->weight, nT * sizeof(double)))) return rc;
    if ((rc = grow(ctx, out->tree_job, nT * sizeof(int32_t)))) return rc;
    if (out->has_length && (rc = grow(ctx, out->length, nM * sizeof(double)))) return rc;
    if (out->has_support && (rc = grow(ctx, out->support, nM * sizeof(double)))) return rc;

    // the per-node arrays, and where each starts in a rank's staging area (the W block of its exchange window)
    struct NodeArray {
        GrowBuf *dst;
        const void *src;  // null: the subtree sizes, computed below
        size_t elem, stage_off;
    };
    std::vector<NodeArray> arrays = {{&out->parent, host->parent.data(), sizeof(int32_t), 0},
                                     {&out->size, nullptr, sizeof(int32_t), 0},
                                     {&out->taxon, host->taxon.data(), sizeof(int32_t), 0}};
    if (out->has_length) arrays.push_back({&out->length, host->length.data(), sizeof(double), 0});
    if (out->has_support) arrays.push_back({&out->support, host->support.data(), sizeof(double), 0});
    size_t stage_bytes = 0;
    for (NodeArray &a : arrays) {
        a.stage_off = stage_bytes;
        stage_bytes += (nM * a.elem + 255) & ~static_cast<size_t>(255);
    }
    ShardState &sh = ctx->shard;
    // worth it from a few million nodes (below that the barriers cost more than the link saves); tests lower the bar
    const char *min_nodes_env = std::getenv("SCS_SHARED_UPLOAD_MIN_NODES");
    const size_t min_nodes = min_nodes_env ? static_cast<size_t>(std::atoll(min_nodes_env)) : (1u << 22);
    const bool share = cooperative && sh.connected && sh.world > 1 && nM > min_nodes &&
                       sh.layout.total > sh.layout.W && stage_bytes <= sh.layout.total - sh.layout.W;
    const int G = share ? sh.world : 1, me = share ? sh.rank : 0;
    auto slice_begin = [&](int p) { return static_cast<int64_t>(M * static_cast<__int128>(p) / G); };
    const int64_t lo = slice_begin(me), hi = slice_begin(me + 1);

    // subtree sizes (pre-order: the nodes of a subtree are a contiguous run) of the trees that overlap this rank's
    // slice, per tree, over the host threads
    const int64_t *offs = host->node_offsets.data();
    const int64_t t_lo = hi > lo ? std::upper_bound(offs, offs + T + 1, lo) - offs - 1 : 0;
    const int64_t t_hi = hi > lo ? std::lower_bound(offs, offs + T + 1, hi) - offs : 0;
    const int64_t size_base = T > 0 && hi > lo ? offs[t_lo] : 0;
    std::vector<int32_t> size(static_cast<size_t>(hi > lo ? offs[t_hi] - size_base : 1));
#pragma omp parallel for schedule(dynamic, 16) num_threads(scs_host_threads()) if (M > (1 << 15))
    for (int64_t t = t_lo; t < t_hi; ++t) {
        const int64_t base = offs[t], count = offs[t + 1] - base;
        int32_t *sz = size.data() + (base - size_base);
        for (int64_t k = 0; k < count; ++k) sz[k] = 1;
        for (int64_t k = count - 1; k >= 1; --k) sz[host->parent[base + k]] += sz[k];
    }

    // Pageable memory goes to the device through two pinned chunks: the host threads fill one while the other is on
    // the wire (a plain cudaMemcpy from pageable memory stages serially through the driver at a fraction of the
    // link's rate: 0.26 s for the 0.8 GB of the 50 000-taxon workload).
    constexpr size_t kChunk = 32u << 20;
    const bool staged = nM * sizeof(double) > (8u << 20);
    unsigned char *pin = nullptr;
    cudaEvent_t sent[2] = {nullptr, nullptr};
    if (staged) {
        void *pin_v;
        if ((rc = reserve_pinned(ctx, 2 * kChunk, &pin_v))) return rc;
        pin = static_cast<unsigned char *>(pin_v);
        for (cudaEvent_t &e : sent) SCS_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    int turn = 0;
    const int copy_threads = scs_host_threads() > 0 ? scs_host_threads() : 1;
    auto put = [&](void *dst, const void *src, size_t bytes) -> int {
        if (bytes == 0) return SCS_OK;
        ctx->h2d_bytes += static_cast<int64_t>(bytes);
        if (!staged || bytes < (1u << 20)) {
            SCS_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
            return SCS_OK;
        }
        for (size_t at = 0; at < bytes; at += kChunk, turn ^= 1) {
            const size_t len = bytes - at < kChunk ? bytes - at : kChunk;
            unsigned char *stage = pin + static_cast<size_t>(turn) * kChunk;
            SCS_CUDA(ctx, cudaEventSynchronize(sent[turn]));  // the chunk's previous copy has left it
            const long long pieces = static_cast<long long>((len + (1u << 20) - 1) >> 20);
#pragma omp parallel for schedule(static) num_threads(copy_threads)
            for (long long p = 0; p < pieces; ++p) {
                const size_t o = static_cast<size_t>(p) << 20;
                std::memcpy(stage + o, static_cast<const unsigned char *>(src) + at + o, len - o < (1u << 20) ? len - o : (1u << 20));
            }
            SCS_CUDA(ctx, cudaMemcpyAsync(static_cast<unsigned char *>(dst) + at, stage, len, cudaMemcpyHostToDevice, ctx->stream));
            SCS_CUDA(ctx, cudaEventRecord(sent[turn], ctx->stream));
        }
        return SCS_OK;
    };
    if ((rc = put(out->tree_off.ptr, host->node_offsets.data(), (nT + 1) * sizeof(int64_t)))) return rc;
    if ((rc = put(out->leaf_off.ptr, host->leaf_offsets.data(), (nT + 1) * sizeof(int64_t)))) return rc;
    if ((rc = put(out->weight.ptr, host->weight.data(), nT * sizeof(double)))) return rc;
    SCS_CUDA(ctx, cudaMemsetAsync(out->tree_job.ptr, 0, nT * sizeof(int32_t), ctx->stream));
    const size_t n_mine = static_cast<size_t>(hi - lo);
    auto source_of = [&](const NodeArray &a) -> const unsigned char * {
        if (a.src) return static_cast<const unsigned char *>(a.src) + static_cast<size_t>(lo) * a.elem;
        return reinterpret_cast<const unsigned char *>(size.data() + (lo - size_base));
    };
    if (!share) {
        for (const NodeArray &a : arrays)
            if ((rc = put(a.dst->ptr, source_of(a), n_mine * a.elem))) return rc;
    } else {
        // nobody may still be reading this rank's W block (the last shared node of a previous build)
        if ((rc = shard_barrier(ctx))) return rc;
        unsigned char *stage = sh.window + sh.layout.W;
        for (const NodeArray &a : arrays)
            if ((rc = put(stage + a.stage_off + static_cast<size_t>(lo) * a.elem, source_of(a), n_mine * a.elem))) return rc;
        if ((rc = shard_barrier(ctx))) return rc;  // every rank's slice is in its window
        for (int step = 0; step < G; ++step) {
            const int p = (me + step) % G;  // staggered so that the ranks do not all read one peer
            const size_t b = static_cast<size_t>(slice_begin(p)), e = static_cast<size_t>(slice_begin(p + 1));
            for (const NodeArray &a : arrays) {
                if (e == b) continue;
                SCS_CUDA(ctx, cudaMemcpyAsync(static_cast<unsigned char *>(a.dst->ptr) + b * a.elem,
                                              sh.peer[p] + sh.layout.W + a.stage_off + b * a.elem, (e - b) * a.elem,
                                              cudaMemcpyDeviceToDevice, ctx->stream));
            }
        }
        if ((rc = shard_barrier(ctx))) return rc;  // the windows are free again
    }
    // `size` is a local: the copies above must have read it before it goes away
    SCS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (cudaEvent_t e : sent)
        if (e) cudaEventDestroy(e);
    if (share) {
        // a peer that timed out in one of the barriers leaves its mark in the window header
        unsigned int error = 0;
        SCS_CUDA(ctx, cudaMemcpy(&error, sh.window + offsetof(ShardHeader, error), sizeof(error), cudaMemcpyDeviceToHost));
        if (error) return fail(ctx, SCS_ERR_PEER, "device forest: a wait for a peer GPU timed out");
    }
    return SCS_OK;
}

int devforest_tours(scs_ctx *ctx, const DevForest &f, int weighting, const int32_t *taxon_vertex_dev, DevTours *tours,
                    int32_t *flags_dev) {
    int rc;
    const size_t nL = static_cast<size_t>(f.leaves), nT = static_cast<size_t>(f.trees), nM = static_cast<size_t>(f.nodes);
    if ((rc = grow(ctx, tours->leaf_taxon, (nL + 1) * sizeof(int32_t)))) return rc;
    if ((rc = grow(ctx, tours->adj_depth, (nL + 1) * sizeof(int32_t)))) return rc;
    if ((rc = grow(ctx, tours->adj_val, (nL + 1) * sizeof(double)))) return rc;
    if ((rc = grow(ctx, tours->root_depth, (nT + 1) * sizeof(int32_t)))) return rc;
    if ((rc = grow(ctx, tours->depth_s, (nM + 1) * sizeof(int32_t)))) return rc;
    if ((rc = grow(ctx, tours->val_s, (nM + 1) * sizeof(double)))) return rc;
    if (f.trees == 0) return SCS_OK;
    df_tours<<<ceil_div(f.trees * 32, 256), 256, 0, ctx->stream>>>(
        f.trees, f.tree_off.as<int64_t>(), f.leaf_off.as<int64_t>(), f.parent.as<int32_t>(), f.taxon.as<int32_t>(),
        f.has_length ? f.length.as<double>() : nullptr, f.has_support ? f.support.as<double>() : nullptr, weighting,
        taxon_vertex_dev, tours->depth_s.as<int32_t>(), tours->val_s.as<double>(), tours->leaf_taxon.as<int32_t>(),
        tours->adj_depth.as<int32_t>(), tours->adj_val.as<double>(), tours->root_depth.as<int32_t>(), flags_dev);
    SCS_LAUNCHED(ctx, "df_tours");
    return SCS_OK;
}

int devforest_restrict(scs_ctx *ctx, const DevForest &src, int J, const int32_t *job_tree_begin, const int32_t *job_parts,
                       const int32_t *job_part_base, int num_parts, const int32_t *part_newjob, int new_jobs,
                       const int32_t *owner_host, int num_taxa, DevForest *dst, DevJobInfo *info, uint8_t *present_host) {
    if (J <= 0 || !job_tree_begin || !job_parts || !job_part_base || !part_newjob || !owner_host || !dst || !info || !present_host)
        return fail(ctx, SCS_ERR_INVALID, "devforest_restrict: bad argument");
    // cells: one per (part, tree) of every job that was split
    std::vector<int32_t> cell_base(static_cast<size_t>(J) + 1, 0);
    int64_t cells = 0;
    for (int j = 0; j < J; ++j) {
        cell_base[j] = static_cast<int32_t>(cells);
        cells += static_cast<int64_t>(job_parts[j]) * (job_tree_begin[j + 1] - job_tree_begin[j]);
        if (cells >= (1ll << 31) - 8) return fail(ctx, SCS_ERR_INVALID, "device forest: too many (tree, part) cells");
    }
    cell_base[J] = static_cast<int32_t>(cells);
    dst->has_length = src.has_length;
    dst->has_support = src.has_support;
    const size_t nJ = static_cast<size_t>(J), nP = static_cast<size_t>(num_parts > 0 ? num_parts : 1), nX = static_cast<size_t>(num_taxa),
                 nC = static_cast<size_t>(cells), nN = static_cast<size_t>(new_jobs > 0 ? new_jobs : 1);
    std::memset(present_host, 0, nX);
    if (cells == 0 || new_jobs == 0) {
        dst->trees = dst->nodes = dst->leaves = 0;
        for (int j = 0; j < new_jobs; ++j) info[j] = DevJobInfo{0, 0, 0, 0, 0, 0};
        return SCS_OK;
    }
    int rc;
    // ---- plan tables + owner: one pinned staging block, one copy -----------------------------------------------------
    const size_t plan_bytes = (4 * (nJ + 1) + nP + nX) * sizeof(int32_t) + 2048;
    void *pin_v;
    if ((rc = reserve_pinned(ctx, plan_bytes + nX + sizeof(DevJobInfo) * nN + 4096, &pin_v))) return rc;
    unsigned char *pin = static_cast<unsigned char *>(pin_v);
    unsigned char *plan_dev_raw;
    if ((rc = reserve_as(ctx, SLOT_MED_STAGE, plan_bytes, &plan_dev_raw))) return rc;
    size_t at = 0;
    auto stage = [&](const int32_t *srcp, size_t count) {
        const size_t here = at;
        std::memcpy(pin + here, srcp, count * sizeof(int32_t));
        at += (count * sizeof(int32_t) + 255) & ~static_cast<size_t>(255);
        return here;
    };
    const size_t o_tb = stage(job_tree_begin, nJ + 1), o_parts = stage(job_parts, nJ), o_pbase = stage(job_part_base, nJ),
                 o_cbase = stage(cell_base.data(), nJ + 1), o_newjob = stage(part_newjob, nP), o_owner = stage(owner_host, nX);
    SCS_CUDA(ctx, cudaMemcpyAsync(plan_dev_raw, pin, at, cudaMemcpyHostToDevice, ctx->stream));
    ctx->h2d_bytes += static_cast<int64_t>(at);
    RestrictPlan plan;
    plan.job_tree_begin = reinterpret_cast<const int32_t *>(plan_dev_raw + o_tb);
    plan.job_parts = reinterpret_cast<const int32_t *>(plan_dev_raw + o_parts);
    plan.job_part_base = reinterpret_cast<const int32_t *>(plan_dev_raw + o_pbase);
    plan.job_cell_base = reinterpret_cast<const int32_t *>(plan_dev_raw + o_cbase);
    plan.part_newjob = reinterpret_cast<const int32_t *>(plan_dev_raw + o_newjob);
    plan.owner = reinterpret_cast<const int32_t *>(plan_dev_raw + o_owner);
    plan.J = J;

    // ---- which (tree, part) pairs exist ---------------------------------------------------------------------------------
    // workspace A: hist | flag | pair_index [cells + 1 each] ; pair arrays sized by cells as an upper bound would be
    // wasteful for many-way splits, so the pair count is read back first
    unsigned char *wa;
    if ((rc = reserve_as(ctx, SLOT_MED_STATE, 3 * ((nC + 2) * sizeof(int32_t) + 256) + 1024, &wa))) return rc;
    int32_t *hist = carve<int32_t>(wa, nC + 1);
    int32_t *flag = carve<int32_t>(wa, nC + 1);
    int32_t *pair_index = carve<int32_t>(wa, nC + 2);
    SCS_CUDA(ctx, cudaMemsetAsync(hist, 0, (nC + 1) * sizeof(int32_t), ctx->stream));
    const int64_t T = src.trees;
    df_count<<<ceil_div(T * 32, 256), 256, 0, ctx->stream>>>(T, src.tree_off.as<int64_t>(), src.taxon.as<int32_t>(),
                                                            src.tree_job.as<int32_t>(), plan, hist);
    SCS_LAUNCHED(ctx, "df_count");
    const int ncells = static_cast<int>(cells);
    df_flag_cells<<<ceil_div(ncells, 256), 256, 0, ctx->stream>>>(ncells, hist, flag);
    SCS_LAUNCHED(ctx, "df_flag_cells");
    if ((rc = exclusive_scan(ctx, ncells, flag, pair_index))) return rc;
    int32_t *pin_counts = reinterpret_cast<int32_t *>(pin + plan_bytes);
    SCS_CUDA(ctx, cudaMemcpyAsync(pin_counts, pair_index + ncells, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    SCS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const int pairs = pin_counts[0];
    if (pairs == 0) {
        dst->trees = dst->nodes = dst->leaves = 0;
        for (int j = 0; j < new_jobs; ++j) info[j] = DevJobInfo{0, 0, 0, 0, 0, 0};
        return SCS_OK;
    }
    const size_t nQ = static_cast<size_t>(pairs);
    // workspace B: per pair
    unsigned char *wb;
    if ((rc = reserve_as(ctx, SLOT_MED_VEC, 9 * ((nQ + 2) * sizeof(int32_t) + 256) + 2 * (nN + 2) * sizeof(int32_t) +
                                               (nN + 1) * sizeof(unsigned long long) + nX + 4096, &wb)))
        return rc;
    int32_t *pair_tree = carve<int32_t>(wb, nQ + 1);
    int32_t *pair_part = carve<int32_t>(wb, nQ + 1);
    int32_t *pair_tips = carve<int32_t>(wb, nQ + 1);
    int32_t *pair_need = carve<int32_t>(wb, nQ + 1);
    int32_t *scratch_off = carve<int32_t>(wb, nQ + 2);
    int32_t *pair_nodes = carve<int32_t>(wb, nQ + 1);
    int32_t *node_off32 = carve<int32_t>(wb, nQ + 2);
    int32_t *leaf_off32 = carve<int32_t>(wb, nQ + 2);
    int32_t *pair_count_dev = carve<int32_t>(wb, 64);
    int32_t *job_trees = carve<int32_t>(wb, nN + 1);
    int32_t *job_tree_begin_new = carve<int32_t>(wb, nN + 2);
    unsigned long long *job_visits = carve<unsigned long long>(wb, nN + 1);
    uint8_t *present = carve<uint8_t>(wb, nX + 1);
    SCS_CUDA(ctx, cudaMemcpyAsync(pair_count_dev, pair_index + ncells, sizeof(int32_t), cudaMemcpyDeviceToDevice, ctx->stream));
    df_make_pairs<<<ceil_div(ncells, 256), 256, 0, ctx->stream>>>(ncells, hist, pair_index, src.tree_off.as<int64_t>(), plan,
                                                                 pair_tree, pair_part, pair_tips, pair_need);
    SCS_LAUNCHED(ctx, "df_make_pairs");
    if ((rc = exclusive_scan(ctx, pairs, pair_need, scratch_off))) return rc;
    SCS_CUDA(ctx, cudaMemcpyAsync(pin_counts + 1, scratch_off + pairs, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    SCS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const size_t scratch_entries = static_cast<size_t>(pin_counts[1]);
    // workspace C: P | R | U per (pair, node)
    unsigned char *wc;
    if ((rc = reserve_as(ctx, SLOT_BASIS, 2 * (scratch_entries * sizeof(int32_t) + 256) + scratch_entries + 1024, &wc))) return rc;
    int32_t *P_all = carve<int32_t>(wc, scratch_entries);
    int32_t *R_all = carve<int32_t>(wc, scratch_entries);
    uint8_t *U_all = carve<uint8_t>(wc, scratch_entries);
    const int pair_blocks = ceil_div(static_cast<int64_t>(pairs) * 32, 256);
    df_mark<<<pair_blocks, 256, 0, ctx->stream>>>(pair_count_dev, pair_tree, pair_part, scratch_off, src.tree_off.as<int64_t>(),
                                                 src.parent.as<int32_t>(), src.size.as<int32_t>(), src.taxon.as<int32_t>(),
                                                 plan.owner, P_all, R_all, U_all, pair_nodes);
    SCS_LAUNCHED(ctx, "df_mark");
    if ((rc = exclusive_scan(ctx, pairs, pair_nodes, node_off32))) return rc;
    if ((rc = exclusive_scan(ctx, pairs, pair_tips, leaf_off32))) return rc;

    // ---- the new forest: at most 2 nodes per kept tip (restricted trees branch everywhere) ------------------------------
    const size_t cap_nodes = 2 * static_cast<size_t>(src.leaves) + 16;
    if ((rc = grow(ctx, dst->tree_off, (nQ + 1) * sizeof(int64_t)))) return rc;
    if ((rc = grow(ctx, dst->leaf_off, (nQ + 1) * sizeof(int64_t)))) return rc;
    if ((rc = grow(ctx, dst->parent, cap_nodes * sizeof(int32_t)))) return rc;
    if ((rc = grow(ctx, dst->size, cap_nodes * sizeof(int32_t)))) return rc;
    if ((rc = grow(ctx, dst->taxon, cap_nodes * sizeof(int32_t)))) return rc;
    if ((rc = grow(ctx, dst->weight, nQ * sizeof(double)))) return rc;
    if ((rc = grow(ctx, dst->tree_job, nQ * sizeof(int32_t)))) return rc;
    if (dst->has_length && (rc = grow(ctx, dst->length, cap_nodes * sizeof(double)))) return rc;
    if (dst->has_support && (rc = grow(ctx, dst->support, cap_nodes * sizeof(double)))) return rc;
    SCS_CUDA(ctx, cudaMemsetAsync(job_trees, 0, (nN + 1) * sizeof(int32_t), ctx->stream));
    SCS_CUDA(ctx, cudaMemsetAsync(job_visits, 0, (nN + 1) * sizeof(unsigned long long), ctx->stream));
    SCS_CUDA(ctx, cudaMemsetAsync(present, 0, nX + 1, ctx->stream));
    df_write<<<pair_blocks, 256, 0, ctx->stream>>>(
        pair_count_dev, pair_tree, pair_part, pair_tips, scratch_off, node_off32, plan.part_newjob, src.tree_off.as<int64_t>(),
        src.parent.as<int32_t>(), src.size.as<int32_t>(), src.taxon.as<int32_t>(),
        src.has_length ? src.length.as<double>() : nullptr, src.has_support ? src.support.as<double>() : nullptr,
        src.weight.as<double>(), R_all, dst->parent.as<int32_t>(), dst->size.as<int32_t>(), dst->taxon.as<int32_t>(),
        dst->has_length ? dst->length.as<double>() : nullptr, dst->has_support ? dst->support.as<double>() : nullptr,
        dst->weight.as<double>(), dst->tree_job.as<int32_t>(), present, job_trees, job_visits);
    SCS_LAUNCHED(ctx, "df_write");
    df_widen<<<ceil_div(pairs + 1, 256), 256, 0, ctx->stream>>>(pairs, node_off32, dst->tree_off.as<int64_t>());
    SCS_LAUNCHED(ctx, "df_widen");
    df_widen<<<ceil_div(pairs + 1, 256), 256, 0, ctx->stream>>>(pairs, leaf_off32, dst->leaf_off.as<int64_t>());
    SCS_LAUNCHED(ctx, "df_widen");
    if ((rc = exclusive_scan(ctx, new_jobs, job_trees, job_tree_begin_new))) return rc;
    DevJobInfo *info_dev;
    if ((rc = reserve_as(ctx, SLOT_NODE_STATS, nN + 1, &info_dev))) return rc;
    df_job_info<<<ceil_div(new_jobs, 256), 256, 0, ctx->stream>>>(new_jobs, job_trees, job_tree_begin_new, node_off32, leaf_off32,
                                                                 job_visits, info_dev);
    SCS_LAUNCHED(ctx, "df_job_info");
    unsigned char *pin_present = pin + plan_bytes + 256;
    DevJobInfo *pin_info = reinterpret_cast<DevJobInfo *>(pin + plan_bytes + 256 + ((nX + 255) & ~static_cast<size_t>(255)));
    SCS_CUDA(ctx, cudaMemcpyAsync(pin_present, present, nX, cudaMemcpyDeviceToHost, ctx->stream));
    SCS_CUDA(ctx, cudaMemcpyAsync(pin_info, info_dev, sizeof(DevJobInfo) * new_jobs, cudaMemcpyDeviceToHost, ctx->stream));
    SCS_CUDA(ctx, cudaMemcpyAsync(pin_counts + 2, node_off32 + pairs, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    SCS_CUDA(ctx, cudaMemcpyAsync(pin_counts + 3, leaf_off32 + pairs, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    SCS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->d2h_bytes += static_cast<int64_t>(nX + sizeof(DevJobInfo) * new_jobs + 16);
    std::memcpy(present_host, pin_present, nX);
    std::memcpy(info, pin_info, sizeof(DevJobInfo) * new_jobs);
    dst->trees = pairs;
    dst->nodes = pin_counts[2];
    dst->leaves = pin_counts[3];
    return SCS_OK;
}

int devforest_fetch_trees(scs_ctx *ctx, const DevForest &forest, int count, const int64_t *first_node,
                          const int64_t *tree_nodes, int32_t *parent_out, int32_t *taxon_out) {
    if (count <= 0) return SCS_OK;
    std::vector<FetchItem> items(static_cast<size_t>(count));
    int64_t total = 0;
    for (int i = 0; i < count; ++i) {
        if (first_node[i] < 0 || first_node[i] + tree_nodes[i] > forest.nodes)
            return fail(ctx, SCS_ERR_INVALID, "devforest_fetch_trees: tree outside the forest");
        items[i] = FetchItem{first_node[i], total, tree_nodes[i]};
        total += tree_nodes[i];
    }
    FetchItem *items_dev;
    int32_t *packed;
    int rc;
    if ((rc = reserve_as(ctx, SLOT_MED_ROW_NODE, static_cast<size_t>(count) + 1, &items_dev))) return rc;
    if ((rc = reserve_as(ctx, SLOT_MED_TREE_NODE, 2 * static_cast<size_t>(total) + 8, &packed))) return rc;
    SCS_CUDA(ctx, cudaMemcpyAsync(items_dev, items.data(), sizeof(FetchItem) * count, cudaMemcpyHostToDevice, ctx->stream));
    df_fetch<<<count, 128, 0, ctx->stream>>>(count, items_dev, forest.parent.as<int32_t>(), forest.taxon.as<int32_t>(), packed,
                                            packed + total);
    SCS_LAUNCHED(ctx, "df_fetch");
    SCS_CUDA(ctx, cudaMemcpyAsync(parent_out, packed, sizeof(int32_t) * total, cudaMemcpyDeviceToHost, ctx->stream));
    SCS_CUDA(ctx, cudaMemcpyAsync(taxon_out, packed + total, sizeof(int32_t) * total, cudaMemcpyDeviceToHost, ctx->stream));
    SCS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // also keeps `items` alive until the copy has read it
    ctx->d2h_bytes += static_cast<int64_t>(2 * sizeof(int32_t) * total);
    return SCS_OK;
}

}  // namespace scs
