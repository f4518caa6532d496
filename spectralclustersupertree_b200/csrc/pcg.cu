// Proper-cluster-graph build on the GPU.
//
// Replaces _proper_cluster_graph_edges + _dfs_pcg_weights of the reference
// (/root/reference/src/sc_supertree/scs.py:495-583, 586-663).
//
// Layout.  Source trees arrive as leaf tours (flatten.py / forest.cpp): leaves in depth-first order
// and, for consecutive leaves i, i+1, the depth and weighting value of their LCA.  The LCA of the
// leaves at tour positions p < q is the shallowest entry of adj[p..q-1].  Seen from one leaf p, that running
// minimum is a staircase: it changes once per ancestor of the leaf, and all leaves between two steps share
// their LCA with p.  pcg_tour_links stores, per tour entry, the nearest strictly shallower entry on either
// side, so a row CTA recovers the staircase of a tree with ~depth dependent reads.
//
// Kernel shape.  One CTA owns one row `a` of W (one taxon) and keeps the row's accumulators in
// shared memory (8 B weight + 2/4 B count per column).  It visits the trees containing `a` in
// input order, 32 at a time: two threads per tree walk the leaf's chains into (boundary, term) segments in
// shared memory; then every thread takes leaves q of the tree that share a non-root LCA with `a` (pairs
// separated by the root are no proper clusters, scs.py:570-579, and are never touched), finds q's segment
// (5-step search, conflict-free) and adds fl(val(LCA) * w_t) to column taxon(q) (scs.py:644-658).  Within a
// tree every leaf is a distinct column, so threads never collide; one barrier separates consecutive trees,
// so every W entry is summed in tree input order with separately rounded multiply and add
// (scs.py:655-657) -- bit-identical to the reference, no atomics on W, and W[a][b] == W[b][a] bit for
// bit.  The finished row is written once, coalesced, together with its adjacency bits (C > 0,
// scs.py:651-652), max-graph bits (C == max(occ_a, occ_b), scs.py:302-305) and its row sum (the degree the
// spectral step needs), so the co-occurrence matrix never has to be written to HBM unless the caller asks.
//
// Rows wider than shared memory (n > ~22k columns) are split into column chunks (gridDim.y);
// each chunk CTA walks the same trees and keeps only its columns.

#include "common.cuh"

namespace scs {

namespace {

constexpr int kRowThreads = 512;  // 2 CTAs/SM at n = 10^4 (100 KB of row accumulators each): 32 warps/SM
constexpr int kWarps = kRowThreads / 32;
constexpr size_t kRowsStaticSmem = 16 * 1024;  // room left for the row kernel's static shared memory (TreeBatch)

// ---- index: leaf -> tree, occurrences, taxon -> leaves ------------------------------------
// Batched build (medium.cu): the trees of several recursion nodes are concatenated, tree t belongs to node
// tree_node[t], and the rows of all nodes form one global row space (row = nodes[b].row_base + vertex id).
// `batch.nodes == nullptr` is the single-node build: one node of n vertices, row == vertex id.
struct BatchView {
    const MedNode *nodes = nullptr;
    const int32_t *tree_node = nullptr;  // [T]
    const int32_t *row_node = nullptr;   // [rows]
};

__global__ void pcg_index_leaves(int n, int T, int64_t L, const int64_t *__restrict__ leaf_offsets,
                                 const int32_t *__restrict__ leaf_taxon, int32_t *__restrict__ leaf_tree,
                                 int32_t *__restrict__ occ, int32_t *__restrict__ bad, BatchView batch) {
    int64_t g = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (g >= L) return;
    int lo = 0, hi = T;  // largest t with leaf_offsets[t] <= g
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (leaf_offsets[mid] <= g) lo = mid; else hi = mid;
    }
    leaf_tree[g] = lo;
    int a = leaf_taxon[g];
    int base = 0;
    if (batch.nodes) {
        const int b = batch.tree_node[lo];
        if (b < 0) return;  // a tree of a sub-problem that is not in this batch
        const MedNode &nd = batch.nodes[b];
        n = nd.n;
        base = nd.row_base;
    }
    if (a < 0 || a >= n) { *bad = 1; return; }
    atomicAdd(&occ[base + a], 1);
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
                                 const int32_t *__restrict__ leaf_tree, const int32_t *__restrict__ row_ptr,
                                 int32_t *__restrict__ cursor, int32_t *__restrict__ inv, BatchView batch) {
    int64_t g = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (g >= L) return;
    int a = leaf_taxon[g];
    int base = 0;
    if (batch.nodes) {
        const int b = batch.tree_node[leaf_tree[g]];
        if (b < 0) return;
        const MedNode &nd = batch.nodes[b];
        n = nd.n;
        base = nd.row_base;
    }
    if (a < 0 || a >= n) return;  // flagged by pcg_index_leaves
    a += base;
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

// ---- per-tree links: nearest strictly shallower tour entry on either side ----------------------------
// adj entry i of a tree is the LCA of its leaves i and i + 1.  For a leaf at tour position p the LCA with
// the leaves to its right is the running minimum of adj[p..]: it changes exactly at the chain
// p -> next_right[p] -> next_right[next_right[p]] ... (next_right[i] = first j > i with a strictly smaller
// depth; equal depths inside a range are the same node), one step per ancestor of the leaf, and likewise to
// the left.  pcg_tour_links stores, per entry, both links and the weighting value, so that a row CTA gets
// the whole "staircase" of (leaf range, value) segments of a tree from ~depth dependent 16-byte reads
// instead of one range-minimum query per leaf pair.  Entries at root depth are marked: every pair they
// separate has the root as LCA and is no proper cluster (scs.py:570-579).
struct __align__(16) LinkEntry {
    int32_t next_right, next_left;  // entry index within the tree, kNone, or kRootLevel for both
    double val;
};
constexpr int32_t kNone = -1;
constexpr int32_t kRootLevel = -2;

// One CTA per tree.  Block minima over 32 and 1024 entries (shared memory) bound every search to
// O(32 + 32 + k / 1024) probes, independent for every entry.
__global__ void __launch_bounds__(256)
pcg_tour_links(int n, const int64_t *__restrict__ leaf_offsets, const int32_t *__restrict__ adj_depth,
               const double *__restrict__ adj_val, const int32_t *__restrict__ root_depth,
               LinkEntry *__restrict__ links) {
    extern __shared__ int32_t link_smem[];
    const int t = blockIdx.x;
    const int64_t tb = leaf_offsets[t];
    const int m = static_cast<int>(leaf_offsets[t + 1] - tb) - 1;  // adj entries of this tree
    if (m <= 0 || m >= n) return;  // nothing to link / a repeated taxon (the row kernel flags it)
    const int32_t *D = adj_depth + tb;
    const int nb1 = (m + 31) >> 5, nb2 = (nb1 + 31) >> 5;
    int32_t *B1 = link_smem, *B2 = link_smem + nb1;
    const int lane = threadIdx.x & 31;
    for (int i0 = (threadIdx.x >> 5) << 5; i0 < m; i0 += blockDim.x) {  // a warp per 32-entry block
        int v = i0 + lane < m ? D[i0 + lane] : INT32_MAX;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, off));
        if (lane == 0) B1[i0 >> 5] = v;
    }
    __syncthreads();
    for (int b0 = (threadIdx.x >> 5) << 5; b0 < nb1; b0 += blockDim.x) {
        int v = b0 + lane < nb1 ? B1[b0 + lane] : INT32_MAX;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, off));
        if (lane == 0) B2[b0 >> 5] = v;
    }
    __syncthreads();
    const int rd = root_depth[t];
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
        const int d = D[i];
        LinkEntry e;
        e.val = adj_val[tb + i];
        if (d == rd) {
            e.next_right = e.next_left = kRootLevel;
            links[tb + i] = e;
            continue;
        }
        // ---- to the right: own block, blocks of the own superblock, superblocks, then back down
        int j = i + 1;
        {
            const int end1 = min(m, ((i >> 5) + 1) << 5);
            while (j < end1 && D[j] >= d) ++j;
            if (j == end1 && j < m) {
                int b = j >> 5;
                const int endb = min(nb1, ((b >> 5) + 1) << 5);
                while (b < endb && B1[b] >= d) ++b;
                if (b == endb && b < nb1) {
                    int s = b >> 5;
                    while (s < nb2 && B2[s] >= d) ++s;
                    if (s < nb2) {
                        b = s << 5;
                        while (B1[b] >= d) ++b;
                    } else {
                        b = nb1;
                    }
                }
                if (b < nb1) {
                    j = b << 5;
                    while (D[j] >= d) ++j;
                } else {
                    j = m;
                }
            }
        }
        e.next_right = j < m ? j : kNone;
        // ---- to the left, mirrored
        j = i - 1;
        {
            const int end1 = (i >> 5) << 5;  // first entry of the own block
            while (j >= end1 && D[j] >= d) --j;
            if (j < end1 && j >= 0) {
                int b = j >> 5;
                const int endb = (b >> 5) << 5;  // first block of the own superblock
                while (b >= endb && B1[b] >= d) --b;
                if (b < endb && b >= 0) {
                    int s = b >> 5;
                    while (s >= 0 && B2[s] >= d) --s;
                    if (s >= 0) {
                        b = (s << 5) + 31;
                        while (B1[b] >= d) --b;
                    } else {
                        b = -1;
                    }
                }
                if (b >= 0) {
                    j = min(m - 1, (b << 5) + 31);
                    while (D[j] >= d) --j;
                } else {
                    j = -1;
                }
            }
        }
        e.next_left = j >= 0 ? j : kNone;
        links[tb + i] = e;
    }
}

// ---- buckets: every tree's leaves, grouped by the warp that owns their column ------------------------
// Column c of a row lives in column chunk y = c / cols_per_chunk (one CTA each) and, inside the chunk, belongs to
// warp (c - y * cols_per_chunk) % W of that CTA.  bucket (t, y * W + w) lists the leaves of tree t whose columns
// that warp owns, as entries {tour position, accumulator slot}.  A warp that walks its own buckets tree by tree
// touches columns nobody else touches: W is summed in tree input order without a single barrier between trees,
// and a chunk CTA reads only its own share of every tree.
struct BucketShape {
    int cols_per_chunk, warps_log2, buckets;  // buckets = chunks * warps
};

__device__ __forceinline__ void bucket_of(const BucketShape &bs, int c, int &bucket, int &slot) {
    const int y = c / bs.cols_per_chunk;
    const int lc = c - y * bs.cols_per_chunk;
    bucket = (y << bs.warps_log2) + (lc & ((1 << bs.warps_log2) - 1));
    slot = lc >> bs.warps_log2;
}

// One warp per tree.  The leaves of a tree are distinct columns, so "sorted by (bucket, slot)" is a rank in a bitmap:
// bit  bucket * key_stride + slot  is set for every leaf (key_stride: the slots of a bucket rounded up to whole words, so
// every bucket starts on a word boundary), an exclusive scan of the words' popcounts gives every set bit its rank, and
// the rank is the leaf's place among the tree's entries.  The starts of the tree's buckets fall out of the same scan
// (no counting pass, no scan over all cells).  Entries of a bucket ascend by slot, i.e. by column: a row CTA that
// only wants the columns beyond its own (triangle mode) finds where they start with a binary search.
template <typename EntryT>
__global__ void __launch_bounds__(256)
pcg_bucket_sorted(int n, int T, BucketShape bs, int key_stride, int words, const int64_t *__restrict__ leaf_offsets,
                  const int32_t *__restrict__ leaf_taxon, int32_t *__restrict__ bucket_ptr, EntryT *__restrict__ entries,
                  int32_t *__restrict__ start16, int32_t *__restrict__ bad, BatchView batch) {
    extern __shared__ uint32_t bucket_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t = blockIdx.x * (blockDim.x >> 5) + warp;
    if (t >= T) return;  // whole warps leave; only warp-level synchronisation below
    uint32_t *bm = bucket_smem + static_cast<size_t>(warp) * 2 * words;
    int32_t *pre = reinterpret_cast<int32_t *>(bm + words);
    const int64_t tb = leaf_offsets[t];
    const int k = static_cast<int>(leaf_offsets[t + 1] - tb);
    const size_t cell0 = static_cast<size_t>(t) * bs.buckets;
    if (batch.nodes) {
        const int b = batch.tree_node[t];
        if (b < 0) {  // a tree of a sub-problem that is not in this batch: empty buckets
            for (int j = lane; j < bs.buckets; j += 32) bucket_ptr[cell0 + j] = static_cast<int32_t>(tb);
            if (t == T - 1 && lane == 0) bucket_ptr[cell0 + bs.buckets] = static_cast<int32_t>(tb);
            return;
        }
        n = batch.nodes[b].n;
    }
    for (int w = lane; w < words; w += 32) bm[w] = 0u;
    __syncwarp();
    for (int i = lane; i < k; i += 32) {
        const int c = leaf_taxon[tb + i];
        if (c < 0 || c >= n) continue;  // flagged by pcg_index_leaves
        int bucket, slot;
        bucket_of(bs, c, bucket, slot);
        const int key = bucket * key_stride + slot;
        atomicOr(&bm[key >> 5], 1u << (key & 31));
    }
    __syncwarp();
    int carry = 0;
    for (int w0 = 0; w0 < words; w0 += 32) {
        const int w = w0 + lane;
        const int v = w < words ? __popc(bm[w]) : 0;
        int inc = v;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int o = __shfl_up_sync(0xffffffffu, inc, off);
            if (lane >= off) inc += o;
        }
        if (w < words) pre[w] = carry + inc - v;
        carry += __shfl_sync(0xffffffffu, inc, 31);
    }
    __syncwarp();
    for (int i = lane; i < k; i += 32) {
        const int c = leaf_taxon[tb + i];
        if (c < 0 || c >= n) continue;
        int bucket, slot;
        bucket_of(bs, c, bucket, slot);
        const int key = bucket * key_stride + slot;
        const int rank = pre[key >> 5] + __popc(bm[key >> 5] & ((1u << (key & 31)) - 1u));
        const uint32_t pos = static_cast<uint32_t>(i);
        if (sizeof(EntryT) == 4) entries[tb + rank] = static_cast<EntryT>((pos << 16) | static_cast<uint32_t>(slot));
        else entries[tb + rank] = static_cast<EntryT>((static_cast<unsigned long long>(pos) << 32) | static_cast<uint32_t>(slot));
        if (start16) {
            // triangle mode: for the row of this leaf's own taxon, where each warp's entries beyond the diagonal start
            // in the chunk that holds the diagonal -- the rank of (bucket, first slot whose column exceeds c)
            const int y = c / bs.cols_per_chunk;
            const int la = c - y * bs.cols_per_chunk;
            const int nw = 1 << bs.warps_log2;
            for (int w = 0; w < nw; ++w) {
                const int smin = la >= w ? ((la - w) >> bs.warps_log2) + 1 : 0;
                const int key2 = ((y << bs.warps_log2) + w) * key_stride + smin;
                start16[(tb + i) * nw + w] =
                    static_cast<int32_t>(tb) + pre[key2 >> 5] + __popc(bm[key2 >> 5] & ((1u << (key2 & 31)) - 1u));
            }
        }
    }
    // a repeated or invalid taxon leaves the tail of the tree's entries unranked: harmless values there (the
    // result is rejected anyway), and the flag
    for (int i = carry + lane; i < k; i += 32) entries[tb + i] = EntryT(0);
    if (carry != k && lane == 0) *bad = 1;
    for (int j = lane; j < bs.buckets; j += 32)
        bucket_ptr[cell0 + j] = static_cast<int32_t>(tb) + pre[(j * key_stride) >> 5];
    if (t == T - 1 && lane == 0) bucket_ptr[cell0 + bs.buckets] = static_cast<int32_t>(tb) + k;
}

// ---- the row kernel -----------------------------------------------------------------------
constexpr int kTreeBatch = 32;  // (row, tree) incidences staged together: one per lane
constexpr int kChain = 15;      // chain steps per side kept per tree; longer chains continue in further rounds
constexpr int kSlots = 2 * kChain + 2;  // == 32: boundaries of one tree, ascending
static_assert(kSlots == 32, "the segment search below is a 5-step binary search over 32 boundaries");
constexpr int32_t kFar = 1 << 30;

struct TreeHeader {
    int64_t base;  // first leaf of the tree in the tour arrays
    double weight;
    int leaves, position, tree, pad;
};

// What a row CTA knows about the trees of the current batch.  For tree e, the leaf at tour position q lies in
// segment idx iff bound[e][idx] < q <= bound[e][idx + 1]; the segment's pairs (a, leaf q) all have the same
// LCA and get term[e][idx].  Segments [low, kChain) are to the left of the row's own leaf, (kChain, high) to its
// right; the middle segment kChain is the part already handled (at first the leaf of the row itself); what
// lies outside [low, high) is separated from the row's leaf by the root, or not reached yet.
struct __align__(16) TreeBatch {
    TreeHeader header[kTreeBatch];
    double term[kTreeBatch][kSlots];
    int32_t bound[kTreeBatch][kSlots];
    int32_t pivot[kTreeBatch][kSlots / 4];      // bound[4j + 3]: the first level of the segment search
    int32_t low[kTreeBatch], high[kTreeBatch];  // first / last boundary in use
    int32_t resume[kTreeBatch][2];              // chain entry to continue from, or kNone
};

// The staircase of one leaf -- what TreeBatch holds for a (row, tree) incidence -- depends only on the leaf, not on
// the row CTA that uses it.  pcg_leaf_stairs walks both chains of EVERY leaf of the node once, two threads per leaf,
// all leaves at the same time (the walk is a chain of dependent 16-byte reads: latency-bound inside a row CTA, where
// 64 threads walked while 448 waited; throughput-bound here), and a row CTA copies the finished staircases of its 32
// trees with coalesced loads.  Chains longer than kChain steps a side continue in the row kernel as before.
struct __align__(16) LeafStairs {
    double term[kSlots];
    int32_t bound[kSlots];
    int32_t pivot[kSlots / 4];
    int32_t low, high, resume0, resume1;
    int32_t leaves, position, tree, pad;  // leaves == 0: a tree with a repeated taxon (flagged), skipped
};
static_assert(sizeof(LeafStairs) == 448, "row CTAs copy a staircase as 28 int4");
constexpr int kStairsInt4 = sizeof(LeafStairs) / 16;

constexpr int kStairsLeaves = 64;  // leaves per CTA of pcg_leaf_stairs (two threads each)

__global__ void __launch_bounds__(2 * kStairsLeaves)
pcg_leaf_stairs(int n, int64_t L, const int64_t *__restrict__ leaf_offsets, const int32_t *__restrict__ leaf_tree,
                const LinkEntry *__restrict__ links, const double *__restrict__ tree_weight,
                LeafStairs *__restrict__ stairs, int32_t *__restrict__ bad, BatchView batch) {
    __shared__ LeafStairs staged[kStairsLeaves];  // written piecemeal by the walkers, copied out in one piece
    const int64_t g0 = static_cast<int64_t>(blockIdx.x) * kStairsLeaves;
    const int64_t g = g0 + (threadIdx.x >> 1);
    const int side = threadIdx.x & 1;
    bool active = g < L;
    int t = 0;
    if (active) {
        t = leaf_tree[g];
        if (batch.nodes) {
            const int b = batch.tree_node[t];
            if (b < 0) active = false;  // a tree of a sub-problem that is not in this batch
            else n = batch.nodes[b].n;
        }
    }
    if (active) {
        const int64_t base = leaf_offsets[t];
        int leaves = static_cast<int>(leaf_offsets[t + 1] - base);
        const int position = static_cast<int>(g - base);
        LeafStairs &out = staged[threadIdx.x >> 1];
        if (leaves > n) {  // more leaves than taxa: a taxon is repeated
            *bad = 1;
            leaves = 0;
        }
        const double weight = tree_weight[t];
        const LinkEntry *lk = links + base;
        int c = 0;
        if (side == 1) {
            out.leaves = leaves;
            out.position = position;
            out.tree = t;
            out.pad = 0;
            out.bound[kChain + 1] = position;
            int i = (leaves >= 2 && position <= leaves - 2) ? position : kNone;
            while (i >= 0 && c < kChain) {
                const LinkEntry r = lk[i];
                if (r.next_right == kRootLevel) { i = kNone; break; }
                out.term[kChain + 1 + c] = __dmul_rn(r.val, weight);
                out.bound[kChain + 2 + c] = r.next_right < 0 ? leaves - 1 : r.next_right;
                ++c;
                i = r.next_right;
            }
            out.high = kChain + 1 + c;
            out.resume1 = i;
            for (int k = kChain + 2 + c; k < kSlots; ++k) out.bound[k] = kFar;
            for (int j = kSlots / 8; j < kSlots / 4; ++j) out.pivot[j] = out.bound[4 * j + 3];
        } else {
            out.bound[kChain] = position - 1;
            int i = (leaves >= 2 && position >= 1 && position <= leaves - 1) ? position - 1 : kNone;
            while (i >= 0 && c < kChain - 1) {  // boundary 0 stays a pad: the search starts from it
                const LinkEntry r = lk[i];
                if (r.next_left == kRootLevel) { i = kNone; break; }
                out.term[kChain - 1 - c] = __dmul_rn(r.val, weight);
                out.bound[kChain - 1 - c] = r.next_left;  // kNone == -1: the segment starts at the first leaf
                ++c;
                i = r.next_left;
            }
            out.low = kChain - c;
            out.resume0 = i;
            for (int k = kChain - c - 1; k >= 0; --k) out.bound[k] = -kFar;
            for (int j = 0; j < kSlots / 8; ++j) out.pivot[j] = out.bound[4 * j + 3];
        }
    }
    __syncthreads();
    const int64_t left = L - g0;
    const int count = left < kStairsLeaves ? static_cast<int>(left) : kStairsLeaves;
    const int4 *src = reinterpret_cast<const int4 *>(staged);
    int4 *dst = reinterpret_cast<int4 *>(stairs + g0);
    for (int k = threadIdx.x; k < count * kStairsInt4; k += 2 * kStairsLeaves) dst[k] = src[k];
}

// Walk one side of the chain of tree e (thread-serial: one dependent 16-byte read per ancestor).
__device__ void walk_chain(TreeBatch &tb, int e, int side, bool first, const LinkEntry *__restrict__ links) {
    const TreeHeader h = tb.header[e];
    const LinkEntry *lk = links + h.base;
    int32_t *bound = tb.bound[e];
    double *term = tb.term[e];
    int c = 0;
    if (side == 1) {
        int i;
        if (first) {
            bound[kChain + 1] = h.position;
            i = (h.leaves >= 2 && h.position <= h.leaves - 2) ? h.position : kNone;
        } else {
            bound[kChain + 1] = bound[tb.high[e]];
            i = tb.resume[e][1];
        }
        while (i >= 0 && c < kChain) {
            const LinkEntry r = lk[i];
            if (r.next_right == kRootLevel) { i = kNone; break; }
            term[kChain + 1 + c] = __dmul_rn(r.val, h.weight);
            bound[kChain + 2 + c] = r.next_right < 0 ? h.leaves - 1 : r.next_right;
            ++c;
            i = r.next_right;
        }
        tb.high[e] = kChain + 1 + c;
        tb.resume[e][1] = i;
        for (int idx = kChain + 2 + c; idx < kSlots; ++idx) bound[idx] = kFar;
        for (int j = kSlots / 8; j < kSlots / 4; ++j) tb.pivot[e][j] = bound[4 * j + 3];
    } else {
        int i;
        if (first) {
            bound[kChain] = h.position - 1;
            i = (h.leaves >= 2 && h.position >= 1 && h.position <= h.leaves - 1) ? h.position - 1 : kNone;
        } else {
            bound[kChain] = bound[tb.low[e]];
            i = tb.resume[e][0];
        }
        while (i >= 0 && c < kChain - 1) {  // boundary 0 stays a pad: the search below starts from it
            const LinkEntry r = lk[i];
            if (r.next_left == kRootLevel) { i = kNone; break; }
            term[kChain - 1 - c] = __dmul_rn(r.val, h.weight);
            bound[kChain - 1 - c] = r.next_left;  // kNone == -1: the segment starts at the first leaf
            ++c;
            i = r.next_left;
        }
        tb.low[e] = kChain - c;
        tb.resume[e][0] = i;
        for (int idx = kChain - c - 1; idx >= 0; --idx) bound[idx] = -kFar;
        for (int j = 0; j < kSlots / 8; ++j) tb.pivot[e][j] = bound[4 * j + 3];
    }
}

template <typename EntryT>
__device__ __forceinline__ void unpack(EntryT en, int &pos, int &slot) {
    if (sizeof(EntryT) == 4) {
        pos = static_cast<int>(static_cast<uint32_t>(en) >> 16);
        slot = static_cast<int>(static_cast<uint32_t>(en) & 0xffffu);
    } else {
        pos = static_cast<int>(static_cast<unsigned long long>(en) >> 32);
        slot = static_cast<int>(static_cast<unsigned long long>(en) & 0xffffffffull);
    }
}

// count += 1 in one shared-memory operation: 16-bit counters are added to as halves of their 32-bit word (a
// counter never exceeds the number of trees, so no carry crosses into the upper half)
__device__ __forceinline__ void bump(uint16_t *counts, int slot) {
    atomicAdd(reinterpret_cast<unsigned int *>(counts) + (slot >> 1), 1u << ((slot & 1) << 4));
}
__device__ __forceinline__ void bump(int32_t *counts, int slot) { atomicAdd(counts + slot, 1); }

// One warp, one tree: every entry of the warp's bucket gets its segment and, if the segment is live, its term.
// Segment search over the tree's 32 ascending boundaries in two levels: 7 pivots (bound[3], bound[7], ...) held
// in registers pick a block of four, one 16-byte shared load fetches the block (the whole warp reads one
// 128-byte row: a single wavefront).
template <typename CountT, typename EntryT>
__device__ __forceinline__ void visit_bucket(const TreeBatch &tb, int e, int ptr, int end, int lane, int slot0,
                                             const EntryT *__restrict__ entries, double *accW, CountT *accC) {
    if (ptr >= end) return;
    const int low = tb.low[e], high = tb.high[e];
    if (low == kChain && high == kChain + 1) return;  // nothing but root-separated pairs
    const int4 *bound4 = reinterpret_cast<const int4 *>(tb.bound[e]);
    const double *term = tb.term[e];
    const int4 pa = *reinterpret_cast<const int4 *>(&tb.pivot[e][0]);
    const int4 pb = *reinterpret_cast<const int4 *>(&tb.pivot[e][4]);
    EntryT en = ptr + lane < end ? entries[ptr + lane] : EntryT(0);
    for (int i = ptr; i < end; i += 32) {
        const bool live = i + lane < end;
        const EntryT cur = en;
        if (i + 32 + lane < end) en = entries[i + 32 + lane];
        int q, slot;
        unpack(cur, q, slot);
        slot += slot0;  // the warp's slots start at slot0
        const int blk = (pa.x < q) + (pa.y < q) + (pa.z < q) + (pa.w < q) + (pb.x < q) + (pb.y < q) + (pb.z < q);
        const int4 b = bound4[blk];
        const int idx = 4 * blk - 1 + (b.x < q) + (b.y < q) + (b.z < q) + (b.w < q);  // bound[0] < q always
        if (live && idx >= low && idx < high && idx != kChain) {
            accW[slot] = __dadd_rn(accW[slot], term[idx]);
            bump(accC, slot);
        }
    }
}

// One warp, trees [first, stop) of the batch at once.  A warp's share of a tree is short (a tree's leaves spread over
// the 16 warps of every chunk: tens of entries), so going tree by tree exposes one global-memory latency per tree and
// leaves most lanes idle.  Here the warp's shares of the trees are concatenated in tree order and taken 32 items at a
// time, four such groups in flight: item j belongs to the tree found by a search over the running totals (held one per
// lane), its entry is loaded, its segment looked up in that tree's staircase.  Two items of a group that hit the same
// column come from different trees; they are applied one after the other in lane order = tree order, so every W entry
// still receives its terms in tree input order, each with one rounded multiply and one rounded add.
template <typename CountT, typename EntryT>
__device__ __forceinline__ void visit_trees(const TreeBatch &tb, int first, int stop, int my_ptr, int my_end, int lane,
                                            int slot0, const EntryT *__restrict__ entries, double *accW, CountT *accC) {
    constexpr unsigned kAll = 0xffffffffu;
    int cnt = 0;
    if (lane >= first && lane < stop) {
        cnt = my_end - my_ptr;
        if (tb.low[lane] == kChain && tb.high[lane] == kChain + 1) cnt = 0;  // nothing but root-separated pairs
    }
    int inc = cnt;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const int o = __shfl_up_sync(kAll, inc, off);
        if (lane >= off) inc += o;
    }
    const int total = __shfl_sync(kAll, inc, 31);
    const int base_of_mine = my_ptr - (inc - cnt);  // entry index of item j of tree `lane` = base_of_mine + j
    constexpr int kDepth = 4;
    for (int j0 = 0; j0 < total; j0 += 32 * kDepth) {
        EntryT en[kDepth];
        int tree[kDepth];
#pragma unroll
        for (int d = 0; d < kDepth; ++d) {
            const int j = j0 + 32 * d + lane;
            int e = 0;  // the first tree whose running total exceeds j
#pragma unroll
            for (int step = 16; step > 0; step >>= 1) {
                const int v = __shfl_sync(kAll, inc, e + step - 1);
                if (v <= j) e += step;
            }
            e = j < total ? e : 0;  // (e <= 31 wherever j < total)
            const int at = __shfl_sync(kAll, base_of_mine, e) + j;
            tree[d] = e;
            en[d] = j < total ? entries[at] : EntryT(0);
        }
#pragma unroll
        for (int d = 0; d < kDepth; ++d) {
            if (j0 + 32 * d >= total) break;  // warp-uniform
            const int e = tree[d];
            int q, slot;
            unpack(en[d], q, slot);
            const int4 pa = *reinterpret_cast<const int4 *>(&tb.pivot[e][0]);
            const int4 pb = *reinterpret_cast<const int4 *>(&tb.pivot[e][4]);
            const int blk = (pa.x < q) + (pa.y < q) + (pa.z < q) + (pa.w < q) + (pb.x < q) + (pb.y < q) + (pb.z < q);
            const int4 b = reinterpret_cast<const int4 *>(tb.bound[e])[blk];
            const int idx = 4 * blk - 1 + (b.x < q) + (b.y < q) + (b.z < q) + (b.w < q);  // bound[0] < q always
            const bool hit = j0 + 32 * d + lane < total && idx >= tb.low[e] && idx < tb.high[e] && idx != kChain;
            const double term = hit ? tb.term[e][idx] : 0.0;
            // the items of one tree are distinct columns; the trees of the group one after the other, in order
            const int e_first = __shfl_sync(kAll, e, 0), e_last = __reduce_max_sync(kAll, e);
            if (e_first == e_last) {
                if (hit) {
                    accW[slot0 + slot] = __dadd_rn(accW[slot0 + slot], term);
                    bump(accC, slot0 + slot);
                }
            } else {
                for (int t = e_first; t <= e_last; ++t) {
                    if (hit && e == t) {
                        accW[slot0 + slot] = __dadd_rn(accW[slot0 + slot], term);
                        bump(accC, slot0 + slot);
                    }
                    __syncwarp();
                }
            }
        }
    }
}

// blockDim.x = 32 * W threads (W = 1 << bs.warps_log2); dynamic shared memory: W * stride accumulators.
// kTri: the CTA of row a accumulates only the columns beyond a (W is symmetric bit for bit: every pair is then
// visited once instead of twice) and writes that part of the row and its mirror image; pcg_mirror_bits and
// pcg_degree_rows complete the bit rows and sum the rows.  The entries of a bucket ascend by column, so each warp finds where its share of a
// tree starts with a binary search (one lane per tree of the batch, while the first warps walk the chains).
template <typename CountT, bool kWriteC, typename EntryT, bool kTri, bool kHalf>
__global__ void __launch_bounds__(kRowThreads, 2)
pcg_rows_kernel(int n, int row0, int words_per_row, BucketShape bs, int stride,
                const int64_t *__restrict__ leaf_offsets, const LinkEntry *__restrict__ links,
                const double *__restrict__ tree_weight, const int32_t *__restrict__ leaf_tree,
                const int32_t *__restrict__ row_ptr, const int32_t *__restrict__ inv_sorted,
                const int32_t *__restrict__ occ, const int32_t *__restrict__ bucket_ptr,
                const EntryT *__restrict__ entries, const LeafStairs *__restrict__ stairs,
                const int32_t *__restrict__ start16,
                double *__restrict__ W, int32_t *__restrict__ C, uint32_t *__restrict__ adj_bits,
                uint32_t *__restrict__ max_bits, double *__restrict__ degree_part, int32_t *__restrict__ bad,
                BatchView view) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int nwarps = 1 << bs.warps_log2, nthreads = nwarps << 5;
    const int slots_total = stride << bs.warps_log2;
    double *accW = reinterpret_cast<double *>(smem_raw);
    CountT *accC = reinterpret_cast<CountT *>(accW + slots_total);
    __shared__ TreeBatch tb;
    __shared__ double warp_sum[kWarps];

    const int a = row0 + blockIdx.x;  // W / C point at the row block: its first row is row0
    // batched: `a` is a row of the global row space; the node it belongs to says how wide the row is and where
    // its outputs go (occ, row_ptr, degree are indexed by the global row)
    const int degree_stride = n;  // rows of the (global) row space: degree_part is [chunk][row]
    int occ_base = 0;
    int a_loc = a;  // the row's vertex id inside its node
    size_t w_row = static_cast<size_t>(blockIdx.x) * n;
    size_t bits_row = static_cast<size_t>(a) * words_per_row;
    if (view.nodes) {
        const MedNode &nd = view.nodes[view.row_node[a]];
        n = nd.n;
        occ_base = nd.row_base;
        a_loc = a - nd.row_base;
        w_row = static_cast<size_t>(nd.w_off) + static_cast<size_t>(a_loc) * n;
        bits_row = static_cast<size_t>(nd.bit_off) + static_cast<size_t>(a_loc) * nd.words;
    }
    const int cols_per_chunk = bs.cols_per_chunk;
    const int col0 = blockIdx.y * cols_per_chunk;
    const int ncols = min(cols_per_chunk, n - col0);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int my_bucket = (blockIdx.y << bs.warps_log2) + warp;
    const int slot0 = warp * stride;  // the slots of this warp's columns start here
    // Partial rows (kTri): the columns this CTA computes, as up to two ranges of the chunk's local columns.
    //   half_mode 0 (one GPU): the columns beyond the diagonal, (a, n);
    //   half_mode 1 (the row block of a node shared out over several GPUs): the cyclic half window (a, a + h] with
    //   h = (n - 1) / 2, one more for the rows of the first half when n is even -- of every pair {a, c} exactly
    //   one of the two rows has the other in its window, every row does half its pairs, and nothing depends on how
    //   the rows are dealt to the ranks; the owners exchange the other half afterwards (pcg_fetch_transposed).
    constexpr bool half_mode = kHalf;  // compiled separately: the one-GPU kernel carries none of the window logic
    static_assert(kTri || !kHalf, "half windows are partial rows");
    int r1b = 0, r1e = 0, r2b = 0, r2e = 0;
    if (kTri) {
        int e1 = n, e2 = 0;
        if (half_mode) {
            const int last = a_loc + (n - 1) / 2 + (((n & 1) == 0 && a_loc < n / 2) ? 1 : 0);  // inclusive
            if (last < n) e1 = last + 1;
            else e2 = last - n + 1;
        }
        r1b = max(a_loc + 1, col0) - col0;
        r1e = max(min(e1, col0 + ncols) - col0, r1b);
        r2b = 0;
        r2e = max(min(e2, col0 + ncols) - col0, 0);
        if (!half_mode && a_loc >= col0 + ncols) return;  // the whole chunk lies before the diagonal: the mirror's part
    }
    // the slots of this warp inside a local column range [lb, le): column = slot * W + warp
    auto slots_below = [&](int le) { return le > warp ? (le - warp + (1 << bs.warps_log2) - 1) >> bs.warps_log2 : 0; };

    for (int c = tid; c < slots_total; c += nthreads) {
        accW[c] = 0.0;
        accC[c] = 0;
    }

    const int ebase = row_ptr[a];
    const int cnt = row_ptr[a + 1] - ebase;
    const bool diagonal_chunk = kTri && !half_mode && a_loc >= col0;  // this chunk holds the row's own column
    const int s1b = slots_below(r1b), s1e = slots_below(r1e), s2b = slots_below(r2b), s2e = slots_below(r2e);
    const int4 *stairs4 = reinterpret_cast<const int4 *>(stairs);
    for (int e0 = 0; e0 < cnt; e0 += kTreeBatch) {
        const int batch = min(kTreeBatch, cnt - e0);
        __syncthreads();  // the previous batch (and the initialisation) is done with
        // Lane e keeps where this warp's share of tree e starts and ends, and whether the tree's chains need further
        // rounds.  Every warp reads this for itself: nothing here waits for another warp.
        int my_ptr = 0, my_end = 0, my_ptr2 = 0, my_end2 = 0;
        bool my_open = false;
        if (lane < batch) {
            const int g = inv_sorted[ebase + e0 + lane];
            const int4 shape = stairs4[static_cast<size_t>(g) * kStairsInt4 + 26];  // low, high, resume0, resume1
            const int4 info = stairs4[static_cast<size_t>(g) * kStairsInt4 + 27];   // leaves, position, tree
            const size_t cell = static_cast<size_t>(info.z) * bs.buckets + my_bucket;
            const int bucket_begin = bucket_ptr[cell], bucket_end = bucket_ptr[cell + 1];
            if (kTri && half_mode) {
                // the entries of a bucket ascend by slot: the two ranges of wanted slots by binary search
                auto first_at_least = [&](int want) {
                    int lo = bucket_begin, hi = bucket_end;
                    while (lo < hi) {
                        const int mid = (lo + hi) >> 1;
                        int q, sl;
                        unpack(entries[mid], q, sl);
                        if (sl < want) lo = mid + 1; else hi = mid;
                    }
                    return lo;
                };
                my_ptr = s1e > s1b ? first_at_least(s1b) : bucket_begin;
                my_end = s1e > s1b ? first_at_least(s1e) : bucket_begin;
                my_ptr2 = s2e > s2b ? first_at_least(s2b) : bucket_begin;
                my_end2 = s2e > s2b ? first_at_least(s2e) : bucket_begin;
            } else {
                my_end = bucket_end;
                // one GPU, the chunk with the row's own column: the warp's entries beyond the diagonal start where
                // pcg_bucket_sorted found them for this leaf; later chunks lie beyond the diagonal altogether
                my_ptr = diagonal_chunk ? start16[static_cast<size_t>(g) * kWarps + warp] : bucket_begin;
            }
            if (info.x == 0) my_end = my_ptr, my_end2 = my_ptr2;
            my_open = shape.z >= 0 || shape.w >= 0;
        }
        const bool second_range = kTri && half_mode && s2e > s2b;  // warp-uniform
        const uint32_t unfinished = __ballot_sync(0xffffffffu, my_open);
        // the finished staircases of the batch's trees (pcg_leaf_stairs): 28 int4 each, coalesced
        for (int idx = tid; idx < batch * kStairsInt4; idx += nthreads) {
            const int e = idx / kStairsInt4, k = idx - e * kStairsInt4;
            const int g = inv_sorted[ebase + e0 + e];
            const int4 v = stairs4[static_cast<size_t>(g) * kStairsInt4 + k];
            if (k < 16) reinterpret_cast<int4 *>(tb.term[e])[k] = v;
            else if (k < 24) reinterpret_cast<int4 *>(tb.bound[e])[k - 16] = v;
            else if (k < 26) reinterpret_cast<int4 *>(tb.pivot[e])[k - 24] = v;
            else if (k == 26) {
                tb.low[e] = v.x;
                tb.high[e] = v.y;
                tb.resume[e][0] = v.z;
                tb.resume[e][1] = v.w;
            } else {
                TreeHeader h;  // only the continuation rounds of an unfinished chain read it
                h.leaves = v.x;
                h.position = v.y;
                h.tree = v.z;
                h.pad = 0;
                h.base = leaf_offsets[v.z];
                h.weight = tree_weight[v.z];
                tb.header[e] = h;
            }
        }
        __syncthreads();
        // Every warp goes through the trees in input order on its own: no barrier between trees.  A tree whose
        // chain did not fit in kChain steps is continued round by round, CTA-wide, before the trees after it.
        int first = 0;
        while (first < batch) {
            const uint32_t open = unfinished >> first;
            const int stop = open != 0 ? first + __ffs(open) : batch;
            visit_trees(tb, first, stop, my_ptr, my_end, lane, slot0, entries, accW, accC);
            if (second_range) visit_trees(tb, first, stop, my_ptr2, my_end2, lane, slot0, entries, accW, accC);
            if (open != 0) {
                const int eu = stop - 1;
                const int ptr = __shfl_sync(0xffffffffu, my_ptr, eu), end = __shfl_sync(0xffffffffu, my_end, eu);
                const int ptr2 = __shfl_sync(0xffffffffu, my_ptr2, eu), end2 = __shfl_sync(0xffffffffu, my_end2, eu);
                while (tb.resume[eu][0] >= 0 || tb.resume[eu][1] >= 0) {
                    __syncthreads();  // everybody has read resume[] and is done with the tree's segments
                    if (tid < 2) walk_chain(tb, eu, tid, false, links);
                    __syncthreads();
                    visit_bucket(tb, eu, ptr, end, lane, slot0, entries, accW, accC);
                    if (second_range) visit_bucket(tb, eu, ptr2, end2, lane, slot0, entries, accW, accC);
                }
            }
            first = stop;
        }
    }
    __syncthreads();

    // ---- write the finished row --------------------------------------------------------------
    const int wmask = nwarps - 1;
    auto slot_of = [&](int c) { return (c & wmask) * stride + (c >> bs.warps_log2); };
    const int occ_a = occ[a];
    double *Wrow = W + w_row + col0;
    double *Wnode = W + (w_row - static_cast<size_t>(a_loc) * n);  // triangle mode: row 0 of the node's matrix
    for (int c = tid; c < ncols; c += nthreads) {
        // partial rows: only what this CTA computed (and the zero on the diagonal); the rest arrives as a mirror image
        if (kTri && !((c >= r1b && c < r1e) || (c >= r2b && c < r2e) || col0 + c == a_loc)) continue;
        const double v = accW[slot_of(c)];
        Wrow[c] = v;
        if (kTri && !half_mode && col0 + c > a_loc) Wnode[static_cast<size_t>(col0 + c) * n + a_loc] = v;
        if (kWriteC) C[w_row + col0 + c] = static_cast<int32_t>(accC[slot_of(c)]);
    }
    const int word0 = col0 >> 5;
    const int nwords = (ncols + 31) >> 5;
    for (int j = warp; j < nwords; j += nwarps) {
        const int c = (j << 5) + lane;
        bool edge = false, top = false;
        if (c < ncols) {
            const int cc = static_cast<int>(accC[slot_of(c)]);
            edge = cc > 0;
            if (edge && max_bits != nullptr) top = cc == max(occ_a, occ[occ_base + col0 + c]);
        }
        const uint32_t eb = __ballot_sync(0xffffffffu, edge);
        const uint32_t tb2 = __ballot_sync(0xffffffffu, top);
        if (lane == 0) {
            adj_bits[bits_row + word0 + j] = eb;
            if (max_bits != nullptr) max_bits[bits_row + word0 + j] = tb2;
        }
    }
    if (kTri) return;  // the row is complete only when every row CTA is done: pcg_degree_rows sums it (in the same order)
    // row sum in a fixed order (the same for every CTA size): virtual thread v < kRowThreads sums columns
    // v, v + kRowThreads, ...; shuffle tree per virtual warp; then the virtual warps in order
    for (int vw = warp; vw < kWarps; vw += nwarps) {
        double part = 0.0;
        for (int c = (vw << 5) + lane; c < ncols; c += kRowThreads) part += accW[slot_of(c)];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) part += __shfl_down_sync(0xffffffffu, part, off);
        if (lane == 0) warp_sum[vw] = part;
    }
    __syncthreads();
    if (tid == 0) {
        double s = 0.0;
        for (int wi = 0; wi < kWarps; ++wi) s += warp_sum[wi];
        degree_part[static_cast<size_t>(blockIdx.y) * degree_stride + a] = s;
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

// Triangle mode, second half.  The row kernel has written both triangles of W (every finished entry also to its
// mirror position: a scattered 8-byte store per entry, which L2 merges with the stores of the neighbouring rows'
// CTAs).  What is left: the bit matrices -- pcg_mirror_bits, one warp per 32 x 32 tile (I, J), I <= J, transposes the
// tile with 32 ballots (the diagonal word keeps its upper bits) -- and the row sums -- pcg_degree_rows, one warp per
// row, in exactly the order of the full-row kernel (virtual thread v < 512 takes columns v, v + 512, ... of a column
// chunk; shuffle tree per virtual warp; the 16 virtual warps in order; the chunks in order), so W, the bits and the
// degrees do not depend on the mode.
constexpr int kMirrorWarps = 8;

__global__ void __launch_bounds__(kMirrorWarps * 32)
pcg_mirror_bits(int n, int words, int B, int both, uint32_t *__restrict__ adj_bits, uint32_t *__restrict__ max_bits,
                BatchView view) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int J = blockIdx.x;
    if (view.nodes) {
        const int blk = blockIdx.x;
        int lo = 0, hi = B;  // last node with blk_base <= blk
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (view.nodes[mid].blk_base <= blk) lo = mid; else hi = mid;
        }
        const MedNode &nd = view.nodes[lo];
        J = blk - nd.blk_base;
        n = nd.n;
        words = nd.words;
        adj_bits += nd.bit_off;
        if (max_bits) max_bits += nd.bit_off;
    }
    const int I = blockIdx.y * kMirrorWarps + warp;
    if (I > J) return;
    const int srow = (I << 5) + lane, rowJ = (J << 5) + lane;
    for (int which = 0; which < 2; ++which) {
        uint32_t *bits = which == 0 ? adj_bits : max_bits;
        if (!bits) continue;
        const uint32_t mine = srow < n ? bits[static_cast<size_t>(srow) * words + J] : 0u;
        // both: either triangle may hold bits (half-window rows of a shared node): the symmetric closure
        const uint32_t theirs = both && I < J && rowJ < n ? bits[static_cast<size_t>(rowJ) * words + I] : 0u;
        uint32_t turned = 0u, turned_back = 0u;
#pragma unroll
        for (int rp = 0; rp < 32; ++rp) {
            const uint32_t b = __ballot_sync(0xffffffffu, (mine >> rp) & 1u);
            if (lane == rp) turned = b;
        }
        if (both && I < J) {
#pragma unroll
            for (int rp = 0; rp < 32; ++rp) {
                const uint32_t b = __ballot_sync(0xffffffffu, (theirs >> rp) & 1u);
                if (lane == rp) turned_back = b;
            }
            if (srow < n) bits[static_cast<size_t>(srow) * words + J] = mine | turned_back;
        }
        if (rowJ < n) {
            if (I < J) bits[static_cast<size_t>(rowJ) * words + I] = theirs | turned;
            else bits[static_cast<size_t>(rowJ) * words + J] = mine | turned;  // srow == rowJ: its own upper bits
        }
    }
}

__global__ void __launch_bounds__(kMirrorWarps * 32)
pcg_degree_rows(int n, int R, int cols_per_chunk, const double *__restrict__ W, double *__restrict__ degree, BatchView view) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * kMirrorWarps + warp;  // of the (global) row space
    if (row >= R) return;
    const double *Wrow = W + static_cast<size_t>(row) * n;
    if (view.nodes) {
        const MedNode &nd = view.nodes[view.row_node[row]];
        n = nd.n;
        cols_per_chunk = n;
        Wrow = W + nd.w_off + static_cast<size_t>(row - nd.row_base) * n;
    }
    const int nchunks = (n + cols_per_chunk - 1) / cols_per_chunk;
    double total = 0.0;
    for (int y = 0; y < nchunks; ++y) {
        const int col0 = y * cols_per_chunk;
        const int ncols = min(cols_per_chunk, n - col0);
        double s = 0.0;
        for (int vw = 0; vw < kWarps; ++vw) {
            double part = 0.0;
            for (int c = (vw << 5) + lane; c < ncols; c += kRowThreads) part += Wrow[col0 + c];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) part += __shfl_down_sync(0xffffffffu, part, off);
            s += part;  // lane 0 holds the virtual warp's sum; the other lanes add along harmlessly
        }
        total = nchunks == 1 ? s : total + s;
    }
    if (lane == 0) degree[row] = total;
}

// Half-window row blocks, second half: element (a, c) of this rank's rows that lies in the window of row c rather than
// of row a was computed by the owner of row c as (c, a).  One warp per 32 x 32 tile: the source rows are read from
// the owners' blocks (over NVLink for a peer; 256 contiguous bytes per row), turned in shared memory, and stored.
struct PeerBlocks {
    const double *W[kMaxPeers];
    int rows_per_rank;
};

__device__ __forceinline__ bool in_half_window(int n, int from, int to) {  // is column `to` in the window of row `from`
    int d = to - from;
    if (d < 0) d += n;
    const int h = (n - 1) / 2 + (((n & 1) == 0 && from < n / 2) ? 1 : 0);
    return d >= 1 && d <= h;
}

constexpr size_t kFetchSmem = sizeof(double) * kMirrorWarps * 32 * 33;

__global__ void __launch_bounds__(kMirrorWarps * 32)
pcg_fetch_transposed_kernel(int n, int row0, int row1, PeerBlocks peers, double *__restrict__ W_block) {
    extern __shared__ __align__(16) unsigned char fetch_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double(*tile)[33] = reinterpret_cast<double(*)[33]>(fetch_raw) + static_cast<size_t>(warp) * 32;
    const int a0 = row0 + (static_cast<int>(blockIdx.x) << 5);
    const int c0 = (static_cast<int>(blockIdx.y) * kMirrorWarps + warp) << 5;
    if (c0 >= n) return;
    // is anything of this tile the other rows' part?  (corners do not decide it for a cyclic window: test every row)
    bool any = false;
    const int a_mine = a0 + lane;
    for (int r = 0; r < 32; ++r) {
        const int c = c0 + r;
        const bool need = c < n && a_mine < row1 && in_half_window(n, c, a_mine);
        const unsigned mask = __ballot_sync(0xffffffffu, need);
        if (mask) {
            any = true;
            if (need) {
                const int owner = c / peers.rows_per_rank;
                tile[r][lane] = peers.W[owner][static_cast<size_t>(c - owner * peers.rows_per_rank) * n + a_mine];
            }
        }
    }
    if (!any) return;  // warp-uniform
    __syncwarp();
    const int c_mine = c0 + lane;
    for (int ra = 0; ra < 32; ++ra) {
        const int a = a0 + ra;
        if (a >= row1) break;
        if (c_mine < n && in_half_window(n, c_mine, a))
            W_block[static_cast<size_t>(a - row0) * n + c_mine] = tile[lane][ra];
    }
}

template <typename CountT, bool kWriteC, typename EntryT, bool kTri, bool kHalf = false>
int launch_rows(scs_ctx *ctx, int n, int row0, int nrows, int words, BucketShape bs, int stride, int nchunks, size_t smem,
                const int64_t *leaf_offsets, const LinkEntry *links, const double *tree_weight,
                const int32_t *leaf_tree, const int32_t *row_ptr, const int32_t *inv_sorted,
                const int32_t *occ, const int32_t *bucket_ptr, const void *entries, const LeafStairs *stairs,
                const int32_t *start16, double *W, int32_t *C,
                uint32_t *adj_bits, uint32_t *max_bits, double *degree_part, int32_t *bad, BatchView batch = BatchView()) {
    auto kernel = pcg_rows_kernel<CountT, kWriteC, EntryT, kTri, kHalf>;
    // always the same (maximal) opt-in size: contexts on other host threads launch this kernel concurrently
    bool &configured = ctx->rows_configured[(sizeof(CountT) == 2 ? 0 : 2) + (kWriteC ? 1 : 0) +
                                            (sizeof(EntryT) == 4 ? 0 : 4) + (kTri ? 8 : 0) + (kHalf ? 16 : 0)];
    if (!configured) {
        const size_t optin = ctx->smem_optin > 2 * kRowsStaticSmem ? ctx->smem_optin - kRowsStaticSmem : 32 * 1024;
        SCS_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(optin)));
        configured = true;
    }
    dim3 grid(nrows, nchunks);
    if (n >= kProfileMinSize && !batch.nodes) {
        // algorithmic bytes: W + both bit matrices + occ/degree written once, C if asked
        const double out_bytes = 8.0 * nrows * n + (C ? 4.0 * nrows * n : 0.0) + 8.0 * nrows * words + 12.0 * nrows;
        profile_begin(ctx, PROFILE_PCG_ROWS, out_bytes, ctx->pending_units);
    }
    kernel<<<grid, 32 << bs.warps_log2, smem, ctx->stream>>>(
        n, row0, words, bs, stride, leaf_offsets, links, tree_weight, leaf_tree, row_ptr, inv_sorted, occ, bucket_ptr,
        static_cast<const EntryT *>(entries), stairs, start16, W, C, adj_bits, max_bits, degree_part, bad, batch);
    if (n >= kProfileMinSize && !batch.nodes) profile_end(ctx);
    SCS_LAUNCHED(ctx, "pcg_rows_kernel");
    return SCS_OK;
}

}  // namespace

// Long inputs: per-tile sums (kScanTile elements per CTA), a one-CTA scan of the tile sums, then every tile scanned
// again from its offset.  Integer sums: any order gives the same result.
constexpr int kScanTile = 4096;

__global__ void __launch_bounds__(1024) scan_tile_sums(int n, const int32_t *__restrict__ in, int32_t *__restrict__ sums) {
    __shared__ int32_t warp_part[32];
    const int base = blockIdx.x * kScanTile;
    int32_t acc = 0;
    for (int i = threadIdx.x; i < kScanTile; i += 1024)
        if (base + i < n) acc += in[base + i];
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, off);
    if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        int32_t v = warp_part[threadIdx.x];
        for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
        if (threadIdx.x == 0) sums[blockIdx.x] = v;
    }
}

// thread t owns elements [4t, 4t + 4) of the tile
__global__ void __launch_bounds__(1024) scan_tiles(int n, const int32_t *__restrict__ in, const int32_t *__restrict__ tile_offset,
                                                  int32_t *__restrict__ out) {
    __shared__ int32_t warp_part[32];
    const int base = blockIdx.x * kScanTile + threadIdx.x * 4;
    int32_t v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = base + k < n ? in[base + k] : 0;
    const int32_t mine = v[0] + v[1] + v[2] + v[3];
    int32_t inc = mine;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int off = 1; off < 32; off <<= 1) {
        const int32_t o = __shfl_up_sync(0xffffffffu, inc, off);
        if (lane >= off) inc += o;
    }
    if (lane == 31) warp_part[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int32_t w = warp_part[lane];
        for (int off = 1; off < 32; off <<= 1) {
            const int32_t o = __shfl_up_sync(0xffffffffu, w, off);
            if (lane >= off) w += o;
        }
        warp_part[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    int32_t run = tile_offset[blockIdx.x] + (warp ? warp_part[warp - 1] : 0) + inc - mine;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (base + k < n) out[base + k] = run;
        run += v[k];
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 1023) out[n] = tile_offset[gridDim.x];
}

int exclusive_scan(scs_ctx *ctx, int n, const int32_t *in, int32_t *out) {
    if (n <= 2 * kScanTile) {
        exclusive_scan_i32<<<1, 1024, 0, ctx->stream>>>(n, in, out);
        SCS_LAUNCHED(ctx, "exclusive_scan_i32");
        return SCS_OK;
    }
    const int tiles = ceil_div(n, kScanTile);
    int32_t *sums;
    int rc;
    if ((rc = reserve_as(ctx, SLOT_SCAN_SUMS, 2 * static_cast<size_t>(tiles) + 2, &sums))) return rc;
    int32_t *offsets = sums + tiles;  // tiles + 1 entries
    scan_tile_sums<<<tiles, 1024, 0, ctx->stream>>>(n, in, sums);
    SCS_LAUNCHED(ctx, "scan_tile_sums");
    exclusive_scan_i32<<<1, 1024, 0, ctx->stream>>>(tiles, sums, offsets);
    SCS_LAUNCHED(ctx, "exclusive_scan_i32");
    scan_tiles<<<tiles, 1024, 0, ctx->stream>>>(n, in, offsets, out);
    SCS_LAUNCHED(ctx, "scan_tiles");
    return SCS_OK;
}

namespace {

// Sorted buckets of every tree (pcg_bucket_sorted): bucket_ptr[T * buckets + 1], entries[L].
int build_buckets(scs_ctx *ctx, int n, int T, int64_t L, const BucketShape &bs, int stride, bool packed,
                  const int64_t *leaf_offsets, const int32_t *leaf_taxon, int32_t *bucket_ptr, void *entries,
                  int32_t *start16, int32_t *bad, const BatchView &batch) {
    if (T <= 0 || L <= 0) {
        SCS_CUDA(ctx, cudaMemsetAsync(bucket_ptr, 0, sizeof(int32_t) * (static_cast<size_t>(T > 0 ? T : 0) * bs.buckets + 1),
                                      ctx->stream));
        return SCS_OK;
    }
    const int key_stride = (stride + 31) & ~31;
    const int words = bs.buckets * (key_stride >> 5);
    const size_t per_warp = 2 * sizeof(uint32_t) * static_cast<size_t>(words);
    if (per_warp + 1024 > ctx->smem_optin) return fail(ctx, SCS_ERR_INVALID, "graph build: too many taxa for the bucket index");
    int warps = static_cast<int>((40 * 1024) / per_warp);
    warps = warps < 1 ? 1 : (warps > 8 ? 8 : warps);
    const size_t smem = per_warp * warps;
    auto launch = [&](auto kernel, auto *typed) -> int {
        if (smem > 48 * 1024)
            SCS_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        kernel<<<ceil_div(T, warps), warps * 32, smem, ctx->stream>>>(n, T, bs, key_stride, words, leaf_offsets, leaf_taxon,
                                                                    bucket_ptr, typed, start16, bad, batch);
        SCS_LAUNCHED(ctx, "pcg_bucket_sorted");
        return SCS_OK;
    };
    if (packed) return launch(pcg_bucket_sorted<uint32_t>, static_cast<uint32_t *>(entries));
    return launch(pcg_bucket_sorted<unsigned long long>, static_cast<unsigned long long *>(entries));
}

int launch_mirror(scs_ctx *ctx, int n, int words, int cols_per_chunk, int total_blocks, int max_blocks, int B, int R,
                  const double *W, uint32_t *adj_bits, uint32_t *max_bits, double *degree, const BatchView &batch) {
    if (total_blocks <= 0 || R <= 0) return SCS_OK;
    dim3 grid(total_blocks, ceil_div(max_blocks, kMirrorWarps));
    pcg_mirror_bits<<<grid, kMirrorWarps * 32, 0, ctx->stream>>>(n, words, B, 0, adj_bits, max_bits, batch);
    SCS_LAUNCHED(ctx, "pcg_mirror_bits");
    pcg_degree_rows<<<ceil_div(R, kMirrorWarps), kMirrorWarps * 32, 0, ctx->stream>>>(n, R, cols_per_chunk, W, degree, batch);
    SCS_LAUNCHED(ctx, "pcg_degree_rows");
    return SCS_OK;
}

// column chunking of the row kernel: the whole row if it fits in shared memory, else equal chunks of 32-multiples
int chunk_columns(const scs_ctx *ctx, int n, int T) {
    const bool narrow = T < 65536;
    const size_t per_col = sizeof(double) + (narrow ? sizeof(uint16_t) : sizeof(int32_t));
    const size_t budget = ctx->smem_optin > 2 * kRowsStaticSmem ? ctx->smem_optin - kRowsStaticSmem : 32 * 1024;
    const int warps = 16;
    const size_t row_static = sizeof(TreeBatch) + 256;  // + warp_sum, alignment
    const size_t half_sm = ctx->smem_per_sm > 4 * kRowsStaticSmem ? ctx->smem_per_sm / 2 - 1024 - row_static : budget;
    const size_t chunk_budget = half_sm < budget ? half_sm : budget;
    const int max_cols = static_cast<int>(((chunk_budget - 16) / per_col - 2 * warps) / 32 * 32);
    const int padded = scs_bit_words(n) * 32;
    const int nchunks = ceil_div(padded, max_cols);
    return ceil_div(ceil_div(padded, nchunks), 32) * 32;
}

}  // namespace

int pcg_fetch_transposed(scs_ctx *ctx, int n, RowBlock rows, int rows_per_rank, int world, const double *const *peer_W,
                         double *W_block) {
    const int nrows = rows.row1 - rows.row0;
    if (nrows <= 0) return SCS_OK;
    if (world > kMaxPeers || rows_per_rank <= 0) return fail(ctx, SCS_ERR_INVALID, "pcg_fetch_transposed: bad argument");
    PeerBlocks peers;
    for (int r = 0; r < kMaxPeers; ++r) peers.W[r] = r < world ? peer_W[r] : nullptr;
    peers.rows_per_rank = rows_per_rank;
    if (!ctx->mirror_configured) {
        SCS_CUDA(ctx, cudaFuncSetAttribute(pcg_fetch_transposed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           static_cast<int>(kFetchSmem)));
        ctx->mirror_configured = true;
    }
    dim3 grid(ceil_div(nrows, 32), ceil_div(ceil_div(n, 32), kMirrorWarps));
    pcg_fetch_transposed_kernel<<<grid, kMirrorWarps * 32, kFetchSmem, ctx->stream>>>(n, rows.row0, rows.row1, peers, W_block);
    SCS_LAUNCHED(ctx, "pcg_fetch_transposed_kernel");
    return SCS_OK;
}

int pcg_symmetrize_bits(scs_ctx *ctx, int n, uint32_t *adj_bits, uint32_t *max_bits) {
    const int words = scs_bit_words(n);
    dim3 grid(words, ceil_div(words, kMirrorWarps));
    pcg_mirror_bits<<<grid, kMirrorWarps * 32, 0, ctx->stream>>>(n, words, 1, 1, adj_bits, max_bits, BatchView());
    SCS_LAUNCHED(ctx, "pcg_mirror_bits");
    return SCS_OK;
}

int pcg_degree_block(scs_ctx *ctx, int n, int T, RowBlock rows, const double *W_block, double *degree) {
    const int nrows = rows.row1 - rows.row0;
    if (nrows <= 0) return SCS_OK;
    pcg_degree_rows<<<ceil_div(nrows, kMirrorWarps), kMirrorWarps * 32, 0, ctx->stream>>>(n, nrows, chunk_columns(ctx, n, T), W_block,
                                                                                        degree + rows.row0, BatchView());
    SCS_LAUNCHED(ctx, "pcg_degree_rows");
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
                                                                   scalars, BatchView());
        SCS_LAUNCHED(ctx, "pcg_index_leaves");
    }
    if ((rc = exclusive_scan(ctx, n, occ, row_ptr))) return rc;
    if (L > 0) {
        pcg_fill_inverse<<<ceil_div(L, 256), 256, 0, ctx->stream>>>(n, L, leaf_taxon, leaf_tree, row_ptr, cursor, inv,
                                                                   BatchView());
        SCS_LAUNCHED(ctx, "pcg_fill_inverse");
        if (nrows > 0) {
            pcg_sort_inverse<<<nrows, 128, 0, ctx->stream>>>(row0, row_ptr, inv, inv_sorted);
            SCS_LAUNCHED(ctx, "pcg_sort_inverse");
        }
    }
    LinkEntry *links;
    if ((rc = reserve_as(ctx, SLOT_LINKS, static_cast<size_t>(L > 0 ? L : 1), &links))) return rc;
    if (L > 0 && T > 0) {
        // block minima of a tree's tour: a tree has at most n leaves (distinct taxa; checked in the kernel)
        const int nb1 = ceil_div(n, 32), nb2 = ceil_div(nb1, 32);
        const size_t link_smem = sizeof(int32_t) * (static_cast<size_t>(nb1) + nb2 + 2);
        if (link_smem > 48 * 1024)
            SCS_CUDA(ctx, cudaFuncSetAttribute(pcg_tour_links, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                               static_cast<int>(link_smem)));
        pcg_tour_links<<<T, 256, link_smem, ctx->stream>>>(n, leaf_offsets, adj_depth, adj_val, root_depth, links);
        SCS_LAUNCHED(ctx, "pcg_tour_links");
    }

    // column chunking: the whole row if it fits in shared memory, else equal chunks of 32-multiples; inside a
    // chunk the columns are dealt out to the CTA's warps (slot = warp * stride + column / warps)
    const bool narrow = T < 65536;
    const size_t per_col = sizeof(double) + (narrow ? sizeof(uint16_t) : sizeof(int32_t));
    BucketShape bs;
    bs.warps_log2 = 4;
    const int warps = 1 << bs.warps_log2;
    // two CTAs per SM hide each other's latencies: a chunk gets at most half an SM's shared memory (a chunk CTA
    // reads only its own share of every tree, so more chunks cost little); slots = warps * stride <= cols + 2 * warps
    // (stride rounded up and made odd)
    bs.cols_per_chunk = chunk_columns(ctx, n, T);
    const int nchunks = ceil_div(n, bs.cols_per_chunk);
    bs.buckets = nchunks * warps;
    const int stride = ceil_div(bs.cols_per_chunk, warps) | 1;  // odd: the write-out reads a column per lane
    const size_t smem = static_cast<size_t>(stride) * warps * per_col + 16;
    if ((rc = reserve_as(ctx, SLOT_DEGREE_PART, static_cast<size_t>(n) * nchunks, &degree_part))) return rc;
    if (nrows == 0) return SCS_OK;

    // buckets: every tree's entries sorted by (bucket, column)
    const bool packed = n <= 65535 && !ctx->wide_entries;  // tour positions and slots fit 16 bits each
    const size_t cells = static_cast<size_t>(T) * bs.buckets;
    if (cells + 1 >= (1ull << 31)) return fail(ctx, SCS_ERR_INVALID, "pcg_build: too many tree buckets");
    int32_t *bucket_ptr;
    void *entries;
    if ((rc = reserve_as(ctx, SLOT_BUCKET_PTR, cells + 2, &bucket_ptr))) return rc;
    if ((rc = reserve(ctx, SLOT_ENTRIES, (static_cast<size_t>(L) + 1) * (packed ? 4 : 8), &entries))) return rc;
    // every pair once (upper triangle + mirror) unless the rows are one rank's block of a sharded node (the other
    // triangle lives on the peers) or the caller wants the co-occurrence matrix as well
    const bool tri = !rows.sharded() && C == nullptr && !ctx->full_rows;
    // ... or, for a row block, the cyclic half window of every row (the caller completes the block from the peers')
    const bool half = rows.sharded() && rows.half_window && C == nullptr && !ctx->full_rows;
    int32_t *start16 = nullptr;
    if (tri && (rc = reserve_as(ctx, SLOT_START16, (static_cast<size_t>(L) + 1) * kWarps, &start16))) return rc;
    if ((rc = build_buckets(ctx, n, T, L, bs, stride, packed, leaf_offsets, leaf_taxon, bucket_ptr, entries, start16, scalars,
                            BatchView())))
        return rc;
    // the staircase of every leaf, once for the whole node
    LeafStairs *stairs;
    if ((rc = reserve_as(ctx, SLOT_BUCKET_COUNT, static_cast<size_t>(L) + 1, &stairs))) return rc;
    if (L > 0) {
        pcg_leaf_stairs<<<ceil_div(L, kStairsLeaves), 2 * kStairsLeaves, 0, ctx->stream>>>(n, L, leaf_offsets, leaf_tree, links, tree_weight, stairs,
                                                                      scalars, BatchView());
        SCS_LAUNCHED(ctx, "pcg_leaf_stairs");
    }

#define SCS_ROWS(CT, WC, ET, TRI, HALF)                                                                                  \
    launch_rows<CT, WC, ET, TRI, HALF>(ctx, n, row0, nrows, words, bs, stride, nchunks, smem, leaf_offsets, links,       \
                                       tree_weight,                                                                      \
                                 leaf_tree, row_ptr, inv_sorted, occ, bucket_ptr, entries, stairs, start16, W, C, adj_bits, \
                                 max_bits, degree_part, scalars)
#define SCS_ROWS_E(CT, WC, TRI, HALF) \
    (packed ? SCS_ROWS(CT, WC, uint32_t, TRI, HALF) : SCS_ROWS(CT, WC, unsigned long long, TRI, HALF))
    if (half) rc = narrow ? SCS_ROWS_E(uint16_t, false, true, true) : SCS_ROWS_E(int32_t, false, true, true);
    else if (tri) rc = narrow ? SCS_ROWS_E(uint16_t, false, true, false) : SCS_ROWS_E(int32_t, false, true, false);
    else if (narrow) rc = C ? SCS_ROWS_E(uint16_t, true, false, false) : SCS_ROWS_E(uint16_t, false, false, false);
    else rc = C ? SCS_ROWS_E(int32_t, true, false, false) : SCS_ROWS_E(int32_t, false, false, false);
#undef SCS_ROWS_E
#undef SCS_ROWS
    if (rc) return rc;
    if (half) return SCS_OK;  // W, the bits and the degrees are completed by the caller once every rank is this far
    if (tri) {
        // the other triangle of W and of the bit matrices, and the row sums (into `degree`, or a scratch if unwanted)
        return launch_mirror(ctx, n, words, bs.cols_per_chunk, words, words, 1, n, W, adj_bits, max_bits,
                             degree ? degree : degree_part, BatchView());
    }
    if (degree) {
        pcg_sum_degree_parts<<<ceil_div(nrows, 256), 256, 0, ctx->stream>>>(n, row0, row1, nchunks, degree_part, degree);
        SCS_LAUNCHED(ctx, "pcg_sum_degree_parts");
    }
    return SCS_OK;
}


// The same pipeline over a batch of nodes (see BatchView): index, inverse lists, links, buckets and ONE launch of
// the row kernel with a CTA per row of the global row space.  Every node fits one column chunk.
int pcg_build_batch(scs_ctx *ctx, int B, int blocks, int R, int T, int64_t L, int max_n, int max_trees, const MedNode *nodes_dev,
                    const int32_t *tree_node, const int32_t *row_node, const int64_t *leaf_offsets,
                    const int32_t *leaf_taxon, const int32_t *adj_depth, const double *adj_val,
                    const int32_t *root_depth, const double *tree_weight, double *W, int32_t *occ, uint32_t *adj_bits,
                    uint32_t *max_bits, double *degree, int32_t *bad_dev) {
    if (R <= 0 || T < 0 || L < 0 || L >= (1ll << 31) || max_n <= 0 || max_n > 65535 || !nodes_dev || !tree_node ||
        !row_node || !W || !occ || !adj_bits || !degree || !bad_dev)
        return fail(ctx, SCS_ERR_INVALID, "pcg_build_batch: bad argument");
    BatchView batch;
    batch.nodes = nodes_dev;
    batch.tree_node = tree_node;
    batch.row_node = row_node;
    int32_t *leaf_tree, *row_ptr, *cursor, *inv, *inv_sorted;
    int rc;
    if ((rc = reserve_as(ctx, SLOT_LEAF_TREE, static_cast<size_t>(L) + 1, &leaf_tree))) return rc;
    if ((rc = reserve_as(ctx, SLOT_ROW_PTR, static_cast<size_t>(R) + 1, &row_ptr))) return rc;
    if ((rc = reserve_as(ctx, SLOT_CURSOR, static_cast<size_t>(R), &cursor))) return rc;
    if ((rc = reserve_as(ctx, SLOT_INV, static_cast<size_t>(L) + 1, &inv))) return rc;
    if ((rc = reserve_as(ctx, SLOT_INV_SORTED, static_cast<size_t>(L) + 1, &inv_sorted))) return rc;
    SCS_CUDA(ctx, cudaMemsetAsync(occ, 0, sizeof(int32_t) * R, ctx->stream));
    SCS_CUDA(ctx, cudaMemsetAsync(cursor, 0, sizeof(int32_t) * R, ctx->stream));
    if (L > 0) {
        pcg_index_leaves<<<ceil_div(L, 256), 256, 0, ctx->stream>>>(max_n, T, L, leaf_offsets, leaf_taxon, leaf_tree, occ,
                                                                   bad_dev, batch);
        SCS_LAUNCHED(ctx, "pcg_index_leaves");
    }
    if ((rc = exclusive_scan(ctx, R, occ, row_ptr))) return rc;
    if (L > 0) {
        pcg_fill_inverse<<<ceil_div(L, 256), 256, 0, ctx->stream>>>(max_n, L, leaf_taxon, leaf_tree, row_ptr, cursor, inv,
                                                                   batch);
        SCS_LAUNCHED(ctx, "pcg_fill_inverse");
        pcg_sort_inverse<<<R, 128, 0, ctx->stream>>>(0, row_ptr, inv, inv_sorted);
        SCS_LAUNCHED(ctx, "pcg_sort_inverse");
    }
    LinkEntry *links;
    if ((rc = reserve_as(ctx, SLOT_LINKS, static_cast<size_t>(L > 0 ? L : 1), &links))) return rc;
    if (L > 0 && T > 0) {
        // a tree of a node has at most max_n leaves (more: a repeated taxon, flagged by the row kernel)
        const int nb1 = ceil_div(max_n + 1, 32), nb2 = ceil_div(nb1, 32);
        const size_t link_smem = sizeof(int32_t) * (static_cast<size_t>(nb1) + nb2 + 2);
        pcg_tour_links<<<T, 256, link_smem, ctx->stream>>>(max_n + 1, leaf_offsets, adj_depth, adj_val, root_depth, links);
        SCS_LAUNCHED(ctx, "pcg_tour_links");
    }
    const bool narrow = max_trees < 65536;
    const size_t per_col = sizeof(double) + (narrow ? sizeof(uint16_t) : sizeof(int32_t));
    BucketShape bs;
    bs.warps_log2 = 4;
    const int warps = 1 << bs.warps_log2;
    bs.cols_per_chunk = ceil_div(max_n, 32) * 32;  // one chunk per row
    bs.buckets = warps;
    const int stride = ceil_div(bs.cols_per_chunk, warps) | 1;
    const size_t smem = static_cast<size_t>(stride) * warps * per_col + 16;
    const size_t cells = static_cast<size_t>(T) * bs.buckets;
    if (cells + 1 >= (1ull << 31)) return fail(ctx, SCS_ERR_INVALID, "pcg_build_batch: too many tree buckets");
    int32_t *bucket_ptr;
    void *entries;
    if ((rc = reserve_as(ctx, SLOT_BUCKET_PTR, cells + 2, &bucket_ptr))) return rc;
    if ((rc = reserve(ctx, SLOT_ENTRIES, (static_cast<size_t>(L) + 1) * 4, &entries))) return rc;
    int32_t *start16 = nullptr;
    if (!ctx->full_rows && (rc = reserve_as(ctx, SLOT_START16, (static_cast<size_t>(L) + 1) * kWarps, &start16))) return rc;
    if ((rc = build_buckets(ctx, max_n, T, L, bs, stride, true, leaf_offsets, leaf_taxon, bucket_ptr, entries, start16, bad_dev,
                            batch)))
        return rc;
    LeafStairs *stairs;
    if ((rc = reserve_as(ctx, SLOT_BUCKET_COUNT, static_cast<size_t>(L) + 1, &stairs))) return rc;
    if (L > 0) {
        pcg_leaf_stairs<<<ceil_div(L, kStairsLeaves), 2 * kStairsLeaves, 0, ctx->stream>>>(max_n, L, leaf_offsets, leaf_tree, links, tree_weight,
                                                                      stairs, bad_dev, batch);
        SCS_LAUNCHED(ctx, "pcg_leaf_stairs");
    }
    if (!ctx->full_rows) {
        // every pair once: the upper triangles, then the mirror (which also writes the row sums into `degree`)
        if (narrow)
            rc = launch_rows<uint16_t, false, uint32_t, true>(ctx, R, 0, R, 0, bs, stride, 1, smem, leaf_offsets, links,
                                                              tree_weight, leaf_tree, row_ptr, inv_sorted, occ, bucket_ptr,
                                                              entries, stairs, start16, W, nullptr, adj_bits, max_bits, degree,
                                                              bad_dev, batch);
        else
            rc = launch_rows<int32_t, false, uint32_t, true>(ctx, R, 0, R, 0, bs, stride, 1, smem, leaf_offsets, links,
                                                             tree_weight, leaf_tree, row_ptr, inv_sorted, occ, bucket_ptr,
                                                             entries, stairs, start16, W, nullptr, adj_bits, max_bits, degree,
                                                             bad_dev, batch);
        if (rc) return rc;
        return launch_mirror(ctx, 0, 0, 0, blocks, ceil_div(max_n, 32), B, R, W, adj_bits, max_bits, degree, batch);
    }
    // the row kernel writes the row sums straight into `degree` (one chunk: degree_part[0][row])
    if (narrow)
        rc = launch_rows<uint16_t, false, uint32_t, false>(ctx, R, 0, R, 0, bs, stride, 1, smem, leaf_offsets, links, tree_weight,
                                                    leaf_tree, row_ptr, inv_sorted, occ, bucket_ptr, entries, stairs, start16,
                                                           W, nullptr, adj_bits, max_bits, degree, bad_dev, batch);
    else
        rc = launch_rows<int32_t, false, uint32_t, false>(ctx, R, 0, R, 0, bs, stride, 1, smem, leaf_offsets, links, tree_weight,
                                                   leaf_tree, row_ptr, inv_sorted, occ, bucket_ptr, entries, stairs, start16,
                                                          W, nullptr, adj_bits, max_bits, degree, bad_dev, batch);
    return rc;
}

}  // namespace scs
