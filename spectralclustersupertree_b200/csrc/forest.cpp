// Host-side source-tree store of libscs_b200.so: flat trees, restriction to a taxon subset, and
// flattening into the leaf tours the kernels consume.
//
// The reference keeps cogent3 PhyloNode objects and, at every recursion node, restricts every
// source tree to the taxa of a component with PhyloNode.get_sub_tree(names, ignore_missing=True,
// as_rooted=True) (/root/reference/src/sc_supertree/scs.py:411-455).  Here a forest is a set of
// flat arrays -- nodes in depth-first pre-order with a parent index -- so the same restriction is
// one linear pass per tree with the same arithmetic:
//   * tips outside the subset vanish; an internal node left with one child is merged into that
//     child, whose length becomes  length(node) + length(child)  (missing if either is missing),
//     applied bottom-up exactly in that operand order;
//   * a root left with one child is replaced by its first branching descendant, whose own length
//     is dropped;
//   * trees left with fewer than two tips are dropped together with their weight (scs.py:447-448).
// Tours follow the reference's length_function top-down (scs.py:555-567, 628).

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <utility>
#include <vector>

#include "scs_b200.h"

#include "forest.hpp"

#include <malloc.h>
#include <omp.h>

static int g_host_threads = 0;
static thread_local const char *g_forest_error = "";

int scs_host_threads() { return g_host_threads > 0 ? g_host_threads : omp_get_max_threads(); }

namespace {

constexpr int64_t kParallelNodes = 1 << 15;  // forests with more nodes are processed by all host threads

}  // namespace

extern "C" {

static void tune_allocator_once() {
    static bool done = false;
    if (done) return;
    done = true;
    // forests of tens of MB are allocated and freed at every recursion node: keep the freed blocks in
    // the process heap instead of returning them to the OS, so re-use does not page-fault again
    mallopt(M_MMAP_THRESHOLD, 32 << 20);  // glibc caps this at 32 MB
    mallopt(M_TRIM_THRESHOLD, 1 << 30);
}

int scs_set_host_threads(int threads) {
    if (threads < 0) return SCS_ERR_INVALID;
    g_host_threads = threads;
    return SCS_OK;
}

static int forest_create(int T, const int64_t *node_offsets, const int32_t *parent, const double *length,
                         const double *support, const int32_t *taxon, const double *tree_weight, int num_taxa, bool view,
                         scs_forest **out);

int scs_forest_create(int T, const int64_t *node_offsets, const int32_t *parent, const double *length,
                      const double *support, const int32_t *taxon, const double *tree_weight, int num_taxa,
                      scs_forest **out) {
    return forest_create(T, node_offsets, parent, length, support, taxon, tree_weight, num_taxa, false, out);
}

int scs_forest_create_view(int T, const int64_t *node_offsets, const int32_t *parent, const double *length,
                           const double *support, const int32_t *taxon, const double *tree_weight, int num_taxa,
                           scs_forest **out) {
    return forest_create(T, node_offsets, parent, length, support, taxon, tree_weight, num_taxa, true, out);
}

static int forest_create(int T, const int64_t *node_offsets, const int32_t *parent, const double *length,
                         const double *support, const int32_t *taxon, const double *tree_weight, int num_taxa, bool view,
                         scs_forest **out) {
    if (!out) return SCS_ERR_INVALID;
    *out = nullptr;
    if (T < 0 || num_taxa < 0 || !node_offsets) return SCS_ERR_INVALID;
    const int64_t M = node_offsets[T];
    if (M < 0 || (M > 0 && (!parent || !taxon))) return SCS_ERR_INVALID;
    if (T > 0 && !tree_weight) return SCS_ERR_INVALID;
    tune_allocator_once();
    const bool trace = std::getenv("SCS_FOREST_TRACE") != nullptr;
    const double t_begin = trace ? omp_get_wtime() : 0.0;
    scs_forest *f = new (std::nothrow) scs_forest();
    if (!f) return SCS_ERR_INVALID;
    f->num_taxa = num_taxa;
    f->node_offsets.assign(node_offsets, node_offsets + T + 1);
    const int copy_threads = scs_host_threads() > 0 ? scs_host_threads() : 1;
    if (view) {
        // the per-node arrays stay the caller's (a gigabyte at 50 000 taxa x 5 000 trees: copying it costs more
        // than validating it); arrays the caller does not have are still made here
        f->parent.adopt(parent, parent + M);
        f->taxon.adopt(taxon, taxon + M);
        if (length) f->length.adopt(length, length + M);
        if (support) f->support.adopt(support, support + M);
    } else {
        f->parent.assign_parallel(parent, parent + M, copy_threads);
        f->taxon.assign_parallel(taxon, taxon + M, copy_threads);
        if (length) f->length.assign_parallel(length, length + M, copy_threads);
        if (support) f->support.assign_parallel(support, support + M, copy_threads);
    }
    if (!length) f->length.assign_parallel(static_cast<size_t>(M), std::nan(""), copy_threads);
    if (!support) f->support.assign_parallel(static_cast<size_t>(M), std::nan(""), copy_threads);
    f->weight.assign(tree_weight, tree_weight + T);
    f->source.resize(T);
    f->branching.assign(static_cast<size_t>(T), 1);
    std::vector<int64_t> tips(static_cast<size_t>(T), 0);
    bool all_ok = true, repeated = false;
    g_forest_error = "";
    const double t_arrays = trace ? omp_get_wtime() : 0.0;
    // validation and shape of every tree: independent, over the host threads
#pragma omp parallel if (M > kParallelNodes) num_threads(scs_host_threads())
    {
        std::vector<int32_t> kids, tip_taxa;
        // a taxon may label one tip of a tree only: the graph kernels give every leaf of a tree its own column
        // and add to it without atomics (pcg.cu, small.cu)
        std::vector<int32_t> seen_in(static_cast<size_t>(num_taxa > 0 ? num_taxa : 1), -1);
#pragma omp for schedule(dynamic, 8)
        for (int t = 0; t < T; ++t) {
            f->source[t] = t;
            const int64_t base = node_offsets[t], count = node_offsets[t + 1] - base;
            // Pass 1: parent indices, children per node, tips (pre-order: a node is a tip iff the next node is not its
            // child), tip taxa in range, internal nodes unnamed.  Which nodes are tips is as good as random to a branch
            // predictor (two mispredictions per node with `if (tip)`), so the pass accumulates flags with selects and
            // compacts the tips' taxa; pass 2 looks for repeated taxa among those; pass 3: does every internal node branch.
            const int32_t *par = parent + base;
            const int32_t *tax = taxon + base;
            unsigned bad = !(count >= 1 && par[0] == -1);
            const size_t room = static_cast<size_t>(count > 0 ? count : 0) + 1;
            kids.assign(room, 0);  // slot `count` takes the increments of out-of-range parents
            if (tip_taxa.size() < room) tip_taxa.resize(room);
            int32_t *kid = kids.data(), *tip_taxon = tip_taxa.data();
            const uint32_t taxa = static_cast<uint32_t>(num_taxa);
            int64_t tip_count = 0;
            for (int64_t k = 0; k < count; ++k) {
                const int32_t p = par[k];
                const bool p_ok = static_cast<uint64_t>(static_cast<int64_t>(p)) < static_cast<uint64_t>(k);  // 0 <= p < k
                bad |= static_cast<unsigned>(k >= 1) & static_cast<unsigned>(!p_ok);
                kid[p_ok ? static_cast<int64_t>(p) : count] += 1;
                const int32_t next_parent = k + 1 < count ? par[k + 1] : -2;
                const bool tip = static_cast<int64_t>(next_parent) != k;
                const int32_t x = tax[k];
                const bool named = static_cast<uint32_t>(x) < taxa;  // 0 <= x < num_taxa
                bad |= static_cast<unsigned>((tip & !named) | (!tip & (x != -1)));
                tip_taxon[tip_count] = x;
                tip_count += tip;
            }
            const bool ok = bad == 0;
            unsigned twice_flag = 0;
            if (ok) {
                int32_t *seen = seen_in.data();
                for (int64_t i = 0; i < tip_count; ++i) {
                    const int32_t x = tip_taxon[i];
                    twice_flag |= static_cast<unsigned>(seen[x] == t);
                    seen[x] = t;
                }
            }
            const bool twice = twice_flag != 0;
            if (!ok || twice) {
#pragma omp atomic write
                all_ok = false;
                if (twice) {
#pragma omp atomic write
                    repeated = true;
                }
                continue;
            }
            int64_t unary = 0;  // internal nodes with fewer than two children
            for (int64_t k = 0; k < count; ++k) unary += static_cast<int64_t>(tax[k] < 0) & static_cast<int64_t>(kid[k] < 2);
            tips[t] = count > 1 ? tip_count : 0;  // a lone tip appears in no tour
            f->branching[t] = unary == 0 ? 1 : 0;
        }
    }
    if (!all_ok) {
        delete f;
        g_forest_error = repeated ? "a taxon labels more than one tip of a source tree"
                                  : "a source tree is not a valid pre-order tree (parent indices / tip taxa)";
        return SCS_ERR_INPUT;
    }
    f->leaf_offsets.resize(static_cast<size_t>(T) + 1);
    f->leaf_offsets[0] = 0;
    for (int t = 0; t < T; ++t) f->leaf_offsets[t + 1] = f->leaf_offsets[t] + tips[t];
    if (trace)
        std::fprintf(stderr, "[scs forest] %d trees, %lld nodes, %s: arrays %.3f ms, validation %.3f ms (%d host threads)\n", T,
                     static_cast<long long>(M), view ? "view" : "copy", 1e3 * (t_arrays - t_begin),
                     1e3 * (omp_get_wtime() - t_arrays), scs_host_threads());
    *out = f;
    return SCS_OK;
}

const char *scs_forest_last_error(void) { return g_forest_error; }

int scs_forest_destroy(scs_forest *f) {
    delete f;
    return SCS_OK;
}

int scs_forest_num_trees(const scs_forest *f) { return f ? f->num_trees() : 0; }
int64_t scs_forest_num_nodes(const scs_forest *f) { return f ? f->node_offsets.back() : 0; }
int64_t scs_forest_num_leaves(const scs_forest *f) { return f ? f->leaf_offsets.back() : 0; }
int scs_forest_num_taxa(const scs_forest *f) { return f ? f->num_taxa : 0; }

int scs_forest_tree_info(const scs_forest *f, int t, int64_t *num_nodes, double *weight, int32_t *source) {
    if (!f || t < 0 || t >= f->num_trees()) return SCS_ERR_INVALID;
    if (num_nodes) *num_nodes = f->node_offsets[t + 1] - f->node_offsets[t];
    if (weight) *weight = f->weight[t];
    if (source) *source = f->source[t];
    return SCS_OK;
}

int scs_forest_tree(const scs_forest *f, int t, int32_t *parent, double *length, double *support, int32_t *taxon) {
    if (!f || t < 0 || t >= f->num_trees()) return SCS_ERR_INVALID;
    const int64_t base = f->node_offsets[t], count = f->node_offsets[t + 1] - base;
    if (parent) std::memcpy(parent, f->parent.data() + base, sizeof(int32_t) * count);
    if (length) std::memcpy(length, f->length.data() + base, sizeof(double) * count);
    if (support) std::memcpy(support, f->support.data() + base, sizeof(double) * count);
    if (taxon) std::memcpy(taxon, f->taxon.data() + base, sizeof(int32_t) * count);
    return SCS_OK;
}

/* present[x] = 1 iff taxon x is a tip of some tree (scs.py:708-725); returns how many are. */
int scs_forest_taxa(const scs_forest *f, uint8_t *present) {
    if (!f || !present) return SCS_ERR_INVALID;
    std::memset(present, 0, f->num_taxa);
    int n = 0;
    for (int32_t x : f->taxon)
        if (x >= 0 && !present[x]) {
            present[x] = 1;
            ++n;
        }
    return n;
}

int scs_forest_induce(const scs_forest *f, const uint8_t *keep, scs_forest **out) {
    if (!f || !keep || !out) return SCS_ERR_INVALID;
    *out = nullptr;
    scs_forest *g = new (std::nothrow) scs_forest();
    if (!g) return SCS_ERR_INVALID;
    g->num_taxa = f->num_taxa;
    const int T = f->num_trees();
    const int64_t M = f->node_offsets.back();
    // pass 1 (trees in parallel): which nodes survive, and their index in the restricted tree
    FlatArray<int32_t> new_index;  // written for every node of every kept tree in pass 1
    new_index.resize_uninitialized(M > 0 ? M : 1);
    std::vector<int32_t> kept_nodes(T + 1, 0), kept_tips(T + 1, 0);
    const bool threaded = M > kParallelNodes;
#pragma omp parallel if (threaded) num_threads(scs_host_threads())
    {
        std::vector<int32_t> cnt, live_children;
#pragma omp for schedule(dynamic, 4)
        for (int t = 0; t < T; ++t) {
            const int64_t base = f->node_offsets[t], count = f->node_offsets[t + 1] - base;
            const int32_t *par = f->parent.data() + base;
            const int32_t *tax = f->taxon.data() + base;
            int32_t *idx = new_index.data() + base;
            cnt.assign(count, 0);
            live_children.assign(count, 0);
            for (int64_t k = 0; k < count; ++k)
                if (tax[k] >= 0 && keep[tax[k]]) cnt[k] = 1;
            for (int64_t k = count - 1; k >= 1; --k) cnt[par[k]] += cnt[k];
            if (cnt[0] < 2) continue;  // scs.py:447-448: the tree is dropped
            for (int64_t k = 1; k < count; ++k)
                if (cnt[k] > 0) live_children[par[k]] += 1;
            int32_t next = 0, tips = 0;
            for (int64_t k = 0; k < count; ++k) {
                // retained: kept tips and nodes that still branch
                const bool retained = cnt[k] > 0 && (tax[k] >= 0 || live_children[k] >= 2);
                idx[k] = retained ? next++ : -1;
                tips += retained && tax[k] >= 0;
            }
            kept_nodes[t] = next;
            kept_tips[t] = tips;
        }
    }
    // output layout
    std::vector<int64_t> out_base(T + 1, 0);
    std::vector<int32_t> out_tree(T + 1, -1);
    int kept_trees = 0;
    for (int t = 0; t < T; ++t) {
        out_base[t + 1] = out_base[t] + kept_nodes[t];
        if (kept_nodes[t] == 0) continue;
        out_tree[t] = kept_trees++;
        g->node_offsets.push_back(out_base[t + 1]);
        g->leaf_offsets.push_back(g->leaf_offsets.back() + kept_tips[t]);
        g->weight.push_back(f->weight[t]);
        g->source.push_back(f->source[t]);
        g->branching.push_back(1);
    }
    const int64_t M_out = out_base[T];
    g->parent.resize_uninitialized(M_out);
    g->length.resize_uninitialized(M_out);
    g->support.resize_uninitialized(M_out);
    g->taxon.resize_uninitialized(M_out);
    // pass 2 (trees in parallel): write the restricted trees
#pragma omp parallel for schedule(dynamic, 4) if (threaded) num_threads(scs_host_threads())
    for (int t = 0; t < T; ++t) {
        if (kept_nodes[t] == 0) continue;
        const int64_t base = f->node_offsets[t], count = f->node_offsets[t + 1] - base;
        const int32_t *par = f->parent.data() + base;
        const int32_t *tax = f->taxon.data() + base;
        const double *len = f->length.data() + base;
        const double *sup = f->support.data() + base;
        const int32_t *idx = new_index.data() + base;
        const int64_t ob = out_base[t];
        for (int64_t k = 0; k < count; ++k) {
            const int32_t j = idx[k];
            if (j < 0) continue;
            if (j == 0) {  // first retained node in pre-order: the new root, its own length is dropped
                g->parent[ob] = -1;
                g->length[ob] = std::nan("");
            } else {
                double acc = len[k];
                int64_t a = par[k];
                while (idx[a] < 0) {     // merged unary ancestors, bottom-up
                    acc = len[a] + acc;  // NaN (missing) propagates like None
                    a = par[a];
                }
                g->parent[ob + j] = idx[a];
                g->length[ob + j] = acc;
            }
            g->support[ob + j] = sup[k];
            g->taxon[ob + j] = tax[k];
        }
    }
    *out = g;
    return SCS_OK;
}

}  // extern "C" (the batch restriction below is internal to the library)

namespace {

// Which nodes of a tree survive when only tips with owner[taxon] == job are kept: one bottom-up pass
// (pre-order puts every node after its parent) for the kept-tip counts and the number of children that
// still carry kept tips, one top-down pass for the new indices.  Returns the number of retained nodes
// (0: fewer than two tips are left and the tree is dropped, scs.py:447-448); idx[k] = new index or -1.
inline int32_t mark_retained(const int32_t *par, const int32_t *tax, int64_t count, const int32_t *owner, int32_t job,
                             std::vector<int32_t> &cnt, std::vector<int32_t> &live, std::vector<int32_t> &idx,
                             int32_t *tips_out) {
    *tips_out = 0;
    if (count < 3) return 0;
    cnt.assign(count, 0);
    live.assign(count, 0);
    int32_t kept = 0;
    for (int64_t k = count - 1; k >= 1; --k) {
        int32_t c = cnt[k];
        if (tax[k] >= 0) {
            c = owner[tax[k]] == job;
            cnt[k] = c;
            kept += c;
        }
        if (c) {
            cnt[par[k]] += c;
            live[par[k]] += 1;
        }
    }
    if (kept < 2) return 0;
    *tips_out = kept;
    idx.resize(count);
    int32_t next = 0;
    for (int64_t k = 0; k < count; ++k) {
        const bool retained = cnt[k] > 0 && (tax[k] >= 0 || live[k] >= 2);
        idx[k] = retained ? next++ : -1;
    }
    return next;
}

// Restricted trees are first written to per-thread staging (their sizes are only known once they are
// built), then copied to their place in the output forests.
struct Staging {
    // grown by doubling, never value-initialised: every element is written right after it is claimed
    int32_t *parent = nullptr, *taxon = nullptr;
    double *length = nullptr, *support = nullptr;
    size_t size = 0, capacity = 0;
    Staging() = default;
    Staging(const Staging &) = delete;
    Staging &operator=(const Staging &) = delete;
    Staging(Staging &&o) noexcept { *this = std::move(o); }
    Staging &operator=(Staging &&o) noexcept {
        release();
        parent = o.parent, taxon = o.taxon, length = o.length, support = o.support;
        size = o.size, capacity = o.capacity;
        o.parent = o.taxon = nullptr;
        o.length = o.support = nullptr;
        o.size = o.capacity = 0;
        return *this;
    }
    ~Staging() { release(); }
    void release() {
        std::free(parent);
        std::free(taxon);
        std::free(length);
        std::free(support);
        parent = taxon = nullptr;
        length = support = nullptr;
        size = capacity = 0;
    }
    void clear() { size = 0; }
    // room for `count` more nodes; returns the index of the first
    size_t claim(size_t count) {
        if (size + count > capacity) {
            size_t want = capacity ? capacity : (1u << 16);
            while (want < size + count) want *= 2;
            parent = static_cast<int32_t *>(std::realloc(parent, want * sizeof(int32_t)));
            taxon = static_cast<int32_t *>(std::realloc(taxon, want * sizeof(int32_t)));
            length = static_cast<double *>(std::realloc(length, want * sizeof(double)));
            support = static_cast<double *>(std::realloc(support, want * sizeof(double)));
            capacity = want;
        }
        const size_t at = size;
        size += count;
        return at;
    }
};

struct StagedTree {
    int32_t nodes = 0, tips = 0, thread = 0;  // thread < 0: copied verbatim from the source forest at `offset`
    int64_t offset = 0;  // in the thread's staging (or the source forest)
};

// The staging buffers are kept between calls (a wave of the recursion re-uses what the previous one
// touched instead of page-faulting fresh memory); a concurrent caller simply uses buffers of its own.
std::vector<Staging> g_staging;
omp_lock_t g_staging_lock;
bool g_staging_lock_ready = false;

}  // namespace

int scs_forest_induce_batch(scs_induce_job *jobs, int count, const int32_t *owner, uint8_t *present) {
    if (count <= 0) return SCS_OK;
    if (!jobs || !owner || !present) return SCS_ERR_INVALID;
    // work items: (run of jobs with the same source forest, tree of that forest); the children of one
    // recursion node are consecutive jobs, so a source tree is restricted to all of them while it is in cache
    std::vector<int> run_first, run_last;  // jobs [first, last) share a source
    std::vector<int64_t> run_item{0};      // first item of the run
    std::vector<int64_t> job_tree{0};      // index of (job, tree 0) in `staged`
    int64_t nodes = 0;
    for (int j = 0; j < count; ++j) {
        if (!jobs[j].src) return SCS_ERR_INVALID;
        job_tree.push_back(job_tree.back() + jobs[j].src->num_trees());
        if (j == 0 || jobs[j].src != jobs[j - 1].src) {
            run_first.push_back(j);
            run_last.push_back(j + 1);
            run_item.push_back(run_item.back() + jobs[j].src->num_trees());
            nodes += jobs[j].src->node_offsets.back();
        } else {
            run_last.back() = j + 1;
        }
    }
    const int64_t items = run_item.back();
    const int runs = static_cast<int>(run_first.size());
    std::vector<int32_t> item_run(items);
    for (int r = 0; r < runs; ++r)
        for (int64_t i = run_item[r]; i < run_item[r + 1]; ++i) item_run[i] = r;
    std::vector<StagedTree> staged(job_tree.back());
    const bool threaded = nodes > kParallelNodes;
    const int threads = threaded ? scs_host_threads() : 1;

#pragma omp critical(scs_staging_init)
    if (!g_staging_lock_ready) {
        omp_init_lock(&g_staging_lock);
        g_staging_lock_ready = true;
    }
    std::vector<Staging> own;
    const bool shared_pool = omp_test_lock(&g_staging_lock) != 0;
    std::vector<Staging> &pool = shared_pool ? g_staging : own;
    if (static_cast<int>(pool.size()) < threads) pool.resize(threads);

    static const bool trace = std::getenv("SCS_DRIVER_TRACE") != nullptr;
    const double t_begin = omp_get_wtime();
    // pass 1: build every restricted tree in staging
#pragma omp parallel if (threaded) num_threads(threads)
    {
        const int me = omp_get_thread_num();
        Staging &st = pool[me];
        st.clear();
        std::vector<int32_t> cnt, live, idx, hist, pair;
#pragma omp for schedule(dynamic, 8)
        for (int64_t i = 0; i < items; ++i) {
            const int r = item_run[i];
            const scs_forest *f = jobs[run_first[r]].src;
            const int t = static_cast<int>(i - run_item[r]);
            const int64_t base = f->node_offsets[t], cnt_nodes = f->node_offsets[t + 1] - base;
            const int32_t *par = f->parent.data() + base;
            const int32_t *tax = f->taxon.data() + base;
            const double *len = f->length.data() + base;
            const double *sup = f->support.data() + base;
            const int run_jobs = run_last[r] - run_first[r];
            if (run_jobs == 2 && cnt_nodes >= 3) {
                // The two sides of a bipartition (the usual run) in two passes over the tree instead of seven:
                // bottom-up, per node and side the kept tips below it, the children that carry some, and how
                // many nodes survive; top-down, both restricted trees written at once.
                const int32_t j0 = run_first[r];
                pair.assign(static_cast<size_t>(cnt_nodes) * 4, 0);  // per node: tips[2], live children[2]
                int32_t tips2[2] = {0, 0}, kept2[2] = {0, 0}, tree_tips = 0;
                for (int64_t k = cnt_nodes - 1; k >= 0; --k) {
                    int32_t *me4 = pair.data() + 4 * k;
                    const bool tip = tax[k] >= 0;
                    if (tip) {
                        ++tree_tips;
                        const int32_t o = owner[tax[k]] - j0;
                        if (o == 0 || o == 1) {
                            me4[o] = 1;
                            tips2[o] += 1;
                        }
                    }
                    for (int h = 0; h < 2; ++h) {
                        const int32_t c = me4[h];
                        if (!c) continue;
                        kept2[h] += tip || me4[2 + h] >= 2;
                        if (k > 0) {
                            int32_t *up4 = pair.data() + 4 * static_cast<int64_t>(par[k]);
                            up4[h] += c;
                            up4[2 + h] += 1;
                        }
                    }
                }
                int32_t *o_par[2] = {nullptr, nullptr}, *o_tax[2] = {nullptr, nullptr};
                double *o_len[2] = {nullptr, nullptr}, *o_sup[2] = {nullptr, nullptr};
                bool build[2] = {false, false};
                for (int h = 0; h < 2; ++h) {
                    StagedTree &out = staged[job_tree[j0 + h] + t];
                    if (tips2[h] < 2) continue;  // scs.py:447-448: the tree is dropped
                    if (tips2[h] == tree_tips && f->branching[t]) {
                        out.nodes = static_cast<int32_t>(cnt_nodes);  // kept whole: copied from the source in pass 2
                        out.tips = tips2[h];
                        out.thread = -1;
                        out.offset = base;
                        for (int64_t k = 0; k < cnt_nodes; ++k)
                            if (tax[k] >= 0) present[tax[k]] = 1;
                        continue;
                    }
                    out.nodes = kept2[h];
                    out.tips = tips2[h];
                    out.thread = me;
                    const size_t at = st.claim(static_cast<size_t>(kept2[h]));
                    out.offset = static_cast<int64_t>(at);
                    build[h] = true;
                }
                for (int h = 0; h < 2; ++h) {  // pointers only now: the second claim may have moved the buffers
                    if (!build[h]) continue;
                    const size_t at = static_cast<size_t>(staged[job_tree[j0 + h] + t].offset);
                    o_par[h] = st.parent + at;
                    o_tax[h] = st.taxon + at;
                    o_len[h] = st.length + at;
                    o_sup[h] = st.support + at;
                }
                if (!build[0] && !build[1]) continue;
                idx.assign(static_cast<size_t>(cnt_nodes) * 2, -1);  // new index per node and side
                int32_t next2[2] = {0, 0};
                for (int64_t k = 0; k < cnt_nodes; ++k) {
                    const int32_t *me4 = pair.data() + 4 * k;
                    const bool tip = tax[k] >= 0;
                    for (int h = 0; h < 2; ++h) {
                        if (!build[h] || !me4[h] || !(tip || me4[2 + h] >= 2)) continue;
                        const int32_t q = next2[h]++;
                        idx[2 * k + h] = q;
                        if (q == 0) {  // first retained node in pre-order: the new root, its own length is dropped
                            o_par[h][0] = -1;
                            o_len[h][0] = std::nan("");
                        } else {
                            double acc = len[k];
                            int64_t a = par[k];
                            while (idx[2 * a + h] < 0) {  // merged unary ancestors, bottom-up
                                acc = len[a] + acc;       // NaN (missing) propagates like None
                                a = par[a];
                            }
                            o_par[h][q] = idx[2 * a + h];
                            o_len[h][q] = acc;
                        }
                        o_sup[h][q] = sup[k];
                        o_tax[h][q] = tax[k];
                        if (tip) present[tax[k]] = 1;  // racing writers all store 1
                    }
                }
                continue;
            }
            // how many tips of this tree each job of the run keeps: most (job, tree) pairs need no second look
            hist.assign(static_cast<size_t>(run_jobs), 0);
            int32_t tree_tips = 0;
            for (int64_t k = 0; k < cnt_nodes; ++k) {
                if (tax[k] < 0) continue;
                ++tree_tips;
                const int32_t o = owner[tax[k]] - run_first[r];
                if (o >= 0 && o < run_jobs) hist[o] += 1;
            }
            for (int j = run_first[r]; j < run_last[r]; ++j) {
                StagedTree &out = staged[job_tree[j] + t];
                const int32_t mine = hist[j - run_first[r]];
                if (mine < 2 || cnt_nodes < 3) continue;  // scs.py:447-448: the tree is dropped
                if (mine == tree_tips && f->branching[t]) {
                    // every tip stays and there is no unary node to merge: the restricted tree is the tree itself
                    // (its root loses its length); copied straight from the source in pass 2
                    out.nodes = static_cast<int32_t>(cnt_nodes);
                    out.tips = mine;
                    out.thread = -1;
                    out.offset = base;
                    for (int64_t k = 0; k < cnt_nodes; ++k)
                        if (tax[k] >= 0) present[tax[k]] = 1;
                    continue;
                }
                int32_t tips = 0;
                const int32_t kept = mark_retained(par, tax, cnt_nodes, owner, j, cnt, live, idx, &tips);
                if (kept == 0) continue;
                out.nodes = kept;
                out.tips = tips;
                out.thread = me;
                const size_t at = st.claim(static_cast<size_t>(kept));
                out.offset = static_cast<int64_t>(at);
                int32_t *o_par = st.parent + at;
                int32_t *o_tax = st.taxon + at;
                double *o_len = st.length + at;
                double *o_sup = st.support + at;
                for (int64_t k = 0; k < cnt_nodes; ++k) {
                    const int32_t q = idx[k];
                    if (q < 0) continue;
                    if (q == 0) {  // first retained node in pre-order: the new root, its own length is dropped
                        o_par[0] = -1;
                        o_len[0] = std::nan("");
                    } else {
                        double acc = len[k];
                        int64_t a = par[k];
                        while (idx[a] < 0) {     // merged unary ancestors, bottom-up
                            acc = len[a] + acc;  // NaN (missing) propagates like None
                            a = par[a];
                        }
                        o_par[q] = idx[a];
                        o_len[q] = acc;
                    }
                    o_sup[q] = sup[k];
                    o_tax[q] = tax[k];
                    if (tax[k] >= 0) present[tax[k]] = 1;  // racing writers all store 1
                }
            }
        }
    }
    // layout of every output forest
    const double t_pass1 = omp_get_wtime();
    int rc = SCS_OK;
    std::vector<int64_t> dest(staged.size(), 0);
    // every job lays out its own forest: independent, so spread over the host threads as well (deep waves have
    // thousands of jobs and hundreds of thousands of little trees)
#pragma omp parallel for schedule(dynamic, 16) if (threaded && count >= 64) num_threads(threads)
    for (int j = 0; j < count; ++j) {
        scs_forest *g = new (std::nothrow) scs_forest();
        jobs[j].out = g;
        if (!g) {
#pragma omp atomic write
            rc = SCS_ERR_INVALID;
            continue;
        }
        const scs_forest *f = jobs[j].src;
        g->num_taxa = f->num_taxa;
        int64_t at = 0;
        const int T = f->num_trees();
        int kept_trees = 0;
        for (int t = 0; t < T; ++t) kept_trees += staged[job_tree[j] + t].nodes != 0;
        g->node_offsets.reserve(static_cast<size_t>(kept_trees) + 1);
        g->leaf_offsets.reserve(static_cast<size_t>(kept_trees) + 1);
        g->weight.reserve(static_cast<size_t>(kept_trees));
        g->source.reserve(static_cast<size_t>(kept_trees));
        g->branching.assign(static_cast<size_t>(kept_trees), 1);
        for (int t = 0; t < T; ++t) {
            const StagedTree &tree = staged[job_tree[j] + t];
            if (tree.nodes == 0) continue;
            dest[job_tree[j] + t] = at;
            at += tree.nodes;
            g->node_offsets.push_back(at);
            g->leaf_offsets.push_back(g->leaf_offsets.back() + tree.tips);
            g->weight.push_back(f->weight[t]);
            g->source.push_back(f->source[t]);
        }
        g->parent.resize_uninitialized(at);
        g->length.resize_uninitialized(at);
        g->support.resize_uninitialized(at);
        g->taxon.resize_uninitialized(at);
    }
    const double t_layout = omp_get_wtime();
    if (rc != SCS_OK) {
        for (int j = 0; j < count; ++j) {
            delete jobs[j].out;
            jobs[j].out = nullptr;
        }
    } else {
        // pass 2: staging -> output forests
        const int64_t total = static_cast<int64_t>(staged.size());
        std::vector<int32_t> tree_job(total);
        for (int j = 0; j < count; ++j)
            for (int64_t i = job_tree[j]; i < job_tree[j + 1]; ++i) tree_job[i] = j;
#pragma omp parallel for schedule(dynamic, 32) if (threaded) num_threads(threads)
        for (int64_t i = 0; i < total; ++i) {
            const StagedTree &tree = staged[i];
            if (tree.nodes == 0) continue;
            scs_forest *g = jobs[tree_job[i]].out;
            const size_t n = static_cast<size_t>(tree.nodes);
            if (tree.thread < 0) {
                const scs_forest *f = jobs[tree_job[i]].src;
                std::memcpy(g->parent.data() + dest[i], f->parent.data() + tree.offset, n * sizeof(int32_t));
                std::memcpy(g->taxon.data() + dest[i], f->taxon.data() + tree.offset, n * sizeof(int32_t));
                std::memcpy(g->length.data() + dest[i], f->length.data() + tree.offset, n * sizeof(double));
                std::memcpy(g->support.data() + dest[i], f->support.data() + tree.offset, n * sizeof(double));
                g->length[dest[i]] = std::nan("");  // the root's own length is dropped
                continue;
            }
            const Staging &st = pool[tree.thread];
            std::memcpy(g->parent.data() + dest[i], st.parent + tree.offset, n * sizeof(int32_t));
            std::memcpy(g->taxon.data() + dest[i], st.taxon + tree.offset, n * sizeof(int32_t));
            std::memcpy(g->length.data() + dest[i], st.length + tree.offset, n * sizeof(double));
            std::memcpy(g->support.data() + dest[i], st.support + tree.offset, n * sizeof(double));
        }
    }
    if (shared_pool) omp_unset_lock(&g_staging_lock);
    if (trace)
        std::fprintf(stderr, "[scs induce] jobs %d items %lld trees %zu: pass1 %.2f ms, layout %.2f ms, pass2 %.2f ms\n", count,
                     static_cast<long long>(items), staged.size(), 1e3 * (t_pass1 - t_begin), 1e3 * (t_layout - t_pass1),
                     1e3 * (omp_get_wtime() - t_layout));
    return rc;
}

extern "C" {

/* The loop scs.py:139-155 in one call: out[c] = the forest restricted to the taxa x with part[x] == c,
 * c = 0..count-1 (part[x] < 0 or >= count: the taxon is dropped); present[x] (may be NULL) is set to 1
 * for every taxon left in some restricted tree. */
int scs_forest_induce_parts(const scs_forest *f, const int32_t *part, int count, scs_forest **out, uint8_t *present) {
    if (!f || !part || count < 0 || !out) return SCS_ERR_INVALID;
    std::vector<scs_induce_job> jobs(static_cast<size_t>(count));
    for (int c = 0; c < count; ++c) {
        jobs[c].src = f;
        out[c] = nullptr;
    }
    std::vector<uint8_t> scratch;
    if (!present) {
        scratch.assign(f->num_taxa ? f->num_taxa : 1, 0);
        present = scratch.data();
    }
    const int rc = scs_forest_induce_batch(jobs.data(), count, part, present);
    if (rc) return rc;
    for (int c = 0; c < count; ++c) out[c] = jobs[c].out;
    return SCS_OK;
}

/* weighting: 0 one, 1 branch, 2 depth, 3 bootstrap (order of the reference's docs, scs.py:41-50).
 * local_id[x] = vertex id of global taxon x at this recursion node.
 * Returns SCS_ERR_INPUT if bootstrap weighting meets an internal node without support that is the
 * LCA of a leaf pair (the reference raises TypeError there: None * weight, scs.py:655-657). */
int scs_forest_tours(const scs_forest *f, int weighting, const int32_t *local_id, int64_t *leaf_offsets,
                     int32_t *leaf_taxon, int32_t *adj_depth, double *adj_val, int32_t *root_depth,
                     double *tree_weight) {
    if (!f || !local_id || !leaf_offsets || weighting < 0 || weighting > 3) return SCS_ERR_INVALID;
    const int T = f->num_trees();
    if (f->leaf_offsets.back() > 0 && (!leaf_taxon || !adj_depth || !adj_val)) return SCS_ERR_INVALID;
    if (T > 0 && (!root_depth || !tree_weight)) return SCS_ERR_INVALID;
    int status = SCS_OK;
    leaf_offsets[0] = 0;
    const bool threaded = f->node_offsets.back() > kParallelNodes;
#pragma omp parallel if (threaded) num_threads(scs_host_threads())
    {
        std::vector<int32_t> depth;
        std::vector<double> val;
#pragma omp for schedule(dynamic, 4)
        for (int t = 0; t < T; ++t) {
            const int64_t base = f->node_offsets[t], count = f->node_offsets[t + 1] - base;
            const int32_t *par = f->parent.data() + base;
            const int32_t *tax = f->taxon.data() + base;
            const double *len = f->length.data() + base;
            const double *sup = f->support.data() + base;
            int64_t o = f->leaf_offsets[t];
            leaf_offsets[t + 1] = f->leaf_offsets[t + 1];
            root_depth[t] = 0;
            tree_weight[t] = f->weight[t];
            if (count <= 1) continue;  // a lone tip has no sides (scs.py:570)
            depth.assign(count, 0);
            val.assign(count, 0.0);  // the value handed to the root's children is 0 (scs.py:577)
            for (int64_t k = 1; k < count; ++k) {
                if (tax[k] >= 0) continue;
                const int32_t p = par[k];
                depth[k] = depth[p] + 1;
                switch (weighting) {
                case 0: val[k] = 1.0; break;
                case 1: val[k] = val[p] + (std::isnan(len[k]) ? 1.0 : len[k]); break;
                case 2: val[k] = val[p] + 1.0; break;
                default: val[k] = sup[k]; break;
                }
            }
            for (int64_t k = 0; k < count; ++k) {
                if (tax[k] < 0) continue;
                leaf_taxon[o] = local_id[tax[k]];
                if (k + 1 < count) {
                    const int32_t lca = par[k + 1];  // the next pre-order node hangs off the LCA with the next tip
                    adj_depth[o] = depth[lca];
                    adj_val[o] = val[lca];
                    if (lca != 0 && std::isnan(val[lca])) {
#pragma omp atomic write
                        status = SCS_ERR_INPUT;
                    }
                } else {
                    adj_depth[o] = -1;
                    adj_val[o] = 0.0;
                }
                ++o;
            }
        }
    }
    return status;
}

/* Number of ordered leaf pairs the row kernel visits: sum over trees of k (k - 1). */
int64_t scs_forest_pair_visits(const scs_forest *f) {
    if (!f) return 0;
    int64_t total = 0;
    for (int t = 0; t < f->num_trees(); ++t) {
        const int64_t k = f->leaf_offsets[t + 1] - f->leaf_offsets[t];
        total += k * (k - 1);
    }
    return total;
}

/* One recursion node straight from a forest (scs.py:102-134): the vertices are the taxa present,
 * numbered in increasing global id; taxa_out[n] receives their global ids and part_out[n] the
 * component index or spectral side of each.  taxa_out / part_out must hold num_taxa entries. */
int scs_forest_split(scs_ctx *ctx, const scs_forest *f, int weighting, int contract_edges, uint64_t seed,
                     int32_t *n_out, int32_t *taxa_out, int32_t *part_out, scs_node_stats *stats) {
    if (!ctx || !f || !n_out || !taxa_out || !part_out) return SCS_ERR_INVALID;
    std::vector<uint8_t> present(f->num_taxa ? f->num_taxa : 1);
    const int n = scs_forest_taxa(f, present.data());
    std::vector<int32_t> local(f->num_taxa ? f->num_taxa : 1, -1);
    int next = 0;
    for (int x = 0; x < f->num_taxa; ++x)
        if (present[x]) {
            local[x] = next;
            taxa_out[next++] = x;
        }
    *n_out = n;
    if (n == 0) return SCS_ERR_INVALID;
    const int T = f->num_trees();
    const int64_t L = f->leaf_offsets.back();
    std::vector<int64_t> leaf_offsets(T + 1);
    std::vector<int32_t> leaf_taxon(L + 1), adj_depth(L + 1), root_depth(T + 1);
    std::vector<double> adj_val(L + 1), tree_weight(T + 1);
    int rc = scs_forest_tours(f, weighting, local.data(), leaf_offsets.data(), leaf_taxon.data(), adj_depth.data(),
                              adj_val.data(), root_depth.data(), tree_weight.data());
    if (rc) return rc;
    return scs_node_split_host(ctx, n, T, L, leaf_offsets.data(), leaf_taxon.data(), adj_depth.data(),
                               adj_val.data(), root_depth.data(), tree_weight.data(), contract_edges, seed, part_out,
                               stats);
}

}  // extern "C"
