/* TEST INFRASTRUCTURE ONLY -- plain C twin of oracle/scs_oracle.py:pcg_dense.
 *
 * Restates the reference's proper-cluster-graph build on dense arrays:
 *   _proper_cluster_graph_edges   /root/reference/src/sc_supertree/scs.py:495-583
 *   _dfs_pcg_weights              /root/reference/src/sc_supertree/scs.py:586-663
 *
 * Trees arrive as child lists with nodes numbered in depth-first pre-order, so the tips below
 * a node are a contiguous run of the tree's tip sequence -- the same lists the reference
 * concatenates at scs.py:660-663.  For every internal node other than the root, every pair of
 * tips taken from two different children receives  W += value(node) * tree_weight  (one
 * rounded multiply, one rounded add, as Python evaluates scs.py:655-657) and  C += 1 ; trees
 * are processed in input order, so the floating-point summation order is the reference's.
 * occ[a] counts the trees holding tip a below a child of the root (scs.py:580-581).
 *
 * Not part of the product: only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
 * legs load this library.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

enum { MODE_ONE = 0, MODE_BRANCH = 1, MODE_DEPTH = 2, MODE_BOOTSTRAP = 3 };

/* value handed down by length_function (scs.py:555-567) */
static double node_value(int mode, double above, double own)
{
    switch (mode) {
    case MODE_ONE: return 1.0;
    case MODE_DEPTH: return above + 1.0;
    case MODE_BRANCH: return above + own;
    default: return own; /* bootstrap: the node's own support */
    }
}

int pcg_oracle_dense(int64_t n, int64_t num_trees, const int64_t *roots, const int64_t *child_ptr,
                     const int64_t *child_idx, const int32_t *tip_taxon, const double *own,
                     const double *tree_weight, int mode, double *W, int32_t *C, int32_t *occ)
{
    int64_t total_nodes = 0;
    for (int64_t t = 0; t < num_trees; ++t) {
        /* nodes of tree t occupy [roots[t], next root) in pre-order */
        (void)t;
    }
    /* scratch sized for the largest tree */
    int64_t max_nodes = 0;
    for (int64_t t = 0; t < num_trees; ++t) {
        int64_t begin = roots[t];
        int64_t end = begin;
        /* walk to the end of the tree: the last node of a pre-order numbering is reached by
           following last children */
        int64_t x = begin;
        while (child_ptr[x + 1] > child_ptr[x]) x = child_idx[child_ptr[x + 1] - 1];
        end = x + 1;
        if (end - begin > max_nodes) max_nodes = end - begin;
        total_nodes += end - begin;
    }
    double *value = (double *)malloc(sizeof(double) * (size_t)(max_nodes + 1));
    int64_t *lo = (int64_t *)malloc(sizeof(int64_t) * (size_t)(max_nodes + 1));
    int64_t *hi = (int64_t *)malloc(sizeof(int64_t) * (size_t)(max_nodes + 1));
    int32_t *tips = (int32_t *)malloc(sizeof(int32_t) * (size_t)(max_nodes + 1));
    if (!value || !lo || !hi || !tips) return -1;

    for (int64_t t = 0; t < num_trees; ++t) {
        const int64_t base = roots[t];
        int64_t last = base;
        while (child_ptr[last + 1] > child_ptr[last]) last = child_idx[child_ptr[last + 1] - 1];
        const int64_t count = last + 1 - base;
        const double w = tree_weight[t];

        /* pre-order pass: values top-down (scs.py:628) and tip runs */
        int64_t num_tips = 0;
        value[0] = 0.0; /* "length" handed to the root's children is 0 (scs.py:577) */
        for (int64_t k = 0; k < count; ++k) {
            const int64_t x = base + k;
            lo[k] = num_tips;
            if (child_ptr[x + 1] == child_ptr[x]) {
                if (tip_taxon[x] < 0 || tip_taxon[x] >= n) return -2;
                tips[num_tips++] = tip_taxon[x];
            } else {
                for (int64_t e = child_ptr[x]; e < child_ptr[x + 1]; ++e) {
                    const int64_t c = child_idx[e] - base;
                    if (c <= k || c >= count) return -3; /* not pre-order */
                    const int is_tip = child_ptr[child_idx[e] + 1] == child_ptr[child_idx[e]];
                    value[c] = is_tip ? value[k] : node_value(mode, value[k], own[child_idx[e]]);
                }
            }
        }
        /* hi[k]: in pre-order, a subtree ends where the next sibling-or-ancestor-sibling starts */
        for (int64_t k = count - 1; k >= 0; --k) {
            const int64_t x = base + k;
            if (child_ptr[x + 1] == child_ptr[x]) hi[k] = lo[k] + 1;
            else hi[k] = hi[child_idx[child_ptr[x + 1] - 1] - base];
        }
        /* occurrences: every tip below a child of the root (a lone tip has no sides) */
        if (child_ptr[base + 1] > child_ptr[base])
            for (int64_t i = 0; i < num_tips; ++i) occ[tips[i]] += 1;

        /* pair updates, children before parents is not required on dense arrays: each pair is
           touched exactly once per tree, so any node order gives the same sums */
        for (int64_t k = 1; k < count; ++k) { /* k = 0 is the root: its pairs are not proper clusters */
            const int64_t x = base + k;
            const int64_t nchild = child_ptr[x + 1] - child_ptr[x];
            if (nchild < 2) continue;
            if (mode == MODE_BOOTSTRAP && isnan(value[k])) return -4; /* None * w -> TypeError there */
            const double term = value[k] * w;
            for (int64_t i = 1; i < nchild; ++i) {
                const int64_t ci = child_idx[child_ptr[x] + i] - base;
                for (int64_t j = 0; j < i; ++j) {
                    const int64_t cj = child_idx[child_ptr[x] + j] - base;
                    for (int64_t p = lo[ci]; p < hi[ci]; ++p) {
                        const int64_t a = tips[p];
                        for (int64_t q = lo[cj]; q < hi[cj]; ++q) {
                            const int64_t b = tips[q];
                            W[a * n + b] = W[a * n + b] + term;
                            W[b * n + a] = W[b * n + a] + term;
                            C[a * n + b] += 1;
                            C[b * n + a] += 1;
                        }
                    }
                }
            }
        }
    }
    (void)total_nodes;
    free(value); free(lo); free(hi); free(tips);
    return 0;
}

/* Rows [row_lo, row_hi) of the same W / C, computed tip-wise instead of node-wise: for every tree and every tip
 * whose taxon is in the row block, every other tip q of the tree meets it at exactly one internal node -- walk
 * up from the tip; at ancestor k (the root excluded: its pairs are no proper clusters, scs.py:570-579) the tips
 * of k outside the child just left get  W += value(k) * tree_weight  and  C += 1 .  Trees in input order, so
 * every entry is the same ordered sum as in pcg_oracle_dense (bit for bit); memory is (row_hi - row_lo) * n, so
 * blocks of a 50 000-taxon matrix can be checked without its 20 GB.  W_rows / C_rows: row a at (a - row_lo) * n. */
int pcg_oracle_rows(int64_t n, int64_t num_trees, const int64_t *roots, const int64_t *child_ptr,
                    const int64_t *child_idx, const int32_t *tip_taxon, const double *own,
                    const double *tree_weight, int mode, int64_t row_lo, int64_t row_hi, double *W_rows,
                    int32_t *C_rows)
{
    int64_t max_nodes = 0;
    for (int64_t t = 0; t < num_trees; ++t) {
        int64_t x = roots[t];
        while (child_ptr[x + 1] > child_ptr[x]) x = child_idx[child_ptr[x + 1] - 1];
        if (x + 1 - roots[t] > max_nodes) max_nodes = x + 1 - roots[t];
    }
    double *value = (double *)malloc(sizeof(double) * (size_t)(max_nodes + 1));
    int64_t *lo = (int64_t *)malloc(sizeof(int64_t) * (size_t)(max_nodes + 1));
    int64_t *hi = (int64_t *)malloc(sizeof(int64_t) * (size_t)(max_nodes + 1));
    int64_t *up = (int64_t *)malloc(sizeof(int64_t) * (size_t)(max_nodes + 1));
    int64_t *tip_node = (int64_t *)malloc(sizeof(int64_t) * (size_t)(max_nodes + 1));
    int32_t *tips = (int32_t *)malloc(sizeof(int32_t) * (size_t)(max_nodes + 1));
    if (!value || !lo || !hi || !up || !tip_node || !tips) return -1;
    for (int64_t t = 0; t < num_trees; ++t) {
        const int64_t base = roots[t];
        int64_t last = base;
        while (child_ptr[last + 1] > child_ptr[last]) last = child_idx[child_ptr[last + 1] - 1];
        const int64_t count = last + 1 - base;
        const double w = tree_weight[t];
        int64_t num_tips = 0;
        value[0] = 0.0;
        up[0] = -1;
        for (int64_t k = 0; k < count; ++k) {
            const int64_t x = base + k;
            lo[k] = num_tips;
            if (child_ptr[x + 1] == child_ptr[x]) {
                if (tip_taxon[x] < 0 || tip_taxon[x] >= n) return -2;
                tip_node[num_tips] = k;
                tips[num_tips++] = tip_taxon[x];
            } else {
                for (int64_t e = child_ptr[x]; e < child_ptr[x + 1]; ++e) {
                    const int64_t c = child_idx[e] - base;
                    if (c <= k || c >= count) return -3;
                    const int is_tip = child_ptr[child_idx[e] + 1] == child_ptr[child_idx[e]];
                    value[c] = is_tip ? value[k] : node_value(mode, value[k], own[child_idx[e]]);
                    up[c] = k;
                }
            }
        }
        for (int64_t k = count - 1; k >= 0; --k) {
            const int64_t x = base + k;
            if (child_ptr[x + 1] == child_ptr[x]) hi[k] = lo[k] + 1;
            else hi[k] = hi[child_idx[child_ptr[x + 1] - 1] - base];
        }
        for (int64_t p = 0; p < num_tips; ++p) {
            const int64_t a = tips[p];
            if (a < row_lo || a >= row_hi) continue;
            double *Wa = W_rows + (a - row_lo) * n;
            int32_t *Ca = C_rows + (a - row_lo) * n;
            int64_t below = tip_node[p];
            for (int64_t k = up[below]; k > 0; below = k, k = up[k]) { /* k == 0 is the root */
                if (mode == MODE_BOOTSTRAP && isnan(value[k])) return -4;
                const double term = value[k] * w;
                for (int64_t q = lo[k]; q < lo[below]; ++q) { Wa[tips[q]] = Wa[tips[q]] + term; Ca[tips[q]] += 1; }
                for (int64_t q = hi[below]; q < hi[k]; ++q) { Wa[tips[q]] = Wa[tips[q]] + term; Ca[tips[q]] += 1; }
            }
        }
    }
    free(value); free(lo); free(hi); free(up); free(tip_node); free(tips);
    return 0;
}
