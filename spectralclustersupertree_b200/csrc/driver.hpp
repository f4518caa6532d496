// The result of a supertree build (shared by the two recursion drivers: driver.cu keeps the source trees on the host,
// devdriver.cu keeps them on the device).
#pragma once

#include <cstddef>
#include <cstdint>
#include <vector>

#include "scs_b200.h"

struct scs_supertree {
    std::vector<int32_t> parent;  // parent[i] < i; -1 for the root
    std::vector<int32_t> taxon;   // global taxon id for tips, -1 for internal nodes
    struct Record {
        std::vector<int32_t> taxa, part;
        scs_node_stats stats;
        int32_t wave = 0;  // wave of the breadth-first recursion that processed the node
    };
    std::vector<Record> records;  // one per recursion node that reached the GPU (if requested)
    int64_t nodes_small = 0, nodes_large = 0, nodes_medium = 0, nodes_rerun = 0, waves = 0;
    int64_t pair_visits = 0;
    double seconds[4] = {0, 0, 0, 0};  // large-node splits, small-node batches, restriction, tours
    double medium_seconds = 0.0;       // medium-node batches (without their tours)
    std::vector<int32_t> wave_tasks, wave_max_n;  // per wave: sub-problems in it, largest taxon count
    std::vector<double> wave_seconds;             // per wave: 3 numbers (GPU splits, restriction, everything)
    int64_t shared_prefix = 0;  // sharded build: output nodes [0, shared_prefix) are identical on every rank
    int64_t shared_records = 0;  // recursion nodes processed while every rank still walked the same frontier
};

struct scs_forest;
struct scs_device_forest;

namespace scs {

// Seed of a recursion node's Lanczos start vector: a function of the node itself (its smallest taxon and its size),
// not of the output slot it happens to fill, so that both drivers, every rank of a cooperative build and every
// dealing of the sub-problems start the eigensolver of a given node from the same vector.
inline uint64_t node_seed(uint64_t seed, int32_t first_taxon, size_t taxa) {
    return seed + static_cast<uint64_t>(first_taxon) * 0x9E3779B97F4A7C15ull + static_cast<uint64_t>(taxa);
}

// The whole recursion with the source trees resident in HBM (devdriver.cu): restriction and tour flattening on the
// device, a few small host round trips per wave for the bookkeeping of the output tree.  One GPU, or rank `rank` of
// `world` in a cooperative build (exchange windows connected: large nodes row-sharded over the GPUs, smaller
// sub-problems dealt out over the ranks).
// scs_device_forest_create; cooperative: the ranks of a connected shard group share the upload (devforest.cuh)
int device_forest_create(scs_ctx *ctx, const scs_forest *forest, int weighting, bool cooperative, scs_device_forest **out);

// the same into the device forest the context keeps between builds (nothing to destroy)
int device_forest_refresh(scs_ctx *ctx, const scs_forest *forest, int weighting, bool cooperative,
                          const scs_device_forest **out);

int run_device_driver(scs_ctx *ctx, const scs_device_forest *forest, int contract_edges, uint64_t seed, bool record, int rank,
                      int world, scs_supertree *out);

}  // namespace scs
