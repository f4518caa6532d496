// The flat source-tree store shared by forest.cpp (operations) and driver.cu (the recursion).
#pragma once

#include <cstdint>
#include <vector>

struct scs_forest {
    int num_taxa = 0;
    std::vector<int64_t> node_offsets{0};  // [T + 1]
    std::vector<int32_t> parent;           // index within the tree, -1 for the root; parent < child
    std::vector<double> length;            // NaN = missing
    std::vector<double> support;           // NaN = missing
    std::vector<int32_t> taxon;            // tips: global taxon id, internal nodes: -1
    std::vector<double> weight;            // [T]
    std::vector<int32_t> source;           // [T] index of the tree in the forest first created
    std::vector<int64_t> leaf_offsets{0};  // [T + 1] tips that appear in tours (a lone tip has none)
    int num_trees() const { return static_cast<int>(weight.size()); }
};


// Host threads used by the forest operations and the recursion driver (0 = OpenMP's default).
int scs_host_threads();
