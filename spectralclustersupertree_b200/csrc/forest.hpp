// The flat source-tree store shared by forest.cpp (operations) and driver.cu (the recursion).
#pragma once

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

// A plain array that is NOT value-initialised when sized: the per-node arrays of a restricted forest
// are written exactly once, by all host threads, so a serial zero-fill (and its serial page faults)
// would cost as much as the restriction itself.
template <typename T>
class FlatArray {
  public:
    FlatArray() = default;
    FlatArray(const FlatArray &) = delete;
    FlatArray &operator=(const FlatArray &) = delete;
    ~FlatArray() {
        if (owned_) std::free(data_);
    }
    void resize_uninitialized(size_t n) {
        if (owned_) std::free(data_);
        data_ = n ? static_cast<T *>(std::malloc(n * sizeof(T))) : nullptr;
        size_ = data_ ? n : 0;
        owned_ = true;
    }
    // refer to the caller's array instead of copying it (scs_forest_create_view: the caller keeps it alive and
    // unchanged for as long as the forest exists); never written through
    void adopt(const T *first, const T *last) {
        if (owned_) std::free(data_);
        data_ = const_cast<T *>(first);
        size_ = static_cast<size_t>(last - first);
        owned_ = false;
    }
    void assign(const T *first, const T *last) {
        resize_uninitialized(static_cast<size_t>(last - first));
        if (size_) std::memcpy(data_, first, size_ * sizeof(T));
    }
    void assign(size_t n, T value) {
        resize_uninitialized(n);
        for (size_t i = 0; i < size_; ++i) data_[i] = value;
    }
    // the same two, spread over the host threads (large inputs: the copy and its page faults are the cost)
    void assign_parallel(const T *first, const T *last, int threads) {
        resize_uninitialized(static_cast<size_t>(last - first));
        const long long n = static_cast<long long>(size_);
        const long long chunk = 1ll << 18;
#pragma omp parallel for schedule(static) num_threads(threads) if (n > (4ll << 20))
        for (long long at = 0; at < n; at += chunk) {
            const long long len = n - at < chunk ? n - at : chunk;
            std::memcpy(data_ + at, first + at, static_cast<size_t>(len) * sizeof(T));
        }
    }
    void assign_parallel(size_t count, T value, int threads) {
        resize_uninitialized(count);
        const long long n = static_cast<long long>(size_);
#pragma omp parallel for schedule(static) num_threads(threads) if (n > (4ll << 20))
        for (long long i = 0; i < n; ++i) data_[i] = value;
    }
    T *data() { return data_; }
    const T *data() const { return data_; }
    size_t size() const { return size_; }
    T &operator[](size_t i) { return data_[i]; }
    const T &operator[](size_t i) const { return data_[i]; }
    const T *begin() const { return data_; }
    const T *end() const { return data_ + size_; }

  private:
    T *data_ = nullptr;
    size_t size_ = 0;
    bool owned_ = true;
};

struct scs_forest {
    int num_taxa = 0;
    std::vector<int64_t> node_offsets{0};  // [T + 1]
    FlatArray<int32_t> parent;             // index within the tree, -1 for the root; parent < child
    FlatArray<double> length;              // NaN = missing
    FlatArray<double> support;             // NaN = missing
    FlatArray<int32_t> taxon;              // tips: global taxon id, internal nodes: -1
    std::vector<double> weight;            // [T]
    std::vector<int32_t> source;           // [T] index of the tree in the forest first created
    std::vector<int64_t> leaf_offsets{0};  // [T + 1] tips that appear in tours (a lone tip has none)
    std::vector<uint8_t> branching;        // [T] 1: every internal node has at least two children (always true for
                                           // a restricted tree); such a tree restricted to ALL its tips is itself
    int num_trees() const { return static_cast<int>(weight.size()); }
};


// Host threads used by the forest operations and the recursion driver (0 = OpenMP's default).
int scs_host_threads();

// Many restrictions at once (the children of every node of one wave of the recursion).  Job j keeps the
// tips x of `src` with owner[x] == j: the sub-problems of a wave have disjoint taxon sets, so one
// owner[] array (global taxon id -> job) serves the whole wave.  All (job, tree) pairs are spread over the
// host threads, which balances a wave of two huge restrictions as well as one of a thousand small ones.
// `out` receives the restricted forest; present[x] is set to 1 for every taxon that is a tip of a kept
// tree (the caller clears it), so the taxa of the child need no further scan (scs.py:708-725).
struct scs_induce_job {
    const scs_forest *src = nullptr;
    scs_forest *out = nullptr;
};
int scs_forest_induce_batch(scs_induce_job *jobs, int count, const int32_t *owner, uint8_t *present);
