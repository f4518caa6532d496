// Host-only timing of the batched forest restriction (scs_forest_induce_batch, the call the recursion driver
// makes once per wave) on a synthetic recursion: the taxa of every node are halved by taxon id, wave after
// wave; prints the best time per wave over 12 repetitions.  Built and run by tools/restrict_harness.py.
#include <chrono>
#include <cstdio>
#include <fstream>
#include <vector>
#include <algorithm>
#include "scs_b200.h"
#include "forest.hpp"
template <typename T> std::vector<T> slurp(const char *path) {
    std::ifstream f(path, std::ios::binary | std::ios::ate);
    size_t bytes = f.tellg(); f.seekg(0);
    std::vector<T> v(bytes / sizeof(T)); f.read(reinterpret_cast<char *>(v.data()), bytes); return v;
}
struct Task { scs_forest *f; std::vector<int32_t> taxa; };
int main(int argc, char **argv) {
    int threads = argc > 1 ? atoi(argv[1]) : 8;
    scs_set_host_threads(threads);
    auto hdr = slurp<int64_t>("/tmp/rp/hdr.bin");
    auto off = slurp<int64_t>("/tmp/rp/off.bin"); auto par = slurp<int32_t>("/tmp/rp/par.bin");
    auto len = slurp<double>("/tmp/rp/len.bin"); auto sup = slurp<double>("/tmp/rp/sup.bin");
    auto tax = slurp<int32_t>("/tmp/rp/tax.bin"); auto w = slurp<double>("/tmp/rp/w.bin");
    int T = hdr[0], n = hdr[2];
    std::vector<double> best; for (int rep = 0; rep < 12; ++rep) {
        scs_forest *root = nullptr;
        scs_forest_create(T, off.data(), par.data(), len.data(), sup.data(), tax.data(), w.data(), n, &root);
        std::vector<Task> wave(1); wave[0].f = root; wave[0].taxa.resize(n);
        for (int i = 0; i < n; ++i) wave[0].taxa[i] = i;
        std::vector<int32_t> owner(n, -1); std::vector<uint8_t> present(n, 0);
        double total = 0; std::vector<double> per;
        while (!wave.empty()) {
            std::vector<scs_induce_job> jobs; std::vector<std::pair<int,int>> src;  // (task, half)
            for (size_t t = 0; t < wave.size(); ++t) {
                Task &task = wave[t];
                if (task.taxa.size() <= 2 || scs_forest_num_trees(task.f) < 2) continue;
                size_t half = task.taxa.size() / 2;
                for (int h = 0; h < 2; ++h) {
                    int job = jobs.size();
                    for (size_t i = h ? half : 0; i < (h ? task.taxa.size() : half); ++i) owner[task.taxa[i]] = job;
                    scs_induce_job j; j.src = task.f; jobs.push_back(j); src.push_back({(int)t, h});
                }
            }
            if (jobs.empty()) break;
            auto t0 = std::chrono::steady_clock::now();
            scs_forest_induce_batch(jobs.data(), jobs.size(), owner.data(), present.data());
            double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            total += dt; per.push_back(dt);
            std::vector<Task> next;
            for (size_t j = 0; j < jobs.size(); ++j) {
                Task &task = wave[src[j].first]; size_t half = task.taxa.size() / 2;
                Task c; c.f = jobs[j].out;
                if (src[j].second == 0) c.taxa.assign(task.taxa.begin(), task.taxa.begin() + half);
                else c.taxa.assign(task.taxa.begin() + half, task.taxa.end());
                for (int32_t x : c.taxa) { owner[x] = -1; present[x] = 0; }
                next.push_back(std::move(c));
            }
            for (Task &t : wave) scs_forest_destroy(t.f);
            wave.swap(next);
        }
        for (Task &t : wave) scs_forest_destroy(t.f);
        if (best.empty()) best = per; else for (size_t i = 0; i < per.size() && i < best.size(); ++i) best[i] = std::min(best[i], per[i]);
        if (rep < 11) continue;
        total = 0; for (double x : best) total += x; per = best;
        printf("best of 12: total %.1f ms over %zu waves:", 1e3 * total, per.size());
        for (double x : per) printf(" %.1f", 1e3 * x);
        printf("\n");
    }
}
