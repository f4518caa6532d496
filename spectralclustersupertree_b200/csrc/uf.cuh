// Lock-free union-find on vertex ids (shared by components.cu and the batched medium-node path, medium.cu).
// Roots are always hooked larger-under-smaller, so the final root of a component is its smallest vertex id
// whatever the interleaving: labels are deterministic.
#pragma once

#include <cstdint>

namespace scs {

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

}  // namespace scs
