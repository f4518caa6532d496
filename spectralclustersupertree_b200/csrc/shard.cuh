// Row-sharded recursion nodes over several GPUs of one NVLink/NVSwitch box: device-side helpers.
//
// Every rank (one process per GPU) owns an *exchange window*: one cudaMalloc'd block, exported with
// cudaIpcGetMemHandle and mapped by every peer.  Ranks push what the others need (vector slices,
// adjacency-bit rows, degrees) straight into the peers' windows with ordinary stores over NVLink and
// then raise a flag there; consumers spin on a flag in their OWN memory.  A flag holds the number of
// the last synchronisation point its writer has passed ("epoch"); every rank walks the same sequence
// of synchronisation points (the kernels are deterministic and the host code is the same), so
// "flag[src] >= epoch" is the whole protocol.  See shard.cu for the layout and the host side.
#pragma once

#include "common.cuh"

namespace scs {

// window header (first 4 KB of the window)
struct ShardHeader {
    unsigned long long arrived[kMaxPeers];  // arrived[src] = last epoch rank `src` signalled to this rank
    unsigned int error;                     // latched by a spin loop that ran out of time
    unsigned int ticket;                    // CTA counter of the fused matvec (local use only)
};

struct PeerTable {
    unsigned char *window[kMaxPeers];
    int rank, world;
};

#ifdef __CUDACC__
__device__ __forceinline__ void peer_signal(unsigned long long *flag, unsigned long long epoch) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag), "l"(epoch) : "memory");
}

__device__ __forceinline__ unsigned long long peer_peek(const unsigned long long *flag) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory");
    return v;
}

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Spin until flag >= epoch.  Bounded: after timeout_ns the error word is latched and the wait gives
// up, so a lost peer turns into an error status on the host instead of a hung GPU.
__device__ __forceinline__ bool peer_wait(const unsigned long long *flag, unsigned long long epoch,
                                          unsigned long long timeout_ns, unsigned int *error) {
    if (peer_peek(flag) >= epoch) return true;
    const unsigned long long t0 = global_ns();
    unsigned int spins = 0;
    while (peer_peek(flag) < epoch) {
        if (*reinterpret_cast<volatile unsigned int *>(error)) return false;
        if ((++spins & 0x3ffu) == 0 && global_ns() - t0 > timeout_ns) {
            atomicExch(error, 1u);
            return false;
        }
        __nanosleep(64);
    }
    return true;
}

// Tell every peer (and ourselves) that this rank passed `epoch`, then wait until all of them did.
// Called by ONE warp; lane p talks to peer p.
__device__ __forceinline__ void peer_signal_and_wait(const PeerTable &peers, unsigned long long epoch,
                                                     unsigned long long timeout_ns) {
    const int lane = threadIdx.x & 31;
    ShardHeader *mine = reinterpret_cast<ShardHeader *>(peers.window[peers.rank]);
    __threadfence_system();
    if (lane < peers.world) {
        ShardHeader *theirs = reinterpret_cast<ShardHeader *>(peers.window[lane]);
        peer_signal(&theirs->arrived[peers.rank], epoch);
        peer_wait(&mine->arrived[lane], epoch, timeout_ns, &mine->error);
    }
    __syncwarp();
}
#endif

}  // namespace scs
