// Peer-memory exchange of small vectors between the ranks of a data-parallel group (CUDA IPC mappings over NVLink / NVSwitch,
// set up in dcnr_comm_create).  Used by the SyncBN kernels: one CTA stores its vector into every rank's staging slot
// (peer-to-peer stores), publishes an epoch flag on every rank, spins on its own flags until every rank has published, then
// reads all vectors locally -- an all-gather of a few KB in the middle of a kernel, no NCCL call, no extra launch.
#pragma once

#include <stdint.h>

namespace dcnr {

constexpr int kP2pSlots = 4, kP2pCap = 1040;          // 2 * 512 + 2 doubles per rank and slot: hidden_dim up to 512
struct P2pView {           // one rank's staging allocation as mapped into this process
    double *data;          // [kP2pSlots][world][kP2pCap]
    uint32_t *flags;       // [kP2pSlots][world]
    uint32_t *epoch;       // 1 (only the owner's is used)
};

// device array [world] of the views of `comm`, or NULL when the ranks could not map each other (NCCL is used then)
const P2pView *comm_p2p_views(const void *comm);

#ifdef __CUDACC__
// Epochs advance identically on all ranks (the same sequence of exchanges runs everywhere), live on the device (so a captured
// CUDA graph replays correctly) and select one of kP2pSlots slots: a rank can start exchange e + 2 only after every rank has
// written e + 1, i.e. after every rank has finished reading e, so two slots would already be enough.
struct P2pExchange {
    const P2pView *views;
    int rank, world, slot;
    uint32_t e;
    __device__ __forceinline__ P2pExchange(const P2pView *v, int r, int w) : views(v), rank(r), world(w) {
        e = *v[r].epoch + 1u;
        slot = (int)(e % kP2pSlots);
    }
    // where this rank's vector goes on rank r (a peer-to-peer store for r != rank)
    __device__ __forceinline__ double *send_slot(int r) const { return views[r].data + ((size_t)slot * world + rank) * kP2pCap; }
    // rank r's vector on this rank, valid after publish_and_wait()
    __device__ __forceinline__ const volatile double *recv_slot(int r) const {
        return views[rank].data + ((size_t)slot * world + r) * kP2pCap;
    }
    // all threads of the CTA (blockDim.x >= world): makes the stores above visible, signals every rank, waits for every rank
    __device__ __forceinline__ void publish_and_wait() const {
        __threadfence_system();
        __syncthreads();
        if ((int)threadIdx.x < world) {
            volatile uint32_t *theirs = views[threadIdx.x].flags + (size_t)slot * world + rank;
            *theirs = e;
            volatile uint32_t *here = views[rank].flags + (size_t)slot * world + threadIdx.x;
            uint32_t spins = 0;
            while (*here != e)
                if (++spins > (1u << 28)) __trap();                   // a rank that never arrives must not hang the GPU
        }
        __threadfence_system();
        __syncthreads();
    }
    // after the last read of recv_slot (all threads)
    __device__ __forceinline__ void finish() const {
        __syncthreads();
        if (threadIdx.x == 0) *views[rank].epoch = e;
    }
};
#endif

}  // namespace dcnr
