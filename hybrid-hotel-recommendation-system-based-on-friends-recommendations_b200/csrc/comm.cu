// Data-parallel communicator of the library: one NCCL communicator per process (one process per GPU,
// NVLink 5 / NVSwitch underneath), used INSIDE the stream-ordered whole-model calls for the exchanges
// the reference's single-device semantics need under data parallelism:
//   * BatchNorm batch statistics over the GLOBAL batch (forward moments, backward sum(dy), sum(dy*xhat)):
//     all-gather of a few KB per layer, folded in rank order on every rank -> identical on all ranks;
//   * the gradient all-reduce and the all-to-all of the row-sharded embedding exchange (host API).
// The BatchNorm exchanges -- 16 per training step, ~4 KB each, pure latency -- do not go through NCCL when the ranks can
// map each other's memory (CUDA IPC over NVLink / NVSwitch, set up once in dcnr_comm_create): ONE kernel stores this rank's
// vector into every peer's staging buffer (peer-to-peer stores), publishes a per-slot epoch flag, spins on the local flags
// of all ranks and copies the gathered vectors out (k_p2p_allgather below).  An NCCL all-gather of the same 4 KB costs
// ~20-30 us of launch + protocol latency; this costs one small kernel.
// NCCL is reached through dlopen (the copy PyTorch already loaded, else DCNR_NCCL_LIB, else
// libnccl.so.2), so libdcnr_sm100a.so has no link-time dependency and single-GPU use never touches it.
#include <dlfcn.h>

#include "kernels.cuh"
#include "p2p.cuh"

namespace dcnr {

// minimal NCCL ABI (stable since 2.x): see nccl.h
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclInt8 = 0, ncclUint8 = 1, ncclFloat32 = 7, ncclFloat64 = 8 };
enum { ncclSum = 0 };

struct NcclApi {
    int (*GetUniqueId)(ncclUniqueId *);
    int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
    int (*CommDestroy)(ncclComm_t);
    int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t);
    int (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t);
    int (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t);
    int (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t);
    int (*GroupStart)();
    int (*GroupEnd)();
    const char *(*GetErrorString)(int);
    bool ok = false;
};

static NcclApi *nccl() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api.ok ? &api : nullptr;
    tried = true;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);            // the copy the host process already uses
    const char *env = getenv("DCNR_NCCL_LIB");
    if (h == nullptr && env != nullptr) h = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
    if (h == nullptr) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (h == nullptr) {
        set_error("NCCL not found (dlopen libnccl.so.2 failed: %s)", dlerror());
        return nullptr;
    }
#define DCNR_SYM(name)                                                         \
    *reinterpret_cast<void **>(&api.name) = dlsym(h, "nccl" #name);            \
    if (api.name == nullptr) {                                                 \
        set_error("NCCL symbol nccl" #name " missing");                        \
        return nullptr;                                                        \
    }
    DCNR_SYM(GetUniqueId) DCNR_SYM(CommInitRank) DCNR_SYM(CommDestroy) DCNR_SYM(AllReduce) DCNR_SYM(AllGather)
    DCNR_SYM(Send) DCNR_SYM(Recv) DCNR_SYM(GroupStart) DCNR_SYM(GroupEnd) DCNR_SYM(GetErrorString)
#undef DCNR_SYM
    api.ok = true;
    return &api;
}

// Peer-memory staging of one rank (p2p.cuh): kP2pSlots x world vectors of kP2pCap doubles, then kP2pSlots x world epoch
// flags, then this rank's epoch counter.  Every rank maps every rank's staging buffer (its own included) at creation time.
struct Comm {
    ncclComm_t comm;
    int rank, world;
    void *p2p_local = nullptr;            // this rank's staging allocation (cudaMalloc: exportable by CUDA IPC)
    void *p2p_peer[kMaxWorld] = {};       // every rank's allocation as mapped into this process
    P2pView *p2p_views = nullptr;         // device array [world] of the views above (NULL: no peer path, NCCL is used)
    P2pView *p2p_mapped = nullptr;        // the same array, kept while dcnr_comm_set_peer_memory(comm, 0) parks the peer path
};

static size_t p2p_bytes(int world) {
    return (size_t)kP2pSlots * world * kP2pCap * sizeof(double) + (size_t)kP2pSlots * world * sizeof(uint32_t) + 256;
}
static P2pView p2p_view(void *base, int world) {
    P2pView v;
    v.data = reinterpret_cast<double *>(base);
    v.flags = reinterpret_cast<uint32_t *>(v.data + (size_t)kP2pSlots * world * kP2pCap);
    v.epoch = v.flags + (size_t)kP2pSlots * world;
    return v;
}

// all[r * n + i] = rank r's mine[i].  One CTA (p2p.cuh has the protocol).
__global__ void __launch_bounds__(256)
k_p2p_allgather(const P2pView *__restrict__ views, const double *__restrict__ mine, double *__restrict__ all, int n,
                int rank, int world) {
    const P2pExchange x(views, rank, world);
    for (int r = 0; r < world; ++r) {
        double *dst = x.send_slot(r);
        for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = mine[i];
    }
    x.publish_and_wait();
    for (int r = 0; r < world; ++r) {
        const volatile double *src = x.recv_slot(r);
        for (int i = threadIdx.x; i < n; i += blockDim.x) all[(size_t)r * n + i] = src[i];
    }
    x.finish();
}

#define DCNR_NCCL_CHECK(expr)                                                             \
    do {                                                                                  \
        int _r = (expr);                                                                  \
        if (_r != ncclSuccess) {                                                          \
            set_error("NCCL: %s failed: %s", #expr, nccl()->GetErrorString(_r));          \
            return DCNR_ERR_CUDA;                                                         \
        }                                                                                 \
    } while (0)

const P2pView *comm_p2p_views(const void *c) { return c ? static_cast<const Comm *>(c)->p2p_views : nullptr; }
int comm_world(const void *c) { return c ? static_cast<const Comm *>(c)->world : 1; }
int comm_rank(const void *c) { return c ? static_cast<const Comm *>(c)->rank : 0; }

// small all-gather of doubles (the SyncBN exchanges): peer-memory kernel when the ranks mapped each other, NCCL otherwise
int comm_allgather_f64(const void *c, const double *send, double *recv, int n, cudaStream_t stream) {
    const Comm *cm = static_cast<const Comm *>(c);
    DCNR_REQUIRE(cm != nullptr, "no communicator");
    if (cm->p2p_views != nullptr && n <= kP2pCap) {
        k_p2p_allgather<<<1, 256, 0, stream>>>(cm->p2p_views, send, recv, n, cm->rank, cm->world);
        DCNR_LAUNCHED();
        return DCNR_OK;
    }
    return comm_allgather(c, send, recv, (int64_t)n * 8, stream);
}

int comm_allgather(const void *c, const void *send, void *recv, int64_t bytes_per_rank, cudaStream_t stream) {
    const Comm *cm = static_cast<const Comm *>(c);
    DCNR_REQUIRE(cm != nullptr && nccl() != nullptr, "no communicator");
    DCNR_NCCL_CHECK(nccl()->AllGather(send, recv, (size_t)bytes_per_rank, ncclInt8, cm->comm, stream));
    return DCNR_OK;
}

}  // namespace dcnr

using namespace dcnr;

extern "C" int dcnr_comm_unique_id(uint8_t *id_host) {
    DCNR_REQUIRE(id_host != nullptr, "null argument");
    NcclApi *a = nccl();
    if (a == nullptr) return DCNR_ERR_CUDA;
    ncclUniqueId id;
    DCNR_NCCL_CHECK(a->GetUniqueId(&id));
    memcpy(id_host, id.internal, 128);
    return DCNR_OK;
}

extern "C" int dcnr_comm_create(const uint8_t *id_host, int32_t rank, int32_t world, void **comm_out) {
    DCNR_REQUIRE(id_host != nullptr && comm_out != nullptr && world >= 1 && rank >= 0 && rank < world, "bad argument");
    NcclApi *a = nccl();
    if (a == nullptr) return DCNR_ERR_CUDA;
    ncclUniqueId id;
    memcpy(id.internal, id_host, 128);
    Comm *c = new Comm{nullptr, rank, world};
    int r = a->CommInitRank(&c->comm, world, id, rank);
    if (r != ncclSuccess) {
        set_error("NCCL: ncclCommInitRank failed: %s", a->GetErrorString(r));
        delete c;
        return DCNR_ERR_CUDA;
    }
    // Peer-memory path for the small exchanges: every rank exports its staging buffer (CUDA IPC), the handles travel through
    // one NCCL all-gather, every rank maps all of them.  Any failure simply leaves the NCCL path in place.
    if (world > 1 && world <= kMaxWorld) {
        bool ok = cudaMalloc(&c->p2p_local, p2p_bytes(world)) == cudaSuccess &&
                  cudaMemset(c->p2p_local, 0, p2p_bytes(world)) == cudaSuccess;
        cudaIpcMemHandle_t mine, *all_h = nullptr;
        char *dev_h = nullptr;
        ok = ok && cudaIpcGetMemHandle(&mine, c->p2p_local) == cudaSuccess;
        ok = ok && cudaMalloc(&dev_h, sizeof(mine) * (world + 1)) == cudaSuccess;
        if (ok) {
            all_h = new cudaIpcMemHandle_t[world];
            ok = cudaMemcpy(dev_h, &mine, sizeof(mine), cudaMemcpyHostToDevice) == cudaSuccess &&
                 a->AllGather(dev_h, dev_h + sizeof(mine), sizeof(mine), ncclInt8, c->comm, (cudaStream_t)0) == ncclSuccess &&
                 cudaStreamSynchronize((cudaStream_t)0) == cudaSuccess &&
                 cudaMemcpy(all_h, dev_h + sizeof(mine), sizeof(mine) * world, cudaMemcpyDeviceToHost) == cudaSuccess;
        }
        P2pView views[kMaxWorld];
        for (int r = 0; ok && r < world; ++r) {
            if (r == rank) c->p2p_peer[r] = c->p2p_local;
            else ok = cudaIpcOpenMemHandle(&c->p2p_peer[r], all_h[r], cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
            if (ok) views[r] = p2p_view(c->p2p_peer[r], world);
        }
        // every rank must agree on the path: a rank that failed to map its peers reports it through a second tiny all-gather
        int *dev_ok = reinterpret_cast<int *>(dev_h);
        int host_ok[kMaxWorld + 1] = {ok ? 1 : 0};
        bool agreed = dev_h != nullptr && cudaMemcpy(dev_ok, host_ok, sizeof(int), cudaMemcpyHostToDevice) == cudaSuccess &&
                      a->AllGather(dev_ok, dev_ok + 1, sizeof(int), ncclInt8, c->comm, (cudaStream_t)0) == ncclSuccess &&
                      cudaStreamSynchronize((cudaStream_t)0) == cudaSuccess &&
                      cudaMemcpy(host_ok + 1, dev_ok + 1, sizeof(int) * world, cudaMemcpyDeviceToHost) == cudaSuccess;
        for (int r = 0; agreed && r < world; ++r) agreed = host_ok[1 + r] == 1;
        if (agreed && cudaMalloc(&c->p2p_views, sizeof(P2pView) * world) == cudaSuccess)
            agreed = cudaMemcpy(c->p2p_views, views, sizeof(P2pView) * world, cudaMemcpyHostToDevice) == cudaSuccess;
        if (!agreed) {
            if (c->p2p_views != nullptr) cudaFree(c->p2p_views);
            c->p2p_views = nullptr;
        }
        c->p2p_mapped = c->p2p_views;
        if (dev_h != nullptr) cudaFree(dev_h);
        delete[] all_h;
        cudaGetLastError();               // a failed IPC probe is not an error of this call
    }
    *comm_out = c;
    return DCNR_OK;
}

extern "C" int dcnr_comm_destroy(void *comm) {
    if (comm == nullptr) return DCNR_OK;
    Comm *c = static_cast<Comm *>(comm);
    for (int r = 0; r < c->world && r < kMaxWorld; ++r)
        if (c->p2p_peer[r] != nullptr && c->p2p_peer[r] != c->p2p_local) cudaIpcCloseMemHandle(c->p2p_peer[r]);
    if (c->p2p_mapped != nullptr) cudaFree(c->p2p_mapped);
    if (c->p2p_local != nullptr) cudaFree(c->p2p_local);
    if (nccl() != nullptr && c->comm != nullptr) nccl()->CommDestroy(c->comm);
    delete c;
    return DCNR_OK;
}

extern "C" int dcnr_comm_info(const void *comm, int32_t *rank, int32_t *world) {
    if (rank) *rank = comm_rank(comm);
    if (world) *world = comm_world(comm);
    return DCNR_OK;
}

extern "C" int dcnr_comm_uses_peer_memory(const void *comm) {
    return comm != nullptr && static_cast<const Comm *>(comm)->p2p_views != nullptr ? 1 : 0;
}

extern "C" int dcnr_comm_set_peer_memory(void *comm, int32_t enable) {
    DCNR_REQUIRE(comm != nullptr, "dcnr_comm_set_peer_memory: null communicator");
    Comm *c = static_cast<Comm *>(comm);
    c->p2p_views = enable ? c->p2p_mapped : nullptr;
    return DCNR_OK;
}

extern "C" int dcnr_comm_allreduce_f32(void *comm, float *buf, int64_t count, dcnr_stream_t stream) {
    Comm *c = static_cast<Comm *>(comm);
    DCNR_REQUIRE(c != nullptr && buf != nullptr && count >= 0, "bad argument");
    if (c->world == 1 || count == 0) return DCNR_OK;
    DCNR_NCCL_CHECK(nccl()->AllReduce(buf, buf, (size_t)count, ncclFloat32, ncclSum, c->comm, as_stream(stream)));
    return DCNR_OK;
}

extern "C" int dcnr_comm_allgather(void *comm, const void *send, void *recv, int64_t bytes_per_rank, dcnr_stream_t stream) {
    DCNR_REQUIRE(comm != nullptr && send != nullptr && recv != nullptr && bytes_per_rank >= 0, "bad argument");
    return comm_allgather(comm, send, recv, bytes_per_rank, as_stream(stream));
}

// Variable all-to-all: rank r sends send[send_off[p] .. +send_bytes[p]) to every peer p and receives
// recv_bytes[p] bytes from p at recv_off[p] (all four arrays are HOST arrays of `world` entries).
extern "C" int dcnr_comm_alltoallv(void *comm, const void *send, const int64_t *send_bytes_host, const int64_t *send_off_host,
                                   void *recv, const int64_t *recv_bytes_host, const int64_t *recv_off_host,
                                   dcnr_stream_t stream) {
    Comm *c = static_cast<Comm *>(comm);
    DCNR_REQUIRE(c != nullptr && send_bytes_host && send_off_host && recv_bytes_host && recv_off_host, "bad argument");
    NcclApi *a = nccl();
    DCNR_REQUIRE(a != nullptr, "no communicator");
    DCNR_NCCL_CHECK(a->GroupStart());
    // an error inside the group must not leave it open: remember the first failure, stop queueing, ALWAYS call GroupEnd
    int failed = ncclSuccess;
    const char *what = "";
    for (int p = 0; p < c->world && failed == ncclSuccess; ++p) {
        if (send_bytes_host[p] > 0) {
            failed = a->Send(static_cast<const char *>(send) + send_off_host[p], (size_t)send_bytes_host[p], ncclInt8, p, c->comm,
                             as_stream(stream));
            what = "ncclSend";
        }
        if (failed == ncclSuccess && recv_bytes_host[p] > 0) {
            failed = a->Recv(static_cast<char *>(recv) + recv_off_host[p], (size_t)recv_bytes_host[p], ncclInt8, p, c->comm,
                             as_stream(stream));
            what = "ncclRecv";
        }
    }
    const int ended = a->GroupEnd();
    if (failed != ncclSuccess) {
        set_error("NCCL: %s failed inside the all-to-all group: %s", what, a->GetErrorString(failed));
        return DCNR_ERR_CUDA;
    }
    if (ended != ncclSuccess) {
        set_error("NCCL: ncclGroupEnd failed: %s", a->GetErrorString(ended));
        return DCNR_ERR_CUDA;
    }
    return DCNR_OK;
}
