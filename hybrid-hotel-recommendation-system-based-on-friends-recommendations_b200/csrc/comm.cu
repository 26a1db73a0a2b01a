// Data-parallel communicator of the library: one NCCL communicator per process (one process per GPU,
// NVLink 5 / NVSwitch underneath), used INSIDE the stream-ordered whole-model calls for the exchanges
// the reference's single-device semantics need under data parallelism:
//   * BatchNorm batch statistics over the GLOBAL batch (forward moments, backward sum(dy), sum(dy*xhat)):
//     all-gather of a few KB per layer, folded in rank order on every rank -> identical on all ranks;
//   * the gradient all-reduce and the all-to-all of the row-sharded embedding exchange (host API).
// NCCL is reached through dlopen (the copy PyTorch already loaded, else DCNR_NCCL_LIB, else
// libnccl.so.2), so libdcnr_sm100a.so has no link-time dependency and single-GPU use never touches it.
#include <dlfcn.h>

#include "kernels.cuh"

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

struct Comm {
    ncclComm_t comm;
    int rank, world;
};

#define DCNR_NCCL_CHECK(expr)                                                             \
    do {                                                                                  \
        int _r = (expr);                                                                  \
        if (_r != ncclSuccess) {                                                          \
            set_error("NCCL: %s failed: %s", #expr, nccl()->GetErrorString(_r));          \
            return DCNR_ERR_CUDA;                                                         \
        }                                                                                 \
    } while (0)

int comm_world(const void *c) { return c ? static_cast<const Comm *>(c)->world : 1; }
int comm_rank(const void *c) { return c ? static_cast<const Comm *>(c)->rank : 0; }

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
    *comm_out = c;
    return DCNR_OK;
}

extern "C" int dcnr_comm_destroy(void *comm) {
    if (comm == nullptr) return DCNR_OK;
    Comm *c = static_cast<Comm *>(comm);
    if (nccl() != nullptr && c->comm != nullptr) nccl()->CommDestroy(c->comm);
    delete c;
    return DCNR_OK;
}

extern "C" int dcnr_comm_info(const void *comm, int32_t *rank, int32_t *world) {
    if (rank) *rank = comm_rank(comm);
    if (world) *world = comm_world(comm);
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
