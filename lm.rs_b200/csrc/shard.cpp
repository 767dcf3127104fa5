// shard.cpp -- NCCL plumbing behind shard.h (dlopen'ed; ABI subset of nccl.h 2.27).
#include "shard.h"

#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>

namespace {
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclUint8 = 1, ncclFloat = 7, ncclSum = 0 };
struct Api {
    void* h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
} api;
thread_local std::string g_serr;
int sfail(const std::string& m) { g_serr = m; return 1; }
int load() {
    if (api.h) return 0;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        api.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (api.h) break;
    }
    if (!api.h) return sfail(std::string("cannot dlopen libnccl.so.2: ") + dlerror());
#define SYM(f)                                                         \
    *(void**)(&api.f) = dlsym(api.h, "nccl" #f);                       \
    if (!api.f) return sfail("libnccl lacks nccl" #f);
    SYM(GetUniqueId) SYM(CommInitRank) SYM(CommDestroy) SYM(AllReduce) SYM(AllGather) SYM(GetErrorString)
#undef SYM
    return 0;
}
int ck(ncclResult_t r, const char* what) {
    if (r == 0) return 0;
    return sfail(std::string(what) + ": " + (api.GetErrorString ? api.GetErrorString(r) : "nccl error"));
}
}  // namespace

const char* shard_error() { return g_serr.c_str(); }

int shard_unique_id(void* out128) {
    if (load()) return 1;
    ncclUniqueId id;
    if (ck(api.GetUniqueId(&id), "ncclGetUniqueId")) return 1;
    memcpy(out128, &id, 128);
    return 0;
}
int shard_init(Shard& s, int rank, int world, const void* unique_id, int) {
    if (load()) return 1;
    ncclUniqueId id;
    memcpy(&id, unique_id, 128);
    // This communicator carries two all-reduces of a few KB per block (fallback data path) or a single 64-byte all-gather
    // (peer mode): NVLink-SHARP multicast buffers buy nothing here and are one more thing to tear down at exit
    setenv("NCCL_NVLS_ENABLE", "0", 0);
    ncclComm_t c;
    if (ck(api.CommInitRank(&c, world, id, rank), "ncclCommInitRank")) return 1;
    s.rank = rank; s.world = world; s.comm = c;
    return 0;
}
int shard_allreduce(Shard& s, float* buf, size_t count, cudaStream_t stream) {
    return ck(api.AllReduce(buf, buf, count, ncclFloat, ncclSum, (ncclComm_t)s.comm, stream), "ncclAllReduce");
}
int shard_allgather_logits(Shard& s, float* logits, size_t per_rank, cudaStream_t stream) {
    return ck(api.AllGather(logits + (size_t)s.rank * per_rank, logits, per_rank, ncclFloat, (ncclComm_t)s.comm, stream),
              "ncclAllGather");
}
// `bytes` from every rank, in rank order (device buffers): carries the IPC handles of the peer-exchange buffers at load time
int shard_allgather_bytes(Shard& s, const void* d_in, void* d_out, size_t bytes, cudaStream_t stream) {
    return ck(api.AllGather(d_in, d_out, bytes, ncclUint8, (ncclComm_t)s.comm, stream), "ncclAllGather");
}
void shard_destroy(Shard& s) {
    if (s.comm && api.CommDestroy) api.CommDestroy((ncclComm_t)s.comm);
    s.comm = nullptr;
}
