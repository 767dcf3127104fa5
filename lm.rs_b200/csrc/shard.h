// shard.h -- multi-GPU exchange for the row-sharded forward (one process per GPU).  The reference has no
// distributed code at all; the sharding scheme is the one BASELINE.json's north_star defines:
// two sums of the dim-sized residual contribution per block + one logits gather per token.  Default data path: the peer
// exchange fused into the kernels (common.cuh: every GPU pushes its partial into all peers' buffers over NVLink); NCCL
// only bootstraps it (all-gather of the IPC handles) and remains as the fallback data path (LMRS_B200_PEER=0).
// NCCL is loaded lazily with dlopen so that the single-GPU library has no NCCL dependency.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

struct Shard {
    int rank = 0, world = 1;
    void* comm = nullptr;   // ncclComm_t
};
int shard_unique_id(void* out128);
int shard_init(Shard& s, int rank, int world, const void* unique_id, int dim);
int shard_allreduce(Shard& s, float* buf, size_t count, cudaStream_t stream);
int shard_allgather_logits(Shard& s, float* logits, size_t per_rank, cudaStream_t stream);
int shard_allgather_bytes(Shard& s, const void* d_in, void* d_out, size_t bytes, cudaStream_t stream);
void shard_destroy(Shard& s);
const char* shard_error();
