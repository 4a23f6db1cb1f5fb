// Declarations for the multi-GPU plumbing (comm.cu).
#pragma once
#include <cstring>
#include <string>

#include "common.cuh"

namespace zk {
int comm_unique_id(uint8_t out[128], std::string& err);
int comm_init(zkaes_ctx* ctx, int rank, int nranks, const uint8_t unique_id[128]);
void comm_destroy(zkaes_ctx* ctx);
// every rank contributes bytes_per_rank bytes; recv_dev receives nranks * bytes_per_rank bytes in rank order
int comm_all_gather(zkaes_ctx* ctx, const void* send_dev, void* recv_dev, size_t bytes_per_rank);
int comm_broadcast(zkaes_ctx* ctx, void* buf, size_t bytes, int root);
void shard_range(size_t n, int rank, int nranks, size_t* start, size_t* count);
}  // namespace zk
