// Declarations for the multi-GPU plumbing (comm.cu).
#pragma once
#include <cstring>
#include <string>
#include <vector>

#include "common.cuh"

namespace zk {
int comm_unique_id(uint8_t out[128], std::string& err);
int comm_init(zkaes_ctx* ctx, int rank, int nranks, const uint8_t unique_id[128]);
void comm_destroy(zkaes_ctx* ctx);
// every rank contributes bytes_per_rank bytes; recv_dev receives nranks * bytes_per_rank bytes in rank order
int comm_all_gather(zkaes_ctx* ctx, const void* send_dev, void* recv_dev, size_t bytes_per_rank);
int comm_broadcast(zkaes_ctx* ctx, void* buf, size_t bytes, int root);
// point-to-point (NCCL send / recv on the context's stream): a coset's helper rank ships a finished |K|- or |H|-sized transform to the
// coset's owner.  Matching is by issue order per (sender, receiver) pair; callers iterate one global task list on every rank.
int comm_group_start(zkaes_ctx* ctx);  // ncclGroupStart / ncclGroupEnd: the sends and receives in between progress concurrently
int comm_group_end(zkaes_ctx* ctx);
int comm_send(zkaes_ctx* ctx, const void* buf, size_t bytes, int peer);
int comm_recv(zkaes_ctx* ctx, void* buf, size_t bytes, int peer);
void shard_range(size_t n, int rank, int nranks, size_t* start, size_t* count);

// ---- work split of the prover's coset evaluations (rounds 2 and 3) over the ranks ------------------------------------------------------
// `ncoset` cosets, each `ntask` forward transforms (needed by the coset's owner in task order 0..ntask-1) plus `own_extra` transforms' worth
// of work that stays with the owner (the inverse transform and the pointwise products).  owner[j] = j mod nranks, as before; but with
// more ranks than cosets the other ranks used to idle (and with fewer, the rank owning an extra coset was the straggler), so forward
// transforms are handed to the least loaded rank, last-needed task first, while that shortens the critical path.  A rank that
// hands work out never takes work in (helpers -> owners only): the send / recv graph has no cycle, so stream-ordered NCCL
// point-to-point calls cannot deadlock.  Deterministic: every rank derives the same plan.
struct CosetPlan {
    int nranks = 1, ncoset = 0, ntask = 0;
    std::vector<int> owner;  // [ncoset]
    std::vector<int> exec;   // [ncoset * ntask]: the rank that computes forward transform p of coset j
    int executor(int j, int p) const { return exec[(size_t)j * ntask + p]; }
};
inline CosetPlan coset_plan(int nranks, int ncoset, int ntask, double own_extra, double xfer = 0.15) {
    CosetPlan pl;
    pl.nranks = nranks;
    pl.ncoset = ncoset;
    pl.ntask = ntask;
    pl.owner.resize(ncoset);
    pl.exec.resize((size_t)ncoset * ntask);
    std::vector<double> load(nranks, 0.0);
    std::vector<int> local(ncoset, ntask);  // forward transforms of coset j still with its owner
    std::vector<char> gives(nranks, 0), takes(nranks, 0);
    for (int j = 0; j < ncoset; ++j) {
        pl.owner[j] = j % nranks;
        load[pl.owner[j]] += ntask + own_extra;
        for (int p = 0; p < ntask; ++p) pl.exec[(size_t)j * ntask + p] = pl.owner[j];
    }
    for (;;) {
        int R = -1, S = -1;
        for (int r = 0; r < nranks; ++r) {
            if (takes[r]) continue;
            bool has = false;
            for (int j = 0; j < ncoset; ++j) has = has || (pl.owner[j] == r && local[j] > 0);
            if (has && (R < 0 || load[r] > load[R])) R = r;
        }
        if (R < 0) break;
        for (int r = 0; r < nranks; ++r)
            if (r != R && !gives[r] && (S < 0 || load[r] < load[S])) S = r;
        if (S < 0 || load[S] + 1.0 + xfer >= load[R]) break;
        int j = -1;
        for (int c = ncoset - 1; c >= 0; --c)  // the owner's LAST coset first: a helper reaches it when it has finished its own
            if (pl.owner[c] == R && local[c] > 0 && (j < 0 || local[c] > local[j])) j = c;
        const int p = --local[j];  // last-needed task first
        pl.exec[(size_t)j * ntask + p] = S;
        load[R] -= 1.0;
        load[S] += 1.0 + xfer;
        gives[R] = 1;
        takes[S] = 1;
    }
    return pl;
}
}  // namespace zk
