// Collectives over the context's NCCL communicator (no-ops when nranks == 1).  Internal.
#pragma once
#include "jues_common.h"
#include <functional>

namespace jues {
// in-place all-gather: rank r contributed full[r*count .. (r+1)*count)
void all_gather_inplace(jues_ctx* ctx, double* full, size_t count_per_rank);
void all_reduce_sum(jues_ctx* ctx, double* buf, size_t count);
double all_reduce_scalar(jues_ctx* ctx, double x);
// personalised exchange: block [send_off[d], +send_cnt[d]) of `send` goes to rank d and lands at
// [recv_off[me-as-seen-by-d] ...) there, i.e. this rank receives recv_cnt[r] elements from rank r at
// recv + recv_off[r].  Arrays have nranks entries; send and recv must not overlap.  With one rank: a copy.
void all_to_all_v(jues_ctx* ctx, const double* send, const size_t* send_off, const size_t* send_cnt,
                  double* recv, const size_t* recv_off, const size_t* recv_cnt);   // sum of a host scalar over ranks (blocking)
// single-process multi-GPU: run fn(member, is_leader) on every member of the leader's group (dist.cu)
int group_run(jues_ctx* lead, const std::function<int(jues_ctx*, bool)>& fn);
// true when an entry point called on `ctx` must fan out over the group
inline bool is_group_call(const jues_ctx* ctx) { return ctx && ctx->group && !ctx->in_group_call && ctx->group->size() > 1; }
// equal slabs of the virtual extent (v is a multiple of 2*nranks): [b0, b0+vs)
inline void slab_of(const jues_ctx* ctx, int64_t v, int64_t* b0, int64_t* vs) {
    *vs = v / ctx->nranks;
    *b0 = *vs * ctx->rank;
}
}  // namespace jues
