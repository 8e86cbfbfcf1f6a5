// Multi-GPU plumbing: one process per GPU, NCCL loaded at run time (dlopen) so that the library
// has no link-time dependency on a particular NCCL build.
#include "api_util.h"

namespace jues {
void dist_teardown(jues_ctx* ctx) { (void)ctx; }
}  // namespace jues

extern "C" int jues_b200_nccl_unique_id(unsigned char id_out[128]) {
    (void)id_out;
    return JUES_B200_ENCCL;
}
extern "C" int jues_b200_init_dist(jues_ctx* ctx, int rank, int nranks, const unsigned char id[128]) {
    (void)id;
    if (!ctx) return JUES_B200_EINVAL;
    ctx->rank = rank;
    ctx->nranks = nranks;
    if (nranks == 1) return JUES_B200_OK;
    ctx->last_error = "multi-GPU not implemented yet";
    return JUES_B200_ENCCL;
}
