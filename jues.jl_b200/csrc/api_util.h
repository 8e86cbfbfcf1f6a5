// Helpers shared by the extern "C" translation units.
#pragma once
#include "jues_common.h"

namespace jues {

extern std::string g_init_error;

// An entry point that does not fan out over a single-process multi-GPU group runs on the leader alone:
// for its duration the leader is a one-rank context (no collective may wait for members that are not there).
struct SoloGuard {
    jues_ctx* ctx;
    int nranks, rank;
    bool active;
    explicit SoloGuard(jues_ctx* c) : ctx(c), nranks(c->nranks), rank(c->rank),
                                      active(c->group != nullptr && !c->in_group_call) {
        if (active) { ctx->nranks = 1; ctx->rank = 0; }
    }
    ~SoloGuard() { if (active) { ctx->nranks = nranks; ctx->rank = rank; } }
};

// exception -> status code conversion at the ABI boundary
#define JUES_API_BEGIN(ctx)                                   \
    if (!(ctx)) return JUES_B200_EINVAL;                      \
    try {                                                     \
        cudaError_t sd__ = cudaSetDevice((ctx)->device);      \
        if (sd__ != cudaSuccess) {                            \
            (ctx)->last_error = cudaGetErrorString(sd__);     \
            return JUES_B200_ECUDA;                           \
        }                                                     \
        ::jues::SoloGuard solo__(ctx);

#define JUES_API_END(ctx)                                     \
        return JUES_B200_OK;                                  \
    } catch (const ::jues::Error& e__) {                      \
        (ctx)->last_error = e__.what();                       \
        cudaGetLastError();                                   \
        return e__.code;                                      \
    } catch (const std::bad_alloc&) {                         \
        (ctx)->last_error = "host allocation failed";         \
        return JUES_B200_ENOMEM;                              \
    } catch (const std::exception& e__) {                     \
        (ctx)->last_error = e__.what();                       \
        return JUES_B200_ECUDA;                               \
    }

// testing hook: JUES_B200_GEMM_CFG=<n> forces a GEMM tile configuration
inline int ctx_force_cfg(jues_ctx*) {
    const char* s = getenv("JUES_B200_GEMM_CFG");
    return s ? atoi(s) : -1;
}

// deterministic pseudo-random fill in [-1,1) (tensor_ops.cu)
void fill_pattern(jues_ctx* ctx, double* p, size_t n, unsigned long long seed);

// NCCL teardown (dist.cu)
void dist_teardown(jues_ctx* ctx);

}  // namespace jues
