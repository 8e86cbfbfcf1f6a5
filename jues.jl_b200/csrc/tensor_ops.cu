// Element-wise, permutation and reduction kernels (HBM-bound side of the path).
#include "tensor_ops.h"
#include "api_util.h"

namespace jues {

namespace {
__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull;
    unsigned long long z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__global__ void fill_pattern_kernel(double* __restrict__ p, size_t n, unsigned long long seed) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        unsigned long long h = splitmix64(seed * 0x100000001B3ull ^ i);
        p[i] = 2.0 * ((double)(h >> 11) * (1.0 / 9007199254740992.0)) - 1.0;
    }
}
}  // namespace

int ew_grid(jues_ctx* ctx, size_t n, int threads) {
    size_t blocks = (n + threads - 1) / threads;
    size_t cap = (size_t)ctx->sm_count * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

void fill_pattern(jues_ctx* ctx, double* p, size_t n, unsigned long long seed) {
    fill_pattern_kernel<<<ew_grid(ctx, n, 256), 256, 0, ctx->stream>>>(p, n, seed);
    JUES_CUDA(cudaGetLastError());
    ctx->stats.aux_launches++;
}

}  // namespace jues
