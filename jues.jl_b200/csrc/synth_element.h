// One element of the counter-based synthetic ERI generator (jues.jl_b200/synth.py: counter_eri_element),
// written with the narrowest integer types that are exact for n < 65536 basis functions: the pair indices fit
// 32 bits, the pair-of-pairs index needs ONE 32x32->64 multiply, and the 53-bit integer is converted to double
// as two exact 32-bit conversions.  (The generator is integer-ALU bound; the all-64-bit form took ~68 integer
// operations per element.)  Shared by the CUDA kernel (tensor_ops.cu) and by a host build that is compared
// bit for bit with the numpy generator on the CPU (tests/host/synth_element_test.cpp).
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define JUES_HD __host__ __device__ __forceinline__
#else
#define JUES_HD inline
#endif

namespace jues {

JUES_HD uint64_t synth_splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// pair index hi(hi+1)/2 + lo of two indices below 65536: fits 32 bits
JUES_HD uint32_t synth_pair32(uint32_t a, uint32_t b) {
    const uint32_t hi = a > b ? a : b, lo = a > b ? b : a;
    return hi * (hi + 1u) / 2u + lo;
}

// value for the pair indices P = pair(mu, nu), Q = pair(lam, sig)
JUES_HD double synth_value(uint32_t P, uint32_t Q, uint64_t seed, double scale) {
    const uint32_t h2 = P > Q ? P : Q, l2 = P > Q ? Q : P;
    const uint64_t K = (uint64_t)h2 * (uint64_t)(h2 + 1u) / 2u + l2;       // one wide multiply
    const uint64_t m = synth_splitmix64(seed ^ K) >> 11;                    // 53 bits
    const double u = ((double)(uint32_t)(m >> 32) * 4294967296.0 + (double)(uint32_t)m) * (1.0 / 9007199254740992.0);
    return scale * (2.0 * u - 1.0);
}

}  // namespace jues
