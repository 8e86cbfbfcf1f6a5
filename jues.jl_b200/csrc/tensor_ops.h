// Element-wise / permutation / reduction kernels.  Internal.
#pragma once
#include "jues_common.h"

namespace jues {
int ew_grid(jues_ctx* ctx, size_t n, int threads);
}  // namespace jues
