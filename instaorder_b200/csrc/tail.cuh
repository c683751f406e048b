// Stem max-pool and fused pool + FC tail (declarations).
#pragma once
#include <algorithm>

#include "common.cuh"

namespace io {
int maxpool_launch(const void* x, void* y, int b, int h, int w, int c, cudaStream_t stream);
int tail_launch(const void* feat, int hw, int pairs, const float* fcw, const float* fcb, int k_total, float* logits,
                cudaStream_t stream);
}  // namespace io
