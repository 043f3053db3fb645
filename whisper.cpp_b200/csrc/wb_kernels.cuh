// wb_kernels.cuh -- launchers of the non-GEMM kernels (see wb_kernels.cu for the reference lines each follows).
#pragma once
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include "wb_quant.cuh"

namespace wb {

void f32_to_f16(const float * src, __half * dst, int64_t n, cudaStream_t st);

// file-layout 32-blocks (18/22/34 B) -> planar arrays inside `dst` (same total bytes, 64 B slack); fills `out`
bool repack_block32(int wtype, const uint8_t * file_blocks_dev, uint8_t * dst, int N, int K, QMat * out, cudaStream_t st);

} // namespace wb
