// wb_kernels.cu -- bandwidth-bound kernels of the encode/decode path.
#include "wb_kernels.cuh"
#include "wb_common.h"

namespace wb {

__global__ void k_f32_to_f16(const float * __restrict__ s, __half * __restrict__ d, int64_t n) {
    int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for (; i < n; i += stride) d[i] = __float2half_rn(s[i]);
}
void f32_to_f16(const float * src, __half * dst, int64_t n, cudaStream_t st) {
    if (n <= 0) return;
    const int blocks = (int) ((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
    k_f32_to_f16<<<blocks, 256, 0, st>>>(src, dst, n); count_launch();
}

// one thread per 32-block: gather the unaligned file block with byte loads, scatter to the planar arrays
template <int WT>
__global__ void k_repack32(const uint8_t * __restrict__ src, uint8_t * __restrict__ qs, uint32_t * __restrict__ qh,
                           __half * __restrict__ d, int64_t nblk) {
    constexpr int BS = (WT == WT_Q4_0) ? 18 : (WT == WT_Q5_0 ? 22 : 34);
    constexpr int QS = (WT == WT_Q8_0) ? 32 : 16;
    int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nblk) return;
    const uint8_t * b = src + i * BS;
    uint16_t dh = (uint16_t) b[0] | ((uint16_t) b[1] << 8);
    d[i] = __ushort_as_half(dh);
    int off = 2;
    if (WT == WT_Q5_0) { qh[i] = (uint32_t) b[2] | ((uint32_t) b[3] << 8) | ((uint32_t) b[4] << 16) | ((uint32_t) b[5] << 24); off = 6; }
    for (int j = 0; j < QS; ++j) qs[i * QS + j] = b[off + j];
}

bool repack_block32(int wtype, const uint8_t * src, uint8_t * dst, int N, int K, QMat * out, cudaStream_t st) {
    const int64_t nblk = (int64_t) N * (K / 32);
    const int QS = wt_qs_bytes(wtype);
    uint8_t * qs = dst;                                   // 16-byte aligned (cudaMalloc base)
    uint8_t * p  = dst + nblk * QS;
    uint32_t * qh = nullptr;
    if (wtype == WT_Q5_0) { qh = reinterpret_cast<uint32_t *>(p); p += nblk * 4; }
    __half * d = reinterpret_cast<__half *>(p);
    const int blocks = (int) ((nblk + 255) / 256);
    switch (wtype) {
        case WT_Q4_0: k_repack32<WT_Q4_0><<<blocks, 256, 0, st>>>(src, qs, qh, d, nblk); break;
        case WT_Q5_0: k_repack32<WT_Q5_0><<<blocks, 256, 0, st>>>(src, qs, qh, d, nblk); break;
        case WT_Q8_0: k_repack32<WT_Q8_0><<<blocks, 256, 0, st>>>(src, qs, qh, d, nblk); break;
        default: set_error("repack_block32: bad type %d", wtype); return false;
    }
    count_launch();
    out->type = wtype; out->N = N; out->K = K; out->base = nullptr; out->qs = qs; out->qh = qh; out->d = d;
    return cudaGetLastError() == cudaSuccess;
}

} // namespace wb
