// wb_debug.cu -- wb200_dbg_* entry points: drive single kernels from host buffers so the parity tests can pin
// each one against the oracle in isolation.  Not used by the product path.
#include <cstring>
#include <vector>
#include "wb_common.h"
#include "wb_gemm.cuh"
#include "wb_kernels.cuh"

using namespace wb;

extern "C" {

// C[n][m] (f32, token-major) = sum_k W[m][k] * X[n][k] ; W given in FILE layout (f16 rows or ggml quant blocks),
// X given as f32 and rounded to f16 on the way in (as the engine does).  flags bit0: apply gelu; bit1: m-major output.
__attribute__((visibility("default")))
int wb200_dbg_gemm(int wtype, int M, int N, int K, const void * w_file, const float * x, const float * bias,
                   float * out, int BN, int flags) {
    cudaStream_t st = 0;
    DevBuf<uint8_t> wraw; DevBuf<__half> xh; DevBuf<float> xf, dout, dbias;
    const size_t wbytes = (size_t) ((double) M * K * wt_bpw(wtype) + 0.5);
    if (!wraw.alloc(wbytes) || !xf.alloc((size_t) N * K) || !xh.alloc((size_t) N * K) || !dout.alloc((size_t) M * N)) return -1;
    WB_CUDA_OKV(cudaMemcpy(wraw.p, w_file, wbytes, cudaMemcpyHostToDevice), -2);
    WB_CUDA_OKV(cudaMemcpy(xf.p, x, (size_t) N * K * 4, cudaMemcpyHostToDevice), -2);
    if (bias) { if (!dbias.alloc(M)) return -1; WB_CUDA_OKV(cudaMemcpy(dbias.p, bias, (size_t) M * 4, cudaMemcpyHostToDevice), -2); }
    f32_to_f16(xf.p, xh.p, (int64_t) N * K, st);

    GemmDesc g;
    g.M = M; g.N = N; g.K = K; g.BN = BN;
    DevBuf<uint8_t> planar;
    if (wtype == WT_F16) {
        g.A.type = WT_F16; g.A.base = wraw.p;
        if (!make_tmap_f16(&g.tmA, wraw.p, K, M, 1, 1, K, 0, 0, 128)) return -3;
    } else if (wt_is_block32(wtype)) {
        if (!planar.alloc(wbytes + 64)) return -1;
        if (!repack_block32(wtype, wraw.p, planar.p, M, K, &g.A, st)) return -4;
    } else if (wt_is_kquant(wtype)) {
        g.A.type = wtype; g.A.N = M; g.A.K = K; g.A.base = wraw.p;
    } else { set_error("wb200_dbg_gemm: unsupported wtype %d", wtype); return -5; }
    if (!make_tmap_f16(&g.tmB, xh.p, K, N, 1, 1, K, 0, 0, BN)) return -3;
    g.ep.bias_m = bias ? dbias.p : nullptr;
    g.ep.act = (flags & 1) ? 1 : 0;
    g.ep.out = dout.p; g.ep.out_f16 = 0;
    g.ep.out_mmajor = (flags & 2) ? 1 : 0;
    g.ep.ldo = (flags & 2) ? N : M;
    cudaError_t e = gemm_launch(g, st);
    if (e != cudaSuccess) { set_error("gemm_launch: %s", cudaGetErrorString(e)); return -6; }
    e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { set_error("gemm sync: %s", cudaGetErrorString(e)); return -7; }
    WB_CUDA_OKV(cudaMemcpy(out, dout.p, (size_t) M * N * 4, cudaMemcpyDeviceToHost), -2);
    return 0;
}

} // extern "C"
