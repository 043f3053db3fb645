// wb_dtw_kernel.cuh -- cross-attention WEIGHTS of the alignment heads for DTW token timestamps: softmax(f16(q) . K^T * 64^-1/4) over the
// audio positions, per (alignment head, text token).  This is what the reference's non-flash decoder graph keeps as aheads_cross_QKs
// (src/whisper.cpp:2722-2741: KQ = mul_mat(Kcross F16, Q -> F16), soft_max_ext(KQ, scale = KQscale)).
#pragma once
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

namespace wb {

struct DtwQkArgs {
    const float  * q;            // captured cross-attention queries [n_sel_layers][n_tokens][d]  (before the 64^-1/4 scale)
    const __half * k_cross;      // cross K of this state's slot: layer l at k_cross + l * layer_stride, [Tp][d], already scaled by 64^-1/4
    int64_t layer_stride;        // elements between layers
    const int * head_layer_slot; // per alignment head: index into the captured layers
    const int * head_layer;      // per alignment head: text layer
    const int * head_index;      // per alignment head: head inside the layer
    int n_heads, n_tokens, n_audio_ctx, d;
    float scale;                 // 64^-1/4
    float * out;                 // [n_heads][n_audio_ctx][n_tokens]  (tokens fastest: the layout of the reference's host copy)
};

bool dtw_qk_launch(const DtwQkArgs & a, cudaStream_t st);
// the same phases walked on the host (test hook; pointers are host pointers)
void dtw_qk_emulated(const DtwQkArgs & a);

} // namespace wb
