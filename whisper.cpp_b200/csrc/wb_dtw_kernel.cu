// wb_dtw_kernel.cu -- see wb_dtw_kernel.cuh.  One CTA per (alignment head, token); written as block-synchronous phases (functions of
// (tid, nthreads) reading only what earlier phases wrote) like wb_vad.cu, so that the host can walk the same arithmetic (dtw_qk_emulated).
#include <cmath>
#include <vector>
#include "wb_common.h"
#include "wb_dtw_kernel.cuh"

namespace wb {

#define DTW_HD __host__ __device__ __forceinline__
constexpr int DTW_THREADS = 256;
constexpr int DTW_MAX_CTX = 1536;

struct DtwScratch { float q16[64]; float s[DTW_MAX_CTX]; float red[DTW_THREADS]; float stat[2]; };

DTW_HD void dq_load(int tid, int nt, DtwScratch & S, const DtwQkArgs & a, int e, int t) {
    const float * q = a.q + ((size_t) a.head_layer_slot[e] * a.n_tokens + t) * a.d + (size_t) a.head_index[e] * 64;
    for (int i = tid; i < 64; i += nt) S.q16[i] = __half2float(__float2half_rn(q[i]));
}
DTW_HD void dq_scores(int tid, int nt, DtwScratch & S, const DtwQkArgs & a, int e) {
    const __half * K = a.k_cross + (size_t) a.head_layer[e] * a.layer_stride + (size_t) a.head_index[e] * 64;
    float mx = -INFINITY;
    for (int j = tid; j < a.n_audio_ctx; j += nt) {
        const __half * kr = K + (size_t) j * a.d;
        float acc = 0.0f;
        for (int i = 0; i < 64; ++i) acc = fmaf(__half2float(kr[i]), S.q16[i], acc);
        const float v = acc * a.scale;
        S.s[j] = v; mx = fmaxf(mx, v);
    }
    S.red[tid] = mx;
}
DTW_HD void dq_max(int tid, int nt, DtwScratch & S) {
    if (tid == 0) { float m = -INFINITY; for (int i = 0; i < nt; ++i) m = fmaxf(m, S.red[i]); S.stat[0] = m; }
}
DTW_HD void dq_exp(int tid, int nt, DtwScratch & S, const DtwQkArgs & a) {
    const float m = S.stat[0];
    float sum = 0.0f;
    for (int j = tid; j < a.n_audio_ctx; j += nt) { const float p = expf(S.s[j] - m); S.s[j] = p; sum += p; }
    S.red[tid] = sum;
}
DTW_HD void dq_sum(int tid, int nt, DtwScratch & S) {
    if (tid == 0) { double s = 0.0; for (int i = 0; i < nt; ++i) s += (double) S.red[i]; S.stat[1] = (float) (1.0 / s); }
}
DTW_HD void dq_store(int tid, int nt, DtwScratch & S, const DtwQkArgs & a, int e, int t) {
    const float inv = S.stat[1];
    for (int j = tid; j < a.n_audio_ctx; j += nt) a.out[((size_t) e * a.n_audio_ctx + j) * a.n_tokens + t] = S.s[j] * inv;
}

__global__ void __launch_bounds__(DTW_THREADS) k_dtw_qk(const DtwQkArgs a) {
    __shared__ DtwScratch S;
    const int e = blockIdx.x, t = blockIdx.y, tid = threadIdx.x, nt = blockDim.x;
    dq_load(tid, nt, S, a, e, t);   __syncthreads();
    dq_scores(tid, nt, S, a, e);    __syncthreads();
    dq_max(tid, nt, S);             __syncthreads();
    dq_exp(tid, nt, S, a);          __syncthreads();
    dq_sum(tid, nt, S);             __syncthreads();
    dq_store(tid, nt, S, a, e, t);
}

bool dtw_qk_launch(const DtwQkArgs & a, cudaStream_t st) {
    if (a.n_audio_ctx > DTW_MAX_CTX || a.n_heads <= 0 || a.n_tokens <= 0) { set_error("dtw: bad shape (n_audio_ctx %d, heads %d, tokens %d)", a.n_audio_ctx, a.n_heads, a.n_tokens); return false; }
    k_dtw_qk<<<dim3((unsigned) a.n_heads, (unsigned) a.n_tokens), DTW_THREADS, 0, st>>>(a);
    count_launch();
    WB_CUDA_OK(cudaGetLastError());
    return true;
}

void dtw_qk_emulated(const DtwQkArgs & a) {
    DtwScratch * Sp = new DtwScratch(); DtwScratch & S = *Sp;
    const int nt = DTW_THREADS;
    for (int e = 0; e < a.n_heads; ++e) for (int t = 0; t < a.n_tokens; ++t) {
        for (int tid = 0; tid < nt; ++tid) dq_load(tid, nt, S, a, e, t);
        for (int tid = 0; tid < nt; ++tid) dq_scores(tid, nt, S, a, e);
        for (int tid = 0; tid < nt; ++tid) dq_max(tid, nt, S);
        for (int tid = 0; tid < nt; ++tid) dq_exp(tid, nt, S, a);
        for (int tid = 0; tid < nt; ++tid) dq_sum(tid, nt, S);
        for (int tid = 0; tid < nt; ++tid) dq_store(tid, nt, S, a, e, t);
    }
    delete Sp;
}

} // namespace wb

// host-only test hook: q [n_sel][n_tokens][d] f32, k [n_layers][Tp][d] f16 bits, heads as (layer slot, layer, head) triples
extern "C" __attribute__((visibility("default"))) int wb200_dbg_dtw_qk(const float * q, const uint16_t * k, int Tp, int d, const int * triples, int n_heads,
                                                                       int n_tokens, int n_audio_ctx, float scale, float * out) {
    if (!q || !k || !triples || !out || n_audio_ctx > wb::DTW_MAX_CTX) return -1;
    std::vector<int> ls((size_t) n_heads), ly((size_t) n_heads), hd((size_t) n_heads);
    for (int e = 0; e < n_heads; ++e) { ls[e] = triples[3 * e]; ly[e] = triples[3 * e + 1]; hd[e] = triples[3 * e + 2]; }
    wb::DtwQkArgs a;
    a.q = q; a.k_cross = reinterpret_cast<const __half *>(k); a.layer_stride = (int64_t) Tp * d;
    a.head_layer_slot = ls.data(); a.head_layer = ly.data(); a.head_index = hd.data();
    a.n_heads = n_heads; a.n_tokens = n_tokens; a.n_audio_ctx = n_audio_ctx; a.d = d; a.scale = scale; a.out = out;
    wb::dtw_qk_emulated(a);
    return 0;
}
