// wb_vad.cu -- the Silero VAD network on the device (graph: whisper_vad_build_graph, src/whisper.cpp:4545-4679).
//
//   k_vad_features   one CTA per 512-sample window (grid-stride): reflect pad, STFT as a 258x256 conv at hop 128,
//                    magnitude, the four 3-tap conv layers, and the input half of the LSTM gates  ->  gi[window][512].
//                    Every window is independent, so the whole clip is one launch.
//   k_vad_lstm       the recurrence of the whole clip in ONE single-CTA launch: thread j owns gate row j of W_hh
//                    (64 coefficients in registers, 64 in shared memory), h lives in shared memory, c in registers;
//                    two block barriers per window.  The 1x1 output conv + sigmoid run at the end, one window per thread.
//
// Numerics follow the reference's CPU graph: every conv input is rounded to F16 (ggml_conv_1d goes through an F16
// im2col), weights are the file's F16/F32 values, all sums are f32.  Each kernel is written as a sequence of
// block-synchronous PHASES -- functions of (tid, nthreads) that read only what earlier phases wrote -- so that the host
// can walk the identical arithmetic thread by thread (vad_forward_emulated, a test hook) and pin it against the
// reference on a machine without a GPU.
#include <algorithm>
#include <cmath>
#include <cstring>
#include "wb_common.h"
#include "wb_vad.h"

namespace wb {

#define VAD_HD __host__ __device__ __forceinline__

constexpr int VAD_THREADS = 512;
constexpr int VAD_XS      = VAD_WIN + 2 * VAD_REFLECT;       // 640
constexpr int VAD_ROWS    = 2 * VAD_BINS;                    // 258 basis rows (real | imaginary)
constexpr int VAD_GATES   = 4 * VAD_HID;                     // 512
constexpr int VAD_WREG    = 64;                              // W_hh coefficients per row kept in registers
constexpr int VAD_BATCH   = 16384;                           // windows per launch pair (8.7 minutes of audio)

struct VadScratch {
    float xs[VAD_XS];                  // padded window, values as the F16 im2col sees them
    float st[VAD_ROWS * VAD_FRAMES];   // STFT conv output [frame][row]
    float mag[VAD_BINS * VAD_FRAMES];  // [bin][frame], F16-rounded
    float a0[128 * 4], a1[64 * 2], a2[64], a3[128];   // conv activations [channel][time]; a3 stays f32 (F32 mat-vec follows)
};

VAD_HD float r16(float v)  { return __half2float(__float2half_rn(v)); }
VAD_HD float h2f(__half v) { return __half2float(v); }
// separately rounded product / sum (the reference computes them as distinct graph nodes; no contraction into an FMA)
VAD_HD float mul_rn(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fmul_rn(a, b);
#else
    volatile float r = a * b; return r;
#endif
}
VAD_HD float add_rn(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fadd_rn(a, b);
#else
    volatile float r = a + b; return r;
#endif
}
VAD_HD float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }

// ---- feature phases ---------------------------------------------------------------------------------------------------
// window -> xs: zero fill past the clip (src/whisper.cpp:5159-5172), then ggml_pad_reflect_1d(64, 64)
VAD_HD void ph_load(int tid, int nt, VadScratch & S, const float * pcm, int64_t n_samples, int64_t base) {
    for (int i = tid; i < VAD_XS; i += nt) {
        int j = i - VAD_REFLECT;
        if (j < 0) j = -j;
        if (j >= VAD_WIN) j = 2 * (VAD_WIN - 1) - j;
        const int64_t g = base + j;
        S.xs[i] = r16(g < n_samples ? pcm[g] : 0.0f);
    }
}
VAD_HD void ph_stft(int tid, int nt, VadScratch & S, const VadWeights & W) {
    for (int o = tid; o < VAD_ROWS * VAD_FRAMES; o += nt) {
        const int r = o % VAD_ROWS, t = o / VAD_ROWS;
        const float * x = S.xs + t * VAD_HOP;
        float acc = 0.0f;
        for (int k = 0; k < VAD_NFFT; ++k) acc = fmaf(h2f(W.stft[k * VAD_ROWS + r]), x[k], acc);
        S.st[o] = acc;
    }
}
VAD_HD void ph_mag(int tid, int nt, VadScratch & S) {
    for (int o = tid; o < VAD_BINS * VAD_FRAMES; o += nt) {
        const int c = o / VAD_FRAMES, t = o % VAD_FRAMES;
        const float re = S.st[t * VAD_ROWS + c], im = S.st[t * VAD_ROWS + VAD_BINS + c];
        S.mag[o] = r16(sqrtf(add_rn(mul_rn(re, re), mul_rn(im, im))));
    }
}
// 3-tap conv, padding 1, + bias, ReLU.  in [IC][TIN], out [OC][TOUT], w [(ic*3+k)][OC]
template <int IC, int OC, int TIN, int TOUT, int STRIDE, bool ROUND>
VAD_HD void ph_conv(int tid, int nt, const float * in, float * out, const __half * w, const float * b) {
    for (int o = tid; o < OC * TOUT; o += nt) {
        const int oc = o % OC, t = o / OC;
        float acc = 0.0f;
        for (int ic = 0; ic < IC; ++ic) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int ti = t * STRIDE + k - 1;
                if (ti >= 0 && ti < TIN) acc = fmaf(h2f(w[(ic * 3 + k) * OC + oc]), in[ic * TIN + ti], acc);
            }
        }
        const float v = fmaxf(add_rn(acc, b[oc]), 0.0f);
        out[oc * TOUT + t] = ROUND ? r16(v) : v;
    }
}
// input half of the gate pre-activations: W_ih x + b_ih (src/whisper.cpp:4601-4603)
VAD_HD void ph_gates_in(int tid, int nt, const VadScratch & S, const VadWeights & W, float * gi) {
    for (int j = tid; j < VAD_GATES; j += nt) {
        float acc = 0.0f;
        for (int k = 0; k < VAD_HID; ++k) acc = fmaf(W.w_ih[k * VAD_GATES + j], S.a3[k], acc);
        gi[j] = add_rn(acc, W.b_ih[j]);
    }
}

// ---- recurrence phases ------------------------------------------------------------------------------------------------
// row j of W_hh . h, four interleaved partial sums, + b_hh, + the input half
VAD_HD float ph_gate_pre(int j, const float (&wr)[VAD_WREG], const float * wsm, const float * hs, float g_in, float b_hh) {
    float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
#pragma unroll
    for (int k = 0; k < VAD_WREG; k += 4) {
        a0 = fmaf(wr[k + 0], hs[k + 0], a0); a1 = fmaf(wr[k + 1], hs[k + 1], a1);
        a2 = fmaf(wr[k + 2], hs[k + 2], a2); a3 = fmaf(wr[k + 3], hs[k + 3], a3);
    }
#pragma unroll 4
    for (int k = 0; k < VAD_HID - VAD_WREG; k += 4) {
        a0 = fmaf(wsm[(k + 0) * VAD_GATES + j], hs[VAD_WREG + k + 0], a0); a1 = fmaf(wsm[(k + 1) * VAD_GATES + j], hs[VAD_WREG + k + 1], a1);
        a2 = fmaf(wsm[(k + 2) * VAD_GATES + j], hs[VAD_WREG + k + 2], a2); a3 = fmaf(wsm[(k + 3) * VAD_GATES + j], hs[VAD_WREG + k + 3], a3);
    }
    const float hid = add_rn(add_rn(add_rn(a0, a1), add_rn(a2, a3)), b_hh);
    return add_rn(g_in, hid);
}
// gates -> (c, h) of unit j (src/whisper.cpp:4611-4633)
VAD_HD float ph_cell(int j, const float * ps, float & c) {
    const float i = sigmoid_f(ps[j]), f = sigmoid_f(ps[VAD_HID + j]), g = tanhf(ps[2 * VAD_HID + j]), o = sigmoid_f(ps[3 * VAD_HID + j]);
    c = add_rn(mul_rn(f, c), mul_rn(i, g));
    return mul_rn(o, tanhf(c));
}
// ReLU -> 1x1 conv (F16 im2col of h) -> + bias -> sigmoid (src/whisper.cpp:4666-4669)
VAD_HD float ph_prob(const float * h, const VadWeights & W) {
    float acc = 0.0f;
    for (int k = 0; k < VAD_HID; ++k) acc = fmaf(h2f(W.fin_w[k]), r16(fmaxf(h[k], 0.0f)), acc);
    return sigmoid_f(add_rn(acc, W.fin_b[0]));
}

// ---- kernels ----------------------------------------------------------------------------------------------------------
#define VAD_FEATURE_PHASES(SYNC)                                                                                              \
    ph_load(tid, nt, S, pcm, n_samples, (int64_t) w * VAD_WIN);                                             SYNC;               \
    ph_stft(tid, nt, S, W);                                                                                 SYNC;               \
    ph_mag(tid, nt, S);                                                                                     SYNC;               \
    ph_conv<VAD_BINS, 128, 4, 4, 1, true >(tid, nt, S.mag, S.a0, W.enc_w[0], W.enc_b[0]);                   SYNC;               \
    ph_conv<128,       64, 4, 2, 2, true >(tid, nt, S.a0,  S.a1, W.enc_w[1], W.enc_b[1]);                   SYNC;               \
    ph_conv< 64,       64, 2, 1, 2, true >(tid, nt, S.a1,  S.a2, W.enc_w[2], W.enc_b[2]);                   SYNC;               \
    ph_conv< 64,      128, 1, 1, 1, false>(tid, nt, S.a2,  S.a3, W.enc_w[3], W.enc_b[3]);                   SYNC;               \
    ph_gates_in(tid, nt, S, W, gi + (int64_t) w * VAD_GATES);                                               SYNC;

__global__ void __launch_bounds__(VAD_THREADS) k_vad_features(VadWeights W, const float * __restrict__ pcm, int64_t n_samples, int n_win, float * __restrict__ gi) {
    __shared__ VadScratch S;
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int w = blockIdx.x; w < n_win; w += gridDim.x) {
        VAD_FEATURE_PHASES(__syncthreads())
    }
}

constexpr size_t VAD_LSTM_SMEM = ((size_t) (VAD_HID - VAD_WREG) * VAD_GATES + VAD_HID + VAD_GATES) * sizeof(float);

__global__ void __launch_bounds__(VAD_THREADS, 1) k_vad_lstm(VadWeights W, const float * __restrict__ gi, int n_win, float * __restrict__ state,
                                                             float * __restrict__ hbuf, float * __restrict__ probs) {
    extern __shared__ float vad_sm[];
    float * wsm = vad_sm;                                           // [64][512] rows 64..127 of W_hh^T
    float * hs  = wsm + (VAD_HID - VAD_WREG) * VAD_GATES;           // [128]
    float * ps  = hs + VAD_HID;                                     // [512]
    const int j = threadIdx.x;
    float wr[VAD_WREG];
#pragma unroll
    for (int k = 0; k < VAD_WREG; ++k) wr[k] = W.w_hh[k * VAD_GATES + j];
    for (int k = 0; k < VAD_HID - VAD_WREG; ++k) wsm[k * VAD_GATES + j] = W.w_hh[(VAD_WREG + k) * VAD_GATES + j];
    const float b_hh = W.b_hh[j];
    float c = 0.0f;
    if (j < VAD_HID) { hs[j] = state[j]; c = state[VAD_HID + j]; }
    __syncthreads();
    float g_next = n_win > 0 ? gi[j] : 0.0f;
    for (int s = 0; s < n_win; ++s) {
        const float g_in = g_next;
        if (s + 1 < n_win) g_next = gi[(int64_t) (s + 1) * VAD_GATES + j];
        ps[j] = ph_gate_pre(j, wr, wsm, hs, g_in, b_hh);
        __syncthreads();
        if (j < VAD_HID) {
            const float h = ph_cell(j, ps, c);
            hs[j] = h;
            hbuf[(int64_t) s * VAD_HID + j] = h;
        }
        __syncthreads();
    }
    if (j < VAD_HID) { state[j] = hs[j]; state[VAD_HID + j] = c; }
    __syncthreads();                                                // hbuf written by this CTA is visible to all of its threads
    for (int s = j; s < n_win; s += blockDim.x) probs[s] = ph_prob(hbuf + (int64_t) s * VAD_HID, W);
}

// ---- host side --------------------------------------------------------------------------------------------------------
bool vad_upload(VadModel & m, int device) {
    WB_CUDA_OK(cudaSetDevice(device));
    WB_CUDA_OK(cudaMalloc(&m.dev_blob, m.host_blob.size()));
    WB_CUDA_OK(cudaMemcpy(m.dev_blob, m.host_blob.data(), m.host_blob.size(), cudaMemcpyHostToDevice));
    count_h2d(m.host_blob.size());
    const uint8_t * hb = m.host_blob.data(); const uint8_t * db = (const uint8_t *) m.dev_blob;
    auto mv = [&](const void * p) -> const void * { return p ? db + ((const uint8_t *) p - hb) : nullptr; };
    m.dw.stft = (const __half *) mv(m.hw.stft);
    for (int i = 0; i < 4; ++i) { m.dw.enc_w[i] = (const __half *) mv(m.hw.enc_w[i]); m.dw.enc_b[i] = (const float *) mv(m.hw.enc_b[i]); }
    m.dw.w_ih = (const float *) mv(m.hw.w_ih); m.dw.w_hh = (const float *) mv(m.hw.w_hh);
    m.dw.b_ih = (const float *) mv(m.hw.b_ih); m.dw.b_hh = (const float *) mv(m.hw.b_hh);
    m.dw.fin_w = (const __half *) mv(m.hw.fin_w); m.dw.fin_b = (const float *) mv(m.hw.fin_b);
    return true;
}

void vad_free_device(VadModel & m, int device) {
    if (m.dev_blob) { cudaSetDevice(device); cudaFree(m.dev_blob); m.dev_blob = nullptr; }
}

bool vad_forward_device(const VadModel & m, int device, float * d_state, const float * samples, int n_samples, std::vector<float> & probs) {
    NvtxRange nvtx("wb200.vad");
    if (!m.dev_blob || !d_state) { set_error("vad: the model is not resident on a GPU"); return false; }
    WB_CUDA_OK(cudaSetDevice(device));
    const int n_win = (n_samples + VAD_WIN - 1) / VAD_WIN;
    probs.assign((size_t) n_win, 0.0f);
    if (n_win == 0) return true;
    int n_sm = 0;
    WB_CUDA_OK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device));
    WB_CUDA_OK(ensure_dyn_smem(reinterpret_cast<const void *>(k_vad_lstm), VAD_LSTM_SMEM));
    const int cap = std::min(n_win, VAD_BATCH);
    DevBuf<float> d_pcm, d_gi, d_h, d_p;
    if (!d_pcm.alloc((size_t) cap * VAD_WIN) || !d_gi.alloc((size_t) cap * VAD_GATES) || !d_h.alloc((size_t) cap * VAD_HID) || !d_p.alloc((size_t) cap)) return false;
    cudaStream_t st = nullptr;
    WB_CUDA_OK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    bool ok = true;
    for (int w0 = 0; w0 < n_win && ok; w0 += cap) {
        const int nw = std::min(cap, n_win - w0);
        const int64_t s0 = (int64_t) w0 * VAD_WIN, ns = std::min<int64_t>((int64_t) nw * VAD_WIN, (int64_t) n_samples - s0);
        cudaError_t e = cudaMemcpyAsync(d_pcm.p, samples + s0, (size_t) ns * sizeof(float), cudaMemcpyHostToDevice, st);
        count_h2d((uint64_t) ns * sizeof(float));
        if (e == cudaSuccess) {
            k_vad_features<<<std::min(nw, 4 * n_sm), VAD_THREADS, 0, st>>>(m.dw, d_pcm.p, ns, nw, d_gi.p);
            k_vad_lstm<<<1, VAD_THREADS, VAD_LSTM_SMEM, st>>>(m.dw, d_gi.p, nw, d_state, d_h.p, d_p.p);
            count_launch(2);
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) e = cudaMemcpyAsync(probs.data() + w0, d_p.p, (size_t) nw * sizeof(float), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        count_d2h((uint64_t) nw * sizeof(float));
        if (e != cudaSuccess) { set_error("vad: device pass failed: %s", cudaGetErrorString(e)); ok = false; }
    }
    cudaStreamDestroy(st);
    return ok;
}

// TEST HOOK: the phases above, executed for tid = 0..VAD_THREADS-1 one after the other with a "barrier" between phases.
void vad_forward_emulated(const VadModel & m, float * state, const float * pcm, int n_samples_i, std::vector<float> & probs) {
    const VadWeights & W = m.hw;
    const int64_t n_samples = n_samples_i;
    const int n_win = (int) ((n_samples + VAD_WIN - 1) / VAD_WIN);
    probs.assign((size_t) n_win, 0.0f);
    std::vector<float> gi_all((size_t) n_win * VAD_GATES), hbuf((size_t) n_win * VAD_HID);
    float * gi = gi_all.data();
    const int nt = VAD_THREADS;
    VadScratch * Sp = new VadScratch(); VadScratch & S = *Sp;
#define VAD_ALL_THREADS(stmt) for (int tid = 0; tid < nt; ++tid) { stmt; }
    for (int w = 0; w < n_win; ++w) {
        VAD_ALL_THREADS(ph_load(tid, nt, S, pcm, n_samples, (int64_t) w * VAD_WIN))
        VAD_ALL_THREADS(ph_stft(tid, nt, S, W))
        VAD_ALL_THREADS(ph_mag(tid, nt, S))
        VAD_ALL_THREADS((ph_conv<VAD_BINS, 128, 4, 4, 1, true >(tid, nt, S.mag, S.a0, W.enc_w[0], W.enc_b[0])))
        VAD_ALL_THREADS((ph_conv<128,       64, 4, 2, 2, true >(tid, nt, S.a0,  S.a1, W.enc_w[1], W.enc_b[1])))
        VAD_ALL_THREADS((ph_conv< 64,       64, 2, 1, 2, true >(tid, nt, S.a1,  S.a2, W.enc_w[2], W.enc_b[2])))
        VAD_ALL_THREADS((ph_conv< 64,      128, 1, 1, 1, false>(tid, nt, S.a2,  S.a3, W.enc_w[3], W.enc_b[3])))
        VAD_ALL_THREADS(ph_gates_in(tid, nt, S, W, gi + (int64_t) w * VAD_GATES))
    }
    delete Sp;
    // recurrence: per-"thread" registers live in arrays
    std::vector<float> wsm((size_t) (VAD_HID - VAD_WREG) * VAD_GATES), ps(VAD_GATES), hs(VAD_HID), c(VAD_HID);
    auto * wr = new float[VAD_GATES][VAD_WREG];
    for (int j = 0; j < VAD_GATES; ++j) {
        for (int k = 0; k < VAD_WREG; ++k) wr[j][k] = W.w_hh[k * VAD_GATES + j];
        for (int k = 0; k < VAD_HID - VAD_WREG; ++k) wsm[(size_t) k * VAD_GATES + j] = W.w_hh[(VAD_WREG + k) * VAD_GATES + j];
    }
    for (int j = 0; j < VAD_HID; ++j) { hs[j] = state[j]; c[j] = state[VAD_HID + j]; }
    for (int s = 0; s < n_win; ++s) {
        for (int j = 0; j < VAD_GATES; ++j) ps[j] = ph_gate_pre(j, wr[j], wsm.data(), hs.data(), gi_all[(size_t) s * VAD_GATES + j], W.b_hh[j]);
        for (int j = 0; j < VAD_HID; ++j) { const float h = ph_cell(j, ps.data(), c[j]); hbuf[(size_t) s * VAD_HID + j] = h; }
        for (int j = 0; j < VAD_HID; ++j) hs[j] = hbuf[(size_t) s * VAD_HID + j];
    }
    for (int j = 0; j < VAD_HID; ++j) { state[j] = hs[j]; state[VAD_HID + j] = c[j]; }
    for (int s = 0; s < n_win; ++s) probs[s] = ph_prob(hbuf.data() + (size_t) s * VAD_HID, W);
    delete[] wr;
}

} // namespace wb
