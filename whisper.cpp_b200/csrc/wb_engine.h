// wb_engine.h -- device-side execution of the four Whisper graphs for one `whisper_state`.
//
// Replaces whisper_encode_internal (src/whisper.cpp:2366-2464: conv + encoder + cross graphs) and
// whisper_decode_internal (src/whisper.cpp:2856-2986) together with the ggml scheduler underneath them.
// A state owns: the PCM / mel buffers, the encoder workspaces for up to `cap_win` 30-s windows processed in ONE
// batched pass, the cross-attention KV of each window ("slot"), the paged self-attention KV pool and the decode
// workspaces.  All work of a state is enqueued on its own CUDA stream.
#pragma once
#include <vector>
#include "wb_model.h"
#include "wb_gemm.cuh"

namespace wb {

struct EncPlan;   // tensor maps + launch descriptors for a given (n_ctx, n_win)

struct DecToken { // one row of a decode batch (whisper_batch, src/whisper.cpp:472-480)
    int32_t token, pos, seq, slot; bool want_logits;
};

struct Engine {
    const Model * m = nullptr;
    cudaStream_t  st = nullptr;
    int cap_win = 1;                 // windows that can be encoded/decoded together
    int n_cells = 0;                 // self-KV pool size (cells); GGML_PAD(n_text_ctx,256) * factor (whisper.cpp:3402, 7167-7172)
    bool debug_taps = false;

    // ---- audio front-end
    DevBuf<float> pcm, mel, gmax;
    int n_len = 0, n_len_org = 0, n_mel = 0;
    int pcm_resident = 0;            // samples already in `pcm` (wb200_pcm_upload); pcm_to_mel(nullptr, n) then skips the H2D copy

    // ---- encoder workspaces
    int Tp_max = 0;
    DevBuf<__half> mel_win, h1, xn, qk, vt, P, attn, hfc, enc16;
    DevBuf<float>  x, S, enc32, conv32;
    DevBuf<__half> kv_cross;         // [cap_win][2][Lt][Tp][d]
    int enc_n_ctx = 0, enc_n_win = 0;   // what the last encode produced
    EncPlan * plan = nullptr;

    // ---- decoder
    DevBuf<__half> kv_k, kv_v;       // [Lt][n_cells][d]
    DevBuf<float>  dx, dqkv, dattn, dq2, dh, dlogits, xpart;
    DevBuf<int>    dints, xcnt;      // packed per-step integers: tokens | pos | cells | slot | n_kv | idx[...]
    int * hints = nullptr;           // pinned mirror of dints
    float * hlogits = nullptr;       // pinned [8][n_vocab]
    int ld_idx = 0;

    // device-event timers of the last encode (mel, conv, encoder, cross)
    cudaEvent_t ev[5] = { nullptr, nullptr, nullptr, nullptr, nullptr };
    float last_ms[4] = { 0, 0, 0, 0 };

    ~Engine();
    bool init(const Model * model, int cap_windows);
    bool set_cells(int n);           // (re)allocate the self-KV pool; contents are lost

    // PCM (host) -> mel on device.  Returns false on CUDA error.
    bool pcm_upload(const float * samples, int n_samples);
    bool pcm_to_mel(const float * samples, int n_samples);
    bool set_mel(const float * data, int n_len, int n_mel);
    bool read_mel(std::vector<float> & out);

    // encode `n_win` windows; window w starts at mel frame seeks[w] and fills cross-KV slot w.
    bool encode(const int * seeks, int n_win, int n_ctx);

    // decode a batch (<= 8 rows per pass internally).  idx lists: for row j, cells[j] is where its K/V go and
    // (kv_idx[j*ld .. +n_kv[j]]) the cells it attends to.  logits_out: host [n_rows][n_vocab], rows with want_logits filled.
    bool decode(const DecToken * rows, int n_rows, const int * cells, const int * kv_idx, int ld, const int * n_kv, float * logits_out);
};

} // namespace wb
