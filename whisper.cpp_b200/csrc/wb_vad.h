// wb_vad.h -- Silero voice-activity detection in front of the transcription path (whisper_vad_*, src/whisper.cpp:4367-5515,
// and the sample filter / time mapping of whisper_full, src/whisper.cpp:6669-6829, 7959-8158).
//
// The network (STFT conv -> 4 conv layers -> LSTM cell -> 1x1 conv -> sigmoid, one probability per 512-sample window)
// runs on the GPU as two kernels (wb_vad.cu): the per-window features of ALL windows in parallel, then the LSTM
// recurrence of the whole clip in one single-CTA launch with the recurrent matrix resident in registers + shared
// memory.  Turning probabilities into speech segments, cutting the PCM and mapping times back is host code.
#pragma once
#include <cstdint>
#include <string>
#include <vector>
#include <cuda_fp16.h>
#include "../../include/whisper_b200.h"

namespace wb {

// fixed architecture of the silero-16k graph (the reference hard-codes the same strides, frame count and bias shapes,
// src/whisper.cpp:4545-4591)
constexpr int VAD_WIN = 512, VAD_REFLECT = 64, VAD_NFFT = 256, VAD_HOP = 128, VAD_FRAMES = 4, VAD_BINS = 129, VAD_HID = 128;

// weights in the layout the kernels read (reduction index slowest, output index fastest -> coalesced across threads)
struct VadWeights {
    const __half * stft  = nullptr;                 // [256][258]   basis^T
    const __half * enc_w[4] = {};                   // [ic*3 + k][oc]
    const float  * enc_b[4] = {};                   // [oc]
    const float  * w_ih = nullptr, * w_hh = nullptr; // [128][512]   W^T
    const float  * b_ih = nullptr, * b_hh = nullptr; // [512]
    const __half * fin_w = nullptr;                 // [128]
    const float  * fin_b = nullptr;                 // [1]
};

struct VadModel {
    std::string type, version;
    int n_window = 0, n_context = 0;
    int n_loaded = 0;
    std::vector<uint8_t> host_blob;                 // the re-laid-out weights (host copy; uploaded once)
    void * dev_blob = nullptr;
    VadWeights hw, dw;                              // pointers into host_blob / dev_blob
};

struct VadSegment { int64_t start = 0, end = 0; };  // centiseconds

// whisper_state's record of what the VAD cut out (src/whisper.cpp:829-832, 925-934)
struct VadTimeMap { int64_t processed = 0, original = 0; };
struct VadSegmentInfo { int64_t orig_start = 0, orig_end = 0, vad_start = 0, vad_end = 0; };
struct VadCut {
    bool has_segments = false;
    std::vector<VadSegmentInfo> segments;
    std::vector<VadTimeMap>     table;
    void clear() { has_segments = false; segments.clear(); table.clear(); }
};

// probabilities -> speech segments (whisper_vad_segments_from_probs, src/whisper.cpp:5229-5463)
std::vector<VadSegment> vad_segments_from_probs(const float * probs, int n_probs, int n_window, const whisper_vad_params & p);
// cut the speech out of `samples` and record the time mapping (whisper_vad, src/whisper.cpp:6699-6827)
void vad_cut_samples(const std::vector<VadSegment> & segs, const whisper_vad_params & p, const float * samples, int n_samples,
                     std::vector<float> & filtered, VadCut & cut);
int64_t vad_map_segment_time(int64_t t, const std::vector<VadTimeMap> & table);             // src/whisper.cpp:7959-7999
int64_t vad_map_token_time(int64_t t, const std::vector<VadSegmentInfo> & segs);            // src/whisper.cpp:8094-8130

// device side (wb_vad.cu).  `state` = h[128] | c[128] on the device; probs (host) gets one value per window.
bool vad_upload(VadModel & m, int device);
void vad_free_device(VadModel & m, int device);
bool vad_forward_device(const VadModel & m, int device, float * d_state, const float * samples, int n_samples, std::vector<float> & probs);
// the same arithmetic walked thread by thread on the host: TEST HOOKS ONLY (wb200_dbg_vad_probs, and the host-only VAD context of
// the engine-less test context); no context created through whisper.h reaches it
void vad_forward_emulated(const VadModel & m, float * state, const float * samples, int n_samples, std::vector<float> & probs);

} // namespace wb

struct whisper_vad_context {
    int64_t t_vad_us = 0;
    int n_threads = 4, device = 0;
    wb::VadModel model;
    float * d_state = nullptr;                      // LSTM h | c
    std::vector<float> h_state;                     // the same for the host-only test context (device < 0)
    std::vector<float> probs;
};
struct whisper_vad_segments { std::vector<wb::VadSegment> data; };

namespace wb {
// loader shared by the API and the host-only hook: device < 0 keeps the weights on the host only
whisper_vad_context * vad_load(whisper_model_loader * loader, int device);
}
