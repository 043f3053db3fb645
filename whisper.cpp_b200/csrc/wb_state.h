// wb_state.h -- host-side objects behind the opaque whisper_context / whisper_state handles.
// Mirrors the roles (not the code) of whisper_context / whisper_state / whisper_decoder / whisper_kv_cache in
// src/whisper.cpp:692-717, 783-820, 834-952.
#pragma once
#include <cstdint>
#include <random>
#include <string>
#include <vector>
#include "../../include/whisper_b200.h"
#include "wb_engine.h"

namespace wb {

constexpr int MAX_DECODERS = 8;               // WHISPER_MAX_DECODERS (src/whisper.cpp:142)

// Unified self-attention KV bookkeeping shared by all decoders of a state (src/whisper.cpp:692-717, 1019-1137).
// A cell holds one position of one or more sequences; beam-search reshuffles only touch this metadata.
struct KvCells {
    struct Cell { int pos = -1; uint32_t seqs = 0; };
    uint32_t head = 0, size = 0, n = 0;
    std::vector<Cell> cells;
    void reset(uint32_t n_cells) { size = n_cells; head = 0; n = 0; cells.assign(n_cells, Cell()); }
    void clear() { for (auto & c : cells) c = Cell(); head = 0; }
    bool find_slot(const int * pos, const int * seq, uint32_t n_tokens);
    int  cell_max() const;
    void seq_rm(int seq, int p0, int p1);
    void seq_cp(int src, int dst, int p0, int p1);
};

struct Sequence {
    std::vector<whisper_token_data> tokens;
    int    result_len = 0;
    double sum_logprobs_all = 0, sum_logprobs = 0, avg_logprobs = 0, entropy = 0, score = 0;
};

struct Decoder {
    Sequence sequence;
    int  i_batch = 0, seek_delta = 0;
    bool failed = false, completed = false, has_ts = false;
    std::vector<float> probs, logits, logprobs;
    std::vector<std::pair<double, int>> logits_id;
    std::mt19937 rng;
};

struct Segment {
    int64_t t0 = 0, t1 = 0;
    std::string text;
    float no_speech_prob = 0;
    std::vector<whisper_token_data> tokens;
    bool speaker_turn_next = false;
};

} // namespace wb

struct whisper_state {
    int64_t t_sample_us = 0, t_encode_us = 0, t_decode_us = 0, t_batchd_us = 0, t_prompt_us = 0, t_mel_us = 0;
    int32_t n_sample = 0, n_encode = 0, n_decode = 0, n_batchd = 0, n_prompt = 0, n_fail_p = 0, n_fail_h = 0;

    wb::Engine  eng;
    wb::KvCells kv;
    int kv_self_n_dec = 1;
    wb::Decoder decoders[wb::MAX_DECODERS];
    std::vector<float> logits;                 // [n_tokens][n_vocab] of the last decode (rows flagged want_logits are valid)
    std::vector<wb::Segment> result_all;
    std::vector<whisper_token> prompt_past0, prompt_past1;
    int   lang_id = 0;
    float no_speech_prob = 0.0f;
    int   exp_n_audio_ctx = 0;
    int   slot = 0;                            // cross-KV slot used by this state's decodes
};

struct whisper_context {
    int64_t t_load_us = 0, t_start_us = 0;
    whisper_context_params params;
    wb::Model model;
    wb::Vocab vocab;
    whisper_state * state = nullptr;
    std::string path_model;
};

namespace wb {
int64_t time_us();
// decode one batch through the engine with the reference's KV bookkeeping (whisper_decode_internal, whisper.cpp:2856-2986)
bool decode_batch(whisper_context & ctx, whisper_state & st, const int * tokens, const int * pos, const int * seq,
                  const int8_t * want, int n_tokens);
bool encode_window(whisper_context & ctx, whisper_state & st, int mel_offset);
}
