// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Builds the reference's own src/whisper.cpp (textually included from /root/reference at
// compile time, nothing is copied into this repository) into oracle/_ref/libwhisper_ref.so and
// adds a few `wref_*` taps so the parity tests can read the reference's intermediates:
//
//   mel spectrogram            whisper_state::mel            (src/whisper.cpp:414-420, 3178-3272)
//   conv stem output           whisper_state::embd_conv      (src/whisper.cpp:1982-2042)
//   encoder output             whisper_state::embd_enc       (src/whisper.cpp:2044-2275)
//   cross / self KV caches     whisper_state::kv_cross/self  (src/whisper.cpp:2278-2354, 2567-2599)
//   logits filter + sampler    whisper_process_logits / whisper_sample_token (src/whisper.cpp:6196-6543)
//   block (de)quantisers       ggml_quantize_chunk / type traits (ggml/src/ggml-quants.c)
//   VAD segments / PCM cut     whisper_vad_segments_from_probs, whisper_vad (src/whisper.cpp:5229-5463, 6669-6829)
//
// The whisper.h API of the reference is exported unchanged by the same library.

#ifndef REF_WHISPER_CPP
#error "REF_WHISPER_CPP must point at <reference>/src/whisper.cpp"
#endif

#include REF_WHISPER_CPP

#include <cstdint>
#include <cstring>

extern "C" {

#define WREF_API __attribute__((visibility("default")))

// ---- mel -------------------------------------------------------------------------------------
WREF_API int wref_mel_n_len(struct whisper_state * st)     { return st->mel.n_len; }
WREF_API int wref_mel_n_len_org(struct whisper_state * st) { return st->mel.n_len_org; }
WREF_API int wref_mel_n_mel(struct whisper_state * st)     { return st->mel.n_mel; }
WREF_API int wref_mel_copy(struct whisper_state * st, float * out, int64_t cap) {
    const int64_t n = (int64_t) st->mel.data.size();
    if (n > cap) return -1;
    memcpy(out, st->mel.data.data(), n*sizeof(float));
    return 0;
}

WREF_API struct whisper_state * wref_ctx_state(struct whisper_context * ctx) { return ctx->state; }

// ---- tensors living in the per-graph compute buffers -------------------------------------------
static int64_t tensor_copy_f32(struct ggml_tensor * t, float * out, int64_t cap) {
    if (!t) return -1;
    const int64_t n = ggml_nelements(t);
    if (!out) return n;
    if (n > cap || t->type != GGML_TYPE_F32) return -2;
    ggml_backend_tensor_get(t, out, 0, n*sizeof(float));
    return n;
}

// embd_conv: ne = [n_ctx, n_state] (time fastest)           src/whisper.cpp:2012-2024
WREF_API int64_t wref_embd_conv(struct whisper_state * st, float * out, int64_t cap) { return tensor_copy_f32(st->embd_conv, out, cap); }
// embd_enc : ne = [n_state, n_ctx] (feature fastest)        src/whisper.cpp:2245-2259
WREF_API int64_t wref_embd_enc (struct whisper_state * st, float * out, int64_t cap) { return tensor_copy_f32(st->embd_enc,  out, cap); }

// raw F16 KV caches; returns number of f16 elements
static int64_t kv_copy(struct ggml_tensor * t, uint16_t * out, int64_t cap) {
    const int64_t n = ggml_nelements(t);
    if (!out) return n;
    if (n > cap || t->type != GGML_TYPE_F16) return -2;
    ggml_backend_tensor_get(t, out, 0, n*sizeof(uint16_t));
    return n;
}
WREF_API int64_t wref_kv_cross_k(struct whisper_state * st, uint16_t * out, int64_t cap) { return kv_copy(st->kv_cross.k, out, cap); }
WREF_API int64_t wref_kv_cross_v(struct whisper_state * st, uint16_t * out, int64_t cap) { return kv_copy(st->kv_cross.v, out, cap); }
WREF_API int64_t wref_kv_self_k (struct whisper_state * st, uint16_t * out, int64_t cap) { return kv_copy(st->kv_self.k,  out, cap); }
WREF_API int64_t wref_kv_self_v (struct whisper_state * st, uint16_t * out, int64_t cap) { return kv_copy(st->kv_self.v,  out, cap); }
WREF_API int     wref_kv_self_size(struct whisper_state * st) { return (int) st->kv_self.size; }

// ---- logits filter + greedy sampler on injected logits ----------------------------------------
// history: token ids already in decoder.sequence.tokens; has_ts/seek_delta as whisper_decoder fields.
// logits_in: raw [n_vocab] logits (as state.logits row).  Outputs are the three per-decoder arrays
// after whisper_process_logits, and the greedy whisper_sample_token result.
WREF_API int wref_process_logits(
        struct whisper_context * ctx, struct whisper_state * st, const struct whisper_full_params * params,
        const whisper_token * history, int n_history, int has_ts, int seek_delta, float temperature,
        const float * logits_in, float * logits_out, float * logprobs_out, float * probs_out,
        whisper_token_data * sampled) {
    const int n_vocab = ctx->vocab.n_vocab;
    auto & dec = st->decoders[0];
    dec.sequence.tokens.clear();
    for (int i = 0; i < n_history; ++i) {
        whisper_token_data td = { history[i], 0, 0.0f, 0.0f, 0.0f, 0.0f, -1, -1, -1, 0.0f };
        dec.sequence.tokens.push_back(td);
    }
    dec.has_ts     = has_ts != 0;
    dec.seek_delta = seek_delta;
    dec.i_batch    = 0;
    dec.grammar    = {};
    st->logits.assign(logits_in, logits_in + n_vocab);
    whisper_process_logits(*ctx, *st, dec, *params, temperature);
    if (logits_out)   memcpy(logits_out,   dec.logits.data(),   n_vocab*sizeof(float));
    if (logprobs_out) memcpy(logprobs_out, dec.logprobs.data(), n_vocab*sizeof(float));
    if (probs_out)    memcpy(probs_out,    dec.probs.data(),    n_vocab*sizeof(float));
    if (sampled)      *sampled = whisper_sample_token(*ctx, dec, true);
    return 0;
}

// the same with a grammar (params->grammar_rules): the decoder's parse state is initialised from the rules and advanced over
// `accepted` (whisper_grammar_accept_token, src/whisper.cpp:5901-5923) before the logits are filtered
WREF_API int wref_process_logits_grammar(
        struct whisper_context * ctx, struct whisper_state * st, const struct whisper_full_params * params,
        const whisper_token * history, int n_history, int has_ts, int seek_delta, float temperature,
        const whisper_token * accepted, int n_accepted,
        const float * logits_in, float * logits_out, float * logprobs_out, float * probs_out, whisper_token_data * sampled, int * n_stacks_out) {
    const int n_vocab = ctx->vocab.n_vocab;
    auto & dec = st->decoders[0];
    dec.sequence.tokens.clear();
    for (int i = 0; i < n_history; ++i) {
        whisper_token_data td = { history[i], 0, 0.0f, 0.0f, 0.0f, 0.0f, -1, -1, -1, 0.0f };
        dec.sequence.tokens.push_back(td);
    }
    dec.has_ts = has_ts != 0; dec.seek_delta = seek_delta; dec.i_batch = 0;
    dec.grammar = whisper_grammar_init(params->grammar_rules, params->n_grammar_rules, params->i_start_rule);
    for (int i = 0; i < n_accepted; ++i) whisper_grammar_accept_token(*ctx, dec.grammar, accepted[i]);
    if (n_stacks_out) *n_stacks_out = (int) dec.grammar.stacks.size();
    st->logits.assign(logits_in, logits_in + n_vocab);
    whisper_process_logits(*ctx, *st, dec, *params, temperature);
    if (logits_out)   memcpy(logits_out,   dec.logits.data(),   n_vocab*sizeof(float));
    if (logprobs_out) memcpy(logprobs_out, dec.logprobs.data(), n_vocab*sizeof(float));
    if (probs_out)    memcpy(probs_out,    dec.probs.data(),    n_vocab*sizeof(float));
    if (sampled)      *sampled = whisper_sample_token(*ctx, dec, true);
    dec.grammar = {};
    return 0;
}

// beam-search candidates: k draws of whisper_sample_token_topk (src/whisper.cpp:6545-6618) from the distribution left in
// decoders[0] by the last wref_process_logits call, with decoder.rng = std::mt19937(seed) as whisper_full seeds it (7199)
WREF_API int wref_sample_topk(struct whisper_context * ctx, struct whisper_state * st, int k, int seed, whisper_token_data * out) {
    auto & dec = st->decoders[0];
    dec.rng = std::mt19937(seed);
    const auto r = whisper_sample_token_topk(*ctx, dec, k);
    for (int i = 0; i < k; ++i) out[i] = r[i];
    return 0;
}

// ---- experimental token-level timestamps + max_len wrapping on an injected segment (src/whisper.cpp:8640-8820, 6096-6147) ----
// carry[3] = { t_beg, t_last, tid_last } in/out.  tok_out: per token of the ORIGINAL segment { t0, t1 } and vlen; piece_out: per
// resulting segment { t0, t1, n_tokens }.  Returns the number of segments the injected one was wrapped into.
WREF_API int wref_token_timestamps(struct whisper_context * ctx, struct whisper_state * st, const int * ids, const int * tids, const float * pt,
                                   const float * ptsum, int n, int64_t seg_t0, int64_t seg_t1, const float * energy, int n_energy,
                                   float thold_pt, float thold_ptsum, int max_len, int split_on_word, int64_t * carry,
                                   int64_t * tok_out, float * vlen_out, int64_t * piece_out, int max_pieces) {
    st->result_all.clear();
    st->energy.assign(energy, energy + n_energy);
    st->t_beg = carry[0]; st->t_last = carry[1]; st->tid_last = (whisper_token) carry[2];
    whisper_segment seg = { seg_t0, seg_t1, "", 0.0f, {}, false };
    for (int i = 0; i < n; ++i) {
        whisper_token_data td = { ids[i], tids[i], 0.0f, 0.0f, pt[i], ptsum[i], -1, -1, -1, 0.0f };
        seg.tokens.push_back(td);
    }
    st->result_all.push_back(seg);
    whisper_exp_compute_token_level_timestamps(*ctx, *st, 0, thold_pt, thold_ptsum);
    for (int i = 0; i < n; ++i) { tok_out[2 * i] = st->result_all[0].tokens[i].t0; tok_out[2 * i + 1] = st->result_all[0].tokens[i].t1; vlen_out[i] = st->result_all[0].tokens[i].vlen; }
    int pieces = 1;
    if (max_len > 0) pieces = whisper_wrap_segment(*ctx, *st, max_len, split_on_word != 0);
    for (int i = 0; i < (int) st->result_all.size() && i < max_pieces; ++i) {
        piece_out[3 * i] = st->result_all[i].t0; piece_out[3 * i + 1] = st->result_all[i].t1; piece_out[3 * i + 2] = (int64_t) st->result_all[i].tokens.size();
    }
    carry[0] = st->t_beg; carry[1] = st->t_last; carry[2] = st->tid_last;
    return pieces;
}
WREF_API int wref_signal_energy(const float * pcm, int n, int hw, float * out) {
    const auto e = get_signal_energy(pcm, n, hw);
    memcpy(out, e.data(), (size_t) n * sizeof(float));
    return 0;
}

// ---- self-attention KV bookkeeping (src/whisper.cpp:1019-1137) on a stand-alone cell table ------------------------------
// ops: n_ops x 5 ints { op, a, b, c, d }: 0 find_slot(n_tokens=a, first pos=b, seq=c)  1 seq_rm(seq=a, p0=b, p1=c)
// 2 seq_cp(src=a, dst=b, p0=c, p1=d)  3 clear.  trace: 3 ints per op { return, head, cell_max }; cells_out: 2 ints per cell
// { pos, bit mask of seq ids }.
WREF_API int wref_kv_script(int size, const int * ops, int n_ops, int * trace, int * cells_out) {
    whisper_kv_cache cache;
    cache.size = size; cache.head = 0; cache.n = 0; cache.k = nullptr; cache.v = nullptr;
    cache.cells.clear(); cache.cells.resize(size);
    for (int o = 0; o < n_ops; ++o) {
        const int * q = ops + 5 * o;
        int ret = 1;
        if (q[0] == 0) {
            whisper_batch b = whisper_batch_init(q[1], 1);
            b.n_tokens = q[1];
            for (int i = 0; i < q[1]; ++i) { b.token[i] = 0; b.pos[i] = q[2] + i; b.n_seq_id[i] = 1; b.seq_id[i][0] = q[3]; b.logits[i] = 0; }
            ret = whisper_kv_cache_find_slot(cache, b) ? 1 : 0;
            whisper_batch_free(b);
        } else if (q[0] == 1) whisper_kv_cache_seq_rm(cache, q[1], q[2], q[3]);
        else if (q[0] == 2)   whisper_kv_cache_seq_cp(cache, q[1], q[2], q[3], q[4]);
        else { for (auto & c : cache.cells) { c.pos = -1; c.seq_id.clear(); } cache.head = 0; }      // whisper_kv_cache_clear without the buffer
        trace[3 * o] = ret; trace[3 * o + 1] = (int) cache.head; trace[3 * o + 2] = whisper_kv_cache_cell_max(cache);
    }
    for (int i = 0; i < size; ++i) {
        int m = 0; for (int sid : cache.cells[i].seq_id) m |= 1 << sid;
        cells_out[2 * i] = cache.cells[i].pos; cells_out[2 * i + 1] = m;
    }
    return 0;
}

// the batch of the most recent whisper_decode_internal of this state (tokens, positions, sequence ids, logits flags): read from inside a
// logits_filter_callback it shows exactly what was fed to the decoder for the step being filtered (tests/test_full_scripted_cpu.py)
WREF_API int wref_last_batch(struct whisper_state * st, int * tok, int * pos, int * seq, int8_t * want, int cap) {
    const whisper_batch & b = st->batch;
    if (b.n_tokens > cap) return -1;
    for (int i = 0; i < b.n_tokens; ++i) { tok[i] = b.token[i]; pos[i] = b.pos[i]; seq[i] = b.seq_id[i][0]; want[i] = b.logits[i]; }
    return b.n_tokens;
}

// per token of the last decoded batch: FNV-1a hash of the sorted positions of the self-attention KV cells it attended to, read back from the
// KQ mask the reference built for that decode (src/whisper.cpp:2917-2947) -- the observable result of all KV bookkeeping before it
WREF_API int wref_last_attended(struct whisper_state * st, uint64_t * out, int cap) {
    const whisper_batch & b = st->batch;
    const int n_kv = (int) st->kv_self.n;
    if (b.n_tokens > cap || (int64_t) st->inp_mask.size() < (int64_t) n_kv * b.n_tokens) return -1;
    std::vector<int> pos;
    for (int j = 0; j < b.n_tokens; ++j) {
        pos.clear();
        for (int i = 0; i < n_kv; ++i) if (st->inp_mask[(size_t) j * n_kv + i] == 0.0f) pos.push_back(st->kv_self.cells[i].pos);
        std::sort(pos.begin(), pos.end());
        uint64_t h = 1469598103934665603ull;
        for (int v : pos) for (int k = 0; k < 4; ++k) { h ^= (uint64_t) ((v >> (8 * k)) & 0xff); h *= 1099511628211ull; }
        out[j] = h;
    }
    return b.n_tokens;
}

// ---- DTW token timestamps (src/whisper.cpp:8880-9167) ------------------------------------------------------------------------------
// the alignment-head cross-attention weights the LAST whisper_exp_compute_token_level_timestamps_dtw call copied to the host
// (state->aheads_cross_QKs_data: [n_heads][n_audio_ctx][n_tokens], src/whisper.cpp:9075-9079)
WREF_API int64_t wref_dtw_qks(struct whisper_state * st, float * out, int64_t cap, int * n_tokens, int * n_audio_ctx, int * n_heads) {
    if (!st->aheads_cross_QKs) return -1;
    *n_tokens = (int) st->aheads_cross_QKs->ne[0]; *n_audio_ctx = (int) st->aheads_cross_QKs->ne[1]; *n_heads = (int) st->aheads_cross_QKs->ne[2];
    const int64_t n = (int64_t) st->aheads_cross_QKs_data.size();
    if (!out) return n;
    if (n > cap) return -2;
    memcpy(out, st->aheads_cross_QKs_data.data(), (size_t) n*sizeof(float));
    return n;
}

// (layer, head) pairs in the order aheads_cross_QKs concatenates them (get_alignment_heads_by_layer, src/whisper.cpp:8856-8875)
WREF_API int wref_dtw_heads(struct whisper_context_params cp, int n_text_layer, int n_head, int * out, int cap) {
    int n = 0;
    for (int il = 0; il < n_text_layer; ++il)
        for (uint32_t h : get_alignment_heads_by_layer(cp, il, n_text_layer, n_head)) { if (n >= cap) return -1; out[2*n] = il; out[2*n + 1] = (int) h; ++n; }
    return n;
}

// ---- voice-activity detection (src/whisper.cpp:4367-5515, 6669-6829, 7959-8130) -------------------------------------
// probabilities -> segments with the reference's own whisper_vad_segments_from_probs (it reads only n_window and probs)
WREF_API int wref_vad_segments(const float * probs, int n_probs, struct whisper_vad_params params, int64_t * t0, int64_t * t1, int cap) {
    whisper_vad_context v;
    v.n_window = 512;
    v.probs.assign(probs, probs + n_probs);
    whisper_vad_segments * s = whisper_vad_segments_from_probs(&v, params);
    if (!s) return -1;
    const int n = (int) s->data.size();
    if (n > cap) { whisper_vad_free_segments(s); return -2; }
    for (int i = 0; i < n; ++i) { t0[i] = s->data[i].start; t1[i] = s->data[i].end; }
    whisper_vad_free_segments(s);
    return n;
}
// the static whisper_vad() of whisper_full (params.vad_model_path must name a VAD model): filtered PCM, mapping table,
// per-segment info, and the two time mappings evaluated at the query times q
WREF_API int wref_vad_cut(struct whisper_context * ctx, struct whisper_full_params params, const float * samples, int n_samples,
                          float * filtered, int cap, int64_t * table, int * n_table, int64_t * info, int * n_info,
                          const int64_t * q, int n_q, int64_t * q_seg, int64_t * q_tok) {
    std::vector<float> out;
    whisper_state * st = ctx->state;
    if (!whisper_vad(ctx, st, params, samples, n_samples, out)) return -1;
    if ((int) out.size() > cap) return -2;
    memcpy(filtered, out.data(), out.size()*sizeof(float));
    *n_table = (int) st->vad_mapping_table.size(); *n_info = (int) st->vad_segments.size();
    for (size_t i = 0; i < st->vad_mapping_table.size(); ++i) { table[2*i] = st->vad_mapping_table[i].processed_time; table[2*i + 1] = st->vad_mapping_table[i].original_time; }
    for (size_t i = 0; i < st->vad_segments.size(); ++i) {
        info[4*i] = st->vad_segments[i].orig_start; info[4*i + 1] = st->vad_segments[i].orig_end;
        info[4*i + 2] = st->vad_segments[i].vad_start; info[4*i + 3] = st->vad_segments[i].vad_end;
    }
    for (int i = 0; i < n_q; ++i) {
        q_seg[i] = map_processed_to_original_time(q[i], st->vad_mapping_table);
        q_tok[i] = whisper_map_token_time_segment_aware(q[i], st->vad_segments);
    }
    return (int) out.size();
}

// ---- block quantisers (ggml/src/ggml-quants.c) ------------------------------------------------
WREF_API int64_t wref_row_size(int type, int64_t n_per_row) { return (int64_t) ggml_row_size((ggml_type) type, n_per_row); }
WREF_API int64_t wref_quantize(int type, const float * src, void * dst, int64_t nrows, int64_t n_per_row) {
    return (int64_t) ggml_quantize_chunk((ggml_type) type, src, dst, 0, nrows, n_per_row, nullptr);
}
WREF_API int wref_dequantize(int type, const void * src, float * dst, int64_t n) {
    const auto * tr = ggml_get_type_traits((ggml_type) type);
    if (!tr || !tr->to_float) return -1;
    tr->to_float(src, dst, n);
    return 0;
}
// GELU exactly as the CPU backend evaluates it (F16 table, ggml/src/ggml-cpu/vec.h:988-1001)
WREF_API float wref_fp16_round(float x) { return ggml_fp16_to_fp32(ggml_fp32_to_fp16(x)); }

WREF_API int wref_sizeof_full_params(void)    { return (int) sizeof(struct whisper_full_params); }
WREF_API int wref_sizeof_context_params(void) { return (int) sizeof(struct whisper_context_params); }
WREF_API int wref_sizeof_token_data(void)     { return (int) sizeof(whisper_token_data); }

} // extern "C"
