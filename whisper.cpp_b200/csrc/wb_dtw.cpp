// wb_dtw.cpp -- see wb_dtw.h.  Pinned on CPU by tests/test_dtw_cpu.py: fed with the alignment-head weights the reference itself computed
// (oracle tap wref_dtw_qks), dtw_path + dtw_assign must reproduce the reference's t_dtw of every token.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include "wb_common.h"
#include "wb_dtw.h"
#include "wb_state.h"

namespace wb {

// alignment heads of the released checkpoints (text layer, head), as published with openai-whisper and tabulated in src/whisper.cpp:384-395
namespace {
struct Ah { int l, h; };
const Ah k_tiny_en[]   = { {1,0},{2,0},{2,5},{3,0},{3,1},{3,2},{3,3},{3,4} };
const Ah k_tiny[]      = { {2,2},{3,0},{3,2},{3,3},{3,4},{3,5} };
const Ah k_base_en[]   = { {3,3},{4,7},{5,1},{5,5},{5,7} };
const Ah k_base[]      = { {3,1},{4,2},{4,3},{4,7},{5,1},{5,2},{5,4},{5,6} };
const Ah k_small_en[]  = { {6,6},{7,0},{7,3},{7,8},{8,2},{8,5},{8,7},{9,0},{9,4},{9,8},{9,10},{10,0},{10,1},{10,2},{10,3},{10,6},{10,11},{11,2},{11,4} };
const Ah k_small[]     = { {5,3},{5,9},{8,0},{8,4},{8,7},{8,8},{9,0},{9,7},{9,9},{10,5} };
const Ah k_medium_en[] = { {11,4},{14,1},{14,12},{14,14},{15,4},{16,0},{16,4},{16,9},{17,12},{17,14},{18,7},{18,10},{18,15},{20,0},{20,3},{20,9},{20,14},{21,12} };
const Ah k_medium[]    = { {13,15},{15,4},{15,15},{16,1},{20,0},{23,4} };
const Ah k_large_v1[]  = { {9,19},{11,2},{11,4},{11,17},{22,7},{22,11},{22,17},{23,2},{23,15} };
const Ah k_large_v2[]  = { {10,12},{13,17},{16,11},{16,12},{16,13},{17,15},{17,16},{18,4},{18,11},{18,19},{19,11},{21,2},{21,3},{22,3},{22,9},{22,12},{23,5},{23,7},{23,13},{25,5},{26,1},{26,12},{27,15} };
const Ah k_large_v3[]  = { {7,0},{10,17},{12,18},{13,12},{16,1},{17,14},{19,11},{21,4},{24,1},{25,6} };
const Ah k_large_v3t[] = { {2,4},{2,11},{3,3},{3,6},{3,11},{3,14} };
struct Preset { const Ah * p; int n; };
#define WB_PRESET(a) { a, (int) (sizeof(a) / sizeof(a[0])) }
const Preset k_presets[] = {   // indexed by whisper_alignment_heads_preset - WHISPER_AHEADS_TINY_EN (= 3)
    WB_PRESET(k_tiny_en), WB_PRESET(k_tiny), WB_PRESET(k_base_en), WB_PRESET(k_base), WB_PRESET(k_small_en), WB_PRESET(k_small),
    WB_PRESET(k_medium_en), WB_PRESET(k_medium), WB_PRESET(k_large_v1), WB_PRESET(k_large_v2), WB_PRESET(k_large_v3), WB_PRESET(k_large_v3t) };
} // namespace

bool dtw_resolve_heads(const whisper_context_params & cp, int n_text_layer, int n_head, std::vector<std::pair<int, int>> & out) {
    out.clear();
    const int preset = (int) cp.dtw_aheads_preset;
    if (preset == 0) { set_error("dtw_aheads_preset should be != DTW_AHEADS_NONE"); return false; }
    if (preset == 1) {                                                               // N top-most text layers, every head
        if (cp.dtw_n_top > n_text_layer || cp.dtw_n_top <= 0) { set_error("dtw_n_top must be between 1 and %d for this model", n_text_layer); return false; }
        for (int l = n_text_layer - cp.dtw_n_top; l < n_text_layer; ++l) for (int h = 0; h < n_head; ++h) out.emplace_back(l, h);
        return true;
    }
    std::vector<Ah> list;
    if (preset == 2) {                                                               // custom
        if (cp.dtw_aheads.n_heads == 0) { set_error("dtw_aheads.n_heads should be > 0"); return false; }
        if (!cp.dtw_aheads.heads) { set_error("dtw_aheads.heads unset"); return false; }
        for (size_t i = 0; i < cp.dtw_aheads.n_heads; ++i) list.push_back({ cp.dtw_aheads.heads[i].n_text_layer, cp.dtw_aheads.heads[i].n_head });
    } else {
        const int k = preset - 3;
        if (k < 0 || k >= (int) (sizeof(k_presets) / sizeof(k_presets[0]))) { set_error("unknown dtw_aheads_preset %d", preset); return false; }
        list.assign(k_presets[k].p, k_presets[k].p + k_presets[k].n);
    }
    for (const Ah & a : list) {
        if (a.l >= n_text_layer) { set_error("tried to set alignment head on text layer %d, but model only has %d text layers", a.l + 1, n_text_layer); return false; }
        if (a.h >= n_head)       { set_error("tried to set alignment head on head %d, but model only has %d heads", a.h + 1, n_head); return false; }
        if (a.l < 0 || a.h < 0)  { set_error("tried to set alignment head on a negative layer / head"); return false; }
    }
    for (int l = 0; l < n_text_layer; ++l) for (const Ah & a : list) if (a.l == l) out.emplace_back(a.l, a.h);    // by layer, table order inside
    if (out.empty()) { set_error("no alignment heads selected"); return false; }
    return true;
}

namespace {
// sum of squares of (x - mean) with the grouping of the reference's AVX2 build (ggml_vec_cvar_f32, ggml-cpu/vec.cpp:455-520): eight lanes are
// reduced as ((l0+l4)+(l2+l6)) + ((l1+l5)+(l3+l7)) in f32, groups and the scalar tail accumulate in double
double centered_sumsq(int n, const float * x, float mean, float * y) {
    double sum = 0.0; int i = 0;
    for (; i + 7 < n; i += 8) {
        float q[8];
        for (int k = 0; k < 8; ++k) { const float v = x[i + k] - mean; y[i + k] = v; q[k] = v * v; }
        const float a0 = q[4] + q[0], a1 = q[5] + q[1], a2 = q[6] + q[2], a3 = q[7] + q[3];
        const float b0 = a0 + a2, b1 = a1 + a3;
        sum += (double) (b0 + b1);
    }
    for (; i < n; ++i) { const float v = x[i] - mean; y[i] = v; sum += (double) (v * v); }
    return sum;
}
} // namespace

void dtw_path(const float * qk, int n_tokens, int n_audio_ctx, int n_heads, int n_audio_tokens, int sot_len, int medfilt_width,
              std::vector<int32_t> & tok_idx, std::vector<int32_t> & time_idx) {
    tok_idx.clear(); time_idx.clear();
    const int N = n_tokens - sot_len - 1, M = n_audio_tokens;                        // rows: "not" + text tokens; columns: 20 ms steps
    if (N <= 0 || M <= medfilt_width || n_heads <= 0) return;                          // (the reference asserts filter_width < n_audio_tokens; here: no stamps)
    // 1. per (head, time): normalise over the tokens (ggml_norm, eps 1e-9: double sum -> f32 mean, see centered_sumsq, scale 1/sqrt(var + eps))
    std::vector<float> w((size_t) n_heads * M * n_tokens);
    for (int h = 0; h < n_heads; ++h) for (int j = 0; j < M; ++j) {
        const float * x = qk + ((size_t) h * n_audio_ctx + j) * n_tokens;
        float * y = w.data() + ((size_t) h * M + j) * n_tokens;
        double s = 0.0; for (int i = 0; i < n_tokens; ++i) s += (double) x[i];
        const float mean = (float) s / n_tokens;
        const float var = (float) (centered_sumsq(n_tokens, x, mean, y) / n_tokens);
        const float scale = 1.0f / sqrtf(var + 1e-9f);
        for (int i = 0; i < n_tokens; ++i) y[i] *= scale;
    }
    // 2. median over time (reflect at the ends), 3. mean over heads (double sum), negated -> cost[token][time]
    const int hw = medfilt_width / 2;
    std::vector<float> cost((size_t) N * M), filt((size_t) medfilt_width), med((size_t) n_heads);
    for (int i = 0; i < N; ++i) for (int j = 0; j < M; ++j) {
        for (int h = 0; h < n_heads; ++h) {
            for (int o = -hw; o <= hw; ++o) {
                int idx = j + o;
                if (idx < 0) idx = -idx; else if (idx >= M) idx = 2 * (M - 1) - idx;
                filt[(size_t) (o + hw)] = w[((size_t) h * M + idx) * n_tokens + sot_len + i];
            }
            std::sort(filt.begin(), filt.end());
            med[(size_t) h] = filt[filt.size() / 2];
        }
        double s = 0.0; for (int h = 0; h < n_heads; ++h) s += (double) med[(size_t) h];
        cost[(size_t) i * M + j] = ((float) s / (float) n_heads) * -1.0f;
    }
    // 4. dynamic time warping: D[i][j] = cost + min(diagonal, up, left), strict comparisons pick diagonal, then up, else left
    const float INF = std::numeric_limits<float>::infinity();
    std::vector<float> D((size_t) (N + 1) * (M + 1), INF);
    std::vector<int8_t> T((size_t) (N + 1) * (M + 1), -1);
    auto at = [&](int i, int j) { return (size_t) i * (M + 1) + j; };
    D[at(0, 0)] = 0.0f;
    for (int j = 1; j <= M; ++j) for (int i = 1; i <= N; ++i) {
        const float c0 = D[at(i - 1, j - 1)], c1 = D[at(i - 1, j)], c2 = D[at(i, j - 1)];
        float c; int8_t t;
        if (c0 < c1 && c0 < c2) { c = c0; t = 0; } else if (c1 < c0 && c1 < c2) { c = c1; t = 1; } else { c = c2; t = 2; }
        D[at(i, j)] = cost[(size_t) (i - 1) * M + (j - 1)] + c;
        T[at(i, j)] = t;
    }
    for (int j = 0; j <= M; ++j) T[at(0, j)] = 2;
    for (int i = 0; i <= N; ++i) T[at(i, 0)] = 1;
    // 5. backtrace from the corner
    int i = N, j = M;
    while (i > 0 || j > 0) {
        tok_idx.push_back(i - 1); time_idx.push_back(j - 1);
        const int8_t t = T[at(i, j)];
        if (t == 0) { --i; --j; } else if (t == 1) --i; else --j;
    }
    std::reverse(tok_idx.begin(), tok_idx.end()); std::reverse(time_idx.begin(), time_idx.end());
}

void dtw_assign(const std::vector<int32_t> & tok_idx, const std::vector<int32_t> & time_idx, int seek, int token_eot,
                std::vector<Segment> & segments, int i_segment, int n_segments) {
    // every change of the token index along the path starts a new text token; index 0 is the no-timestamps token and gets nothing
    int32_t last = 0;
    int si = i_segment; size_t ti = 0;
    const int s_end = i_segment + n_segments;
    auto skip = [&]() {                                   // move to the next text token (id < eot), across segments
        while (si < s_end) {
            if (ti >= segments[(size_t) si].tokens.size()) { ++si; ti = 0; continue; }
            if (segments[(size_t) si].tokens[ti].id < token_eot) return true;
            ++ti;
        }
        return false;
    };
    for (size_t k = 0; k < tok_idx.size(); ++k) {
        if (tok_idx[k] == last) continue;
        last = tok_idx[k];
        if (!skip()) return;
        segments[(size_t) si].tokens[ti].t_dtw = (int64_t) time_idx[k] * 2 + seek;   // one path column = 20 ms
        ++ti;
    }
}

} // namespace wb

// host-only test hook (tests/test_dtw_cpu.py): alignment-head weights in, t_dtw per token out
extern "C" WB_EXPORT int wb200_dbg_dtw(const float * qk, int n_tokens, int n_audio_ctx, int n_heads, int n_frames, int sot_len, int seek, int token_eot,
                                       const int * ids, const int * seg_sizes, int n_segments, int64_t * t_dtw_out) {
    using namespace wb;
    if (!qk || !ids || !seg_sizes || !t_dtw_out || n_segments <= 0) return -1;
    std::vector<Segment> segs((size_t) n_segments);
    int k = 0;
    for (int s = 0; s < n_segments; ++s) for (int j = 0; j < seg_sizes[s]; ++j) {
        whisper_token_data td; memset(&td, 0, sizeof(td)); td.id = ids[k++]; td.t0 = td.t1 = td.t_dtw = -1;
        segs[(size_t) s].tokens.push_back(td);
    }
    std::vector<int32_t> ti, tj;
    dtw_path(qk, n_tokens, n_audio_ctx, n_heads, n_frames / 2, sot_len, 7, ti, tj);
    dtw_assign(ti, tj, seek, token_eot, segs, 0, n_segments);
    k = 0;
    for (const Segment & s : segs) for (const whisper_token_data & t : s.tokens) t_dtw_out[k++] = t.t_dtw;
    return (int) ti.size();
}

extern "C" WB_EXPORT int wb200_dbg_dtw_heads(struct whisper_context_params cp, int n_text_layer, int n_head, int * out, int cap) {
    std::vector<std::pair<int, int>> heads;
    if (!wb::dtw_resolve_heads(cp, n_text_layer, n_head, heads)) return -2;
    if ((int) heads.size() > cap) return -1;
    for (size_t i = 0; i < heads.size(); ++i) { out[2 * i] = heads[i].first; out[2 * i + 1] = heads[i].second; }
    return (int) heads.size();
}
