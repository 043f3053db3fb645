// wb_vad.cpp -- host side of the VAD: model file, probabilities -> speech segments, PCM cut + time mapping, whisper_vad_* API.
// Reference behaviour: src/whisper.cpp:4787-5116 (file format), 5229-5463 (segments), 6669-6829 (cut), 7959-8158 (mapping).
#include <algorithm>
#include <climits>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <string>
#include "wb_common.h"
#include "wb_vad.h"

namespace wb {

int64_t time_us();

namespace {

constexpr int SAMPLE_RATE = 16000;

int     cs_to_samples(int64_t cs)   { return (int) ((cs / 100.0) * SAMPLE_RATE + 0.5); }        // src/whisper.cpp:4477-4483
int64_t samples_to_cs(int samples)  { return (int64_t) ((samples / (double) SAMPLE_RATE) * 100.0 + 0.5); }

struct Rd {
    whisper_model_loader * l; bool ok = true;
    bool bytes(void * dst, size_t n) { if (n && l->read(l->context, dst, n) != n) ok = false; return ok; }
    template <typename T> bool pod(T & v) { return bytes(&v, sizeof(T)); }
};

struct RawTensor { int type = 0; int ne[3] = {1, 1, 1}; std::vector<uint8_t> data; bool seen = false; };

} // namespace

// ---- model file ---------------------------------------------------------------------------------------------------------
whisper_vad_context * vad_load(whisper_model_loader * loader, int device) {
    Rd rd{loader};
    uint32_t magic = 0;
    if (!rd.pod(magic) || magic != 0x67676d6c) { set_error("vad: invalid model data (bad magic)"); logf(LOG_ERROR, "%s: invalid model data (bad magic)\n", __func__); return nullptr; }
    std::unique_ptr<whisper_vad_context> vctx(new whisper_vad_context());
    VadModel & m = vctx->model;
    int32_t len = 0;
    if (!rd.pod(len) || len < 0 || len > 256) { set_error("vad: bad model type string"); return nullptr; }
    m.type.resize((size_t) len); rd.bytes(&m.type[0], (size_t) len);
    int32_t ver[3] = {0, 0, 0};
    rd.pod(ver[0]); rd.pod(ver[1]); rd.pod(ver[2]);
    m.version = std::to_string(ver[0]) + "." + std::to_string(ver[1]) + "." + std::to_string(ver[2]);
    int32_t n_window = 0, n_context = 0, n_layers = 0;
    rd.pod(n_window); rd.pod(n_context); rd.pod(n_layers);
    if (!rd.ok || n_layers < 0 || n_layers > 16) { set_error("vad: truncated header"); return nullptr; }
    std::vector<int32_t> cin((size_t) n_layers), cout((size_t) n_layers), ksz((size_t) n_layers);
    for (int i = 0; i < n_layers; ++i) { rd.pod(cin[i]); rd.pod(cout[i]); rd.pod(ksz[i]); }
    int32_t lstm_in = 0, lstm_hid = 0, fin_in = 0, fin_out = 0;
    rd.pod(lstm_in); rd.pod(lstm_hid); rd.pod(fin_in); rd.pod(fin_out);
    if (!rd.ok) { set_error("vad: truncated header"); return nullptr; }
    m.n_window = n_window; m.n_context = n_context;
    logf(LOG_INFO, "%s: model type: %s, version: %s, n_window = %d, encoder layers = %d, lstm %d -> %d\n", __func__, m.type.c_str(), m.version.c_str(),
         n_window, n_layers, lstm_in, lstm_hid);
    // The reference's graph is only meaningful for this one shape (strides, the 4-frame view and the bias reshapes are literals there).
    static const int want_in[4] = {VAD_BINS, 128, 64, 64}, want_out[4] = {128, 64, 64, 128};
    bool shape_ok = n_window == VAD_WIN && n_layers == 4 && lstm_in == VAD_HID && lstm_hid == VAD_HID && fin_in == VAD_HID && fin_out == 1;
    for (int i = 0; shape_ok && i < 4; ++i) shape_ok = cin[i] == want_in[i] && cout[i] == want_out[i] && ksz[i] == 3;
    if (!shape_ok) { set_error("vad: unsupported architecture (expected silero-16k: window 512, convs 129>128>64>64>128 k=3, LSTM 128)"); logf(LOG_ERROR, "%s: %s\n", __func__, last_error()); return nullptr; }

    // expected tensors: name -> (type, ne)
    std::map<std::string, RawTensor> want;
    auto expect = [&](const std::string & name, int type, int n0, int n1, int n2) { RawTensor t; t.type = type; t.ne[0] = n0; t.ne[1] = n1; t.ne[2] = n2; want[name] = std::move(t); };
    expect("_model.stft.forward_basis_buffer", 1, VAD_NFFT, 1, 2 * VAD_BINS);
    for (int i = 0; i < 4; ++i) {
        expect("_model.encoder." + std::to_string(i) + ".reparam_conv.weight", 1, 3, want_in[i], want_out[i]);
        expect("_model.encoder." + std::to_string(i) + ".reparam_conv.bias",   0, want_out[i], 1, 1);
    }
    expect("_model.decoder.rnn.weight_ih", 0, VAD_HID, 4 * VAD_HID, 1);
    expect("_model.decoder.rnn.weight_hh", 0, VAD_HID, 4 * VAD_HID, 1);
    expect("_model.decoder.rnn.bias_ih",   0, 4 * VAD_HID, 1, 1);
    expect("_model.decoder.rnn.bias_hh",   0, 4 * VAD_HID, 1, 1);
    expect("_model.decoder.decoder.2.weight", 1, VAD_HID, 1, 1);
    expect("_model.decoder.decoder.2.bias",   0, 1, 1, 1);

    size_t total = 0;
    while (true) {
        int32_t n_dims = 0, name_len = 0, ttype = 0;
        rd.ok = true;
        rd.pod(n_dims); rd.pod(name_len); rd.pod(ttype);
        if (!rd.ok || loader->eof(loader->context)) break;
        if (n_dims < 0 || n_dims > 4) { set_error("vad: invalid n_dims %d in model file", n_dims); return nullptr; }
        int32_t ne[4] = {1, 1, 1, 1}; int64_t nelem = 1;
        for (int i = 0; i < n_dims; ++i) { rd.pod(ne[i]); nelem *= ne[i]; }
        if (!rd.ok || name_len < 0 || name_len > 512) { set_error("vad: corrupt tensor header"); return nullptr; }
        std::string name((size_t) name_len, '\0');
        rd.bytes(&name[0], (size_t) name_len);
        auto it = want.find(name);
        if (!rd.ok || it == want.end()) { set_error("vad: unknown tensor '%s' in model file", name.c_str()); logf(LOG_ERROR, "%s: %s\n", __func__, last_error()); return nullptr; }
        RawTensor & t = it->second;
        if (ne[0] != t.ne[0] || ne[1] != t.ne[1] || ne[2] != t.ne[2] || ne[3] != 1) {
            set_error("vad: tensor '%s' has wrong shape in model file: got [%d, %d, %d], expected [%d, %d, %d]", name.c_str(), ne[0], ne[1], ne[2], t.ne[0], t.ne[1], t.ne[2]);
            logf(LOG_ERROR, "%s: %s\n", __func__, last_error()); return nullptr;
        }
        if ((ttype != 0 && ttype != 1) || ttype != t.type) { set_error("vad: tensor '%s' has wrong size in model file (type %d, expected %d)", name.c_str(), ttype, t.type); logf(LOG_ERROR, "%s: %s\n", __func__, last_error()); return nullptr; }
        t.data.resize((size_t) nelem * (t.type == 1 ? 2 : 4));
        if (!rd.bytes(t.data.data(), t.data.size())) { set_error("vad: tensor '%s' is truncated", name.c_str()); return nullptr; }
        t.seen = true; total += t.data.size(); ++m.n_loaded;
    }
    logf(LOG_INFO, "%s: model size    = %7.2f MB\n", __func__, total / 1e6);
    if (m.n_loaded == 0) {
        logf(LOG_WARN, "%s: WARN no tensors loaded from model file - assuming empty model for testing\n", __func__);
    } else if (m.n_loaded != (int) want.size()) {
        set_error("vad: not all tensors loaded from model file - expected %zu, got %d", want.size(), m.n_loaded);
        logf(LOG_ERROR, "%s: ERROR %s\n", __func__, last_error()); return nullptr;
    }

    // re-lay the weights out for the kernels: reduction index slowest, output index fastest
    if (m.n_loaded > 0) {
        size_t off = 0;
        auto reserve = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t) 255; return o; };
        const size_t o_stft = reserve((size_t) VAD_NFFT * 2 * VAD_BINS * 2);
        size_t o_w[4], o_b[4];
        for (int i = 0; i < 4; ++i) { o_w[i] = reserve((size_t) want_in[i] * 3 * want_out[i] * 2); o_b[i] = reserve((size_t) want_out[i] * 4); }
        const size_t o_wih = reserve((size_t) VAD_HID * 4 * VAD_HID * 4), o_whh = reserve((size_t) VAD_HID * 4 * VAD_HID * 4);
        const size_t o_bih = reserve((size_t) 4 * VAD_HID * 4), o_bhh = reserve((size_t) 4 * VAD_HID * 4);
        const size_t o_fw = reserve((size_t) VAD_HID * 2), o_fb = reserve(4);
        m.host_blob.assign(off, 0);
        uint8_t * B = m.host_blob.data();
        {   // basis [row][1][k] -> [k][row]
            const uint16_t * src = (const uint16_t *) want["_model.stft.forward_basis_buffer"].data.data(); uint16_t * dst = (uint16_t *) (B + o_stft);
            for (int r = 0; r < 2 * VAD_BINS; ++r) for (int k = 0; k < VAD_NFFT; ++k) dst[(size_t) k * 2 * VAD_BINS + r] = src[(size_t) r * VAD_NFFT + k];
        }
        for (int i = 0; i < 4; ++i) {   // [oc][ic][k] -> [ic*3+k][oc]
            const std::string base = "_model.encoder." + std::to_string(i) + ".reparam_conv.";
            const uint16_t * src = (const uint16_t *) want[base + "weight"].data.data(); uint16_t * dst = (uint16_t *) (B + o_w[i]);
            const int IC = want_in[i], OC = want_out[i];
            for (int oc = 0; oc < OC; ++oc) for (int j = 0; j < IC * 3; ++j) dst[(size_t) j * OC + oc] = src[(size_t) oc * IC * 3 + j];
            memcpy(B + o_b[i], want[base + "bias"].data.data(), (size_t) OC * 4);
        }
        auto transpose_gate = [&](const char * name, size_t o) {   // [gate row][k] -> [k][gate row]
            const float * src = (const float *) want[name].data.data(); float * dst = (float *) (B + o);
            for (int j = 0; j < 4 * VAD_HID; ++j) for (int k = 0; k < VAD_HID; ++k) dst[(size_t) k * 4 * VAD_HID + j] = src[(size_t) j * VAD_HID + k];
        };
        transpose_gate("_model.decoder.rnn.weight_ih", o_wih);
        transpose_gate("_model.decoder.rnn.weight_hh", o_whh);
        memcpy(B + o_bih, want["_model.decoder.rnn.bias_ih"].data.data(), (size_t) 4 * VAD_HID * 4);
        memcpy(B + o_bhh, want["_model.decoder.rnn.bias_hh"].data.data(), (size_t) 4 * VAD_HID * 4);
        memcpy(B + o_fw,  want["_model.decoder.decoder.2.weight"].data.data(), (size_t) VAD_HID * 2);
        memcpy(B + o_fb,  want["_model.decoder.decoder.2.bias"].data.data(), 4);
        VadWeights & w = m.hw;
        w.stft = (const __half *) (B + o_stft);
        for (int i = 0; i < 4; ++i) { w.enc_w[i] = (const __half *) (B + o_w[i]); w.enc_b[i] = (const float *) (B + o_b[i]); }
        w.w_ih = (const float *) (B + o_wih); w.w_hh = (const float *) (B + o_whh);
        w.b_ih = (const float *) (B + o_bih); w.b_hh = (const float *) (B + o_bhh);
        w.fin_w = (const __half *) (B + o_fw); w.fin_b = (const float *) (B + o_fb);
    }

    vctx->device = device;
    if (device >= 0) {
        if (m.n_loaded > 0 && !vad_upload(m, device)) { logf(LOG_ERROR, "%s: %s\n", __func__, last_error()); return nullptr; }
        if (cudaSetDevice(device) != cudaSuccess || cudaMalloc(&vctx->d_state, 2 * VAD_HID * sizeof(float)) != cudaSuccess ||
            cudaMemset(vctx->d_state, 0, 2 * VAD_HID * sizeof(float)) != cudaSuccess) {
            set_error("vad: failed to allocate memory for the VAD state on GPU %d", device); logf(LOG_ERROR, "%s: %s\n", __func__, last_error());
            vad_free_device(m, device); return nullptr;
        }
    }
    return vctx.release();
}

// ---- probabilities -> speech segments -------------------------------------------------------------------------------------
// Restates whisper_vad_segments_from_probs (src/whisper.cpp:5229-5463), itself modelled on silero's get_speech_timestamps:
// a hysteresis detector over the per-window probabilities, optional splitting of over-long speech at the last long-enough
// pause, merging of close neighbours, minimum-length filter, padding; all in samples, reported in centiseconds.
std::vector<VadSegment> vad_segments_from_probs(const float * probs, int n_probs, int n_window, const whisper_vad_params & p) {
    const int rate = SAMPLE_RATE;
    const int min_silence = rate * p.min_silence_duration_ms / 1000;
    const int min_speech  = rate * p.min_speech_duration_ms / 1000;
    const int pad         = rate * p.speech_pad_ms / 1000;
    const int total       = n_probs * n_window;
    const int split_pause = rate * 98 / 1000;                 // a pause this long is a candidate cut point for max_speech
    const int merge_gap   = rate * 200 / 1000;
    int max_speech = INT_MAX / 2;
    if (!(p.max_speech_duration_s > 100000.0f)) {
        const int64_t v = (int64_t) rate * (int64_t) p.max_speech_duration_s - n_window - 2 * pad;
        if (v >= 0 && v <= INT_MAX) max_speech = (int) v;
    }
    const float hi = p.threshold;
    const float lo = std::max(p.threshold - 0.15f, 0.01f);

    struct Span { int start, end; };
    std::vector<Span> spans;
    bool in_speech = false, open = false;      // open: a span has been started and not yet emitted or dropped
    int start = 0;                             // first sample of the span being built
    int pause_at = 0;                          // where the current sub-threshold run began (0 = none)
    int cut_end = 0, cut_next = 0;             // last pause long enough to cut at, and where speech resumed after it

    auto reset_pause = [&]() { pause_at = cut_end = cut_next = 0; };
    for (int i = 0; i < n_probs; ++i) {
        const float pr = probs[i];
        const int   at = n_window * i;
        if (pr >= hi && pause_at) {            // speech resumed
            pause_at = 0;
            if (cut_next < cut_end) cut_next = at;
        }
        if (pr >= hi && !in_speech) { in_speech = true; open = true; start = at; continue; }
        if (in_speech && at - start > max_speech) {
            if (cut_end) {
                spans.push_back({start, cut_end});
                open = true;
                if (cut_next < cut_end) { in_speech = false; open = false; }     // still inside that pause
                else start = cut_next;
                reset_pause();
            } else {
                spans.push_back({start, at});
                reset_pause(); in_speech = false; open = false;
                continue;
            }
        }
        if (pr < lo && in_speech) {
            if (!pause_at) pause_at = at;
            if (at - pause_at > split_pause) cut_end = pause_at;
            if (at - pause_at < min_silence) continue;
            if (pause_at - start > min_speech) spans.push_back({start, pause_at});
            reset_pause(); in_speech = false; open = false;
            continue;
        }
    }
    if (open && total - start > min_speech) spans.push_back({start, total});

    for (size_t i = 0; i + 1 < spans.size(); ) {                       // merge neighbours separated by < 200 ms
        if (spans[i + 1].start - spans[i].end < merge_gap) { spans[i].end = spans[i + 1].end; spans.erase(spans.begin() + (long) i + 1); }
        else ++i;
    }
    spans.erase(std::remove_if(spans.begin(), spans.end(), [&](const Span & s) { return s.end - s.start < min_speech; }), spans.end());

    std::vector<VadSegment> out(spans.size());
    for (size_t i = 0; i < spans.size(); ++i) {
        if (i == 0) spans[i].start = spans[i].start > pad ? spans[i].start - pad : 0;
        if (i + 1 < spans.size()) {
            const int gap = spans[i + 1].start - spans[i].end;
            if (gap < 2 * pad) {                                       // close neighbours share the gap
                spans[i].end += gap / 2;
                spans[i + 1].start = spans[i + 1].start > gap / 2 ? spans[i + 1].start - gap / 2 : 0;
            } else {
                spans[i].end = spans[i].end + pad < total ? spans[i].end + pad : total;
                spans[i + 1].start = spans[i + 1].start > pad ? spans[i + 1].start - pad : 0;
            }
        } else {
            spans[i].end = spans[i].end + pad < total ? spans[i].end + pad : total;
        }
        out[i].start = samples_to_cs(spans[i].start);
        out[i].end   = samples_to_cs(spans[i].end);
        logf(LOG_INFO, "%s: VAD segment %d: start = %.2f, end = %.2f (duration: %.2f)\n", __func__, (int) i, out[i].start / 100.0, out[i].end / 100.0,
             (out[i].end - out[i].start) / 100.0);
    }
    return out;
}

// ---- cut the speech out, remember how to map times back ---------------------------------------------------------------------
void vad_cut_samples(const std::vector<VadSegment> & segs, const whisper_vad_params & p, const float * samples, int n_samples,
                     std::vector<float> & filtered, VadCut & cut) {
    cut.clear();
    filtered.clear();
    if (segs.empty()) return;
    cut.has_segments = true;
    const int n = (int) segs.size();
    const int overlap = (int) (p.samples_overlap * SAMPLE_RATE);
    const int gap     = (int) (0.1 * SAMPLE_RATE);                     // silence inserted between kept segments
    // size of the output as the reference computes it (start is not clamped in this first pass)
    int need = n > 1 ? (n - 1) * gap : 0;
    for (int i = 0; i < n; ++i) {
        const int s = cs_to_samples(segs[i].start);
        int e = cs_to_samples(segs[i].end);
        if (i < n - 1) e += overlap;
        e = std::min(e, n_samples - 1);
        need += e - s;
    }
    filtered.assign((size_t) std::max(need, 0), 0.0f);
    int off = 0;
    for (int i = 0; i < n; ++i) {
        const int s = std::min(cs_to_samples(segs[i].start), n_samples - 1);
        int e = std::min(cs_to_samples(segs[i].end), n_samples - 1);
        const int len_orig = e - s;
        if (i < n - 1) e = std::min(e + overlap, n_samples - 1);
        const int len = e - s;
        if (len <= 0) continue;
        VadSegmentInfo info;
        info.orig_start = segs[i].start; info.orig_end = segs[i].end;
        info.vad_start = samples_to_cs(off); info.vad_end = samples_to_cs(off + len_orig);
        cut.table.push_back({info.vad_start, info.orig_start});
        cut.table.push_back({info.vad_end,   info.orig_end});
        cut.segments.push_back(info);
        if ((size_t) (off + len) > filtered.size()) filtered.resize((size_t) (off + len), 0.0f);
        memcpy(filtered.data() + off, samples + s, (size_t) len * sizeof(float));
        off += len;
        if (i < n - 1) {
            cut.table.push_back({samples_to_cs(off),       info.orig_end});
            cut.table.push_back({samples_to_cs(off + gap), segs[i + 1].start});
            if ((size_t) (off + gap) > filtered.size()) filtered.resize((size_t) (off + gap), 0.0f);
            std::fill(filtered.begin() + off, filtered.begin() + off + gap, 0.0f);
            off += gap;
        }
    }
    std::sort(cut.table.begin(), cut.table.end(), [](const VadTimeMap & a, const VadTimeMap & b) { return a.processed < b.processed; });
    cut.table.erase(std::unique(cut.table.begin(), cut.table.end(), [](const VadTimeMap & a, const VadTimeMap & b) { return a.processed == b.processed; }), cut.table.end());
    logf(LOG_INFO, "%s: Reduced audio from %d to %d samples (%.1f%% reduction)\n", __func__, n_samples, off, 100.0f * (1.0f - (float) off / std::max(n_samples, 1)));
}

int64_t vad_map_segment_time(int64_t t, const std::vector<VadTimeMap> & tab) {
    if (tab.empty()) return t;
    if (t <= tab.front().processed) return tab.front().original;
    if (t >= tab.back().processed)  return tab.back().original;
    auto hi = std::lower_bound(tab.begin(), tab.end(), t, [](const VadTimeMap & e, int64_t v) { return e.processed < v; });
    if (hi->processed == t) return hi->original;
    auto lo = hi - 1;
    const int64_t dp = hi->processed - lo->processed;
    if (dp == 0) return lo->original;
    return lo->original + ((t - lo->processed) * (hi->original - lo->original)) / dp;
}

int64_t vad_map_token_time(int64_t t, const std::vector<VadSegmentInfo> & segs) {
    if (segs.empty()) return t;
    if (t <= segs.front().vad_start) return segs.front().orig_start;
    if (t >= segs.back().vad_end)    return segs.back().orig_end;
    for (size_t i = 0; i < segs.size(); ++i) {
        const VadSegmentInfo & s = segs[i];
        if (t >= s.vad_start && t <= s.vad_end) {
            const int64_t vd = s.vad_end - s.vad_start;
            return vd <= 0 ? s.orig_start : s.orig_start + (t - s.vad_start) * (s.orig_end - s.orig_start) / vd;
        }
        if (i + 1 < segs.size() && t > s.vad_end && t < segs[i + 1].vad_start)       // inside the inserted silence: snap to the nearer edge
            return t <= (s.vad_end + segs[i + 1].vad_start) / 2 ? s.orig_end : segs[i + 1].orig_start;
    }
    return t;
}

} // namespace wb

// ---- C API ----------------------------------------------------------------------------------------------------------------
using namespace wb;
extern "C" {

WB_EXPORT struct whisper_vad_params whisper_vad_default_params(void) {               // src/whisper.cpp:4464-4474
    whisper_vad_params p; p.threshold = 0.5f; p.min_speech_duration_ms = 250; p.min_silence_duration_ms = 100;
    p.max_speech_duration_s = 3.402823466e+38f; p.speech_pad_ms = 30; p.samples_overlap = 0.1f; return p;
}
WB_EXPORT struct whisper_vad_context_params whisper_vad_default_context_params(void) { whisper_vad_context_params p; p.n_threads = 4; p.use_gpu = false; p.gpu_device = 0; return p; }

// use_gpu is ignored: this engine has no CPU path, the network always runs on GPU `gpu_device`.
WB_EXPORT struct whisper_vad_context * whisper_vad_init_with_params(struct whisper_model_loader * loader, struct whisper_vad_context_params params) {
    if (!loader || !loader->read || !loader->eof) return nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) {
        set_error("no CUDA device available: libwhisper_b200 has no CPU fallback (VAD)"); logf(LOG_ERROR, "%s: %s\n", __func__, last_error());
        return nullptr;
    }
    if (params.gpu_device >= ndev) { set_error("vad: gpu_device %d out of range (%d devices)", params.gpu_device, ndev); logf(LOG_ERROR, "%s: %s\n", __func__, last_error()); return nullptr; }
    whisper_vad_context * v = nullptr;
    try { v = vad_load(loader, params.gpu_device < 0 ? 0 : params.gpu_device); } catch (...) { set_error("vad: allocation failed"); v = nullptr; }
    if (v) v->n_threads = params.n_threads;
    return v;
}
WB_EXPORT struct whisper_vad_context * whisper_vad_init_from_file_with_params(const char * path_model, struct whisper_vad_context_params params) {
    if (!path_model) return nullptr;
    logf(LOG_INFO, "%s: loading VAD model from '%s'\n", __func__, path_model);
    std::ifstream fin(path_model, std::ios::binary);
    if (!fin) { set_error("failed to open VAD model '%s'", path_model); logf(LOG_ERROR, "%s: failed to open VAD model '%s'\n", __func__, path_model); return nullptr; }
    whisper_model_loader loader = {};
    loader.context = &fin;
    loader.read  = [](void * c, void * out, size_t n) -> size_t { auto * f = (std::ifstream *) c; f->read((char *) out, (std::streamsize) n); return (size_t) f->gcount(); };
    loader.eof   = [](void * c) -> bool { return ((std::ifstream *) c)->eof(); };
    loader.close = [](void * c) { ((std::ifstream *) c)->close(); };
    return whisper_vad_init_with_params(&loader, params);
}

WB_EXPORT void whisper_vad_reset_state(struct whisper_vad_context * v) {
    if (v && v->device < 0) { v->h_state.assign(2 * VAD_HID, 0.0f); return; }      // host-only test context
    if (!v || !v->d_state) return;
    cudaSetDevice(v->device);
    cudaMemset(v->d_state, 0, 2 * VAD_HID * sizeof(float));
}
WB_EXPORT bool whisper_vad_detect_speech_no_reset(struct whisper_vad_context * v, const float * samples, int n_samples) {
    if (!v || (!samples && n_samples > 0) || n_samples < 0) return false;
    if (v->model.n_loaded == 0) { set_error("vad: the model file held no tensors"); logf(LOG_ERROR, "%s: %s\n", __func__, last_error()); return false; }
    const int64_t t0 = time_us();
    logf(LOG_INFO, "%s: detecting speech in %d samples\n", __func__, n_samples);
    bool ok;
    if (v->device < 0) {
        // Only the engine-less TEST context (wb200_dbg_scripted_context -> vad_filter) owns a host-only VAD context; every context made by
        // whisper_vad_init_* lives on a GPU.  The host walk of the kernels' phases stands in for the device there.
        if (v->h_state.empty()) v->h_state.assign(2 * VAD_HID, 0.0f);
        vad_forward_emulated(v->model, v->h_state.data(), samples, n_samples, v->probs);
        ok = true;
    } else
    ok = vad_forward_device(v->model, v->device, v->d_state, samples, n_samples, v->probs);
    v->t_vad_us += time_us() - t0;
    if (!ok) { logf(LOG_ERROR, "%s: failed to compute VAD graph: %s\n", __func__, last_error()); return false; }
    logf(LOG_INFO, "%s: vad time = %.2f ms processing %d samples\n", __func__, 1e-3f * v->t_vad_us, n_samples);
    return true;
}
WB_EXPORT bool whisper_vad_detect_speech(struct whisper_vad_context * v, const float * samples, int n_samples) {
    whisper_vad_reset_state(v);
    return whisper_vad_detect_speech_no_reset(v, samples, n_samples);
}
WB_EXPORT int     whisper_vad_n_probs(struct whisper_vad_context * v) { return v ? (int) v->probs.size() : 0; }
WB_EXPORT float * whisper_vad_probs  (struct whisper_vad_context * v) { return v ? v->probs.data() : nullptr; }

WB_EXPORT struct whisper_vad_segments * whisper_vad_segments_from_probs(struct whisper_vad_context * v, struct whisper_vad_params params) {
    if (!v) return nullptr;
    logf(LOG_INFO, "%s: detecting speech timestamps using %d probabilities\n", __func__, (int) v->probs.size());
    try {
        std::unique_ptr<whisper_vad_segments> s(new whisper_vad_segments());
        s->data = vad_segments_from_probs(v->probs.data(), (int) v->probs.size(), v->model.n_window, params);
        return s.release();
    } catch (...) { set_error("vad: failed to allocate memory for the segments"); return nullptr; }
}
WB_EXPORT struct whisper_vad_segments * whisper_vad_segments_from_samples(struct whisper_vad_context * v, struct whisper_vad_params params, const float * samples, int n_samples) {
    if (!whisper_vad_detect_speech(v, samples, n_samples)) { logf(LOG_ERROR, "%s: failed to detect speech\n", __func__); return nullptr; }
    return whisper_vad_segments_from_probs(v, params);
}
WB_EXPORT int   whisper_vad_segments_n_segments    (struct whisper_vad_segments * s)        { return (int) s->data.size(); }
WB_EXPORT float whisper_vad_segments_get_segment_t0(struct whisper_vad_segments * s, int i) { return (float) s->data[(size_t) i].start; }
WB_EXPORT float whisper_vad_segments_get_segment_t1(struct whisper_vad_segments * s, int i) { return (float) s->data[(size_t) i].end; }
WB_EXPORT void  whisper_vad_free_segments(struct whisper_vad_segments * s) { delete s; }
WB_EXPORT void  whisper_vad_free(struct whisper_vad_context * v) {
    if (!v) return;
    if (v->device >= 0) { if (v->d_state) { cudaSetDevice(v->device); cudaFree(v->d_state); } vad_free_device(v->model, v->device); }
    delete v;
}

} // extern "C"

// ---- host-only test hooks (tests/test_vad_cpu.py); nothing in the API above calls these ---------------------------------------
extern "C" {

// load the model file on the host only and walk the kernels' phases thread by thread: probabilities as the device computes them
WB_EXPORT int wb200_dbg_vad_probs(const char * model_path, const float * samples, int n_samples, int reset_every, float * probs_out, int cap) {
    std::ifstream fin(model_path, std::ios::binary);
    if (!fin) return -1;
    whisper_model_loader loader = {};
    loader.context = &fin;
    loader.read  = [](void * c, void * out, size_t n) -> size_t { auto * f = (std::ifstream *) c; f->read((char *) out, (std::streamsize) n); return (size_t) f->gcount(); };
    loader.eof   = [](void * c) -> bool { return ((std::ifstream *) c)->eof(); };
    loader.close = [](void *) {};
    std::unique_ptr<whisper_vad_context> v(vad_load(&loader, -1));
    if (!v || v->model.n_loaded == 0) return -2;
    float state[2 * VAD_HID] = {0};
    std::vector<float> all, part;
    // reset_every > 0: feed the clip in pieces of that many samples WITHOUT resetting the state (the streaming entry point)
    const int step = reset_every > 0 ? reset_every : std::max(n_samples, 1);
    for (int s0 = 0; s0 < n_samples; s0 += step) {
        vad_forward_emulated(v->model, state, samples + s0, std::min(step, n_samples - s0), part);
        all.insert(all.end(), part.begin(), part.end());
    }
    if ((int) all.size() > cap) return -3;
    memcpy(probs_out, all.data(), all.size() * sizeof(float));
    return (int) all.size();
}

// probabilities -> segments (centiseconds) without a context
WB_EXPORT int wb200_dbg_vad_segments(const float * probs, int n_probs, struct whisper_vad_params params, int64_t * t0, int64_t * t1, int cap) {
    const std::vector<VadSegment> segs = vad_segments_from_probs(probs, n_probs, VAD_WIN, params);
    if ((int) segs.size() > cap) return -1;
    for (size_t i = 0; i < segs.size(); ++i) { t0[i] = segs[i].start; t1[i] = segs[i].end; }
    return (int) segs.size();
}

// segments + PCM -> filtered PCM, mapping table and per-segment info; then maps `n_q` query times both ways
WB_EXPORT int wb200_dbg_vad_cut(const int64_t * t0, const int64_t * t1, int n_seg, struct whisper_vad_params params, const float * samples, int n_samples,
                                float * filtered, int cap, int64_t * table, int * n_table, int64_t * info, int * n_info,
                                const int64_t * q, int n_q, int64_t * q_seg, int64_t * q_tok) {
    std::vector<VadSegment> segs((size_t) n_seg);
    for (int i = 0; i < n_seg; ++i) { segs[i].start = t0[i]; segs[i].end = t1[i]; }
    std::vector<float> out; VadCut cut;
    vad_cut_samples(segs, params, samples, n_samples, out, cut);
    if ((int) out.size() > cap) return -1;
    memcpy(filtered, out.data(), out.size() * sizeof(float));
    *n_table = (int) cut.table.size(); *n_info = (int) cut.segments.size();
    for (size_t i = 0; i < cut.table.size(); ++i) { table[2 * i] = cut.table[i].processed; table[2 * i + 1] = cut.table[i].original; }
    for (size_t i = 0; i < cut.segments.size(); ++i) { info[4 * i] = cut.segments[i].orig_start; info[4 * i + 1] = cut.segments[i].orig_end; info[4 * i + 2] = cut.segments[i].vad_start; info[4 * i + 3] = cut.segments[i].vad_end; }
    for (int i = 0; i < n_q; ++i) { q_seg[i] = vad_map_segment_time(q[i], cut.table); q_tok[i] = vad_map_token_time(q[i], cut.segments); }
    return (int) out.size();
}

} // extern "C"
