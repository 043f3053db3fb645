"""Side-by-side drivers for the product library and the reference library (same whisper.h calls on both)."""
import ctypes as C
import importlib.util
import os
import numpy as np

from wbtest import ROOT, DATA_DIR, bind_whisper_api, load_lib, load_ref, FullParams, TokenData

spec = importlib.util.spec_from_file_location("wb_synth", os.path.join(ROOT, "whisper.cpp_b200", "synth.py"))
synth = importlib.util.module_from_spec(spec)
spec.loader.exec_module(synth)

vp = C.c_void_p
_LOG_CB = C.CFUNCTYPE(None, C.c_int, C.c_char_p, vp)
_quiet = _LOG_CB(lambda level, text, ud: None)


def _p(a):
    return a.ctypes.data_as(vp)


class Side:
    """one library + one context"""

    def __init__(self, L, model_path, is_ref):
        self.L = bind_whisper_api(L)
        self.is_ref = is_ref
        if not os.environ.get("WB200_VERBOSE"):
            L.whisper_log_set.argtypes = [_LOG_CB, vp]
            L.whisper_log_set(_quiet, None)
        cp = L.whisper_context_default_params()
        cp.use_gpu = not is_ref
        self.ctx = L.whisper_init_from_file_with_params(model_path.encode(), cp)
        assert self.ctx, "init failed: " + (L.wb200_last_error().decode() if not is_ref else "reference")
        self.n_vocab = L.whisper_n_vocab(self.ctx)
        self.d = L.whisper_model_n_audio_state(self.ctx)
        self.T = L.whisper_model_n_audio_ctx(self.ctx)
        self.Lt = L.whisper_model_n_text_layer(self.ctx)
        if is_ref:
            L.wref_ctx_state.restype = vp; L.wref_ctx_state.argtypes = [vp]
            for n in ("wref_embd_conv", "wref_embd_enc"):
                getattr(L, n).restype = C.c_int64; getattr(L, n).argtypes = [vp, vp, C.c_int64]
            for n in ("wref_kv_cross_k", "wref_kv_cross_v"):
                getattr(L, n).restype = C.c_int64; getattr(L, n).argtypes = [vp, vp, C.c_int64]
            L.wref_mel_copy.argtypes = [vp, vp, C.c_int64]; L.wref_mel_n_len.argtypes = [vp]
            self.state = L.wref_ctx_state(self.ctx)
        else:
            L.wb200_read_tensor.restype = C.c_int64
            L.wb200_read_tensor.argtypes = [vp, C.c_int, vp, C.c_int64]
            L.wb200_ctx_state.restype = vp; L.wb200_ctx_state.argtypes = [vp]
            self.state = L.wb200_ctx_state(self.ctx)

    def free(self):
        self.L.whisper_free(self.ctx)

    def pcm_to_mel(self, pcm):
        assert self.L.whisper_pcm_to_mel(self.ctx, _p(pcm), len(pcm), 4) == 0

    def encode(self, offset=0):
        assert self.L.whisper_encode(self.ctx, offset, 4) == 0

    def decode(self, tokens, n_past):
        t = np.asarray(tokens, np.int32)
        rc = self.L.whisper_decode(self.ctx, _p(t), len(t), n_past, 4)
        assert rc == 0, (rc, None if self.is_ref else self.L.wb200_last_error())
        lg = self.L.whisper_get_logits(self.ctx)
        row = (len(t) - 1) * self.n_vocab
        return np.ctypeslib.as_array(lg, shape=(len(t) * self.n_vocab,))[row:row + self.n_vocab].copy()

    def full(self, pcm, **kw):
        fp = self.L.whisper_full_default_params(kw.pop("strategy", 0))
        fp.print_progress = False
        for k, v in kw.items():
            if k in ("greedy_best_of",):
                fp.greedy.best_of = v
            elif k == "beam_size":
                fp.beam_search.beam_size = v
            else:
                setattr(fp, k, v)
        rc = self.L.whisper_full(self.ctx, fp, _p(pcm), len(pcm))
        segs = []
        for i in range(self.L.whisper_full_n_segments(self.ctx)):
            toks = [self.L.whisper_full_get_token_id(self.ctx, i, j) for j in range(self.L.whisper_full_n_tokens(self.ctx, i))]
            segs.append((self.L.whisper_full_get_segment_t0(self.ctx, i), self.L.whisper_full_get_segment_t1(self.ctx, i), toks,
                         self.L.whisper_full_get_segment_text(self.ctx, i)))
        return rc, segs


def rel_err(a, b):
    """max |a-b| relative to the rms of b"""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / (np.sqrt((b ** 2).mean()) + 1e-30))


def rms_err(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.sqrt(((a - b) ** 2).mean()) / (np.sqrt((b ** 2).mean()) + 1e-30))


def taps(side):
    """dict of the encode intermediates as float32 arrays in a COMMON layout:
    mel [n_mel][n_len], conv [T][d], enc [T][d], kc/kv [Lt][1536][d]"""
    L = side.L
    T, d, Lt = side.T, side.d, side.Lt
    Tp = (T + 255) // 256 * 256
    out = {}
    if side.is_ref:
        n_len = L.wref_mel_n_len(side.state)
        n_mel = L.whisper_model_n_mels(side.ctx)
        mel = np.empty((n_mel, n_len), np.float32); assert L.wref_mel_copy(side.state, _p(mel), mel.size) == 0
        conv = np.empty((d, T), np.float32); assert L.wref_embd_conv(side.state, _p(conv), conv.size) == conv.size   # ne=[T, d]: d rows of T
        enc = np.empty((T, d), np.float32); assert L.wref_embd_enc(side.state, _p(enc), enc.size) == enc.size
        kc = np.empty((Lt, Tp, d), np.float16); kv = np.empty((Lt, Tp, d), np.float16)
        assert L.wref_kv_cross_k(side.state, _p(kc), kc.size) == kc.size
        assert L.wref_kv_cross_v(side.state, _p(kv), kv.size) == kv.size
        out.update(mel=mel, conv=conv.T.copy(), enc=enc, kc=kc.astype(np.float32), kv=kv.astype(np.float32))
    else:
        n = L.wb200_read_tensor(side.state, 0, None, 0)
        n_mel = L.whisper_model_n_mels(side.ctx)
        mel = np.empty((n_mel, n // n_mel), np.float32); assert L.wb200_read_tensor(side.state, 0, _p(mel), mel.size) == mel.size
        conv = np.empty((T, d), np.float32); enc = np.empty((T, d), np.float32)
        assert L.wb200_read_tensor(side.state, 1, _p(conv), conv.size) == conv.size, L.wb200_last_error()
        assert L.wb200_read_tensor(side.state, 2, _p(enc), enc.size) == enc.size
        kc = np.empty((Lt, Tp, d), np.float32); kv = np.empty((Lt, Tp, d), np.float32)
        assert L.wb200_read_tensor(side.state, 3, _p(kc), kc.size) == kc.size
        assert L.wb200_read_tensor(side.state, 4, _p(kv), kv.size) == kv.size
        out.update(mel=mel, conv=conv, enc=enc, kc=kc, kv=kv)
    return out
