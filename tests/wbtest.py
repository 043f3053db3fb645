"""Shared helpers for the tests: library loading, ctypes mirrors of the ABI structs, numpy (de)quantisers.

The reference library (oracle/_ref) is only ever used as a CHECKER here.
"""
import ctypes as C
import os
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(ROOT, "whisper.cpp_b200", "libwhisper_b200.so")
REF_PATH = os.path.join(ROOT, "oracle", "_ref", "libwhisper_ref.so")

# ggml_type ids (ggml/include/ggml.h:390-405)
F32, F16, Q4_0, Q5_0, Q8_0, Q4_K, Q5_K = 0, 1, 2, 6, 8, 12, 13
BLOCK_BYTES = {Q4_0: (32, 18), Q5_0: (32, 22), Q8_0: (32, 34), Q4_K: (256, 144), Q5_K: (256, 176)}


def load_ref():
    if not os.path.exists(REF_PATH):
        import pytest
        pytest.skip("oracle/_ref/libwhisper_ref.so not built (run __graft_entry__.build() where /root/reference exists)")
    L = C.CDLL(REF_PATH)
    L.wref_quantize.restype = C.c_int64
    L.wref_quantize.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64]
    L.wref_row_size.restype = C.c_int64
    L.wref_row_size.argtypes = [C.c_int, C.c_int64]
    L.wref_dequantize.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int64]
    return L


def load_lib():
    L = C.CDLL(LIB_PATH)
    L.wb200_last_error.restype = C.c_char_p
    return L


def ref_quantize(ref, wtype, w):
    """w: float32 [rows][k] -> bytes in the ggml file layout (ggml_quantize_chunk, ggml-quants.c)"""
    w = np.ascontiguousarray(w, dtype=np.float32)
    rows, k = w.shape
    if wtype == F16:
        return w.astype(np.float16).tobytes()
    nbytes = ref.wref_row_size(wtype, k) * rows
    out = np.empty(nbytes, dtype=np.uint8)
    got = ref.wref_quantize(wtype, w.ctypes.data, out.ctypes.data, rows, k)
    assert got == nbytes
    return out.tobytes()


def ref_dequantize(ref, wtype, raw, rows, k):
    if wtype == F16:
        return np.frombuffer(raw, dtype=np.float16).astype(np.float32).reshape(rows, k)
    out = np.empty(rows * k, dtype=np.float32)
    buf = np.frombuffer(raw, dtype=np.uint8)
    ref.wref_dequantize(wtype, buf.ctypes.data, out.ctypes.data, rows * k)
    return out.reshape(rows, k)


def gelu_ref_f16(x):
    """GELU with the reference CPU semantics (f16 table): ggml/src/ggml-cpu/vec.h:988-1001"""
    x = np.asarray(x, dtype=np.float32)
    xh = x.astype(np.float16).astype(np.float32)
    g = 0.5 * xh * (1.0 + np.tanh(np.float32(0.79788456080286535587989211986876) * xh * (1.0 + np.float32(0.044715) * xh * xh)))
    g = g.astype(np.float16).astype(np.float32)
    return np.where(x <= -10, 0.0, np.where(x >= 10, x, g)).astype(np.float32)


# ----------------------------------------------------------------------------------------------------------------------
# ctypes mirrors of the by-value ABI structs (include/whisper_b200.h == reference include/whisper.h:116-151, 487-591)
# ----------------------------------------------------------------------------------------------------------------------
class Aheads(C.Structure):
    _fields_ = [("n_heads", C.c_size_t), ("heads", C.c_void_p)]


class ContextParams(C.Structure):
    _fields_ = [("use_gpu", C.c_bool), ("flash_attn", C.c_bool), ("gpu_device", C.c_int),
                ("dtw_token_timestamps", C.c_bool), ("dtw_aheads_preset", C.c_int), ("dtw_n_top", C.c_int),
                ("dtw_aheads", Aheads), ("dtw_mem_size", C.c_size_t)]


class TokenData(C.Structure):
    _fields_ = [("id", C.c_int32), ("tid", C.c_int32), ("p", C.c_float), ("plog", C.c_float), ("pt", C.c_float),
                ("ptsum", C.c_float), ("t0", C.c_int64), ("t1", C.c_int64), ("t_dtw", C.c_int64), ("vlen", C.c_float)]


class VadParams(C.Structure):
    _fields_ = [("threshold", C.c_float), ("min_speech_duration_ms", C.c_int), ("min_silence_duration_ms", C.c_int),
                ("max_speech_duration_s", C.c_float), ("speech_pad_ms", C.c_int), ("samples_overlap", C.c_float)]


class Greedy(C.Structure):
    _fields_ = [("best_of", C.c_int)]


class BeamSearch(C.Structure):
    _fields_ = [("beam_size", C.c_int), ("patience", C.c_float)]


class FullParams(C.Structure):
    _fields_ = [("strategy", C.c_int), ("n_threads", C.c_int), ("n_max_text_ctx", C.c_int), ("offset_ms", C.c_int),
                ("duration_ms", C.c_int), ("translate", C.c_bool), ("no_context", C.c_bool), ("no_timestamps", C.c_bool),
                ("single_segment", C.c_bool), ("print_special", C.c_bool), ("print_progress", C.c_bool),
                ("print_realtime", C.c_bool), ("print_timestamps", C.c_bool), ("token_timestamps", C.c_bool),
                ("thold_pt", C.c_float), ("thold_ptsum", C.c_float), ("max_len", C.c_int), ("split_on_word", C.c_bool),
                ("max_tokens", C.c_int), ("debug_mode", C.c_bool), ("audio_ctx", C.c_int), ("tdrz_enable", C.c_bool),
                ("suppress_regex", C.c_char_p), ("initial_prompt", C.c_char_p), ("carry_initial_prompt", C.c_bool),
                ("prompt_tokens", C.POINTER(C.c_int32)), ("prompt_n_tokens", C.c_int), ("language", C.c_char_p),
                ("detect_language", C.c_bool), ("suppress_blank", C.c_bool), ("suppress_nst", C.c_bool),
                ("temperature", C.c_float), ("max_initial_ts", C.c_float), ("length_penalty", C.c_float),
                ("temperature_inc", C.c_float), ("entropy_thold", C.c_float), ("logprob_thold", C.c_float),
                ("no_speech_thold", C.c_float), ("greedy", Greedy), ("beam_search", BeamSearch),
                ("new_segment_callback", C.c_void_p), ("new_segment_callback_user_data", C.c_void_p),
                ("progress_callback", C.c_void_p), ("progress_callback_user_data", C.c_void_p),
                ("encoder_begin_callback", C.c_void_p), ("encoder_begin_callback_user_data", C.c_void_p),
                ("abort_callback", C.c_void_p), ("abort_callback_user_data", C.c_void_p),
                ("logits_filter_callback", C.c_void_p), ("logits_filter_callback_user_data", C.c_void_p),
                ("grammar_rules", C.c_void_p), ("n_grammar_rules", C.c_size_t), ("i_start_rule", C.c_size_t),
                ("grammar_penalty", C.c_float), ("vad", C.c_bool), ("vad_model_path", C.c_char_p),
                ("vad_params", VadParams)]


def bind_whisper_api(L):
    """declare the signatures of the whisper.h entry points used by the tests on either library"""
    vp = C.c_void_p
    L.whisper_context_default_params.restype = ContextParams
    L.whisper_full_default_params.restype = FullParams
    L.whisper_full_default_params.argtypes = [C.c_int]
    L.whisper_init_from_file_with_params.restype = vp
    L.whisper_init_from_file_with_params.argtypes = [C.c_char_p, ContextParams]
    L.whisper_init_from_file_with_params_no_state.restype = vp
    L.whisper_init_from_file_with_params_no_state.argtypes = [C.c_char_p, ContextParams]
    L.whisper_init_state.restype = vp
    L.whisper_init_state.argtypes = [vp]
    L.whisper_free.argtypes = [vp]
    L.whisper_free_state.argtypes = [vp]
    L.whisper_pcm_to_mel.argtypes = [vp, vp, C.c_int, C.c_int]
    L.whisper_set_mel.argtypes = [vp, vp, C.c_int, C.c_int]
    L.whisper_encode.argtypes = [vp, C.c_int, C.c_int]
    L.whisper_decode.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int]
    L.whisper_get_logits.restype = C.POINTER(C.c_float)
    L.whisper_get_logits.argtypes = [vp]
    L.whisper_full.argtypes = [vp, FullParams, vp, C.c_int]
    L.whisper_full_with_state.argtypes = [vp, vp, FullParams, vp, C.c_int]
    L.whisper_full_parallel.argtypes = [vp, FullParams, vp, C.c_int, C.c_int]
    L.whisper_full_n_segments.argtypes = [vp]
    L.whisper_full_n_segments_from_state.argtypes = [vp]
    L.whisper_full_n_tokens.argtypes = [vp, C.c_int]
    L.whisper_full_n_tokens_from_state.argtypes = [vp, C.c_int]
    L.whisper_full_get_token_id.argtypes = [vp, C.c_int, C.c_int]
    L.whisper_full_get_token_id_from_state.argtypes = [vp, C.c_int, C.c_int]
    L.whisper_full_get_token_data.restype = TokenData
    L.whisper_full_get_token_data.argtypes = [vp, C.c_int, C.c_int]
    L.whisper_full_get_token_p.restype = C.c_float
    L.whisper_full_get_token_p.argtypes = [vp, C.c_int, C.c_int]
    L.whisper_full_get_segment_text.restype = C.c_char_p
    L.whisper_full_get_segment_text.argtypes = [vp, C.c_int]
    L.whisper_full_get_segment_text_from_state.restype = C.c_char_p
    L.whisper_full_get_segment_text_from_state.argtypes = [vp, C.c_int]
    for n in ("whisper_full_get_segment_t0", "whisper_full_get_segment_t1"):
        getattr(L, n).restype = C.c_int64
        getattr(L, n).argtypes = [vp, C.c_int]
        getattr(L, n + "_from_state").restype = C.c_int64
        getattr(L, n + "_from_state").argtypes = [vp, C.c_int]
    L.whisper_full_get_segment_no_speech_prob.restype = C.c_float
    L.whisper_full_get_segment_no_speech_prob.argtypes = [vp, C.c_int]
    L.whisper_tokenize.argtypes = [vp, C.c_char_p, C.POINTER(C.c_int32), C.c_int]
    L.whisper_token_to_str.restype = C.c_char_p
    L.whisper_token_to_str.argtypes = [vp, C.c_int32]
    L.whisper_lang_str.restype = C.c_char_p
    L.whisper_lang_str_full.restype = C.c_char_p
    L.whisper_lang_id.argtypes = [C.c_char_p]
    L.whisper_print_system_info.restype = C.c_char_p
    L.whisper_version.restype = C.c_char_p
    L.whisper_model_type_readable.restype = C.c_char_p
    L.whisper_model_type_readable.argtypes = [vp]
    L.whisper_get_timings.restype = C.POINTER(C.c_float * 5)
    L.whisper_get_timings.argtypes = [vp]
    for n in ("whisper_n_len", "whisper_n_vocab", "whisper_n_text_ctx", "whisper_n_audio_ctx", "whisper_is_multilingual",
              "whisper_model_n_vocab", "whisper_model_n_audio_ctx", "whisper_model_n_audio_state", "whisper_model_n_audio_head",
              "whisper_model_n_audio_layer", "whisper_model_n_text_ctx", "whisper_model_n_text_state", "whisper_model_n_text_head",
              "whisper_model_n_text_layer", "whisper_model_n_mels", "whisper_model_ftype", "whisper_model_type",
              "whisper_token_eot", "whisper_token_sot", "whisper_token_solm", "whisper_token_prev", "whisper_token_nosp",
              "whisper_token_not", "whisper_token_beg", "whisper_token_translate", "whisper_token_transcribe",
              "whisper_full_lang_id", "whisper_print_timings", "whisper_reset_timings"):
        getattr(L, n).argtypes = [vp]
    L.whisper_token_lang.argtypes = [vp, C.c_int]
    return L


def read_wav_f32(path):
    import wave
    with wave.open(path, "rb") as w:
        assert w.getframerate() == 16000 and w.getsampwidth() == 2
        data = np.frombuffer(w.readframes(w.getnframes()), dtype=np.int16)
        if w.getnchannels() == 2:
            data = data.reshape(-1, 2).mean(axis=1)
    return (data.astype(np.float32) / 32768.0).astype(np.float32)


DATA_DIR = os.path.join(ROOT, "oracle", "_ref", "data")


def parse_model_header(path):
    """hparams + mel filters of a legacy ggml model file (src/whisper.cpp:1496-1586)"""
    import struct
    with open(path, "rb") as f:
        magic, = struct.unpack("<I", f.read(4))
        assert magic == 0x67676d6c
        hp = struct.unpack("<11i", f.read(44))
        n_mel, n_fft = struct.unpack("<2i", f.read(8))
        filt = np.frombuffer(f.read(n_mel * n_fft * 4), dtype=np.float32).reshape(n_mel, n_fft).copy()
    return hp, filt
