"""CPU tests (no GPU): pin the NumPy restatement (oracle/ref_numpy.py) against the UNMODIFIED reference build
(oracle/_ref/libwhisper_ref.so) -- the reference ships no numeric golden vectors for this path."""
import ctypes as C
import os
import numpy as np
import pytest

from wbtest import (F16, Q4_0, Q5_0, Q8_0, Q4_K, Q5_K, ref_quantize, ref_dequantize, bind_whisper_api, read_wav_f32,
                    DATA_DIR, parse_model_header, ContextParams)
from oracle import ref_numpy as rn


@pytest.mark.parametrize("wtype", [Q4_0, Q5_0, Q8_0, Q4_K, Q5_K, F16])
def test_block_decoders_match_ggml(ref, wtype):
    rng = np.random.default_rng(wtype)
    w = (rng.standard_normal((16, 512)) * 0.1).astype(np.float32)
    w[3, :64] = 0.0
    raw = ref_quantize(ref, wtype, w)
    want = ref_dequantize(ref, wtype, raw, 16, 512)
    got = rn.dequantize(wtype, raw, 16, 512)
    assert np.array_equal(got, want)           # bit-exact: same f32 operations


def test_gelu_table_semantics(ref):
    ref.wref_fp16_round.restype = C.c_float
    ref.wref_fp16_round.argtypes = [C.c_float]
    x = np.linspace(-12, 12, 4001).astype(np.float32)
    y = rn.gelu(x)
    assert y[0] == 0.0 and y[-1] == x[-1]
    # the table output is an f16 value
    assert np.array_equal(y[(x > -10) & (x < 10)].astype(np.float16).astype(np.float32), y[(x > -10) & (x < 10)])


def _ref_ctx(ref, model):
    bind_whisper_api(ref)
    cp = ref.whisper_context_default_params()
    cp.use_gpu = False
    ctx = ref.whisper_init_from_file_with_params(os.path.join(DATA_DIR, model).encode(), cp)
    assert ctx
    return ctx


def test_log_mel_float64_restatement_vs_reference(ref):
    ref.wref_ctx_state.restype = C.c_void_p
    ref.wref_ctx_state.argtypes = [C.c_void_p]
    ref.wref_mel_copy.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
    ref.wref_mel_n_len.argtypes = [C.c_void_p]
    ctx = _ref_ctx(ref, "for-tests-ggml-tiny.en.bin")
    pcm = read_wav_f32(os.path.join(DATA_DIR, "jfk.wav"))
    assert ref.whisper_pcm_to_mel(ctx, pcm.ctypes.data, len(pcm), 4) == 0
    st = ref.wref_ctx_state(ctx)
    n_len = ref.wref_mel_n_len(st)
    assert n_len == (len(pcm) + 480000) // 160
    mel = np.empty((80, n_len), np.float32)
    assert ref.wref_mel_copy(st, mel.ctypes.data, mel.size) == 0
    _, filt = parse_model_header(os.path.join(DATA_DIR, "for-tests-ggml-tiny.en.bin"))
    want = rn.log_mel(pcm, filt)
    assert want.shape == mel.shape
    # reference computes the FFT in f32: agreement to ~1e-4 except in bins far below the frame energy
    assert np.abs(mel - want).max() < 2e-3
    assert np.abs(mel - want).mean() < 2e-5
    ref.whisper_free(ctx)


def test_q8_0_dot_restatement_matches_exact_math_within_quantisation_noise(ref):
    rng = np.random.default_rng(7)
    w = (rng.standard_normal((64, 256)) * 0.05).astype(np.float32)
    x = rng.standard_normal((3, 256)).astype(np.float32)
    for wtype in (Q4_0, Q5_0, Q8_0):
        raw = ref_quantize(ref, wtype, w)
        wd = rn.dequantize(wtype, raw, 64, 256)
        y = rn.mul_mat_q(wtype, raw, 64, 256, x)
        exact = x.astype(np.float64) @ wd.astype(np.float64).T
        rel = np.abs(y - exact).max() / np.abs(exact).max()
        assert rel < 2e-2      # the reference's own int8 activation noise
