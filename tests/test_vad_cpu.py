"""CPU: voice-activity detection (whisper_vad_*, params.vad) against the reference compiled in oracle/_ref.
  * network arithmetic: the kernels' phases walked thread by thread on the host (wb200_dbg_vad_probs) vs the reference's ggml graph
    on a synthetic silero-16k model -- same F16 rounding points, f32 sums in a different order.  A sum that lands next to an F16
    rounding boundary can round the other way (one F16 ulp = 5e-4 of that activation), so the bound on a probability is 1.5e-3 with a
    median below 5e-5 (measured: max 5e-4, median 4e-6);
  * probabilities -> speech segments: identical (integer arithmetic) on random and adversarial probability tracks;
  * PCM cut, mapping table, segment/token time mapping: identical to the reference's static whisper_vad() + getters."""
import ctypes as C
import os
import numpy as np
import pytest

from wbtest import DATA_DIR, FullParams, VadParams, bind_whisper_api, read_wav_f32
from vad_synth import write_vad_model, speechy_audio

vp = C.c_void_p
i64p = C.POINTER(C.c_int64)


class VadCtxParams(C.Structure):
    _fields_ = [("n_threads", C.c_int), ("use_gpu", C.c_bool), ("gpu_device", C.c_int)]


def bind_vad(R):
    R.whisper_vad_default_params.restype = VadParams
    R.whisper_vad_default_context_params.restype = VadCtxParams
    R.whisper_vad_init_from_file_with_params.restype = vp
    R.whisper_vad_init_from_file_with_params.argtypes = [C.c_char_p, VadCtxParams]
    for fn in ("whisper_vad_detect_speech", "whisper_vad_detect_speech_no_reset"):
        getattr(R, fn).restype = C.c_bool; getattr(R, fn).argtypes = [vp, vp, C.c_int]
    R.whisper_vad_reset_state.argtypes = [vp]
    R.whisper_vad_n_probs.argtypes = [vp]
    R.whisper_vad_probs.restype = C.POINTER(C.c_float); R.whisper_vad_probs.argtypes = [vp]
    R.whisper_vad_segments_from_samples.restype = vp
    R.whisper_vad_segments_from_samples.argtypes = [vp, VadParams, vp, C.c_int]
    R.whisper_vad_segments_n_segments.argtypes = [vp]
    R.whisper_vad_segments_get_segment_t0.restype = C.c_float; R.whisper_vad_segments_get_segment_t0.argtypes = [vp, C.c_int]
    R.whisper_vad_segments_get_segment_t1.restype = C.c_float; R.whisper_vad_segments_get_segment_t1.argtypes = [vp, C.c_int]
    R.whisper_vad_free_segments.argtypes = [vp]
    R.whisper_vad_free.argtypes = [vp]


def ref_probs(R, vctx, pcm, pieces=None):
    out = []
    if pieces is None:
        assert R.whisper_vad_detect_speech(vctx, pcm.ctypes.data_as(vp), len(pcm))
        return np.ctypeslib.as_array(R.whisper_vad_probs(vctx), (R.whisper_vad_n_probs(vctx),)).copy()
    R.whisper_vad_reset_state(vctx)
    for s0 in range(0, len(pcm), pieces):
        part = np.ascontiguousarray(pcm[s0:s0 + pieces])
        assert R.whisper_vad_detect_speech_no_reset(vctx, part.ctypes.data_as(vp), len(part))
        out.append(np.ctypeslib.as_array(R.whisper_vad_probs(vctx), (R.whisper_vad_n_probs(vctx),)).copy())
    return np.concatenate(out)


def need(ref, *names):
    for n in names:
        if not hasattr(ref, n):
            pytest.skip("oracle/_ref predates %s (rebuild with make -C oracle)" % n)


@pytest.mark.parametrize("seed,gain", [(1, 1.0), (2, 1.6)])
def test_vad_network_phases_match_reference_graph(lib, ref, tmp_path, seed, gain):
    bind_vad(ref)
    path = write_vad_model(str(tmp_path / "vad.bin"), seed=seed, gain=gain).encode()
    vctx = ref.whisper_vad_init_from_file_with_params(path, ref.whisper_vad_default_context_params())
    assert vctx
    lib.wb200_dbg_vad_probs.argtypes = [C.c_char_p, vp, C.c_int, C.c_int, vp, C.c_int]
    clips = [read_wav_f32(os.path.join(DATA_DIR, "jfk.wav"))[: 16000 * 6 + 137], speechy_audio(7.3, seed)]   # ragged last window
    for pcm in clips:
        pcm = np.ascontiguousarray(pcm, np.float32)
        want = ref_probs(ref, vctx, pcm)
        got = np.empty(len(want) + 8, np.float32)
        n = lib.wb200_dbg_vad_probs(path, pcm.ctypes.data_as(vp), len(pcm), 0, got.ctypes.data_as(vp), len(got))
        assert n == len(want) == (len(pcm) + 511) // 512
        d = np.abs(got[:n] - want)
        print("vad probs: n=%d  range %.3f..%.3f  max|d|=%.2e median %.2e" % (n, want.min(), want.max(), d.max(), np.median(d)))
        assert want.max() - want.min() > 0.05                     # the synthetic network is not stuck
        assert d.max() < 1.5e-3 and np.median(d) < 5e-5
    # streaming entry point: state carried across calls, pieces that are whole windows
    pcm = np.ascontiguousarray(clips[1][: 512 * 150])
    want = ref_probs(ref, vctx, pcm, pieces=512 * 37)
    got = np.empty(len(want) + 8, np.float32)
    n = lib.wb200_dbg_vad_probs(path, pcm.ctypes.data_as(vp), len(pcm), 512 * 37, got.ctypes.data_as(vp), len(got))
    assert n == len(want) and np.abs(got[:n] - want).max() < 1.5e-3
    assert np.abs(want - ref_probs(ref, vctx, pcm)).max() < 1e-6   # and equals the one-shot run on both sides
    ref.whisper_vad_free(vctx)


GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vad_r01.npz")
SILERO = os.path.join(DATA_DIR, "for-tests-silero-v6.2.0-ggml.bin")


def test_vad_real_silero_weights_match_golden_and_reference_kat(lib, ref):
    """the Silero weights the reference ships for tests/test-vad.cpp + samples/jfk.wav: 344 probabilities, 4 segments with the default
    parameters (the reference's own assertions), probabilities as committed in tests/golden/vad_r01.npz (made by make_vad_golden.py)"""
    if not os.path.exists(SILERO):
        pytest.skip("oracle/_ref/data lacks the silero fixture (make -C oracle data)")
    g = np.load(GOLDEN)
    bind_vad(ref)
    pcm = np.ascontiguousarray(read_wav_f32(os.path.join(DATA_DIR, "jfk.wav")), np.float32)
    vctx = ref.whisper_vad_init_from_file_with_params(SILERO.encode(), ref.whisper_vad_default_context_params())
    assert vctx
    want = ref_probs(ref, vctx, pcm)
    ref.whisper_vad_free(vctx)
    assert np.array_equal(want, g["probs"])                                   # the oracle build reproduces the committed fixture
    lib.wb200_dbg_vad_probs.argtypes = [C.c_char_p, vp, C.c_int, C.c_int, vp, C.c_int]
    got = np.empty(400, np.float32)
    n = lib.wb200_dbg_vad_probs(SILERO.encode(), pcm.ctypes.data_as(vp), len(pcm), 0, got.ctypes.data_as(vp), len(got))
    assert n == 344
    d = np.abs(got[:n] - g["probs"])
    print("silero v6.2.0 on jfk.wav: max|d|=%.2e median %.2e" % (d.max(), np.median(d)))
    assert d.max() < 1.5e-3 and np.median(d) < 5e-5
    sig = [vp, C.c_int, VadParams, i64p, i64p, C.c_int]
    lib.wb200_dbg_vad_segments.argtypes = sig
    lib.whisper_vad_default_params.restype = VadParams
    t0 = (C.c_int64 * 16)(); t1 = (C.c_int64 * 16)()
    k = lib.wb200_dbg_vad_segments(got.ctypes.data_as(vp), n, lib.whisper_vad_default_params(), t0, t1, 16)
    assert k == 4 and list(t0[:k]) == g["seg_t0"].tolist() and list(t1[:k]) == g["seg_t1"].tolist()


def test_vad_model_loader_rejects_bad_files(lib, tmp_path):
    lib.wb200_dbg_vad_probs.argtypes = [C.c_char_p, vp, C.c_int, C.c_int, vp, C.c_int]
    pcm = np.zeros(1024, np.float32); out = np.zeros(8, np.float32)
    call = lambda p: lib.wb200_dbg_vad_probs(p.encode(), pcm.ctypes.data_as(vp), len(pcm), 0, out.ctypes.data_as(vp), len(out))
    good = write_vad_model(str(tmp_path / "good.bin"), seed=3)
    assert call(good) == 2
    assert call(write_vad_model(str(tmp_path / "missing.bin"), seed=3, drop={"_model.decoder.rnn.bias_hh"})) == -2
    assert call(write_vad_model(str(tmp_path / "arch.bin"), seed=3, window=256)) == -2
    raw = open(good, "rb").read()
    open(str(tmp_path / "magic.bin"), "wb").write(b"\0\0\0\0" + raw[4:])
    assert call(str(tmp_path / "magic.bin")) == -2
    open(str(tmp_path / "trunc.bin"), "wb").write(raw[: len(raw) // 2])
    assert call(str(tmp_path / "trunc.bin")) == -2
    assert call(str(tmp_path / "nonexistent.bin")) == -1


def _tracks(rng):
    yield np.zeros(0, np.float32)
    yield np.full(40, 0.9, np.float32)
    yield np.full(40, 0.1, np.float32)
    yield np.array([0.9] * 3 + [0.1] * 50, np.float32)                                      # too short to count
    yield np.array(([0.9] * 30 + [0.2] * 2) * 6 + [0.9] * 5, np.float32)                    # pauses shorter than min_silence
    yield np.array(([0.9] * 30 + [0.2] * 5) * 6, np.float32)
    yield np.array([0.1] * 10 + [0.9] * 400 + [0.1] * 10, np.float32)
    yield np.array(([0.9] * 60 + [0.3] * 4 + [0.45] * 3 + [0.9] * 50 + [0.1] * 9) * 5, np.float32)   # hysteresis band 0.35..0.5
    for n in (17, 200, 1000, 3000):
        x = rng.random(n).astype(np.float32)
        yield x
        yield np.clip(np.convolve(x, np.ones(9) / 9, mode="same") * 1.3 - 0.1, 0, 1).astype(np.float32)
        blocks = np.repeat(rng.random(n // 12 + 1), 12)[:n]
        yield np.clip(blocks + 0.08 * rng.standard_normal(n), 0, 1).astype(np.float32)


def test_vad_segments_from_probs_identical(lib, ref):
    need(ref, "wref_vad_segments")
    bind_vad(ref)
    sig = [vp, C.c_int, VadParams, i64p, i64p, C.c_int]
    ref.wref_vad_segments.argtypes = sig; lib.wb200_dbg_vad_segments.argtypes = sig
    rng = np.random.default_rng(5)
    variants = []
    d = ref.whisper_vad_default_params()
    for thr, msp, msi, mx, pad in [(0.5, 250, 100, None, 30), (0.5, 250, 100, 6.0, 30), (0.3, 100, 300, 2.5, 100), (0.7, 500, 50, 1.0, 0),
                                   (0.1, 0, 0, 0.4, 400), (0.5, 64, 2000, 10.0, 30), (0.5, 250, 100, -3.0, 30), (0.5, 250, 100, 200000.0, 30)]:
        p = ref.whisper_vad_default_params()
        p.threshold = thr; p.min_speech_duration_ms = msp; p.min_silence_duration_ms = msi; p.speech_pad_ms = pad
        if mx is not None:
            p.max_speech_duration_s = mx
        variants.append(p)
    assert d.threshold == 0.5 and d.speech_pad_ms == 30
    n_cases = n_segs = 0
    for probs in _tracks(rng):
        for p in variants:
            res = []
            for L_, fn in ((ref, ref.wref_vad_segments), (lib, lib.wb200_dbg_vad_segments)):
                t0 = (C.c_int64 * 4096)(); t1 = (C.c_int64 * 4096)()
                n = fn(probs.ctypes.data_as(vp), len(probs), p, t0, t1, 4096)
                assert n >= 0
                res.append((n, list(t0[:n]), list(t1[:n])))
            assert res[0] == res[1], (len(probs), p.threshold, p.max_speech_duration_s, res[0][:1], res[1][:1])
            n_cases += 1; n_segs += res[0][0]
    assert n_cases > 100 and n_segs > 300
    dl = lib.whisper_vad_default_params
    dl.restype = VadParams
    a, b = dl(), d
    assert [getattr(a, f) for f, _ in VadParams._fields_] == [getattr(b, f) for f, _ in VadParams._fields_]


def test_vad_cut_and_time_mapping_identical(lib, ref, tmp_path):
    need(ref, "wref_vad_cut")
    bind_vad(ref)
    R = bind_whisper_api(ref)
    path = write_vad_model(str(tmp_path / "vad.bin"), seed=4, gain=1.6).encode()
    cp = R.whisper_context_default_params(); cp.use_gpu = False
    rctx = R.whisper_init_from_file_with_params(os.path.join(DATA_DIR, "for-tests-ggml-tiny.bin").encode(), cp)
    assert rctx
    vctx = ref.whisper_vad_init_from_file_with_params(path, ref.whisper_vad_default_context_params())
    cut_sig = [vp, C.c_int, i64p, C.POINTER(C.c_int), i64p, C.POINTER(C.c_int), i64p, C.c_int, i64p, i64p]
    ref.wref_vad_cut.argtypes = [vp, FullParams, vp, C.c_int] + cut_sig
    lib.wb200_dbg_vad_cut.argtypes = [i64p, i64p, C.c_int, VadParams, vp, C.c_int] + cut_sig
    n_multi = 0
    for seed, seconds in [(1, 9.0), (2, 21.5), (3, 14.2)]:
        pcm = np.ascontiguousarray(speechy_audio(seconds, seed))
        probs = ref_probs(ref, vctx, pcm)
        for thr_q, overlap, pad in [(0.5, 0.1, 30), (0.35, 0.0, 0), (0.65, 0.3, 120)]:
            fp = R.whisper_full_default_params(0)
            fp.vad = True; fp.vad_model_path = path
            fp.vad_params.threshold = float(np.quantile(probs, thr_q)); fp.vad_params.samples_overlap = overlap; fp.vad_params.speech_pad_ms = pad
            fp.vad_params.min_speech_duration_ms = 100; fp.vad_params.min_silence_duration_ms = 60
            q = np.concatenate([np.arange(-5, int(seconds * 100) + 50, 7), np.array([0, 1, 10**6])]).astype(np.int64)
            # the reference's own segments for these parameters, through its public API
            segs = ref.whisper_vad_segments_from_samples(vctx, fp.vad_params, pcm.ctypes.data_as(vp), len(pcm))
            ns = ref.whisper_vad_segments_n_segments(segs)
            t0 = (C.c_int64 * max(ns, 1))(*[int(ref.whisper_vad_segments_get_segment_t0(segs, i)) for i in range(ns)])
            t1 = (C.c_int64 * max(ns, 1))(*[int(ref.whisper_vad_segments_get_segment_t1(segs, i)) for i in range(ns)])
            ref.whisper_vad_free_segments(segs)
            res = []
            for which in (0, 1):
                filt = np.full(len(pcm) + 16000 * 4, 7.0, np.float32); tab = (C.c_int64 * 4096)(); info = (C.c_int64 * 4096)()
                nt = C.c_int(-1); ni = C.c_int(-1); qs = np.zeros(len(q), np.int64); qt = np.zeros(len(q), np.int64)
                tail = (filt.ctypes.data_as(vp), len(filt), tab, C.byref(nt), info, C.byref(ni), q.ctypes.data_as(i64p), len(q), qs.ctypes.data_as(i64p), qt.ctypes.data_as(i64p))
                if which == 0:
                    n = ref.wref_vad_cut(rctx, fp, pcm.ctypes.data_as(vp), len(pcm), *tail)
                else:
                    n = lib.wb200_dbg_vad_cut(t0, t1, ns, fp.vad_params, pcm.ctypes.data_as(vp), len(pcm), *tail)
                assert n >= 0
                res.append((n, filt[:n].copy(), list(tab[: 2 * nt.value]), list(info[: 4 * ni.value]), qs, qt))
            a, b = res
            assert a[0] == b[0] and np.array_equal(a[1], b[1]), (seed, thr_q, a[0], b[0])
            assert a[2] == b[2] and a[3] == b[3]
            assert np.array_equal(a[4], b[4]) and np.array_equal(a[5], b[5])
            if ns >= 2:
                n_multi += 1
                assert a[0] < len(pcm)                                       # something was cut out
    assert n_multi >= 3
    ref.whisper_vad_free(vctx)
    R.whisper_free(rctx)
