#!/usr/bin/env python
"""Regenerates tests/golden/vad_r01.npz from the UNMODIFIED reference build (oracle/_ref/libwhisper_ref.so) and the Silero weights
the reference ships for its own tests (models/for-tests-silero-v6.2.0-ggml.bin, copied to oracle/_ref/data by `make -C oracle`):

  probs      whisper_vad_detect_speech on samples/jfk.wav                      (tests/test-vad.cpp:25-33: 344 probabilities)
  seg_t0/t1  whisper_vad_segments_from_probs with whisper_vad_default_params   (tests/test-vad.cpp:35-47: 4 segments)

Run where /root/reference exists:   make -C oracle && python tests/golden/make_vad_golden.py
"""
import ctypes as C
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from wbtest import DATA_DIR, load_ref, read_wav_f32  # noqa: E402
from test_vad_cpu import bind_vad, ref_probs  # noqa: E402

vp = C.c_void_p


def main():
    ref = load_ref()
    bind_vad(ref)
    path = os.path.join(DATA_DIR, "for-tests-silero-v6.2.0-ggml.bin").encode()
    vctx = ref.whisper_vad_init_from_file_with_params(path, ref.whisper_vad_default_context_params())
    assert vctx
    pcm = np.ascontiguousarray(read_wav_f32(os.path.join(DATA_DIR, "jfk.wav")), np.float32)
    probs = ref_probs(ref, vctx, pcm)
    ref.whisper_vad_segments_from_probs.restype = vp
    ref.whisper_vad_segments_from_probs.argtypes = [vp, type(ref.whisper_vad_default_params())]
    segs = ref.whisper_vad_segments_from_probs(vctx, ref.whisper_vad_default_params())
    n = ref.whisper_vad_segments_n_segments(segs)
    t0 = np.array([ref.whisper_vad_segments_get_segment_t0(segs, i) for i in range(n)], np.int64)
    t1 = np.array([ref.whisper_vad_segments_get_segment_t1(segs, i) for i in range(n)], np.int64)
    assert len(probs) == 344 and n == 4, (len(probs), n)          # the reference's own assertions
    out = os.path.join(HERE, "vad_r01.npz")
    np.savez_compressed(out, probs=probs, seg_t0=t0, seg_t1=t1)
    print("wrote", out, "probs", probs.shape, "segments", list(zip(t0.tolist(), t1.tolist())))


if __name__ == "__main__":
    main()
