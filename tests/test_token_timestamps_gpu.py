"""GPU: whisper_full with token_timestamps + max_len through the C ABI (the algorithm itself is pinned bit for bit on CPU by
tests/test_token_timestamps_cpu.py): every token carries t0 <= t1 inside its segment, times do not run backwards, wrapped
segments respect max_len."""
import ctypes as C
import os
import numpy as np
import pytest

from wbtest import DATA_DIR, read_wav_f32, Q5_0, TokenData
from e2e_util import Side, synth

pytestmark = pytest.mark.gpu


def test_token_timestamps_and_max_len(lib, ref, tmp_path):
    path = str(tmp_path / "m.bin")
    synth.write_model(path, "test-2l.en", Q5_0, seed=11, vocab_from=os.path.join(DATA_DIR, "for-tests-ggml-tiny.en.bin"))
    pcm = read_wav_f32(os.path.join(DATA_DIR, "jfk.wav"))
    A = Side(lib, path, False)
    try:
        L = A.L
        L.whisper_full_get_token_data.restype = TokenData
        L.whisper_full_get_token_data.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.whisper_full_get_segment_text.restype = C.c_char_p
        eot = L.whisper_token_eot(A.ctx)
        dur = len(pcm) * 100 // 16000

        def run(max_len):
            fp = L.whisper_full_default_params(0)
            fp.print_progress = False; fp.temperature_inc = 0.0; fp.greedy.best_of = 1
            fp.token_timestamps = True; fp.max_len = max_len; fp.split_on_word = False
            assert L.whisper_full(A.ctx, fp, pcm.ctypes.data_as(C.c_void_p), len(pcm)) == 0, L.wb200_last_error()
            segs = []
            for i in range(L.whisper_full_n_segments(A.ctx)):
                toks = [L.whisper_full_get_token_data(A.ctx, i, j) for j in range(L.whisper_full_n_tokens(A.ctx, i))]
                segs.append((L.whisper_full_get_segment_t0(A.ctx, i), L.whisper_full_get_segment_t1(A.ctx, i),
                             L.whisper_full_get_segment_text(A.ctx, i).decode("utf-8", "replace"), [(t.id, t.t0, t.t1, t.vlen) for t in toks]))
            return segs

        plain = run(0)
        assert len(plain) >= 1
        n_timed = 0
        for s0, s1, text, toks in plain:
            assert s0 <= s1
            for tid, t0, t1, vlen in toks:
                if tid >= eot:
                    continue
                assert 0 <= t0 <= dur + 3000 and 0 <= t1 <= dur + 3000, (t0, t1)   # every text token got a time
                assert vlen > 0.0
                n_timed += 1
        assert n_timed > 0
        wrapped = run(24)
        assert len(wrapped) >= len(plain)
        if any(len(text) > 40 for _, _, text, toks in plain if len(toks) > 3):
            assert len(wrapped) > len(plain)                     # long segments were cut
        assert [t[0] for s in wrapped for t in s[3]] == [t[0] for s in plain for t in s[3]]   # wrapping keeps the token sequence
    finally:
        A.free()
