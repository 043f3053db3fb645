"""CPU: the self-attention KV bookkeeping (KvCells: find_slot / seq_rm / seq_cp / cell_max / clear -- SURVEY.md 8 row a7) against
the reference's whisper_kv_cache_* functions (src/whisper.cpp:1019-1137) on random scripts that mimic greedy decoding, beam
reshuffles and prompt passes.  Exact equality of every intermediate (return value, head, cell_max) and of the final cell table."""
import ctypes as C
import numpy as np
import pytest


def _script(rng, size, n_ops, n_seq):
    ops = []
    pos = [0] * n_seq
    for _ in range(n_ops):
        r = rng.random()
        if r < 0.55:                                   # decode step / prompt of a sequence
            s = int(rng.integers(0, n_seq)); n = int(rng.integers(1, 6)) if rng.random() < 0.8 else int(rng.integers(6, 40))
            ops.append((0, n, pos[s], s, 0)); pos[s] += n
        elif r < 0.75:                                 # drop a sequence (or its tail)
            s = int(rng.integers(-1, n_seq)); p0 = int(rng.integers(-1, 30)); p1 = -1 if rng.random() < 0.6 else p0 + int(rng.integers(1, 20))
            ops.append((1, s, p0, p1, 0))
            if s >= 0 and p0 <= 0 and p1 < 0:
                pos[s] = 0
        elif r < 0.95:                                 # beam reshuffle: copy src -> dst
            a, b = int(rng.integers(0, n_seq)), int(rng.integers(0, n_seq))
            ops.append((2, a, b, -1 if rng.random() < 0.7 else int(rng.integers(0, 10)), -1)); pos[b] = max(pos[b], pos[a])
        else:
            ops.append((3, 0, 0, 0, 0)); pos = [0] * n_seq
    return np.asarray(ops, np.int32)


@pytest.mark.parametrize("size,n_seq,seed", [(64, 3, 0), (256, 5, 1), (3584, 7, 2), (32, 2, 3)])
def test_kv_bookkeeping_matches_reference(lib, ref, size, n_seq, seed):
    if not hasattr(ref, "wref_kv_script"):
        pytest.skip("oracle/_ref predates wref_kv_script (rebuild with make -C oracle)")
    rng = np.random.default_rng(seed)
    ops = _script(rng, size, 400, n_seq)
    out = []
    for fn in (lib.wb200_dbg_kv_script, ref.wref_kv_script):
        fn.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        trace = np.zeros((len(ops), 3), np.int32); cells = np.zeros((size, 2), np.int32)
        assert fn(size, ops.ctypes.data, len(ops), trace.ctypes.data, cells.ctypes.data) == 0
        out.append((trace, cells))
    assert np.array_equal(out[0][0], out[1][0])
    assert np.array_equal(out[0][1], out[1][1])
    assert (out[0][0][:, 0] == 0).any() or size > 64          # small tables also exercise the "no slot" path
