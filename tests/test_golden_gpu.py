"""GPU: libwhisper_b200.so against the committed reference outputs (tests/golden/golden_r01.npz); no reference
code runs here (only the vocabulary stub under oracle/_ref/data is read to write the synthetic model)."""
import os
import sys
import numpy as np
import pytest

from wbtest import ROOT, F16, Q5_0
from e2e_util import Side, taps, rms_err

sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from make_golden import golden_inputs, build_model, N_STEPS  # noqa: E402

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(ROOT, "tests", "golden", "golden_r01.npz"))
#        enc rms, kv rms, logits (top-64 values, in units of the logits' std)
TOL = {"f16": (2e-3, 3e-3, 1e-2), "q5_0": (3e-2, 3.5e-2, 5e-2)}


@pytest.mark.parametrize("wt,tag", [(F16, "f16"), (Q5_0, "q5_0")])
def test_product_against_golden(lib, tmp_path, wt, tag):
    os.environ["WB200_DEBUG_TAPS"] = "1"
    _, pcm_mel, pcm = golden_inputs()
    path = str(tmp_path / "m.bin")
    build_model(path, None, wt)
    A = Side(lib, path, False)
    try:
        if wt == F16:
            A.pcm_to_mel(pcm_mel)
            assert np.abs(taps_mel(A)[:, :128] - G["mel_1s"]).max() < 2e-3
        A.pcm_to_mel(pcm); A.encode(0)
        t = taps(A)
        e_enc, e_kv, e_log = TOL[tag]
        assert rms_err(t["enc"][:32], G[tag + "_enc_head"]) < e_enc
        assert rms_err(t["enc"].astype(np.float64).sum(1), G[tag + "_enc_rowsum"]) < e_enc
        assert rms_err(t["kc"][1, :16], G[tag + "_kc_l1_head"].astype(np.float32)) < e_kv
        assert rms_err(t["kv"][1, :16], G[tag + "_kv_l1_head"].astype(np.float32)) < e_kv
        sot = A.L.whisper_token_sot(A.ctx)
        toks = [sot]; n_past = 0
        for step in range(N_STEPS):
            feed = toks if step == 0 else toks[-1:]
            lg = A.decode(feed, n_past); n_past += len(feed)
            mean, std = G[tag + "_stats"][step]
            ids = G[tag + "_top_ids"][step]; want = G[tag + "_top_vals"][step]
            assert np.sqrt(((lg[ids] - want) ** 2).mean()) / std < e_log, step
            if (want[0] - want[1]) / std > 6 * e_log:
                assert int(lg.argmax()) == int(G[tag + "_next"][step])
            toks.append(int(G[tag + "_next"][step]))          # teacher forcing with the reference's choice
    finally:
        A.free()


def taps_mel(side):
    import ctypes as C
    L = side.L
    n = L.wb200_read_tensor(side.state, 0, None, 0)
    n_mel = L.whisper_model_n_mels(side.ctx)
    mel = np.empty((n_mel, n // n_mel), np.float32)
    assert L.wb200_read_tensor(side.state, 0, mel.ctypes.data_as(C.c_void_p), mel.size) == mel.size
    return mel
