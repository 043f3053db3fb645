"""GPU: libwhisper_b200.so against committed reference outputs AT THE SHAPES THE BENCHMARK RUNS (tests/golden/golden_r02_*.npz, written
by tests/golden/make_golden_large.py from the unmodified reference CPU build): large-v3 width (d = 1280, 20 heads, 128 mel bands,
51866 ids) at 4 + 4 layers and at the full 32 + 32 layers (the very model file bench.py measures), and base.en at full depth.

No reference code runs here.  Tolerances: the reference quantises ACTIVATIONS to int8 in front of every Q5_0 matrix of the ENCODER,
this engine multiplies exact f16 activations on the tensor cores; the error of that reference-side rounding grows with depth like a
random walk (measured by tests/test_oracle_cpu.py on the NumPy restatement).  Decoder steps mirror the integer arithmetic.  The first
decode (3 prompt rows in one call) also carries the reference's F16 accumulators of multi-row flash attention (ggml-cpu/ops.cpp:8585-8600).
"""
import os
import sys
import numpy as np
import pytest

from wbtest import ROOT
from e2e_util import Side, taps, rms_err

sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from make_golden_large import MODELS, ENC_ROWS, KV_ROWS, FIXED_IDS, N_STEPS, model_path, prompt_of, golden_pcm  # noqa: E402

pytestmark = pytest.mark.gpu

#            conv, enc rows, cross K/V, logits single-row step, logits of the multi-row prompt pass (units of the logits' std)
TOL = {"large4": (1e-3, 3e-2, 3.5e-2, 5e-2, 9e-2), "large32": (1e-3, 6e-2, 7e-2, 8e-2, 1.2e-1), "base": (1e-3, 3.5e-2, 4e-2, 5e-2, 9e-2)}


@pytest.mark.parametrize("tag", ["large4", "base", "large32"])
def test_product_against_large_shape_golden(lib, tag):
    os.environ["WB200_DEBUG_TAPS"] = "1"
    G = np.load(os.path.join(ROOT, "tests", "golden", "golden_r02_%s.npz" % tag))
    A = Side(lib, model_path(tag), False)
    try:
        A.pcm_to_mel(golden_pcm()); A.encode(0)
        t = taps(A)
        e_conv, e_enc, e_kv, e_log, e_log0 = TOL[tag]
        assert np.abs(t["mel"][:, :96] - G["mel_head"]).max() < 2e-3
        assert rms_err(t["conv"][ENC_ROWS], G["conv_rows"]) < e_conv
        assert rms_err(t["conv"].astype(np.float64).sum(1), G["conv_rowsum"]) < e_conv
        errs = {"enc": rms_err(t["enc"][ENC_ROWS], G["enc_rows"])}
        assert errs["enc"] < e_enc, errs
        # row sums: every frame of the encoder output contributes one number (scaled by what a row sum of independent errors would be)
        d = t["enc"].shape[1]
        assert np.sqrt(((t["enc"].astype(np.float64).sum(1) - G["enc_rowsum"]) ** 2).mean()) / (float(G["enc_rms"]) * np.sqrt(d)) < e_enc
        for l in (0, A.Lt - 1):
            kc, kv = t["kc"][l, KV_ROWS], t["kv"][l, KV_ROWS]
            gk, gv = G["kc_l%d" % l].astype(np.float32), G["kv_l%d" % l].astype(np.float32)
            assert rms_err(kc, gk) < e_kv and rms_err(kv, gv) < e_kv, (l, rms_err(kc, gk), rms_err(kv, gv))
            assert np.abs(kc[-4:]).max() == 0 and np.abs(kv[-4:]).max() == 0 and np.abs(gk[-4:]).max() == 0       # keys 1500..1503 stay zero
        toks = prompt_of(A); n_past = 0
        worst = 0.0
        for step in range(N_STEPS):
            feed = toks if step == 0 else toks[-1:]
            lg = A.decode(feed, n_past); n_past += len(feed)
            assert np.isfinite(lg).all()
            mean, std = G["stats"][step]
            ids = G["top_ids"][step]; want = G["top_vals"][step]
            e_top = np.sqrt(((lg[ids] - want) ** 2).mean()) / std
            e_fix = np.sqrt(((lg[FIXED_IDS] - G["fixed_vals"][step]) ** 2).mean()) / std
            tol = e_log0 if step == 0 else e_log
            assert e_top < tol and e_fix < tol, (step, e_top, e_fix)
            assert abs(lg.std() / std - 1.0) < 2e-2 and abs(lg.mean() - mean) / std < tol
            worst = max(worst, e_top, e_fix)
            if (want[0] - want[1]) / std > 6 * tol:
                assert int(lg.argmax()) == int(G["next"][step]), step
            toks.append(int(G["next"][step]))             # teacher forcing with the reference's choice
        print("%s: enc rows rms %.2e, worst logits error %.2e of std" % (tag, errs["enc"], worst))
    finally:
        A.free()
