#!/usr/bin/env python
"""Regenerates tests/golden/golden_r01.npz from the UNMODIFIED reference build (oracle/_ref/libwhisper_ref.so).

Run where /root/reference exists (after `make -C oracle`):   python tests/golden/make_golden.py
The fixture pins, for inputs that are fully determined by seeds committed here:
  * the ggml block quantisers / dequantisers (ggml-quants.c) on a seeded matrix, all five formats;
  * the log-mel front end (src/whisper.cpp:3005-3272) on 1 s of seeded audio;
  * encoder output, cross K/V and teacher-forced decoder logits of the synthetic 2-layer model (synth.py, seed 7,
    F16 and Q5_0) on 3 s of seeded audio -- what whisper_encode / whisper_decode return on the reference CPU path.
The reference ships no numeric golden vectors for this path, hence this generated fixture.
"""
import os
import sys
import tempfile
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from wbtest import DATA_DIR, F16, Q4_0, Q5_0, Q8_0, Q4_K, Q5_K, load_ref, ref_quantize, ref_dequantize  # noqa: E402
from e2e_util import Side, synth, taps  # noqa: E402

N_STEPS = 6
TOPK = 64


def golden_inputs():
    """everything seeded; shared by the generator and by the tests"""
    rng = np.random.default_rng(20240901)
    w = (rng.standard_normal((8, 512)) * 0.1).astype(np.float32)
    w[2, :96] = 0.0
    pcm_mel = synth.synth_audio(seed=77, seconds=1.0)
    pcm_e2e = synth.synth_audio(seed=78, seconds=3.0)
    return w, pcm_mel, pcm_e2e


def build_model(path, ref, wt):
    stub = os.path.join(DATA_DIR, "for-tests-ggml-tiny.en.bin")
    synth.write_model(path, "test-2l.en", wt, seed=7, vocab_from=stub)


def run_side(side, pcm):
    side.pcm_to_mel(pcm)
    side.encode(0)
    t = taps(side)
    sot = side.L.whisper_token_sot(side.ctx)
    return t, sot


def main():
    ref = load_ref()
    w, pcm_mel, pcm_e2e = golden_inputs()
    out = {}
    for wt in (Q4_0, Q5_0, Q8_0, Q4_K, Q5_K):
        raw = ref_quantize(ref, wt, w)
        out["q%d_raw" % wt] = np.frombuffer(raw, np.uint8).copy()
        out["q%d_deq" % wt] = ref_dequantize(ref, wt, raw, *w.shape)
    tmp = tempfile.mkdtemp()
    for wt, tag in ((F16, "f16"), (Q5_0, "q5_0")):
        path = os.path.join(tmp, "m-%s.bin" % tag)
        build_model(path, ref, wt)
        B = Side(ref, path, True)
        if wt == F16:
            B.pcm_to_mel(pcm_mel)
            out["mel_1s"] = taps(B)["mel"][:, :128].copy()
        t, sot = run_side(B, pcm_e2e)
        out[tag + "_enc_head"] = t["enc"][:32].copy()                     # first 32 frames
        out[tag + "_enc_rowsum"] = t["enc"].astype(np.float64).sum(1)       # every frame, one number
        out[tag + "_kc_l1_head"] = t["kc"][1, :16].astype(np.float16)
        out[tag + "_kv_l1_head"] = t["kv"][1, :16].astype(np.float16)
        toks = [sot]; n_past = 0; ids = []; vals = []; stats = []; nxt = []
        for step in range(N_STEPS):
            feed = toks if step == 0 else toks[-1:]
            lg = B.decode(feed, n_past); n_past += len(feed)
            top = np.argsort(-lg)[:TOPK]
            ids.append(top.astype(np.int32)); vals.append(lg[top]); stats.append([lg.mean(), lg.std()])
            nxt.append(int(lg.argmax())); toks.append(int(lg.argmax()))
        out[tag + "_top_ids"] = np.stack(ids); out[tag + "_top_vals"] = np.stack(vals).astype(np.float32)
        out[tag + "_stats"] = np.asarray(stats, np.float64); out[tag + "_next"] = np.asarray(nxt, np.int32)
        B.free()
    dst = os.path.join(HERE, "golden_r01.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
