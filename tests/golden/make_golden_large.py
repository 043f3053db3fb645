#!/usr/bin/env python
"""Regenerates tests/golden/golden_r02_<tag>.npz: outputs of the UNMODIFIED reference CPU build (oracle/_ref/libwhisper_ref.so) at the
SHAPES THE BENCHMARK RUNS -- large-v3 width (d = 1280, 20 heads, 128 mel bands, 51866 ids) -- and at full base.en depth.

  large4   large-v3 width, 4 + 4 layers, Q5_0   (seed 3)
  large32  large-v3, 32 + 32 layers, Q5_0       (seed 0, fast_pool: the very file bench.py measures)
  base     base.en, 6 + 6 layers, Q5_0          (seed 4)

Run where /root/reference exists (after `make -C oracle`):   python tests/golden/make_golden_large.py [tags...]
Input: synth_audio(seed 1234, 30 s).  Stored per model: mel head, rows of the conv stem / encoder output,
every row's sum, rows of the cross K/V of the first and last text layer (padded zero rows included), and for a multi-token prompt
pass + 5 teacher-forced single-token steps the top-64 logits, 256 fixed ids and mean / std.  The reference ships no numeric golden
vectors for this path, hence this generated fixture; tests/test_golden_large_gpu.py compares the CUDA path to it.
"""
import os
import sys
import time
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from wbtest import Q5_0, load_ref  # noqa: E402
from e2e_util import Side, synth, taps  # noqa: E402

MODELS = {"large4": ("large-v3-4l", 3, False), "large32": ("large-v3", 0, True), "base": ("base.en", 4, False)}
N_STEPS = 6
TOPK = 64
ENC_ROWS = np.r_[0:8, 700:708, 1492:1500]
KV_ROWS = np.r_[0:4, 748:752, 1496:1504]          # 1500..1503: the zero keys every decoder query attends to
FIXED_IDS = np.random.default_rng(99).integers(0, 51864, 256)


def model_path(tag):
    cfg, seed, fast = MODELS[tag]
    return synth.cached_model(cfg, Q5_0, seed=seed, fast_pool=fast, tag=None if tag == "large32" else "s%d" % seed)


def prompt_of(side):
    L = side.L
    sot = L.whisper_token_sot(side.ctx)
    if L.whisper_is_multilingual(side.ctx):
        return [sot, sot + 1, L.whisper_token_transcribe(side.ctx)]
    return [sot, 1000, 2000]                          # three rows as well: a multi-token pass


def golden_pcm():
    return synth.synth_audio(seed=1234, seconds=30.0)


def main():
    ref = load_ref()
    tags = sys.argv[1:] or list(MODELS)
    pcm = golden_pcm()
    for tag in tags:
        t0 = time.time()
        B = Side(ref, model_path(tag), True)
        B.pcm_to_mel(pcm); B.encode(0)
        t = taps(B)
        out = {"mel_head": t["mel"][:, :96].copy(), "conv_rows": t["conv"][ENC_ROWS].copy(), "conv_rowsum": t["conv"].astype(np.float64).sum(1),
               "enc_rows": t["enc"][ENC_ROWS].copy(), "enc_rowsum": t["enc"].astype(np.float64).sum(1), "enc_rms": np.float64(np.sqrt((t["enc"].astype(np.float64) ** 2).mean()))}
        for l in (0, B.Lt - 1):
            out["kc_l%d" % l] = t["kc"][l, KV_ROWS].astype(np.float16); out["kv_l%d" % l] = t["kv"][l, KV_ROWS].astype(np.float16)
        toks = prompt_of(B); n_past = 0
        ids, vals, fixed, stats, nxt = [], [], [], [], []
        for step in range(N_STEPS):
            feed = toks if step == 0 else toks[-1:]
            lg = B.decode(feed, n_past); n_past += len(feed)
            top = np.argsort(-lg)[:TOPK]
            ids.append(top.astype(np.int32)); vals.append(lg[top]); fixed.append(lg[FIXED_IDS]); stats.append([lg.mean(), lg.std()])
            nxt.append(int(lg.argmax())); toks.append(int(lg.argmax()))
        out.update(top_ids=np.stack(ids), top_vals=np.stack(vals).astype(np.float32), fixed_vals=np.stack(fixed).astype(np.float32),
                   stats=np.asarray(stats, np.float64), next=np.asarray(nxt, np.int32))
        B.free()
        dst = os.path.join(HERE, "golden_r02_%s.npz" % tag)
        np.savez_compressed(dst, **out)
        print("wrote %s (%d bytes) in %.1f s" % (dst, os.path.getsize(dst), time.time() - t0), flush=True)


if __name__ == "__main__":
    main()
