"""debug: first divergence between this engine and the reference in beam search on a conditioned model (per sampling call)"""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from wbtest import DATA_DIR, F16, Q5_0, TokenData, load_lib, load_ref
from e2e_util import Side, synth
vp = C.c_void_p
LOGITS_CB = C.CFUNCTYPE(None, vp, vp, C.POINTER(TokenData), C.c_int, C.POINTER(C.c_float), vp)
wt, seed, attn = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3])
strategy = int(sys.argv[4]) if len(sys.argv) > 4 else 1
path = "/tmp/dbg-cond.bin"
synth.write_model(path, "test-3l.en", wt, seed=seed, vocab_from=os.path.join(DATA_DIR, "for-tests-ggml-tiny.en.bin"), scale=lambda n: synth.conditioned(n, attn, 100.0))
pcm = synth.synth_audio(seed=500 + seed, seconds=60.0)
logs = []
for lib, is_ref in ((load_lib(), False), (load_ref(), True)):
    S = Side(lib, path, is_ref); L = S.L; V = S.n_vocab
    log = []
    def rec(c, st, toks, nt, logits, ud):
        x = np.ctypeslib.as_array(logits, (V,)).copy()
        top = np.argsort(-x)[:4]
        log.append((tuple(toks[k].id for k in range(nt)), [int(t) for t in top], [float(x[t]) for t in top], float(x[np.isfinite(x)].std())))
    cb = LOGITS_CB(rec)
    fp = L.whisper_full_default_params(strategy); fp.print_progress = False; fp.temperature_inc = 0.0; fp.greedy.best_of = 1; fp.n_threads = 1
    if strategy == 1: fp.beam_search.beam_size = 5
    fp.logits_filter_callback = C.cast(cb, vp)
    assert L.whisper_full(S.ctx, fp, pcm.ctypes.data_as(vp), len(pcm)) == 0
    toks = [L.whisper_full_get_token_id(S.ctx, i, j) for i in range(L.whisper_full_n_segments(S.ctx)) for j in range(L.whisper_full_n_tokens(S.ctx, i))]
    logs.append((log, toks)); S.free()
(la, ta), (lb, tb) = logs
print("calls", len(la), len(lb), "tokens", len(ta), len(tb))
for i, (a, b) in enumerate(zip(la, lb)):
    if a[0] != b[0] or a[1][0] != b[1][0]:
        print("first differing call", i)
        for k in range(max(0, i - 6), min(len(la), i + 3)):
            x, y = la[k], lb[k]
            print(k, "hist len", len(x[0]), len(y[0]), "same hist", x[0] == y[0], "last toks", x[0][-3:], y[0][-3:], "top ids", x[1], y[1], "top vals", [round(v, 4) for v in x[2]], [round(v, 4) for v in y[2]], "std %.2f" % y[3])
        break
else:
    print("all recorded calls identical")
k = 0
while k < min(len(ta), len(tb)) and ta[k] == tb[k]: k += 1
print("token prefix", k)
