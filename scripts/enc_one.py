"""driver for profiler runs: two encodes of one 30 s window on the large-v3 Q5_0 benchmark model (whisper.h calls only)"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_pkg
pkg = load_pkg(); synth = pkg.synth
model = synth.cached_model("large-v3", synth.Q5_0, seed=0, fast_pool=True)
eng = pkg.WhisperB200(model, gpu_device=0)
L = eng.L
pcm = synth.synth_audio(seed=1234, seconds=30.0)
assert L.whisper_pcm_to_mel(eng.ctx, pcm.ctypes.data_as(C.c_void_p), len(pcm), 1) == 0
for _ in range(int(os.environ.get("N_ENC", "2"))):
    assert L.whisper_encode(eng.ctx, 0, 1) == 0
print("ok", flush=True)
eng.close()
