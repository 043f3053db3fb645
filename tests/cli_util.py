"""Drive the reference's unmodified whisper-cli (oracle/_ref/whisper-cli-*) and compare what it writes with the same run through the C ABI.
TEST INFRASTRUCTURE shared by tests/test_cli_cpu.py (reference CLI on the reference library, on the CPU) and
tests/test_zz_reference_cli_gpu.py (the same CLI sources linked against libwhisper_b200.so, on the GPU)."""
import ctypes as C
import json
import os
import subprocess
import numpy as np

from wbtest import ROOT, DATA_DIR, TokenData, read_wav_f32
from e2e_util import Side

vp = C.c_void_p
SILERO = os.path.join(DATA_DIR, "for-tests-silero-v6.2.0-ggml.bin")


def _needs_more(b):
    """utf8_trailing_bytes_needed of the reference's examples/common-whisper.cpp:201-227"""
    i = len(b) - 1
    while i >= 0 and (b[i] & 0xC0) == 0x80:
        i -= 1
    if i < 0:
        return 0
    c = b[i]
    expected = 1 if c & 0x80 == 0 else 2 if c & 0xE0 == 0xC0 else 3 if c & 0xF0 == 0xE0 else 4 if c & 0xF8 == 0xF0 else 0
    return max(0, expected - (len(b) - i)) if expected else 0


def _merge(toks):
    """cli.cpp:760-780 writes one JSON entry per run of tokens that completes a UTF-8 sequence, with the id / p of the first one"""
    out = []; j = 0
    while j < len(toks):
        tid, p, text = toks[j]; j += 1
        while j < len(toks) and _needs_more(text) > 0:
            text += toks[j][2]; j += 1
        out.append((tid, p))
    return out


def api_run(lib, path, pcm, vad, is_ref):
    A = Side(lib, path, is_ref)
    try:
        L = A.L
        L.whisper_full_get_token_data.restype = TokenData
        L.whisper_full_get_token_data.argtypes = [vp, C.c_int, C.c_int]
        fp = L.whisper_full_default_params(0)                       # what cli.cpp sets for: -bs 1 -bo 1 -nf -ojf
        fp.print_progress = False; fp.print_realtime = False; fp.n_threads = 4
        fp.greedy.best_of = 1; fp.beam_search.beam_size = 1; fp.temperature_inc = 0.0
        fp.token_timestamps = True; fp.initial_prompt = b""; fp.language = b"en"
        if vad:
            fp.vad = True; fp.vad_model_path = SILERO.encode()
        assert L.whisper_full(A.ctx, fp, pcm.ctypes.data_as(vp), len(pcm)) == 0
        eot = L.whisper_token_eot(A.ctx)
        segs = []
        for i in range(L.whisper_full_n_segments(A.ctx)):
            toks = [L.whisper_full_get_token_data(A.ctx, i, j) for j in range(L.whisper_full_n_tokens(A.ctx, i))]
            segs.append((L.whisper_full_get_segment_t0(A.ctx, i) * 10, L.whisper_full_get_segment_t1(A.ctx, i) * 10,
                         L.whisper_full_get_segment_text(A.ctx, i).decode("utf-8", "replace"), _merge([(t.id, t.p, L.whisper_token_to_str(A.ctx, t.id)) for t in toks])))
        return segs, eot
    finally:
        A.free()




def check_cli_against_api(cli, lib, is_ref, model_path, tmp_path, vad):
    lib.whisper_full_get_segment_t0.restype = C.c_int64; lib.whisper_full_get_segment_t1.restype = C.c_int64
    lib.whisper_token_to_str.restype = C.c_char_p; lib.whisper_token_to_str.argtypes = [vp, C.c_int]
    wav = os.path.join(DATA_DIR, "jfk.wav")
    out = str(tmp_path / "cli_out")
    cmd = [cli, "-m", model_path, "-f", wav, "-bs", "1", "-bo", "1", "-nf", "-np", "-ojf", "-of", out]
    if vad:
        cmd += ["--vad", "--vad-model", SILERO]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, (r.returncode, r.stderr[-800:])
    doc = json.loads(open(out + ".json", encoding="utf-8", errors="replace").read(), strict=False)   # token text may hold raw control characters
    got = doc["transcription"]
    assert len(got) >= 1
    pcm = np.ascontiguousarray(read_wav_f32(wav), np.float32)
    api, eot = api_run(lib, model_path, pcm, vad, is_ref)
    assert len(got) == len(api), (len(got), len(api))
    for c, a in zip(got, api):
        assert (c["offsets"]["from"], c["offsets"]["to"]) == (a[0], a[1])
        assert c["text"] == a[2]
        assert [t["id"] for t in c["tokens"]] == [t[0] for t in a[3]]
        assert np.allclose([t["p"] for t in c["tokens"]], [t[1] for t in a[3]], rtol=2e-5, atol=2e-6)
    if vad:
        assert all(0 <= c["offsets"]["from"] <= c["offsets"]["to"] <= 11100 for c in got)      # times on the ORIGINAL timeline of the 11 s clip
    return got


def run_reference_bench(exe, model_path, threads=4):
    """the reference's examples/bench/bench.cpp: returns {name: (total ms, runs, ms per run)} parsed from whisper_print_timings"""
    import re
    r = subprocess.run([exe, "-m", model_path, "-t", str(threads)], capture_output=True, text=True, timeout=1800)
    assert r.returncode == 0, (r.returncode, r.stderr[-800:])
    out = {}
    for name, total, runs, per in re.findall(r"(\w+) time =\s*([\d.]+) ms /\s*(\d+) runs \(\s*([\d.]+) ms per run\)", r.stderr + r.stdout):
        out[name] = (float(total), int(runs), float(per))
    # what bench.cpp:120-150 executes after whisper_reset_timings: 1 encode, 256 single-token decodes, 64 batches of 5, 16 prompts of 256
    # (batched and prompt "runs" count tokens, src/whisper.cpp:2974-2983)
    assert {k: v[1] for k, v in out.items() if k in ("encode", "decode", "batchd", "prompt")} == {"encode": 1, "decode": 256, "batchd": 64 * 5, "prompt": 16 * 256}, out
    assert all(out[k][0] > 0 for k in ("encode", "decode", "batchd", "prompt"))
    return out


def run_reference_vad_example(exe):
    """examples/vad-speech-segments of the reference on jfk.wav with the Silero fixture: [(start cs, end cs)] it prints"""
    import re
    r = subprocess.run([exe, "-vm", SILERO, "-f", os.path.join(DATA_DIR, "jfk.wav"), "-np"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.returncode, r.stderr[-800:])
    return [(float(a), float(b)) for a, b in re.findall(r"Speech segment \d+: start = ([\d.]+), end = ([\d.]+)", r.stdout)]


# ---- the reference's HTTP server (examples/server/server.cpp) -----------------------------------------------------------------
def _free_port():
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def check_server_against_api(exe, lib, is_ref, model_path):
    """start the server on 127.0.0.1, POST samples/jfk.wav to /inference (greedy, no fallback) as json and verbose_json, and compare with
    the same parameters through the C ABI: texts, segment times (verbose_json turns token timestamps and the 60-character wrap on,
    server.cpp:627-637, 941-942) and token ids."""
    import time
    import requests
    lib.whisper_full_get_segment_t0.restype = C.c_int64; lib.whisper_full_get_segment_t1.restype = C.c_int64
    wav = os.path.join(DATA_DIR, "jfk.wav")
    port = _free_port()
    proc = subprocess.Popen([exe, "-m", model_path, "--host", "127.0.0.1", "--port", str(port), "-t", "4"], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    try:
        base = "http://127.0.0.1:%d" % port
        for _ in range(600):                                        # model load + first CUDA initialisation
            if proc.poll() is not None:
                raise AssertionError("server exited: %s" % proc.stderr.read().decode("utf-8", "replace")[-800:])
            try:
                if requests.get(base + "/", timeout=1).status_code == 200:
                    break
            except requests.RequestException:
                time.sleep(0.2)
        else:
            raise AssertionError("server did not come up")
        common = {"temperature": "0.0", "temperature_inc": "0.0", "best_of": "1"}
        out = {}
        for fmt in ("json", "verbose_json"):
            with open(wav, "rb") as f:
                r = requests.post(base + "/inference", files={"file": ("jfk.wav", f, "audio/wav")}, data=dict(common, response_format=fmt), timeout=900)
            assert r.status_code == 200, (r.status_code, r.text[:400])
            out[fmt] = json.loads(r.content.decode("utf-8", "replace"), strict=False)
    finally:
        proc.terminate()
        try:
            proc.wait(timeout=20)
        except subprocess.TimeoutExpired:
            proc.kill(); proc.wait()

    pcm = np.ascontiguousarray(read_wav_f32(wav), np.float32)

    def api(token_ts):
        A = Side(lib, model_path, is_ref)
        try:
            L = A.L
            L.whisper_full_get_token_data.restype = TokenData; L.whisper_full_get_token_data.argtypes = [vp, C.c_int, C.c_int]
            fp = L.whisper_full_default_params(0)                   # server.cpp:925-965 with the request above
            fp.print_progress = False; fp.print_realtime = False; fp.n_threads = 4
            fp.greedy.best_of = 1; fp.beam_search.beam_size = -1; fp.temperature = 0.0; fp.temperature_inc = 0.0
            fp.initial_prompt = b""; fp.language = b"en"; fp.max_len = 60; fp.token_timestamps = token_ts; fp.no_context = True
            assert L.whisper_full(A.ctx, fp, pcm.ctypes.data_as(vp), len(pcm)) == 0
            eot = L.whisper_token_eot(A.ctx)
            segs = []
            for i in range(L.whisper_full_n_segments(A.ctx)):
                ids = [L.whisper_full_get_token_data(A.ctx, i, j).id for j in range(L.whisper_full_n_tokens(A.ctx, i))]
                segs.append((L.whisper_full_get_segment_text(A.ctx, i).decode("utf-8", "replace"), L.whisper_full_get_segment_t0(A.ctx, i),
                             L.whisper_full_get_segment_t1(A.ctx, i), [t for t in ids if t < eot]))
            return segs
        finally:
            A.free()

    plain = api(False)
    assert out["json"]["text"] == "".join(s[0] + "\n" for s in plain)
    wrapped = api(True)
    vs = out["verbose_json"]["segments"]
    assert len(vs) == len(wrapped), (len(vs), len(wrapped))
    for v, a in zip(vs, wrapped):
        assert v["text"] == a[0]
        assert abs(v["start"] - a[1] * 0.01) < 1e-6 and abs(v["end"] - a[2] * 0.01) < 1e-6
        assert v.get("tokens", []) == a[3]
    return out
