"""Synthetic Silero-VAD model files in the reference's ggml container (layout: models/convert-silero-vad-to-ggml.py:31-185 of the
reference): real Hann-windowed DFT basis, seeded random conv / LSTM / output weights.  TEST INFRASTRUCTURE."""
import struct
import numpy as np


def write_vad_model(path, seed=0, gain=1.0, drop=None, window=512, n_layers=4):
    rng = np.random.default_rng(seed)
    cin, cout = [129, 128, 64, 64], [128, 64, 64, 128]
    k = np.arange(256)
    hann = 0.5 - 0.5 * np.cos(2 * np.pi * k / 256)
    c = np.arange(129)[:, None]
    basis = np.concatenate([np.cos(2 * np.pi * c * k / 256) * hann, -np.sin(2 * np.pi * c * k / 256) * hann]).astype(np.float32)   # [258][256]
    tensors = []
    for i in range(4):
        w = rng.standard_normal((cout[i], cin[i], 3)) * gain * (2.0 / (cin[i] * 3)) ** 0.5
        if i == 0:
            w *= 0.5                                                                  # magnitudes of |x| ~ 0.1 audio are O(1..10)
        tensors.append((f"_model.encoder.{i}.reparam_conv.weight", w.astype(np.float16)))
        tensors.append((f"_model.encoder.{i}.reparam_conv.bias", (rng.standard_normal(cout[i]) * 0.1).astype(np.float32)))
    s = gain * 2.0 / 128 ** 0.5
    tensors.append(("_model.decoder.rnn.weight_ih", rng.uniform(-s, s, (512, 128)).astype(np.float32)))
    tensors.append(("_model.decoder.rnn.weight_hh", rng.uniform(-s, s, (512, 128)).astype(np.float32)))
    tensors.append(("_model.decoder.rnn.bias_ih", rng.uniform(-s, s, 512).astype(np.float32)))
    tensors.append(("_model.decoder.rnn.bias_hh", rng.uniform(-s, s, 512).astype(np.float32)))
    tensors.append(("_model.decoder.decoder.2.weight", (rng.standard_normal(128) * 1.5).astype(np.float16)))
    tensors.append(("_model.decoder.decoder.2.bias", np.zeros((), np.float32)))
    tensors.append(("_model.stft.forward_basis_buffer", basis.astype(np.float16).reshape(258, 1, 256)))
    with open(path, "wb") as f:
        f.write(struct.pack("<i", 0x67676d6c))
        mt = b"silero-16k"
        f.write(struct.pack("<i", len(mt))); f.write(mt)
        f.write(struct.pack("<iii", 5, 1, 2))
        f.write(struct.pack("<ii", window, 64))
        f.write(struct.pack("<i", n_layers))
        for i in range(n_layers):
            f.write(struct.pack("<iii", cin[i % 4], cout[i % 4], 3))
        f.write(struct.pack("<iiii", 128, 128, 128, 1))
        for name, a in tensors:
            if drop and name in drop:
                continue
            shape = list(a.shape)[::-1] if name != "_model.stft.forward_basis_buffer" else [256, 1, 258]
            nb = name.encode()
            f.write(struct.pack("<iii", len(shape), len(nb), 1 if a.dtype == np.float16 else 0))
            for d in shape:
                f.write(struct.pack("<i", d))
            f.write(nb)
            f.write(a.tobytes())
    return path


def speechy_audio(seconds, seed=0):
    """bursts of modulated harmonics separated by near-silence (so that per-window features really differ)"""
    rng = np.random.default_rng(seed)
    n = int(seconds * 16000)
    t = np.arange(n) / 16000.0
    x = np.zeros(n, np.float32)
    pos = 0.0
    while pos < seconds:
        dur = rng.uniform(0.2, 1.6); gap = rng.uniform(0.05, 1.2)
        a, b = int(pos * 16000), min(n, int((pos + dur) * 16000))
        f0 = rng.uniform(90, 260)
        seg = sum(np.sin(2 * np.pi * f0 * h * t[a:b] + rng.uniform(0, 6.28)) / h for h in range(1, 9))
        env = np.sin(np.pi * np.linspace(0, 1, max(b - a, 1))) ** 0.5
        x[a:b] += (0.12 * seg * env).astype(np.float32)
        pos += dur + gap
    x += (rng.standard_normal(n) * 0.003).astype(np.float32)
    return x
