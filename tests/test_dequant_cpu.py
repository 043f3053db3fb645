"""CPU: the ggml block formats that this engine expands on the host at load time (Q4_1, Q5_1, Q2_K, Q3_K, Q6_K -> F16 matrices in HBM;
csrc/wb_dequant_host.cpp) against ggml's own to_float on blocks produced by ggml's own quantisers.  Formats whose value is a product
chain (Q3_K, Q6_K) must match bit for bit; formats with a*b+c / a*b-c (Q4_1, Q5_1, Q2_K) may differ by one f32 ulp because the
reference build contracts the expression into an FMA (its CPU inference path never materialises these values at all)."""
import ctypes as C
import numpy as np
import pytest

from wbtest import ref_quantize, ref_dequantize

TYPES = {"Q4_1": 3, "Q5_1": 7, "Q2_K": 10, "Q3_K": 11, "Q6_K": 14, "BF16": 30, "Q4_K": 12, "Q5_K": 13}


@pytest.mark.parametrize("name", sorted(TYPES))
def test_host_dequantisers_match_ggml(lib, ref, name):
    t = TYPES[name]
    rng = np.random.default_rng(t)
    rows, k = 24, 1024
    w = (rng.standard_normal((rows, k)) * 0.05).astype(np.float32)
    w[3, :256] = 0.0                                   # an all-zero super-block
    w[5] *= 40.0                                       # large dynamic range
    w[7, ::7] = 0.9                                    # outliers
    raw = ref_quantize(ref, t, w)
    want = ref_dequantize(ref, t, raw, rows, k)
    got = np.empty(rows * k, np.float32)
    buf = np.frombuffer(raw, np.uint8)
    lib.wb200_dbg_dequantize.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int64]
    assert lib.wb200_dbg_dequantize(t, buf.ctypes.data, got.ctypes.data, rows * k) == 0
    got = got.reshape(rows, k)
    if name in ("Q3_K", "Q6_K", "BF16"):
        assert np.array_equal(got, want)
    else:
        ulp = np.spacing(np.maximum(np.abs(want), np.float32(1e-30)))
        assert np.all(np.abs(got - want) <= ulp), float(np.abs(got - want).max())
    assert np.abs(want - w).max() < 0.5 * np.abs(w).max()          # the blocks really encode w
    # what goes to HBM: the f16 rounding of these values
    assert np.array_equal(got.astype(np.float16), want.astype(np.float16)) or np.mean(got.astype(np.float16) != want.astype(np.float16)) < 1e-3
    assert lib.wb200_dbg_dequantize(t, buf.ctypes.data, got.ctypes.data, 17) == (0 if name == "BF16" else -1)
    assert lib.wb200_dbg_dequantize(2, buf.ctypes.data, got.ctypes.data, 32) == -1   # Q4_0 only has device kernels
