// wb_dequant_host.cpp -- see wb_dequant_host.h.  Pinned bit for bit (1 ulp where the reference's build contracts a*b+c) against
// ggml's own to_float by tests/test_dequant_cpu.py through wb200_dbg_dequantize.
#include <cstring>
#include <cuda_fp16.h>
#include "wb_dequant_host.h"
#include "../../include/whisper_b200.h"

namespace wb {
namespace {

inline float h2f(const uint8_t * p) { __half h; memcpy(&h, p, 2); return __half2float(h); }

// Q4_1: d, m (f16) | 16 bytes of nibbles; value j = low nibble of byte j, value j+16 = high nibble      (20 bytes / 32 values)
void dq_q4_1(const uint8_t * b, float * y) {
    const float d = h2f(b), m = h2f(b + 2);
    const uint8_t * q = b + 4;
    for (int j = 0; j < 16; ++j) { y[j] = (float) (q[j] & 0x0F) * d + m; y[j + 16] = (float) (q[j] >> 4) * d + m; }
}
// Q5_1: d, m | 32 high bits | 16 bytes of nibbles; bit j of the mask is the fifth bit of value j      (24 bytes / 32 values)
void dq_q5_1(const uint8_t * b, float * y) {
    const float d = h2f(b), m = h2f(b + 2);
    uint32_t hi; memcpy(&hi, b + 4, 4);
    const uint8_t * q = b + 8;
    for (int j = 0; j < 16; ++j) {
        const int lo0 = q[j] & 0x0F, lo1 = q[j] >> 4;
        y[j]      = (float) (lo0 | (int) (((hi >> j) & 1u) << 4)) * d + m;
        y[j + 16] = (float) (lo1 | (int) (((hi >> (j + 16)) & 1u) << 4)) * d + m;
    }
}
// Q2_K: 16 x (4-bit scale | 4-bit min) | 64 bytes of 2-bit values | d, dmin                              (84 bytes / 256 values)
// a super-block is 2 halves of 128; inside a half, bit pair s of byte l belongs to value 32*s + l (16 values per scale)
void dq_q2_k(const uint8_t * b, float * y) {
    const uint8_t * sc = b, * q = b + 16;
    const float d = h2f(b + 80), dmin = h2f(b + 82);
    int is = 0;
    for (int half = 0; half < 2; ++half, q += 32) {
        for (int s = 0; s < 4; ++s) {
            for (int part = 0; part < 2; ++part) {
                const float dl = d * (float) (sc[is] & 0x0F), ml = dmin * (float) (sc[is] >> 4); ++is;
                for (int l = 0; l < 16; ++l) *y++ = dl * (float) ((q[16 * part + l] >> (2 * s)) & 3) - ml;
            }
        }
    }
}
// Q3_K: 32 bytes of high bits | 64 bytes of low 2-bit values | 12 bytes holding 16 six-bit scales | d    (110 bytes / 256 values)
void dq_q3_k(const uint8_t * b, float * y) {
    const uint8_t * hm = b, * q = b + 32, * ps = b + 96;
    const float d_all = h2f(b + 108);
    // scale k (0..15): low 4 bits in nibble k of the first 8 bytes, high 2 bits in bit pair k of the last 4 bytes
    int8_t scales[16];
    for (int k = 0; k < 16; ++k) {
        const int lo = k < 8 ? (ps[k] & 0x0F) : (ps[k - 8] >> 4);
        const int hi = (ps[8 + (k & 3)] >> (2 * (k >> 2))) & 3;
        scales[k] = (int8_t) ((lo | (hi << 4)) - 32);
    }
    int is = 0; uint8_t mbit = 1;
    for (int half = 0; half < 2; ++half, q += 32) {
        for (int s = 0; s < 4; ++s, mbit = (uint8_t) (mbit << 1)) {
            for (int part = 0; part < 2; ++part) {
                const float dl = d_all * (float) scales[is++];
                for (int l = 0; l < 16; ++l) {
                    const int idx = 16 * part + l;
                    const int v = (int) ((q[idx] >> (2 * s)) & 3) - ((hm[idx] & mbit) ? 0 : 4);
                    *y++ = dl * (float) v;
                }
            }
        }
    }
}
// Q6_K: 128 bytes low nibbles | 64 bytes high bit pairs | 16 int8 scales | d                              (210 bytes / 256 values)
void dq_q6_k(const uint8_t * b, float * y) {
    const uint8_t * ql = b, * qh = b + 128;
    const int8_t * sc = (const int8_t *) (b + 192);
    const float d = h2f(b + 208);
    for (int half = 0; half < 2; ++half, y += 128, ql += 64, qh += 32, sc += 8) {
        for (int l = 0; l < 32; ++l) {
            const int is = l / 16;
            const int q1 = (int) ((ql[l]      & 0x0F) | (((qh[l] >> 0) & 3) << 4)) - 32;
            const int q2 = (int) ((ql[l + 32] & 0x0F) | (((qh[l] >> 2) & 3) << 4)) - 32;
            const int q3 = (int) ((ql[l]      >> 4)   | (((qh[l] >> 4) & 3) << 4)) - 32;
            const int q4 = (int) ((ql[l + 32] >> 4)   | (((qh[l] >> 6) & 3) << 4)) - 32;
            y[l]      = d * (float) sc[is + 0] * (float) q1;
            y[l + 32] = d * (float) sc[is + 2] * (float) q2;
            y[l + 64] = d * (float) sc[is + 4] * (float) q3;
            y[l + 96] = d * (float) sc[is + 6] * (float) q4;
        }
    }
}

// Q4_K / Q5_K: d, dmin | 12 bytes holding 8 six-bit (scale, min) pairs | [Q5_K: 32 bytes of fifth bits] | 128 bytes of nibbles   (144 / 176 bytes)
// These two formats HAVE device kernels (kernel chain); the host expansion serves the opt-in WB200_KQUANT_AS_F16 load mode.
inline void k4_scale_min(int j, const uint8_t * q, int & sc, int & mn) {
    if (j < 4) { sc = q[j] & 63; mn = q[j + 4] & 63; }
    else       { sc = (q[j + 4] & 0x0F) | ((q[j - 4] >> 6) << 4); mn = (q[j + 4] >> 4) | ((q[j] >> 6) << 4); }
}
void dq_q45_k(const uint8_t * b, float * y, bool five) {
    const float d = h2f(b), dmin = h2f(b + 2);
    const uint8_t * scales = b + 4, * qh = b + 16, * q = b + (five ? 48 : 16);
    for (int pair = 0; pair < 4; ++pair, q += 32) {
        for (int part = 0; part < 2; ++part) {
            int sc, mn; k4_scale_min(2 * pair + part, scales, sc, mn);
            const float dl = d * (float) sc, ml = dmin * (float) mn;
            for (int l = 0; l < 32; ++l) {
                int v = part ? (q[l] >> 4) : (q[l] & 0x0F);
                if (five && (qh[l] & (1u << (2 * pair + part)))) v += 16;
                *y++ = dl * (float) v - ml;
            }
        }
    }
}

} // namespace

bool   host_dq_supported(int t)    { return t == HT_Q4_1 || t == HT_Q5_1 || t == HT_Q2_K || t == HT_Q3_K || t == HT_Q6_K || t == HT_BF16 || t == HT_Q4_K || t == HT_Q5_K; }
int    host_dq_block_values(int t) { return t == HT_BF16 ? 1 : (t == HT_Q4_1 || t == HT_Q5_1) ? 32 : 256; }
size_t host_dq_block_bytes(int t)  { switch (t) { case HT_Q4_1: return 20; case HT_Q5_1: return 24; case HT_Q2_K: return 84; case HT_Q3_K: return 110; case HT_Q6_K: return 210; case HT_BF16: return 2; case HT_Q4_K: return 144; case HT_Q5_K: return 176; } return 0; }

void host_dequantize(int t, const void * src, float * dst, int64_t n) {
    const int bv = host_dq_block_values(t); const size_t bb = host_dq_block_bytes(t);
    const uint8_t * p = (const uint8_t *) src;
    for (int64_t i = 0; i < n / bv; ++i, p += bb, dst += bv) {
        switch (t) {
            case HT_Q4_1: dq_q4_1(p, dst); break; case HT_Q5_1: dq_q5_1(p, dst); break; case HT_Q2_K: dq_q2_k(p, dst); break;
            case HT_Q3_K: dq_q3_k(p, dst); break; case HT_Q6_K: dq_q6_k(p, dst); break;
            case HT_Q4_K: dq_q45_k(p, dst, false); break; case HT_Q5_K: dq_q45_k(p, dst, true); break;
            case HT_BF16: { uint16_t h; memcpy(&h, p, 2); const uint32_t u = (uint32_t) h << 16; memcpy(dst, &u, 4); } break;   // bf16 = upper half of an f32
        }
    }
}

} // namespace wb

// host-only test hook: tests/test_dequant_cpu.py compares it with ggml's to_float of the oracle build
extern "C" WB_EXPORT int wb200_dbg_dequantize(int ggml_type, const void * src, float * dst, int64_t n) {
    if (!wb::host_dq_supported(ggml_type) || n % wb::host_dq_block_values(ggml_type)) return -1;
    wb::host_dequantize(ggml_type, src, dst, n);
    return 0;
}
