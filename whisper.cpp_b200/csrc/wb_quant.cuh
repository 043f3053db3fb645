// wb_quant.cuh -- weight formats as they live in HBM, and the device-side decoders.
//
// File side (unchanged legacy ggml container): 32-value blocks block_q4_0 / block_q5_0 / block_q8_0
// (ggml/src/ggml-common.h:194-256) and 256-value super-blocks block_q4_K / block_q5_K
// (ggml-common.h:327-356).  Bit-level meaning follows ggml/src/ggml-quants.c:459-567 (dequantize_row_q4_0,
// _q5_0, _q8_0), :880-887 (get_scale_min_k4), :1529-1551 (q4_K), :1731-1756 (q5_K).
//
// HBM side ("planar" layout, built once at load by wb_repack_*): the 18/22/34-byte file blocks are not
// 16-byte aligned, so every 32-block matrix [N rows][K] is split into three dense arrays
//     qs : [N][K/32][QS]   QS = 16 bytes (4/5-bit) or 32 bytes (8-bit)   -> one or two 128-bit loads per block
//     qh : [N][K/32]       uint32, the fifth bits (Q5_0 only)
//     d  : [N][K/32]       f16 block scale
// Total bytes are identical to the file (18 / 22 / 34 per 32 weights).  K-quant super-blocks are 144 / 176
// bytes = multiples of 16 and are kept verbatim.
#pragma once
#include <cstdint>
#include <cuda_fp16.h>

namespace wb {

// numeric ids are ggml_type (ggml/include/ggml.h:390-405)
enum WType : int { WT_F32 = 0, WT_F16 = 1, WT_Q4_0 = 2, WT_Q5_0 = 6, WT_Q8_0 = 8, WT_Q4_K = 12, WT_Q5_K = 13 };

struct QMat {            // one weight matrix resident in HBM
    int   type = WT_F16;
    int   N = 0, K = 0;  // rows (output features), contraction length
    const void * base = nullptr; // F16: half[N][K];  K-quants: super-blocks [N][K/256]
    const uint8_t  * qs = nullptr;
    const uint32_t * qh = nullptr;
    const __half   * d  = nullptr;
    int   layout = 0;    // 0: rows (F16) / planar (32-blocks) / verbatim (K-quants); 1: tile-major records at `base` (see below)
    const __half * f16 = nullptr;  // encoder-side matrices of quantised models: f16 expansion [N][K] made once at load (TMA operand of the
                                   // persistent GEMM; each value is the f16 rounding of the exact d * (q - off)); nullptr = expand per launch
};

// "Tile-major" layout of the decoder matrices (persistent decode kernel, wb_decode_mk.cu).  The matrix is cut into tiles of
// 16 output rows; for every tile and every 32-value block (F16: every 16 values) there is ONE contiguous record holding the
// operands in the register order of the warp-level MMA (lane = 4*g + c reads rows g and g+8 of the tile):
//   Q4_0 (288 B): u32 qs[32 lanes][2] = { word c of row g, word c of row g+8 }            | f16 d[8][2]  = { row g, row g+8 }
//   Q5_0 (352 B): qs as Q4_0 (256 B) | u32 qh[8][2] = { row g, row g+8 } (64 B)            | f16 d[8][2]
//   Q8_0 (544 B): u32 qs[32 lanes][4] = { w c / row g, w c / row g+8, w 4+c / row g, w 4+c / row g+8 } | f16 d[8][2]
//   F16  (512 B per 16 values): u32 a[32 lanes][4] = the m16n8k16 A fragment
// so one warp fetches a record with three fully coalesced loads.  Same bytes per weight as the file; rows >= N are zero.
__host__ __device__ inline int wt_tm_rec_bytes(int t) { return t == WT_Q4_0 ? 288 : t == WT_Q5_0 ? 352 : t == WT_Q8_0 ? 544 : t == WT_F16 ? 512 : 0; }
__host__ __device__ inline int wt_tm_rec_k(int t)     { return t == WT_F16 ? 16 : 32; }
__host__ __device__ inline size_t wt_tm_bytes(int t, int N, int K) { return (size_t) ((N + 15) / 16) * (K / wt_tm_rec_k(t)) * wt_tm_rec_bytes(t); }

__host__ __device__ inline int wt_qs_bytes(int t)    { return t == WT_Q8_0 ? 32 : 16; }
__host__ __device__ inline bool wt_is_block32(int t) { return t == WT_Q4_0 || t == WT_Q5_0 || t == WT_Q8_0; }
__host__ __device__ inline bool wt_is_kquant(int t)  { return t == WT_Q4_K || t == WT_Q5_K; }
// bytes per weight on disk and in HBM
__host__ __device__ inline double wt_bpw(int t) {
    switch (t) { case WT_F32: return 4; case WT_F16: return 2; case WT_Q4_0: return 18.0/32; case WT_Q5_0: return 22.0/32;
                 case WT_Q8_0: return 34.0/32; case WT_Q4_K: return 144.0/256; case WT_Q5_K: return 176.0/256; }
    return 0;
}

#ifdef __CUDACC__

// ---- helpers ------------------------------------------------------------------------------------
// spread the low 4 bits of q to bit 4 of each byte of a word: bit i -> byte i
__device__ __forceinline__ uint32_t spread4_to_bit4(uint32_t q) {
    return (((q & 0xFu) * 0x00204081u) & 0x01010101u) << 4;
}
// 4 unsigned bytes -> two half2 holding (1024 + b) exactly (0x6400 | b); (v - (1024+off)) is exact in f16, and the
// single multiply by d rounds once: the result is the f16 rounding of the exact value d*(q - off).
// (A fused fma(v, d, -(1024+off)*d) would round the constant term and lose up to d/4 -- not used.)
__device__ __forceinline__ void bytes4_to_half2x2(uint32_t b, __half2 dd, __half2 off, uint32_t & lo, uint32_t & hi) {
    const uint32_t p0 = __byte_perm(b, 0x64646464u, 0x4140); // {b0, 0x64, b1, 0x64}
    const uint32_t p1 = __byte_perm(b, 0x64646464u, 0x4342); // {b2, 0x64, b3, 0x64}
    const __half2 h0 = __hmul2(__hsub2(*reinterpret_cast<const __half2 *>(&p0), off), dd);
    const __half2 h1 = __hmul2(__hsub2(*reinterpret_cast<const __half2 *>(&p1), off), dd);
    lo = *reinterpret_cast<const uint32_t *>(&h0);
    hi = *reinterpret_cast<const uint32_t *>(&h1);
}

// ---- decode one 32-block into 32 halves (out[0..3] = 4 x 16 bytes, element order 0..31) ---------
// Rounding: each output is the f16 rounding of the exact product d*(q - off), i.e. exactly what the reference's
// dequantize_row_* (f32, exact) gives after one f16 store.
__device__ __forceinline__ void dequant_q4_0(const uint4 qs, const __half d, uint4 (&out)[4]) {
    const __half2 dd = __half2half2(d);
    const __half2 no = __float2half2_rn(1032.0f);               // 1024 + 8
    const uint32_t w[4] = { qs.x, qs.y, qs.z, qs.w };
    uint32_t lo[8], hi[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        bytes4_to_half2x2( w[i]       & 0x0F0F0F0Fu, dd, no, lo[2*i], lo[2*i+1]);   // elements 4i..4i+3
        bytes4_to_half2x2((w[i] >> 4) & 0x0F0F0F0Fu, dd, no, hi[2*i], hi[2*i+1]);   // elements 16+4i..
    }
    out[0] = make_uint4(lo[0], lo[1], lo[2], lo[3]); out[1] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
    out[2] = make_uint4(hi[0], hi[1], hi[2], hi[3]); out[3] = make_uint4(hi[4], hi[5], hi[6], hi[7]);
}
__device__ __forceinline__ void dequant_q5_0(const uint4 qs, const uint32_t qh, const __half d, uint4 (&out)[4]) {
    const __half2 dd = __half2half2(d);
    const __half2 no = __float2half2_rn(1040.0f);               // 1024 + 16
    const uint32_t w[4] = { qs.x, qs.y, qs.z, qs.w };
    uint32_t lo[8], hi[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t bl = ( w[i]       & 0x0F0F0F0Fu) | spread4_to_bit4(qh >> (4*i));
        const uint32_t bh = ((w[i] >> 4) & 0x0F0F0F0Fu) | spread4_to_bit4(qh >> (16 + 4*i));
        bytes4_to_half2x2(bl, dd, no, lo[2*i], lo[2*i+1]);
        bytes4_to_half2x2(bh, dd, no, hi[2*i], hi[2*i+1]);
    }
    out[0] = make_uint4(lo[0], lo[1], lo[2], lo[3]); out[1] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
    out[2] = make_uint4(hi[0], hi[1], hi[2], hi[3]); out[3] = make_uint4(hi[4], hi[5], hi[6], hi[7]);
}
__device__ __forceinline__ void dequant_q8_0(const uint4 q0, const uint4 q1, const __half d, uint4 (&out)[4]) {
    const __half2 dd = __half2half2(d);
    const __half2 no = __float2half2_rn(1152.0f);               // 1024 + 128 ; bytes are biased by ^0x80
    const uint32_t w[8] = { q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w };
    uint32_t h[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        // (b ^ 0x80) is b + 128 as unsigned; split in two 0..255 bytes needs 0x6400|b: b < 256 occupies 8 mantissa bits: exact
        bytes4_to_half2x2(w[i] ^ 0x80808080u, dd, no, h[2*i], h[2*i+1]);
    }
    out[0] = make_uint4(h[0], h[1], h[2], h[3]);   out[1] = make_uint4(h[4], h[5], h[6], h[7]);
    out[2] = make_uint4(h[8], h[9], h[10], h[11]); out[3] = make_uint4(h[12], h[13], h[14], h[15]);
}

// 6-bit scale / min of sub-block j (0..7) from the 12 packed bytes (ggml-quants.c:880-887)
__device__ __forceinline__ void kq_scale_min(int j, const uint8_t * q, int & sc, int & m) {
    if (j < 4) { sc = q[j] & 63; m = q[j + 4] & 63; }
    else       { sc = (q[j + 4] & 0xF) | ((q[j - 4] >> 6) << 4); m = (q[j + 4] >> 4) | ((q[j] >> 6) << 4); }
}

// decode sub-block j (32 values) of a Q4_K / Q5_K super-block into 32 halves.
// y = d*sc*q - dmin*m  evaluated in f32 then rounded once to f16.
template <bool Q5>
__device__ __forceinline__ void dequant_kq_sub(const uint8_t * __restrict__ blk, int j, uint4 (&out)[4]) {
    const uint4 hdr = *reinterpret_cast<const uint4 *>(blk);          // d, dmin, scales[12]
    const __half2 dm = *reinterpret_cast<const __half2 *>(&hdr.x);
    const uint32_t sw[3] = { hdr.y, hdr.z, hdr.w };
    const uint8_t * sc8 = reinterpret_cast<const uint8_t *>(sw);
    int sc, mn; kq_scale_min(j, sc8, sc, mn);
    const float dl = __low2float(dm) * (float) sc;
    const float ml = __high2float(dm) * (float) mn;
    const uint8_t * qs = blk + 16 + (Q5 ? 32 : 0) + 32 * (j >> 1);
    const uint8_t * qh = blk + 16;
    const int sh = 4 * (j & 1);
    uint32_t h[16];
#pragma unroll
    for (int v = 0; v < 2; ++v) {
        const uint4 q4 = *reinterpret_cast<const uint4 *>(qs + 16 * v);
        uint4 hb = make_uint4(0, 0, 0, 0);
        if (Q5) hb = *reinterpret_cast<const uint4 *>(qh + 16 * v);
        const uint32_t w[4]  = { q4.x, q4.y, q4.z, q4.w };
        const uint32_t hw[4] = { hb.x, hb.y, hb.z, hb.w };
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            uint32_t b = (w[i] >> sh) & 0x0F0F0F0Fu;
            if (Q5) b |= ((hw[i] >> j) & 0x01010101u) << 4;
            float f[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) f[e] = dl * (float) ((b >> (8*e)) & 0xFF) - ml;
            const __half2 a = __floats2half2_rn(f[0], f[1]);
            const __half2 c = __floats2half2_rn(f[2], f[3]);
            h[8*v + 2*i]     = *reinterpret_cast<const uint32_t *>(&a);
            h[8*v + 2*i + 1] = *reinterpret_cast<const uint32_t *>(&c);
        }
    }
    out[0] = make_uint4(h[0], h[1], h[2], h[3]);   out[1] = make_uint4(h[4], h[5], h[6], h[7]);
    out[2] = make_uint4(h[8], h[9], h[10], h[11]); out[3] = make_uint4(h[12], h[13], h[14], h[15]);
}

#endif // __CUDACC__
} // namespace wb
