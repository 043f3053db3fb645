// wb_decode_mk.cu -- one decoder pass (<= 8 rows) as a single persistent cooperative kernel.
//
// What it computes is whisper_build_graph_decoder (src/whisper.cpp:2466-2844) for the rows of one whisper_batch:
//   per layer:  LN -> QKV (+KV append) -> self-attention over the paged cache -> O + residual
//               LN -> Q -> cross-attention over the padded encoder keys       -> O + residual
//               LN -> FC1 + GELU -> FC2 + residual
//   then LN -> logits (token-embedding matrix).
// The arithmetic per op is the one of the stand-alone kernels in wb_kernels.cu (Q8_0 activation blocks + integer block
// dots like the reference CPU path, f16-rounded Q, f32 softmax); this file changes HOW the ops are scheduled:
//   * grid = one CTA of 16 warps per SM, launched cooperatively; a phase boundary is a grid barrier instead of a kernel
//     boundary (the decode step is latency-bound: ~270 dependent phases of a few microseconds each);
//   * every CTA rebuilds the (tiny) quantised activation vector of a GEMV phase in its own shared memory, so LayerNorm and
//     activation quantisation need no phase of their own; attention and FC1 hand their outputs over already quantised;
//   * a GEMV phase hands 16-row tiles round-robin to the CTAs; the warps of a CTA split K and reduce through smem.  Weights
//     are stored tile-major (wb_quant.cuh): a warp fetches the MMA operands of one (tile, block) with three fully coalesced
//     loads -- the planar layout cost 16 sectors per 128 useful bytes and made the phase L1-sector bound;
//   * cross-attention (the HBM-heavy part: 2*n_keys*d f16 per row and layer) is cut into (row, head, 128-key) units that are
//     distributed stream-K style, register-prefetched one unit ahead.
// The kernel is written for a small instruction footprint: the ~20 phase bodies of a layer are executed once per layer by
// every warp, so code that does not fit the instruction cache is fetched from L2 again in every layer (measured: 25 cycles
// per instruction with a 160 KB kernel).  Hence: no 64-bit divisions, approximate reciprocals where they only feed a
// rounding, shared (noinline) phase functions, modest unrolling.
#include <cmath>
#include "wb_decode_mk.cuh"
#include "wb_common.h"
#include "wb_dev.cuh"

namespace wb {

constexpr int MK_THREADS = 512, MK_WARPS = 16, MK_XKEYS = 128, MK_RED = 16 * 16 * 9, MK_PART = 68, MK_XSLOTS = 16;

struct MkSm {
    uint32_t * xq;      // quantised activations of the current GEMV phase: [8][SW] words (int8x4 or half2); q staging in attention phases
    float * xd;         // Q8_0 block scales [8][K/32]
    float * red;        // [2][16 warps][16 rows][9]  split-K partials (double buffered)
    float * part;       // [2][16 warps][68]          attention warp partials: m, l, -, -, o[64]
    float * stat;       // [32] LayerNorm partial sums
    int   * flag;       // [4]
    int SW;
};

#define MK_STAMP() do { if (TRACE) { if (blockIdx.x == 0 && threadIdx.x == 0) a.trace[n_stamp] = clock64(); ++n_stamp; } } while (0)

__device__ __forceinline__ void l2_prefetch(const void * p, uint32_t bytes) {
    if (bytes >= 16) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(p), "r"(bytes & ~15u) : "memory");
}
__device__ __forceinline__ void bar_named(int id, int n) { asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(n) : "memory"); }

// Grid barrier.  Arrival = one atomic on a counter; the LAST CTA to arrive releases everybody by writing one flag per CTA
// (each on its own 128-byte line), and every other CTA polls only its own flag.
__device__ __noinline__ void mk_grid_sync(const MkArgs & a, MkSm & sm, unsigned long long target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned long long old = atomicAdd(a.bar, 1ULL);
        __threadfence();
        sm.flag[2] = (old + 1 == target);
    }
    __syncthreads();
    if (sm.flag[2]) {
        if (threadIdx.x < gridDim.x)
            asm volatile("st.release.gpu.global.u64 [%0], %1;" :: "l"(a.bar + 16 + 16 * threadIdx.x), "l"(target) : "memory");
    } else if (threadIdx.x == 0) {
        const unsigned long long * f = a.bar + 16 + 16 * blockIdx.x;
        const long long t0 = clock64();
        unsigned long long v;
        for (;;) {
            asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
            if (v >= target) break;
            if (clock64() - t0 > (6LL << 30)) { *a.err = 1; __threadfence_system(); __trap(); }    // ~3 s: a CTA never arrived
        }
    }
    __syncthreads();
}

// ---- prologue of a GEMV phase --------------------------------------------------------------------------------------------
// quantised activations in global memory ("actq" format, written by the producing phase): block-32 weight types:
// int8 [8][K] followed by f32 block scales [8][K/32]; F16 weights: half [8][K].
// Q8_0 block of the 4 values of 8 consecutive lanes (quantize_row_q8_0: d = amax/127 stored as f16, q = rint(x * 127/amax);
// the two quotients are formed with a reciprocal multiply / __fdividef: <= 2 ulp from the IEEE quotient, which only matters when
// a product lands within 1e-6 of a rounding boundary)
__device__ __forceinline__ uint32_t q8_block4(const float4 & y, float & d_out) {
    float amax = fmaxf(fmaxf(fabsf(y.x), fabsf(y.y)), fmaxf(fabsf(y.z), fabsf(y.w)));
    amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 1));
    amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 2));
    amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 4));
    d_out = __half2float(__float2half_rn(amax * (1.0f / 127.0f)));
    const float id = (amax != 0.0f) ? __fdividef(127.0f, amax) : 0.0f;
    const uint32_t q0 = (uint32_t) __float2int_rn(y.x * id) & 0xffu, q1 = (uint32_t) __float2int_rn(y.y * id) & 0xffu;
    const uint32_t q2 = (uint32_t) __float2int_rn(y.z * id) & 0xffu, q3 = (uint32_t) __float2int_rn(y.w * id) & 0xffu;
    return q0 | (q1 << 8) | (q2 << 16) | (q3 << 24);
}

// LayerNorm (ggml-cpu/ops.cpp:3698-3765, two passes, then mul/add whisper.cpp:2536-2543) + quantisation of every row of the
// f32 residual stream into shared memory.  Two warps per row; each lane keeps its <= 5 float4 of the row in registers.
template <int WT>
__device__ __noinline__ void mk_load_ln(const MkArgs & a, MkSm & sm, const float * src, int K, const float * __restrict__ ln_w, const float * __restrict__ ln_b) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, t = warp >> 1, hw = warp & 1;
    const int nch = K >> 7;                                      // 128-value chunks per row; this warp takes chunks hw, hw+2, ...
    const int SW = (WT == WT_F16) ? (K / 2 + 4) : (K / 4 + 4);
    sm.SW = SW;
    const bool act = t < a.n_tok;
    const float4 * xr = reinterpret_cast<const float4 *>(src + (size_t) t * K);
    float4 v[5];
    float s = 0.0f;
#pragma unroll
    for (int u = 0; u < 5; ++u) {
        const int j = hw + 2 * u;
        v[u] = (act && j < nch) ? __ldcg(xr + j * 32 + lane) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        s += (v[u].x + v[u].y) + (v[u].z + v[u].w);
    }
    s = warp_sum(s);
    if (lane == 0) sm.stat[warp] = s;
    __syncthreads();
    const float mean = (sm.stat[warp & ~1] + sm.stat[warp | 1]) / K;
    float q = 0.0f;
#pragma unroll
    for (int u = 0; u < 5; ++u) {
        if (hw + 2 * u < nch) {
            v[u].x -= mean; v[u].y -= mean; v[u].z -= mean; v[u].w -= mean;
            q += (v[u].x * v[u].x + v[u].y * v[u].y) + (v[u].z * v[u].z + v[u].w * v[u].w);
        }
    }
    q = warp_sum(q);
    if (lane == 0) sm.stat[16 + warp] = q;
    __syncthreads();
    const float rstd = 1.0f / sqrtf((sm.stat[16 + (warp & ~1)] + sm.stat[16 + (warp | 1)]) / K + a.eps);
    if (act) {
#pragma unroll
        for (int u = 0; u < 5; ++u) {
            const int j = hw + 2 * u;
            if (j < nch) {
                const float4 w = __ldg(reinterpret_cast<const float4 *>(ln_w) + j * 32 + lane), b = __ldg(reinterpret_cast<const float4 *>(ln_b) + j * 32 + lane);
                float4 y;
                y.x = __fadd_rn(__fmul_rn(__fmul_rn(v[u].x, rstd), w.x), b.x);
                y.y = __fadd_rn(__fmul_rn(__fmul_rn(v[u].y, rstd), w.y), b.y);
                y.z = __fadd_rn(__fmul_rn(__fmul_rn(v[u].z, rstd), w.z), b.z);
                y.w = __fadd_rn(__fmul_rn(__fmul_rn(v[u].w, rstd), w.w), b.w);
                const int e0 = j * 128 + lane * 4;
                if (WT == WT_F16) {
                    const __half2 h0 = __floats2half2_rn(y.x, y.y), h1 = __floats2half2_rn(y.z, y.w);
                    sm.xq[t * SW + e0 / 2]     = *reinterpret_cast<const uint32_t *>(&h0);
                    sm.xq[t * SW + e0 / 2 + 1] = *reinterpret_cast<const uint32_t *>(&h1);
                } else {
                    float d;
                    sm.xq[t * SW + e0 / 4] = q8_block4(y, d);
                    if ((lane & 7) == 0) sm.xd[t * (K >> 5) + j * 4 + (lane >> 3)] = d;
                }
            }
        }
    }
    __syncthreads();
}

// copy already-quantised activations (written by the producing phase) into shared memory: two warps per row
template <int WT>
__device__ __noinline__ void mk_load_q(const MkArgs & a, MkSm & sm, const uint8_t * src, int K) {
    const int tid = threadIdx.x, t = tid >> 6, l64 = tid & 63;
    const int SW = (WT == WT_F16) ? (K / 2 + 4) : (K / 4 + 4);
    sm.SW = SW;
    const int cpr = ((WT == WT_F16) ? K * 2 : K) >> 4;           // 16-byte chunks per row
    if (t < a.n_tok) {
        const uint4 * s4 = reinterpret_cast<const uint4 *>(src) + (size_t) t * cpr;
        for (int c0 = l64; c0 < cpr; c0 += 64 * 5) {
            uint4 v[5];
#pragma unroll
            for (int u = 0; u < 5; ++u) if (c0 + 64 * u < cpr) v[u] = __ldcg(s4 + c0 + 64 * u);
#pragma unroll
            for (int u = 0; u < 5; ++u) if (c0 + 64 * u < cpr) *reinterpret_cast<uint4 *>(sm.xq + t * SW + (c0 + 64 * u) * 4) = v[u];
        }
        if (WT != WT_F16) {
            const float * sc = reinterpret_cast<const float *>(src + (size_t) 8 * K) + t * (K >> 5);
            for (int i = l64; i < (K >> 5); i += 64) sm.xd[t * (K >> 5) + i] = __ldcg(sc + i);
        }
    }
    __syncthreads();
}

// one warp quantises the 32 values its lanes hold (one Q8_0 block of row t starting at element e0) into the actq format
template <int WT>
__device__ __forceinline__ void mk_store_q(uint8_t * dst, int K, int t, int e0, int lane, float v) {
    if (WT == WT_F16) { reinterpret_cast<__half *>(dst)[(size_t) t * K + e0 + lane] = __float2half_rn(v); return; }
    const float amax = warp_max(fabsf(v));
    const float id = (amax != 0.0f) ? __fdividef(127.0f, amax) : 0.0f;
    reinterpret_cast<int8_t *>(dst)[(size_t) t * K + e0 + lane] = (int8_t) __float2int_rn(v * id);
    if (lane == 0) reinterpret_cast<float *>(dst + (size_t) 8 * K)[t * (K >> 5) + (e0 >> 5)] = __half2float(__float2half_rn(amax * (1.0f / 127.0f)));
}

struct MkEpi {
    const float * bias = nullptr, * scale = nullptr; int act = 0; const float * res = nullptr; float * out = nullptr;
    uint8_t * outq = nullptr;            // quantised output (actq format) for the next GEMV; needs KS <= 8 (row pairs in one CTA)
    __half * kc = nullptr, * vc = nullptr; int kv_d = 0;
};

// L2 prefetch of the tiles this CTA will own in a later GEMV phase (tile-major: the records of a tile are contiguous)
__device__ __noinline__ void mk_prefetch_w(const QMat & W, int KS) {
    if (threadIdx.x != MK_THREADS - 32) return;                  // one lane of the last warp (idle in the LayerNorm prologue)
    const int n_tiles = (W.N + 15) >> 4, TPC = MK_WARPS / KS;
    const uint32_t tile_bytes = (uint32_t) (W.K / wt_tm_rec_k(W.type)) * wt_tm_rec_bytes(W.type);
    for (int tile0 = blockIdx.x * TPC; tile0 < n_tiles; tile0 += gridDim.x * TPC)
        l2_prefetch(reinterpret_cast<const uint8_t *>(W.base) + (size_t) tile0 * tile_bytes, tile_bytes * (uint32_t) min(TPC, n_tiles - tile0));
}

// y[t][n] = act((W[n,:] . x[t,:] + bias[n]) * scale[n]) + res[t][n]   for the tiles owned by this CTA
template <int WT>
__device__ __noinline__ void mk_gemv(const MkArgs & a, MkSm & sm, const QMat & W, const MkEpi & e, int KS, int & round) {
    constexpr int REC = (WT == WT_Q4_0) ? 288 : (WT == WT_Q5_0 ? 352 : (WT == WT_Q8_0 ? 544 : 512));
    constexpr int QSB = (WT == WT_Q8_0) ? 512 : 256;             // bytes of the qs part of a record
    constexpr int UB = 5;                                        // records whose loads are issued together
    const int N = W.N, n_tok = a.n_tok;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, c = lane & 3;
    const int n_tiles = (N + 15) >> 4, nrec = W.K / ((WT == WT_F16) ? 16 : 32);
    const int TPC = MK_WARPS / KS, tl = warp / KS, kp = warp - tl * KS;
    const int SW = sm.SW;
    const bool tok_ok = g < n_tok;
    const int t0 = min(2 * c, n_tok - 1), t1 = min(2 * c + 1, n_tok - 1);
    for (int tile0 = blockIdx.x * TPC; tile0 < n_tiles; tile0 += gridDim.x * TPC, ++round) {
        const int tile = tile0 + tl;
        float acc[4] = { 0.0f, 0.0f, 0.0f, 0.0f };
        if (tile < n_tiles) {
            const uint8_t * tb = reinterpret_cast<const uint8_t *>(W.base) + (size_t) tile * nrec * REC;
            for (int kb = kp; kb < nrec; kb += KS * UB) {
                uint4 wq[UB]; uint2 wh[UB]; uint32_t wd[UB];
#pragma unroll
                for (int u = 0; u < UB; ++u) {
                    const uint8_t * rec = tb + (size_t) min(kb + u * KS, nrec - 1) * REC;
                    if (WT == WT_F16 || WT == WT_Q8_0) wq[u] = __ldg(reinterpret_cast<const uint4 *>(rec) + lane);
                    else { const uint2 q2 = __ldg(reinterpret_cast<const uint2 *>(rec) + lane); wq[u].x = q2.x; wq[u].y = q2.y; }
                    if (WT == WT_Q5_0) wh[u] = __ldg(reinterpret_cast<const uint2 *>(rec + QSB) + g);
                    if (WT != WT_F16)  wd[u] = __ldg(reinterpret_cast<const uint32_t *>(rec + QSB + (WT == WT_Q5_0 ? 64 : 0)) + g);
                }
#pragma unroll
                for (int u = 0; u < UB; ++u) {
                    const int b = kb + u * KS;
                    if (b < nrec) {
                        const uint32_t b0 = tok_ok ? sm.xq[g * SW + b * 8 + c] : 0u, b1 = tok_ok ? sm.xq[g * SW + b * 8 + 4 + c] : 0u;
                        if (WT == WT_F16) {
                            const uint32_t af[4] = { wq[u].x, wq[u].y, wq[u].z, wq[u].w };
                            mma_f16_16816(acc, af, b0, b1);
                        } else {
                            uint32_t af[4];
                            if (WT == WT_Q8_0) { af[0] = wq[u].x; af[1] = wq[u].y; af[2] = wq[u].z; af[3] = wq[u].w; }
                            else {
                                uint32_t lo0 = wq[u].x & 0x0F0F0F0Fu, hi0 = (wq[u].x >> 4) & 0x0F0F0F0Fu, lo1 = wq[u].y & 0x0F0F0F0Fu, hi1 = (wq[u].y >> 4) & 0x0F0F0F0Fu;
                                if (WT == WT_Q5_0) {
                                    lo0 |= spread4_to_bit4(wh[u].x >> (4 * c)); hi0 |= spread4_to_bit4(wh[u].x >> (16 + 4 * c));
                                    lo1 |= spread4_to_bit4(wh[u].y >> (4 * c)); hi1 |= spread4_to_bit4(wh[u].y >> (16 + 4 * c));
                                    af[0] = __vsub4(lo0, 0x10101010u); af[2] = __vsub4(hi0, 0x10101010u);
                                    af[1] = __vsub4(lo1, 0x10101010u); af[3] = __vsub4(hi1, 0x10101010u);
                                } else {
                                    af[0] = __vsub4(lo0, 0x08080808u); af[2] = __vsub4(hi0, 0x08080808u);
                                    af[1] = __vsub4(lo1, 0x08080808u); af[3] = __vsub4(hi1, 0x08080808u);
                                }
                            }
                            int dd[4]; mma_s8_16832(dd, af, b0, b1);
                            const float dx0 = sm.xd[t0 * nrec + b], dx1 = sm.xd[t1 * nrec + b];
                            const float dw0 = __half2float(__ushort_as_half((unsigned short) (wd[u] & 0xffffu))), dw1 = __half2float(__ushort_as_half((unsigned short) (wd[u] >> 16)));
                            acc[0] = fmaf(dw0 * dx0, (float) dd[0], acc[0]);
                            acc[1] = fmaf(dw0 * dx1, (float) dd[1], acc[1]);
                            acc[2] = fmaf(dw1 * dx0, (float) dd[2], acc[2]);
                            acc[3] = fmaf(dw1 * dx1, (float) dd[3], acc[3]);
                        }
                    }
                }
            }
        }
        float * red = sm.red + (round & 1) * MK_RED;
#pragma unroll
        for (int i = 0; i < 4; ++i) red[(warp * 16 + g + (i >> 1) * 8) * 9 + 2 * c + (i & 1)] = acc[i];
        __syncthreads();
        for (int o = tid; o < TPC * 128; o += MK_THREADS) {
            // plain: 16 consecutive rows per (tile, row-of-batch); quantised output: one warp = 32 consecutive rows of one batch row
            int tl2, t, rl;
            if (e.outq) { const int pr = o >> 8, rem = o & 255; t = rem >> 5; tl2 = pr * 2 + ((rem & 31) >> 4); rl = rem & 15; }
            else        { tl2 = o >> 7; t = (o & 127) >> 4; rl = o & 15; }
            const int row = (tile0 + tl2) * 16 + rl;
            if (t < n_tok && row < N) {
                float v = 0.0f;
                for (int w = 0; w < KS; ++w) v += red[((tl2 * KS + w) * 16 + rl) * 9 + t];
                v = (v + (e.bias ? __ldg(e.bias + row) : 0.0f)) * (e.scale ? __ldg(e.scale + row) : 1.0f);
                if (e.act == 1) v = gelu_ref_f16(v);
                if (e.res) v += __ldcg(e.res + (size_t) t * N + row);
                if (e.out) e.out[(size_t) t * N + row] = v;
                if (e.outq) mk_store_q<WT>(e.outq, N, t, row & ~31, lane, v);
                if (e.kc && row >= e.kv_d) {
                    const size_t cell = a.cell[t];
                    if (row < 2 * e.kv_d) e.kc[cell * e.kv_d + (row - e.kv_d)] = __float2half_rn(v);
                    else                  e.vc[cell * e.kv_d + (row - 2 * e.kv_d)] = __float2half_rn(v);
                }
            }
        }
    }
}

// ---- attention -----------------------------------------------------------------------------------------------------------
// stage the f16-rounded queries of all rows in shared memory (ggml_flash_attn_ext converts Q to f16: ggml-cpu/ops.cpp:8560-8571)
__device__ __noinline__ void mk_stage_q(const MkArgs & a, MkSm & sm, const float * q, int ldq) {
    float * qs = reinterpret_cast<float *>(sm.xq);
    const int tid = threadIdx.x, t = tid >> 6, l64 = tid & 63, d = a.d, n4 = d >> 2;
    if (t < a.n_tok) {
        const float4 * s4 = reinterpret_cast<const float4 *>(q + (size_t) t * ldq);
        float4 v[5];
#pragma unroll
        for (int u = 0; u < 5; ++u) if (l64 + 64 * u < n4) v[u] = __ldcg(s4 + l64 + 64 * u);
#pragma unroll
        for (int u = 0; u < 5; ++u) if (l64 + 64 * u < n4) {
            float4 r;
            r.x = __half2float(__float2half_rn(v[u].x)); r.y = __half2float(__float2half_rn(v[u].y));
            r.z = __half2float(__float2half_rn(v[u].z)); r.w = __half2float(__float2half_rn(v[u].w));
            *reinterpret_cast<float4 *>(qs + t * d + (l64 + 64 * u) * 4) = r;
        }
    }
    __syncthreads();
}

struct KVFrag { uint4 k0, k1; uint32_t v[8]; };

__device__ __forceinline__ float dot16(const uint4 & k0, const uint4 & k1, const float * __restrict__ q) {
    float s = 0.0f;
    const __half2 * h0 = reinterpret_cast<const __half2 *>(&k0), * h1 = reinterpret_cast<const __half2 *>(&k1);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 f = __half22float2(h0[i]), f2 = __half22float2(h1[i]);
        const float2 qa = *reinterpret_cast<const float2 *>(q + 2 * i), qb = *reinterpret_cast<const float2 *>(q + 8 + 2 * i);
        s = fmaf(f.x, qa.x, s); s = fmaf(f.y, qa.y, s); s = fmaf(f2.x, qb.x, s); s = fmaf(f2.y, qb.y, s);
    }
    return s;
}

// one 8-key group of a warp: online-softmax update of (m, l, o0, o1).  sc: this lane's key score (-inf = masked), v: the 8 value rows
__device__ __forceinline__ void attn_update(float sc, const uint32_t (&v)[8], float & m, float & l, float & o0, float & o1) {
    float mg = sc;
    mg = fmaxf(mg, __shfl_xor_sync(0xffffffffu, mg, 4));
    mg = fmaxf(mg, __shfl_xor_sync(0xffffffffu, mg, 8));
    mg = fmaxf(mg, __shfl_xor_sync(0xffffffffu, mg, 16));
    const float mn = fmaxf(m, mg);                               // finite: at least one key of the group is valid
    const float resc = __expf(m - mn);
    const float pk = __expf(sc - mn);                            // exp(-inf) = 0 for masked keys
    float ps = pk;
    ps += __shfl_xor_sync(0xffffffffu, ps, 4);
    ps += __shfl_xor_sync(0xffffffffu, ps, 8);
    ps += __shfl_xor_sync(0xffffffffu, ps, 16);
    l = l * resc + ps; o0 *= resc; o1 *= resc; m = mn;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float pi = __shfl_sync(0xffffffffu, pk, 4 * i);
        const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&v[i]));
        o0 = fmaf(pi, f.x, o0); o1 = fmaf(pi, f.y, o1);
    }
}

// merge `nw` warp partials (m, l, o[64]) at pp; returns this thread's dimension of the merged numerator, M and L by reference
__device__ __forceinline__ float attn_merge(const float * pp, int nw, int dim, float & M, float & Lsum) {
    M = -INFINITY;
    for (int w = 0; w < nw; ++w) M = fmaxf(M, pp[w * MK_PART]);
    Lsum = 0.0f;
    float o = 0.0f;
    for (int w = 0; w < nw; ++w) {
        const float mw = pp[w * MK_PART];
        const float wgt = (mw > -INFINITY) ? __expf(mw - M) : 0.0f;
        Lsum = fmaf(pp[w * MK_PART + 1], wgt, Lsum);
        o = fmaf(pp[w * MK_PART + 4 + dim], wgt, o);
    }
    return o;
}

// self-attention over the paged cache (whisper.cpp:2603-2625).  Item = (row, head); each half of the CTA (8 warps) takes one.
template <int WT>
__device__ __noinline__ void mk_attn_self(const MkArgs & a, MkSm & sm, const MkLayer & L) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, half = warp >> 3, hw = warp & 7;
    const int d = a.d, H = a.n_head, n_pairs = a.n_tok * H;
    const float * qs = reinterpret_cast<const float *>(sm.xq);
    const int kslot = lane >> 2, r = lane & 3;
    for (int p = blockIdx.x * 2 + half; p < n_pairs; p += gridDim.x * 2) {
        const int t = p / H, h = p - t * H;
        const int nk = a.nkv[t];
        const int * cells = a.idx + (size_t) t * a.ld_idx;
        const float * q = qs + t * d + h * 64 + r * 16;
        float m = -INFINITY, l = 0.0f, o0 = 0.0f, o1 = 0.0f;
        for (int k0 = hw * 8; k0 < nk; k0 += 64) {
            const bool ok = k0 + kslot < nk;
            const int cell = ok ? cells[k0 + kslot] : 0;
            uint4 ka = make_uint4(0, 0, 0, 0), kb = ka;
            if (ok) { const uint4 * kp = reinterpret_cast<const uint4 *>(L.kc + (size_t) cell * d + h * 64 + r * 16); ka = __ldcg(kp); kb = __ldcg(kp + 1); }
            uint32_t vv[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int ci = __shfl_sync(0xffffffffu, cell, 4 * i);
                vv[i] = (k0 + i < nk) ? __ldcg(reinterpret_cast<const uint32_t *>(L.vc + (size_t) ci * d + h * 64) + lane) : 0u;
            }
            float sc = dot16(ka, kb, q);
            sc += __shfl_xor_sync(0xffffffffu, sc, 1);
            sc += __shfl_xor_sync(0xffffffffu, sc, 2);
            attn_update(ok ? sc : -INFINITY, vv, m, l, o0, o1);
        }
        float * part = sm.part + (half * 8 + hw) * MK_PART;
        if (lane == 0) { part[0] = m; part[1] = l; }
        *reinterpret_cast<float2 *>(part + 4 + 2 * lane) = make_float2(o0, o1);
        bar_named(1 + half, 256);
        if (hw < 2) {
            float M, Lsum;
            const float o = attn_merge(sm.part + half * 8 * MK_PART, 8, hw * 32 + lane, M, Lsum);
            mk_store_q<WT>(a.actq, d, t, h * 64 + hw * 32, lane, (Lsum > 0.0f) ? __fdividef(o, Lsum) : 0.0f);
        }
        bar_named(1 + half, 256);
    }
}

__device__ __forceinline__ void xkv_load(KVFrag & f, const __half * __restrict__ kb, const __half * __restrict__ vb, int d, int lane) {
    const uint4 * kp = reinterpret_cast<const uint4 *>(kb + (size_t) (lane >> 2) * d + (lane & 3) * 16);
    f.k0 = __ldg(kp); f.k1 = __ldg(kp + 1);
#pragma unroll
    for (int i = 0; i < 8; ++i) f.v[i] = __ldg(reinterpret_cast<const uint32_t *>(vb + (size_t) i * d) + lane);
}

// cross-attention over the n_keys padded encoder positions, zero rows included (whisper.cpp:2688-2705).
// Work unit = (row, head, chunk of 128 keys), pair-major.  The units are cut into equal contiguous ranges ("stream-K"): a CTA
// walks its range chunk by chunk -- warp w owns keys 8w..8w+7 of every chunk and keeps a running (max, sum, out) per (row, head)
// in registers -- and only merges its 16 warps when the (row, head) pair changes.  A pair that is spread over several CTAs is
// finished by the last one to arrive (counter), merging the per-CTA partials in CTA order.
template <int WT>
__device__ __noinline__ void mk_attn_cross(const MkArgs & a, MkSm & sm, const MkLayer & L) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int d = a.d, H = a.n_head, nc = a.n_keys / MK_XKEYS;
    const unsigned I = (unsigned) (a.n_tok * H * nc), G = min(gridDim.x, I);   // G <= I: every participating CTA owns >= 1 unit
    const float * qs = reinterpret_cast<const float *>(sm.xq);
    const bool part_of = blockIdx.x < G;
    const unsigned s0 = part_of ? (blockIdx.x * I) / G : 0u, s1 = part_of ? ((blockIdx.x + 1) * I) / G : 0u;
    // (pair, chunk) of the unit being loaded, advanced incrementally (no divisions in the loop)
    int pl = (int) (s0 / (unsigned) nc), jl = (int) (s0 - (unsigned) pl * nc);
    int tl_ = pl / H, hl = pl - tl_ * H;
    int p = pl, t = tl_, h = hl;                                 // pair of the unit being computed
    auto load_next = [&](KVFrag & f) {
        const size_t off = (size_t) a.slot[tl_] * a.slot_stride + (size_t) (jl * MK_XKEYS + warp * 8) * d + hl * 64;
        xkv_load(f, L.xk + off, L.xv + off, d, lane);
        if (++jl == nc) { jl = 0; if (++hl == H) { hl = 0; ++tl_; } }
    };
    KVFrag cur, nxt;
    unsigned loaded = s0;
    if (s0 < s1) { load_next(cur); ++loaded; }
    int buf = 0;
    unsigned it = s0;
    while (it < s1) {
        const unsigned seg_end = min(s1, (unsigned) (p + 1) * nc);
        const int nseg = (int) (seg_end - it);
        const float * q = qs + t * d + h * 64 + (lane & 3) * 16;
        float m = -INFINITY, l = 0.0f, o0 = 0.0f, o1 = 0.0f;
        for (; it < seg_end; ++it) {
            if (loaded < s1) { load_next(nxt); ++loaded; }
            float sc = dot16(cur.k0, cur.k1, q);
            sc += __shfl_xor_sync(0xffffffffu, sc, 1);
            sc += __shfl_xor_sync(0xffffffffu, sc, 2);
            attn_update(sc * a.kq_scale, cur.v, m, l, o0, o1);
            cur = nxt;
        }
        float * part = sm.part + (buf * MK_WARPS + warp) * MK_PART;
        if (lane == 0) { part[0] = m; part[1] = l; }
        *reinterpret_cast<float2 *>(part + 4 + 2 * lane) = make_float2(o0, o1);
        __syncthreads();
        if (warp < 2) {                                          // merge the 16 warp partials of this segment
            const int dim = warp * 32 + lane;
            float M, Lsum;
            const float o = attn_merge(sm.part + buf * MK_WARPS * MK_PART, MK_WARPS, dim, M, Lsum);
            if (nseg == nc) {
                mk_store_q<WT>(a.actq, d, t, h * 64 + warp * 32, lane, __fdividef(o, Lsum));
            } else {
                // CTAs that share this pair: first = the one whose range contains unit p*nc, last = the one containing (p+1)*nc - 1
                const unsigned u0 = (unsigned) p * nc, u1 = u0 + nc - 1;
                // (CTA c owns units [c*I/G, (c+1)*I/G): the owner of unit u is ((u+1)*G - 1) / I)
                const unsigned cf = ((u0 + 1) * G - 1) / I, cl = ((u1 + 1) * G - 1) / I;
                const int slot = (int) (blockIdx.x - cf), n_contrib = (int) (cl - cf + 1);
                float * gp = a.xpart + ((size_t) p * MK_XSLOTS + slot) * 66;
                gp[2 + dim] = o;
                if (dim == 0) { gp[0] = M; gp[1] = Lsum; }
                __threadfence();
                bar_named(3, 64);
                if (tid == 0) sm.flag[buf] = (atomicAdd(a.xcnt + p, 1) == n_contrib - 1);
                bar_named(3, 64);
                if (sm.flag[buf]) {
                    __threadfence();
                    const float * p0 = a.xpart + (size_t) p * MK_XSLOTS * 66;
                    float MM = -INFINITY;
                    for (int sidx = 0; sidx < n_contrib; ++sidx) MM = fmaxf(MM, __ldcg(p0 + sidx * 66));
                    float LL = 0.0f, oo = 0.0f;
                    for (int sidx = 0; sidx < n_contrib; ++sidx) {
                        const float wgt = __expf(__ldcg(p0 + sidx * 66) - MM);
                        LL = fmaf(__ldcg(p0 + sidx * 66 + 1), wgt, LL);
                        oo = fmaf(__ldcg(p0 + sidx * 66 + 2 + dim), wgt, oo);
                    }
                    mk_store_q<WT>(a.actq, d, t, h * 64 + warp * 32, lane, __fdividef(oo, LL));
                    if (tid == 0) a.xcnt[p] = 0;
                }
            }
        }
        buf ^= 1;
        ++p; if (++h == H) { h = 0; ++t; }
    }
}

template <int WT, bool TRACE>
__global__ void __launch_bounds__(MK_THREADS, 1)
k_decode_pass(const __grid_constant__ MkArgs a) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    MkSm sm;
    {
        const int KMAX = 4 * a.d;
        const size_t xq_bytes = (WT == WT_F16) ? (size_t) 8 * (KMAX * 2 + 16) : (size_t) 8 * (KMAX + 16);
        uint8_t * p = smem_raw;
        sm.xq = reinterpret_cast<uint32_t *>(p); p += xq_bytes;
        sm.xd = reinterpret_cast<float *>(p);    p += (size_t) 8 * (KMAX / 32) * 4;
        sm.red = reinterpret_cast<float *>(p);   p += (size_t) 2 * MK_RED * 4;
        sm.part = reinterpret_cast<float *>(p);  p += (size_t) 2 * MK_WARPS * MK_PART * 4;
        sm.stat = reinterpret_cast<float *>(p);  p += 32 * 4;
        sm.flag = reinterpret_cast<int *>(p);
        sm.SW = 0;
    }
    unsigned long long target = a.bar_base;
    int round = 0, n_stamp = 0;
    const int d = a.d;
    const bool pf_w = a.prefetch & 1;
    const int KS_LOG = 4;
    MK_STAMP();
    if (pf_w) mk_prefetch_w(a.layers[0].qkv, 16);

    for (int l = 0; l < a.n_layer; ++l) {
        const MkLayer & L = a.layers[l];
        MkEpi e;
        // A: LN + QKV + KV append (whisper.cpp:2536-2599)
        if (pf_w) mk_prefetch_w(L.o, 16);
        mk_load_ln<WT>(a, sm, a.x, d, L.ln0_w, L.ln0_b);
        MK_STAMP();
        e = MkEpi(); e.bias = L.qkv_bias; e.scale = L.qkv_scale; e.out = a.qkv; e.kc = L.kc; e.vc = L.vc; e.kv_d = d;
        mk_gemv<WT>(a, sm, L.qkv, e, 16, round);
        MK_STAMP(); target += gridDim.x; mk_grid_sync(a, sm, target); MK_STAMP();
        // B: self-attention (2603-2625) -> quantised rows for the O projection
        if (pf_w) mk_prefetch_w(L.cq, 16);
        mk_stage_q(a, sm, a.qkv, 3 * d);
        MK_STAMP();
        mk_attn_self<WT>(a, sm, L);
        MK_STAMP(); target += gridDim.x; mk_grid_sync(a, sm, target); MK_STAMP();
        // C: O + residual (2647-2659)
        if (pf_w) mk_prefetch_w(L.co, 16);
        mk_load_q<WT>(a, sm, a.actq, d);
        MK_STAMP();
        e = MkEpi(); e.bias = L.o_bias; e.res = a.x; e.out = a.x;
        mk_gemv<WT>(a, sm, L.o, e, 16, round);
        MK_STAMP(); target += gridDim.x; mk_grid_sync(a, sm, target); MK_STAMP();
        // D: LN + cross Q (2661-2681)
        if (pf_w) mk_prefetch_w(L.fc1, 8);
        mk_load_ln<WT>(a, sm, a.x, d, L.lnc_w, L.lnc_b);
        MK_STAMP();
        e = MkEpi(); e.bias = L.cq_bias; e.out = a.q2;
        mk_gemv<WT>(a, sm, L.cq, e, 16, round);
        MK_STAMP(); target += gridDim.x; mk_grid_sync(a, sm, target); MK_STAMP();
        // E: cross-attention (2688-2705)
        if (pf_w) mk_prefetch_w(L.fc2, 16);
        mk_stage_q(a, sm, a.q2, d);
        MK_STAMP();
        mk_attn_cross<WT>(a, sm, L);
        MK_STAMP(); target += gridDim.x; mk_grid_sync(a, sm, target); MK_STAMP();
        // F: cross O + residual (2754-2766)
        if (pf_w) { if (l + 1 < a.n_layer) mk_prefetch_w(a.layers[l + 1].qkv, 16); else if (a.want_logits) mk_prefetch_w(a.te, KS_LOG); }
        mk_load_q<WT>(a, sm, a.actq, d);
        MK_STAMP();
        e = MkEpi(); e.bias = L.co_bias; e.res = a.x; e.out = a.x;
        mk_gemv<WT>(a, sm, L.co, e, 16, round);
        MK_STAMP(); target += gridDim.x; mk_grid_sync(a, sm, target); MK_STAMP();
        // G: LN + FC1 + GELU (2770-2794) -> quantised rows for FC2 (a CTA owns 32-row pairs of tiles: one Q8_0 block per row)
        mk_load_ln<WT>(a, sm, a.x, d, L.lnm_w, L.lnm_b);
        MK_STAMP();
        e = MkEpi(); e.bias = L.fc1_bias; e.act = 1; e.outq = a.hq;
        mk_gemv<WT>(a, sm, L.fc1, e, 8, round);
        MK_STAMP(); target += gridDim.x; mk_grid_sync(a, sm, target); MK_STAMP();
        // H: FC2 + residual (2797-2806)
        mk_load_q<WT>(a, sm, a.hq, 4 * d);
        MK_STAMP();
        e = MkEpi(); e.bias = L.fc2_bias; e.res = a.x; e.out = a.x;
        mk_gemv<WT>(a, sm, L.fc2, e, 16, round);
        MK_STAMP(); target += gridDim.x; mk_grid_sync(a, sm, target); MK_STAMP();
    }
    if (a.want_logits) {                                         // final LN + logits (2811-2827)
        mk_load_ln<WT>(a, sm, a.x, d, a.lnf_w, a.lnf_b);
        MK_STAMP();
        MkEpi e; e.out = a.logits;
        mk_gemv<WT>(a, sm, a.te, e, KS_LOG, round);
        MK_STAMP();
    }
}

int mk_barriers(int n_layer, bool) { return 8 * n_layer; }
bool mk_supported(int wtype) { return wtype == WT_F16 || wt_is_block32(wtype); }
size_t mk_smem_bytes(int wtype, int d) {
    const size_t KMAX = (size_t) 4 * d;
    const size_t xq = (wtype == WT_F16) ? 8 * (KMAX * 2 + 16) : 8 * (KMAX + 16);
    return xq + 8 * (KMAX / 32) * 4 + 2 * MK_RED * 4 + 2 * MK_WARPS * MK_PART * 4 + 32 * 4 + 16 * 4;
}

template <int WT, bool TRACE>
static bool mk_launch_t(const MkArgs & a, int n_sm, cudaStream_t st) {
    static bool configured = false;
    const size_t smem = mk_smem_bytes(WT, a.d);
    if (!configured) {
        if (cudaFuncSetAttribute(k_decode_pass<WT, TRACE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) {
            set_error("decode megakernel: cannot raise the shared-memory limit"); return false;
        }
        configured = true;
    }
    if (smem > 200 * 1024) { set_error("decode megakernel: %zu bytes of shared memory needed", smem); return false; }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(n_sm); cfg.blockDim = dim3(MK_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeCooperative; at[0].val.cooperative = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, k_decode_pass<WT, TRACE>, a);
    if (e != cudaSuccess) { set_error("decode megakernel launch: %s", cudaGetErrorString(e)); return false; }
    count_launch();
    return true;
}

bool mk_launch(const MkArgs & a, int wtype, int n_sm, cudaStream_t st) {
    if (a.te.layout != 1) { set_error("decode megakernel: weights are not in the tile-major layout"); return false; }
#define WB_MK(T) (a.trace ? mk_launch_t<T, true>(a, n_sm, st) : mk_launch_t<T, false>(a, n_sm, st))
    switch (wtype) {
        case WT_F16:  return WB_MK(WT_F16);
        case WT_Q4_0: return WB_MK(WT_Q4_0);
        case WT_Q5_0: return WB_MK(WT_Q5_0);
        case WT_Q8_0: return WB_MK(WT_Q8_0);
        default: set_error("decode megakernel: unsupported weight type %d", wtype); return false;
    }
#undef WB_MK
}

} // namespace wb
