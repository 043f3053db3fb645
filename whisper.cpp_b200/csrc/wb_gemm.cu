// wb_gemm.cu -- the tcgen05 GEMM used by every dense contraction of the encode path
// (conv stem, encoder QKV/O/MLP, cross K/V, attention scores and PV in the unfused path, decoder prompt).
//
// One CTA computes a 128 (weight rows, TMEM lanes) x BN (activation rows, TMEM columns) tile:
//   warp 0      : TMA producer   -- B tile (and A tile for f16 weights) into a STAGES-deep smem ring
//   warp 1      : MMA issuer     -- one thread issues 4 x tcgen05.mma (K=16) per 64-wide k-block, commits to mbarriers
//   warps 2..9  : quant decoders -- (quantised A only) each thread turns one 32-value block into 32 halves and
//                                   stores them into the 128B-swizzled K-major operand tile;
//                                   after the main loop the same warps run the epilogue (TMEM -> regs -> global)
// Layout facts relied on (guides: blackwell_cuda_programming.md "UMMA", B300_MICROARCH.md "tcgen05"):
//   * K-major SW128 tile: row r at byte r*128, 16-byte chunk c stored at chunk (c ^ (r & 7)); 8-row groups 1024 B apart
//   * a warp may only tcgen05.ld the TMEM lane quarter (warp_id % 4)
#include <algorithm>
#include "wb_gemm.cuh"
#include "wb_ptx.cuh"
#include "wb_common.h"

namespace wb {

static constexpr int GEMM_THREADS  = 320;
static constexpr int A_TILE_BYTES  = 128 * 64 * 2;

template <int BN> struct GemmCfg {
    static constexpr int B_TILE_BYTES = BN * 64 * 2;
    static constexpr int STAGES = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
    static constexpr int SMEM   = STAGES * (A_TILE_BYTES + B_TILE_BYTES) + 1024 /*align slack*/ + 256 /*barriers*/;
};

struct GemmKParams {
    int M, N, K, taps, nkb_per_tap, nb0;
    int a_zsel0, a_zsel1, b_zsel0, b_zsel1, a_rows_per_b0, b1_in_off;
    // quantised A
    const void * a_base; const uint8_t * a_qs; const uint32_t * a_qh; const __half * a_d;
    GemmEpilogue ep;
};

__device__ __forceinline__ float gelu_ref_f16(float x) {
    // ggml CPU: table lookup on the f16-rounded input, table entries are f16(gelu(f32(x16)))  (vec.h:988-1001, ggml-cpu.c:3847).
    // Evaluated as x * sigmoid(2u) = 0.5 x (1 + tanh u) with ex2.approx + rcp.approx (a dozen instructions; tanhf made the FC1 epilogue
    // longer than its main loop).  Against the table: identical for 98.7 % of all f16 inputs, one f16 ulp of a value below 1e-2 in
    // magnitude for the rest (x < -2, where the reference's own 1 + tanhf(u) cancels) -- checked over all 37376 inputs in (-10, 10).
    if (x <= -10.0f) return 0.0f;
    if (x >=  10.0f) return x;
    const float xh = __half2float(__float2half_rn(x));
    const float e  = __expf(-1.5957691216057308f * xh * (1.0f + 0.044715f * xh * xh));
    return __half2float(__float2half_rn(__fdividef(xh, 1.0f + e)));
}

struct RawBlk { uint4 q0, q1; uint32_t qh; __half d; };

template <int WT>
__device__ __forceinline__ void raw_load(const GemmKParams & p, int m, int k, RawBlk & r, bool valid) {
    if (!valid) { r.q0 = r.q1 = make_uint4(0, 0, 0, 0); r.qh = 0; r.d = __float2half(0.0f); return; }
    const int64_t nblk = (int64_t) p.K >> 5;
    const int64_t bi   = (int64_t) m * nblk + (k >> 5);
    if (WT == WT_Q4_0 || WT == WT_Q5_0) {
        r.q0 = __ldg(reinterpret_cast<const uint4 *>(p.a_qs) + bi);
        if (WT == WT_Q5_0) r.qh = __ldg(p.a_qh + bi);
        r.d = p.a_d[bi];
    } else if (WT == WT_Q8_0) {
        r.q0 = __ldg(reinterpret_cast<const uint4 *>(p.a_qs) + 2*bi);
        r.q1 = __ldg(reinterpret_cast<const uint4 *>(p.a_qs) + 2*bi + 1);
        r.d = p.a_d[bi];
    }
}

// Epilogue of one 128 x BN tile held in TMEM columns [tmem_acc, tmem_acc + BN): TMEM -> registers -> bias / scale / GELU / residual -> global.
// Called by the 8 epilogue warps (warp = 2..9): warp & 3 selects the TMEM lane quarter the warp may read, (warp - 2) >> 2 the column half.
template <int BN>
__device__ __forceinline__ void gemm_epilogue_tile(const GemmKParams & p, uint32_t tmem_base, int m0, int n0, int b0, int b1, int warp, int lane) {
    const int q    = warp & 3;                 // TMEM lane quarter this warp may read
    const int hsel = (warp - 2) >> 2;          // which half of the columns
    const int m    = m0 + q * 32 + lane;
    const bool mval = m < p.M;
    const GemmEpilogue & e = p.ep;
    const int mg = m + b0 * p.a_rows_per_b0;   // bias / scale follow the stacked-matrix row
    const float bias = (mval && e.bias_m)  ? e.bias_m[mg]  : 0.0f;
    const float scl  = ((mval && e.scale_m) ? e.scale_m[mg] : 1.0f) * e.alpha;
    const int64_t ooff = (int64_t) b0 * e.out_b0 + (int64_t) b1 * e.out_b1;
    const int64_t roff = (int64_t) b0 * e.res_b0 + (int64_t) b1 * e.res_b1;
    constexpr int CH = BN / 2 / 32;            // 32-column chunks per warp
    if (e.res && !e.out_mmajor && !e.out_f16 && e.act == 0) {
        // f32 residual stream (attention-O / FC2: x += W a): the residual values of chunk c+1 are requested before chunk c is drained, so
        // a warp keeps two chunks (8 KB) in flight -- with one, the 128 KB of a tile took longer to arrive than its main loop runs
        float * o = reinterpret_cast<float *>(e.out) + ooff;
        const float * rp = e.res + roff + m;
        float rn[32];
        auto fetch = [&](int col) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const int n = n0 + col + j;
                rn[j] = (mval && n < p.N) ? __ldcg(rp + (int64_t) n * e.ldr) : 0.0f;
            }
        };
        fetch(hsel * (BN / 2));
        const float bs = bias * scl;
#pragma unroll 1
        for (int c = 0; c < CH; ++c) {
            const int col = hsel * (BN / 2) + c * 32;
            if (n0 + col >= p.N) break;        // warp-uniform
            uint32_t v[32];
            tmem_ld_32x32(tmem_base + ((uint32_t) (q * 32) << 16) + (uint32_t) col, v);
            float f[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = rn[j];
            if (c + 1 < CH) fetch(col + 32);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const int n = n0 + col + j;
                if (mval && n < p.N) o[(int64_t) (n + e.n_row_off) * e.ldo + m] = fmaf(__uint_as_float(v[j]), scl, bs) + f[j];
            }
        }
        return;
    }
    if (e.act == 1 && e.out_f16 && !e.res && !e.out_mmajor && e.hm_rows == 0) {
        // FC1: GELU with the reference's f16 table semantics, a dozen instructions per value (the generic path below spends about twenty,
        // which made this epilogue longer than the K = 1280 main loop): see gelu_ref_f16 for the formula and its accuracy
        __half * o = reinterpret_cast<__half *>(e.out) + ooff;
        const float bs = bias * scl;
#pragma unroll 1
        for (int c = 0; c < CH; ++c) {
            const int col = hsel * (BN / 2) + c * 32;
            if (n0 + col >= p.N) break;
            uint32_t v[32];
            tmem_ld_32x32(tmem_base + ((uint32_t) (q * 32) << 16) + (uint32_t) col, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
                const __half2 h2 = __floats2half2_rn(fmaf(__uint_as_float(v[j]), scl, bs), fmaf(__uint_as_float(v[j + 1]), scl, bs));
                float2 x = __half22float2(h2);
                x.x = fmaxf(x.x, -20.0f); x.y = fmaxf(x.y, -20.0f);            // gelu = 0 below -10 either way; keeps exp finite
                const float e0 = __expf(x.x * fmaf(x.x * x.x, -0.07135481627f, -1.5957691216f));
                const float e1 = __expf(x.y * fmaf(x.y * x.y, -0.07135481627f, -1.5957691216f));
                const __half2 g2 = __floats2half2_rn(__fdividef(x.x, 1.0f + e0), __fdividef(x.y, 1.0f + e1));
                const int n = n0 + col + j;
                if (mval && n < p.N)     o[(int64_t) (n + e.n_row_off) * e.ldo + m]     = __low2half(g2);
                if (mval && n + 1 < p.N) o[(int64_t) (n + 1 + e.n_row_off) * e.ldo + m] = __high2half(g2);
            }
        }
        return;
    }
#pragma unroll 1
    for (int c = 0; c < CH; ++c) {
        const int col = hsel * (BN / 2) + c * 32;
        if (n0 + col >= p.N) break;            // warp-uniform
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t) (q * 32) << 16) + (uint32_t) col, v);
        tmem_ld_wait();
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            float x = (__uint_as_float(v[j]) + bias) * scl;
            if (e.act == 1) x = gelu_ref_f16(x);
            f[j] = x;
        }
        if (e.res) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const int n = n0 + col + j;
                if (mval && n < p.N) f[j] += e.res[roff + (int64_t) n * e.ldr + m];
            }
        }
        if (!e.out_mmajor) {
            if (e.out_f16 && e.hm_rows > 0) {                     // head-major: 64 consecutive features of a key are one 128-byte row
                __half * o = reinterpret_cast<__half *>(e.out) + ooff + (int64_t) (m >> 6) * e.hm_rows * 64 + (m & 63);
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int n = n0 + col + j;
                    if (mval && n < p.N) o[(int64_t) (n + e.n_row_off) * 64] = __float2half_rn(f[j]);
                }
            } else if (e.out_f16) {
                __half * o = reinterpret_cast<__half *>(e.out) + ooff;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int n = n0 + col + j;
                    if (mval && n < p.N) o[(int64_t) (n + e.n_row_off) * e.ldo + m] = __float2half_rn(f[j]);
                }
            } else {
                float * o = reinterpret_cast<float *>(e.out) + ooff;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int n = n0 + col + j;
                    if (mval && n < p.N) o[(int64_t) (n + e.n_row_off) * e.ldo + m] = f[j];
                }
            }
        } else if (mval) {
            const int nb = n0 + col;
            const bool full = (nb + 32 <= p.N);
            if (e.out_f16) {
                __half * o = reinterpret_cast<__half *>(e.out) + ooff + (int64_t) m * e.ldo + nb;
                if (full && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
                    for (int j = 0; j < 32; j += 8) {
                        __half2 h0 = __floats2half2_rn(f[j], f[j+1]),   h1 = __floats2half2_rn(f[j+2], f[j+3]);
                        __half2 h2 = __floats2half2_rn(f[j+4], f[j+5]), h3 = __floats2half2_rn(f[j+6], f[j+7]);
                        uint4 u = make_uint4(*reinterpret_cast<uint32_t *>(&h0), *reinterpret_cast<uint32_t *>(&h1),
                                             *reinterpret_cast<uint32_t *>(&h2), *reinterpret_cast<uint32_t *>(&h3));
                        *reinterpret_cast<uint4 *>(o + j) = u;
                    }
                } else {
                    for (int j = 0; j < 32; ++j) if (nb + j < p.N) o[j] = __float2half_rn(f[j]);
                }
            } else {
                float * o = reinterpret_cast<float *>(e.out) + ooff + (int64_t) m * e.ldo + nb;
                if (full && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4 *>(o + j) = make_float4(f[j], f[j+1], f[j+2], f[j+3]);
                } else {
                    for (int j = 0; j < 32; ++j) if (nb + j < p.N) o[j] = f[j];
                }
            }
        }
    }
}

template <int BN, int WT>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_kernel(const GemmKParams p, const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB) {
    using Cfg = GemmCfg<BN>;
    constexpr int STAGES = Cfg::STAGES;
    constexpr bool QUANT = (WT != WT_F16);

    extern __shared__ uint8_t smem_raw[];
    uint8_t * smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t * sA = smem;
    uint8_t * sB = smem + STAGES * A_TILE_BYTES;
    uint64_t * bars      = reinterpret_cast<uint64_t *>(sB + STAGES * Cfg::B_TILE_BYTES);
    uint64_t * full_bar  = bars;
    uint64_t * empty_bar = bars + STAGES;
    uint64_t * accum_bar = bars + 2*STAGES;
    uint32_t * tmem_slot = reinterpret_cast<uint32_t *>(bars + 2*STAGES + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * 128;
    const int n0 = blockIdx.y * BN;
    const int b0 = blockIdx.z % p.nb0;
    const int b1 = blockIdx.z / p.nb0;
    const int nkb = p.taps * p.nkb_per_tap;
    const int mg0 = m0 + b0 * p.a_rows_per_b0;     // first global A row of this tile (stacked weight matrices)

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmB);
        if (!QUANT) tma_prefetch_desc(&tmA);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], QUANT ? 1 + 8 : 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(accum_bar, 1);
        mbar_fence_init();
    }
    if (warp == 2) tmem_alloc<BN>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            for (int kb = 0; kb < nkb; ++kb) {
                const int tap = kb / p.nkb_per_tap;
                const int k0  = (kb - tap * p.nkb_per_tap) * 64;
                mbar_wait(&empty_bar[s], ph ^ 1);
                const uint32_t tx = Cfg::B_TILE_BYTES + (QUANT ? 0 : A_TILE_BYTES);
                mbar_arrive_expect_tx(&full_bar[s], tx);
                const int sel[4] = { 0, b0, b1 + p.b1_in_off, tap };
                tma_load_4d(sB + s * Cfg::B_TILE_BYTES, &tmB, &full_bar[s], k0, n0, sel[p.b_zsel0], sel[p.b_zsel1]);
                if (!QUANT)
                    tma_load_4d(sA + s * A_TILE_BYTES, &tmA, &full_bar[s], k0, mg0, sel[p.a_zsel0], sel[p.a_zsel1]);
                if (++s == STAGES) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer (single thread)
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_f16(128, BN);
            int s = 0; uint32_t ph = 0;
            for (int kb = 0; kb < nkb; ++kb) {
                mbar_wait(&full_bar[s], ph);
                tc_fence_after();
                const uint64_t adesc = umma_desc_sw128(smem_u32(sA + s * A_TILE_BYTES));
                const uint64_t bdesc = umma_desc_sw128(smem_u32(sB + s * Cfg::B_TILE_BYTES));
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    // +32 bytes per K=16 step inside the swizzle atom: +2 in the (addr >> 4) field
                    umma_f16_ss(tmem_base, adesc + 2*k, bdesc + 2*k, idesc, (kb | k) ? 1u : 0u);
                }
                umma_commit(&empty_bar[s]);
                if (++s == STAGES) { s = 0; ph ^= 1; }
            }
            umma_commit(accum_bar);
        }
    } else {
        // ------------------------------------------------------------------ warps 2..9
        const int pt = threadIdx.x - 64;           // 0..255
        if (QUANT) {
            const int row  = pt >> 1;
            const int half = pt & 1;
            const int m    = mg0 + row;              // global row in the (stacked) weight matrix
            const bool mval = (m0 + row) < p.M;
            uint8_t * dst_row = sA + row * 128;
            const int sw = row & 7;
            int s = 0; uint32_t ph = 0;
            if (WT == WT_Q4_K || WT == WT_Q5_K) {
                constexpr int BLK = (WT == WT_Q4_K) ? 144 : 176;
                const uint8_t * rowp = reinterpret_cast<const uint8_t *>(p.a_base) + (int64_t) m * (p.K >> 8) * BLK;
                for (int kb = 0; kb < nkb; ++kb) {
                    const int k = kb * 64 + half * 32;
                    uint4 o[4];
                    if (mval && k < p.K) {
                        dequant_kq_sub<WT == WT_Q5_K>(rowp + (int64_t) (k >> 8) * BLK, (k & 255) >> 5, o);
                    } else {
                        o[0] = o[1] = o[2] = o[3] = make_uint4(0, 0, 0, 0);
                    }
                    mbar_wait(&empty_bar[s], ph ^ 1);
                    uint8_t * d = dst_row + s * A_TILE_BYTES;
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        *reinterpret_cast<uint4 *>(d + (((half * 4 + c) ^ sw) << 4)) = o[c];
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&full_bar[s]);
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
            } else {
                RawBlk r0, r1, r2;
                raw_load<WT>(p, m, half * 32,      r0, mval && (half * 32      < p.K) && 0 < nkb);
                raw_load<WT>(p, m, 64 + half * 32, r1, mval && (64 + half * 32 < p.K) && 1 < nkb);
                for (int kb = 0; kb < nkb; ++kb) {
                    const int k2 = (kb + 2) * 64 + half * 32;
                    raw_load<WT>(p, m, k2, r2, mval && (kb + 2 < nkb) && (k2 < p.K));
                    uint4 o[4];
                    if (WT == WT_Q4_0)      dequant_q4_0(r0.q0, r0.d, o);
                    else if (WT == WT_Q5_0) dequant_q5_0(r0.q0, r0.qh, r0.d, o);
                    else                    dequant_q8_0(r0.q0, r0.q1, r0.d, o);
                    mbar_wait(&empty_bar[s], ph ^ 1);
                    uint8_t * d = dst_row + s * A_TILE_BYTES;
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        *reinterpret_cast<uint4 *>(d + (((half * 4 + c) ^ sw) << 4)) = o[c];
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&full_bar[s]);
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                    r0 = r1; r1 = r2;
                }
            }
        }

        // ------------------------------------------------------------------ epilogue
        mbar_wait(accum_bar, 0);
        tc_fence_after();
        gemm_epilogue_tile<BN>(p, tmem_base, m0, n0, b0, b1, warp, lane);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc<BN>(tmem_base);
}

// =============================================================================================== persistent f16 x f16 GEMM
// Second generation of the encode-path GEMM.  Both operands are f16 tiles fetched by TMA (quantised weights are expanded ONCE per
// launch into an f16 scratch by k_dequant_f16 -- a 128-row weight tile used to be re-decoded by every one of the up to 375 CTAs that
// shared it), the kernel is PERSISTENT (one CTA per SM walks the tile list, weight-tile index fastest so that concurrently running CTAs
// share activation tiles in L2) and the accumulator is DOUBLE-BUFFERED in tensor memory (2 x BN columns): the epilogue of tile i
// (TMEM -> registers -> bias / GELU / residual -> global) overlaps the main loop of tile i+1.
//   warp 0      : TMA producer   -- A and B tiles into a STAGES-deep smem ring (full / empty mbarriers)
//   warp 1      : MMA issuer     -- one thread, 4 x tcgen05.mma (K = 16) per 64-wide k-block; commits the stage's "empty" barrier and,
//                                   after the last k-block of a tile, the accumulator's "full" barrier
//   warps 2..9  : epilogue       -- wait "full", drain the accumulator (gemm_epilogue_tile), arrive on its "empty" barrier
// CL = 2: CTA PAIRS (thread-block cluster of 2 along the weight-row direction) share the activation tile: each CTA fetches HALF of it and
// multicasts it into both CTAs' shared memory, so a CTA pulls 32 KB instead of 48 KB from L2 per k-block.  (Measured: the 1-CTA kernel
// sits at ~0.8 of what L2 can feed -- 48 KB per 128 x 256 x 64 block at ~42 B/clk/SM is twice the 542 clocks the tensor pipe needs.)  A stage
// is refilled only when BOTH CTAs have consumed it: the MMA thread's commit arrives on the "empty" barrier of both (count 2).
template <int BN, int CL>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm2_kernel(const GemmKParams p, const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int n_mt, int n_nt, int n_tiles) {
    using Cfg = GemmCfg<BN>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t * smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t * sA = smem;
    uint8_t * sB = smem + STAGES * A_TILE_BYTES;
    uint64_t * bars      = reinterpret_cast<uint64_t *>(sB + STAGES * Cfg::B_TILE_BYTES);
    uint64_t * full_bar  = bars;
    uint64_t * empty_bar = bars + STAGES;
    uint64_t * tfull     = bars + 2*STAGES;          // [2] accumulator ready for the epilogue
    uint64_t * tempty    = bars + 2*STAGES + 2;      // [2] accumulator drained
    uint32_t * tmem_slot = reinterpret_cast<uint32_t *>(bars + 2*STAGES + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int nkb = p.taps * p.nkb_per_tap;

    if (warp == 0 && lane == 0) { tma_prefetch_desc(&tmB); tma_prefetch_desc(&tmA); }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], CL); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 8); }
        mbar_fence_init();
    }
    if (warp == 2) tmem_alloc<2 * BN>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();                               // the peer's barriers exist before anything arrives on them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // tile list: with CL = 2 the CTAs of a pair take weight tiles 2q and 2q+1 of the same activation tile (n_mt is the number of PAIRS then)
    const int crank = (CL > 1) ? (int) cluster_ctarank() : 0;
    const int t_first = (CL > 1) ? (int) (blockIdx.x / CL) : (int) blockIdx.x, t_step = (int) gridDim.x / CL;

    if (warp == 0) {
        int s = 0; uint32_t ph = 0;
        for (int t = t_first; t < n_tiles; t += t_step) {
            const int mt = (t % n_mt) * CL + crank, r = t / n_mt, nt = r % n_nt, z = r / n_nt;
            const int b0 = z % p.nb0, b1 = z / p.nb0;
            const int mg0 = mt * 128 + b0 * p.a_rows_per_b0, n0 = nt * BN;
            if (p.ep.res && lane > 0) {
                // the residual tile (f32, 128 features of BN rows) goes to L2 while the tile's main loop runs: the epilogue then adds it at L2
                // latency instead of stalling on HBM with only 32 KB in flight per SM
                const int m0 = mt * 128, mw = min(128, p.M - m0);
                const float * rp = p.ep.res + (int64_t) b0 * p.ep.res_b0 + (int64_t) b1 * p.ep.res_b1 + m0;
                if (mw > 0 && (mw & 3) == 0 && ((reinterpret_cast<uintptr_t>(rp) | ((uintptr_t) p.ep.ldr * 4)) & 15) == 0)
                    for (int rr = lane - 1; rr < BN && n0 + rr < p.N; rr += 31)
                        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(rp + (int64_t) (n0 + rr) * p.ep.ldr), "r"(mw * 4) : "memory");
            }
            if (lane == 0) {
                for (int kb = 0; kb < nkb; ++kb) {
                    const int tap = kb / p.nkb_per_tap;
                    const int k0  = (kb - tap * p.nkb_per_tap) * 64;
                    mbar_wait(&empty_bar[s], ph ^ 1);
                    mbar_arrive_expect_tx(&full_bar[s], Cfg::B_TILE_BYTES + A_TILE_BYTES);
                    const int sel[4] = { 0, b0, b1 + p.b1_in_off, tap };
                    if (CL > 1)                                    // my half of the activation tile, into both CTAs of the pair
                        tma_load_4d_mc(sB + s * Cfg::B_TILE_BYTES + crank * (Cfg::B_TILE_BYTES / 2), &tmB, &full_bar[s], k0, n0 + crank * (BN / 2), sel[p.b_zsel0], sel[p.b_zsel1], (uint16_t) 3);
                    else
                        tma_load_4d(sB + s * Cfg::B_TILE_BYTES, &tmB, &full_bar[s], k0, n0, sel[p.b_zsel0], sel[p.b_zsel1]);
                    tma_load_4d(sA + s * A_TILE_BYTES, &tmA, &full_bar[s], k0, mg0, sel[p.a_zsel0], sel[p.a_zsel1]);
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_f16(128, BN);
            int s = 0; uint32_t ph = 0;
            int i = 0;
            for (int t = t_first; t < n_tiles; t += t_step, ++i) {
                const int acc = i & 1;
                mbar_wait(&tempty[acc], (uint32_t) ((i >> 1) & 1) ^ 1u);       // the epilogue has drained this accumulator (first use: free)
                tc_fence_after();
                const uint32_t tacc = tmem_base + (uint32_t) (acc * BN);
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(&full_bar[s], ph);
                    tc_fence_after();
                    const uint64_t adesc = umma_desc_sw128(smem_u32(sA + s * A_TILE_BYTES));
                    const uint64_t bdesc = umma_desc_sw128(smem_u32(sB + s * Cfg::B_TILE_BYTES));
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_f16_ss(tacc, adesc + 2*k, bdesc + 2*k, idesc, (kb | k) ? 1u : 0u);
                    if (CL > 1) umma_commit_mc(&empty_bar[s], (uint16_t) 3); else umma_commit(&empty_bar[s]);
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
                umma_commit(&tfull[acc]);
            }
        }
    } else {
        int i = 0;
        for (int t = t_first; t < n_tiles; t += t_step, ++i) {
            const int mt = (t % n_mt) * CL + crank, r = t / n_mt, nt = r % n_nt, z = r / n_nt;
            const int b0 = z % p.nb0, b1 = z / p.nb0;
            const int acc = i & 1;
            mbar_wait(&tfull[acc], (uint32_t) ((i >> 1) & 1));
            tc_fence_after();
            gemm_epilogue_tile<BN>(p, tmem_base + (uint32_t) (acc * BN), mt * 128, nt * BN, b0, b1, warp, lane);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();                               // nobody leaves while the peer may still multicast into it or arrive on its barriers
    if (warp == 2) tmem_dealloc<2 * BN>(tmem_base);
}

// expand a quantised matrix (planar 32-blocks or verbatim K-quant super-blocks) to f16 rows [rows][K]; one thread = one 32-value block.
// Each value is the f16 rounding of the exact d * (q - off), i.e. what dequantize_row_* gives after one f16 store (wb_quant.cuh).
template <int WT>
__global__ void __launch_bounds__(256)
k_dequant_f16(const void * base, const uint8_t * qs, const uint32_t * qh, const __half * dd, int64_t row0, int64_t n_blocks, int K, __half * out) {
    const int64_t bi = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    if (bi >= n_blocks) return;
    const int64_t nblk = K >> 5;
    uint4 o[4];
    if (WT == WT_Q4_K || WT == WT_Q5_K) {
        constexpr int BLK = (WT == WT_Q4_K) ? 144 : 176;
        const int64_t row = bi / nblk, kb = bi - row * nblk;
        const uint8_t * sb = reinterpret_cast<const uint8_t *>(base) + ((row0 + row) * (K >> 8) + (kb >> 3)) * BLK;
        dequant_kq_sub<WT == WT_Q5_K>(sb, (int) (kb & 7), o);
    } else {
        const int64_t g = row0 * nblk + bi;
        if (WT == WT_Q4_0)      dequant_q4_0(__ldg(reinterpret_cast<const uint4 *>(qs) + g), dd[g], o);
        else if (WT == WT_Q5_0) dequant_q5_0(__ldg(reinterpret_cast<const uint4 *>(qs) + g), __ldg(qh + g), dd[g], o);
        else                    dequant_q8_0(__ldg(reinterpret_cast<const uint4 *>(qs) + 2*g), __ldg(reinterpret_cast<const uint4 *>(qs) + 2*g + 1), dd[g], o);
    }
    uint4 * dst = reinterpret_cast<uint4 *>(out + bi * 32);
    dst[0] = o[0]; dst[1] = o[1]; dst[2] = o[2]; dst[3] = o[3];
}

// =============================================================================================== host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled get_encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void * p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
    return fn;
}

bool make_tmap_f16(CUtensorMap * out, const void * base, uint64_t k, uint64_t rows, uint64_t z2, uint64_t z3,
                   uint64_t stride_rows, uint64_t stride_z2, uint64_t stride_z3, uint32_t box_rows) {
    PFN_encodeTiled fn = get_encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled entry point not available"); return false; }
    cuuint64_t dims[4]    = { k, rows, z2 ? z2 : 1, z3 ? z3 : 1 };
    cuuint64_t strides[3] = { stride_rows * 2, (stride_z2 ? stride_z2 : stride_rows * rows) * 2,
                              (stride_z3 ? stride_z3 : stride_rows * rows * (z2 ? z2 : 1)) * 2 };
    cuuint32_t box[4]     = { 64, box_rows, 1, 1 };
    cuuint32_t estr[4]    = { 1, 1, 1, 1 };
    if ((reinterpret_cast<uintptr_t>(base) & 15) || (strides[0] & 15) || (strides[1] & 15) || (strides[2] & 15)) {
        set_error("make_tmap_f16: base/strides must be 16-byte aligned (base=%p strides=%llu,%llu,%llu)", base,
                  (unsigned long long) strides[0], (unsigned long long) strides[1], (unsigned long long) strides[2]);
        return false;
    }
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void *>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed: %d", (int) r); return false; }
    return true;
}

template <int BN, int WT>
static cudaError_t launch_t(const GemmDesc & g, const GemmKParams & kp, cudaStream_t st) {
    auto kern = gemm_kernel<BN, WT>;
    {
        const cudaError_t e = ensure_dyn_smem(reinterpret_cast<const void *>(kern), GemmCfg<BN>::SMEM);
        if (e != cudaSuccess) return e;
    }
    dim3 grid((g.M + 127) / 128, (g.N + BN - 1) / BN, g.nb0 * g.nb1);
    kern<<<grid, GEMM_THREADS, GemmCfg<BN>::SMEM, st>>>(kp, g.tmA, g.tmB);
    count_launch();
    return cudaGetLastError();
}

template <int WT>
static cudaError_t launch_bn(const GemmDesc & g, const GemmKParams & kp, cudaStream_t st) {
    switch (g.BN) {
        case 64:  return launch_t<64,  WT>(g, kp, st);
        case 128: return launch_t<128, WT>(g, kp, st);
        case 256: return launch_t<256, WT>(g, kp, st);
    }
    return cudaErrorInvalidValue;
}

static int gemm_n_sm() {
    static int n = 0;
    if (!n) { int dev = 0; cudaGetDevice(&dev); cudaDeviceProp pr; if (cudaGetDeviceProperties(&pr, dev) == cudaSuccess) n = pr.multiProcessorCount; else n = 148; }
    return n;
}

template <int BN, int CL>
static cudaError_t launch2_t(const GemmDesc & g, const GemmKParams & kp, cudaStream_t st) {
    auto kern = gemm2_kernel<BN, CL>;
    const size_t smem = GemmCfg<BN>::SMEM + 64;
    { const cudaError_t e = ensure_dyn_smem(reinterpret_cast<const void *>(kern), smem); if (e != cudaSuccess) return e; }
    const int n_mt = ((g.M + 127) / 128 + CL - 1) / CL, n_nt = (g.N + BN - 1) / BN, n_tiles = n_mt * n_nt * g.nb0 * g.nb1;   // CL = 2: pairs of weight tiles
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned) (CL * std::min(n_tiles, gemm_n_sm() / CL))); cfg.blockDim = dim3(GEMM_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = CL > 1 ? 1 : 0;
    const cudaError_t e = cudaLaunchKernelEx(&cfg, kern, kp, g.tmA, g.tmB, n_mt, n_nt, n_tiles);
    count_launch();
    return e;
}

cudaError_t dequant_to_f16(const QMat & W, int64_t rows, __half * out, cudaStream_t st) {
    const int64_t n_blocks = rows * (W.K >> 5);
    const unsigned grid = (unsigned) ((n_blocks + 255) / 256);
    switch (W.type) {
        case WT_Q4_0: k_dequant_f16<WT_Q4_0><<<grid, 256, 0, st>>>(W.base, W.qs, W.qh, W.d, 0, n_blocks, W.K, out); break;
        case WT_Q5_0: k_dequant_f16<WT_Q5_0><<<grid, 256, 0, st>>>(W.base, W.qs, W.qh, W.d, 0, n_blocks, W.K, out); break;
        case WT_Q8_0: k_dequant_f16<WT_Q8_0><<<grid, 256, 0, st>>>(W.base, W.qs, W.qh, W.d, 0, n_blocks, W.K, out); break;
        case WT_Q4_K: k_dequant_f16<WT_Q4_K><<<grid, 256, 0, st>>>(W.base, W.qs, W.qh, W.d, 0, n_blocks, W.K, out); break;
        case WT_Q5_K: k_dequant_f16<WT_Q5_K><<<grid, 256, 0, st>>>(W.base, W.qs, W.qh, W.d, 0, n_blocks, W.K, out); break;
        default: return cudaErrorInvalidValue;
    }
    count_launch();
    return cudaGetLastError();
}

// expand the rows of g.A that this launch uses into g.a16 (f16 [rows][K]); rows = M x (stacked matrices)
static cudaError_t dequant_launch(const GemmDesc & g, cudaStream_t st) {
    const int64_t rows = g.a_rows_per_b0 ? (int64_t) g.a_rows_per_b0 * g.nb0 : g.M;
    QMat W = g.A; W.K = g.K;
    return dequant_to_f16(W, rows, g.a16, st);
}

cudaError_t gemm_launch(const GemmDesc & g, cudaStream_t st) {
    const double nb = (double) g.nb0 * g.nb1;
    ProfScope prof(PC_GEMM, st, nb * ((double) g.M * g.K * g.taps * wt_bpw(g.A.type) + (double) g.N * g.K * g.taps * 2 + (double) g.M * g.N * (g.ep.out_f16 ? 2 : 4)),
                   nb * 2.0 * g.M * g.N * g.K * g.taps);
    GemmKParams kp;
    kp.M = g.M; kp.N = g.N; kp.K = g.K; kp.taps = g.taps; kp.nkb_per_tap = (g.K + 63) / 64; kp.nb0 = g.nb0;
    kp.a_zsel0 = g.a_zsel[0]; kp.a_zsel1 = g.a_zsel[1]; kp.b_zsel0 = g.b_zsel[0]; kp.b_zsel1 = g.b_zsel[1];
    kp.a_rows_per_b0 = g.a_rows_per_b0; kp.b1_in_off = g.b1_in_off;
    kp.a_base = g.A.base; kp.a_qs = g.A.qs; kp.a_qh = g.A.qh; kp.a_d = g.A.d;
    kp.ep = g.ep;
    if (g.v2 && (g.A.type == WT_F16 || g.a16)) {                  // persistent kernel: both operands f16 through TMA
        if (g.A.type != WT_F16 && !g.a16_keep) { const cudaError_t e = dequant_launch(g, st); if (e != cudaSuccess) return e; }
        if (g.cluster2) {                                          // tmB's box holds BN / 2 rows then
            switch (g.BN) {
                case 128: return launch2_t<128, 2>(g, kp, st);
                case 256: return launch2_t<256, 2>(g, kp, st);
            }
            return cudaErrorInvalidValue;
        }
        switch (g.BN) {
            case 64:  return launch2_t<64, 1>(g, kp, st);
            case 128: return launch2_t<128, 1>(g, kp, st);
            case 256: return launch2_t<256, 1>(g, kp, st);
        }
        return cudaErrorInvalidValue;
    }
    switch (g.A.type) {
        case WT_F16:  return launch_bn<WT_F16>(g, kp, st);
        case WT_Q4_0: return launch_bn<WT_Q4_0>(g, kp, st);
        case WT_Q5_0: return launch_bn<WT_Q5_0>(g, kp, st);
        case WT_Q8_0: return launch_bn<WT_Q8_0>(g, kp, st);
        case WT_Q4_K: return launch_bn<WT_Q4_K>(g, kp, st);
        case WT_Q5_K: return launch_bn<WT_Q5_K>(g, kp, st);
    }
    return cudaErrorInvalidValue;
}

} // namespace wb
