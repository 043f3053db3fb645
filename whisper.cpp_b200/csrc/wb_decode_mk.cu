// wb_decode_mk.cu -- one decoder pass (<= 64 rows) as a single persistent cooperative kernel.
//
// What it computes is whisper_build_graph_decoder (src/whisper.cpp:2466-2844) for the rows of one whisper_batch:
//   per layer:  LN -> QKV (+KV append) -> self-attention over the paged cache -> O + residual
//               LN -> Q -> cross-attention over the padded encoder keys       -> O + residual
//               LN -> FC1 + GELU -> FC2 + residual
//   then LN -> logits (token-embedding matrix).
// The arithmetic per op is the one of the stand-alone kernels in wb_kernels.cu (Q8_0 activation blocks + integer block
// dots like the reference CPU path, f16-rounded Q, f32 softmax); this file changes HOW the ops are scheduled:
//   * grid = one CTA of 16 warps per SM, launched cooperatively; a phase boundary is a grid barrier instead of a kernel
//     boundary (the decode step is a chain of ~400 dependent phases of a few microseconds each);
//   * rows handed from one phase to the next are quantised once by their producer ("actq" rows in global memory): a
//     distributed LayerNorm -> Q8_0 phase, attention epilogues, and a quantise phase after FC1;
//   * a GEMV phase is cut by row group first: a CTA stages the 16 rows of its group with TMA bulk copies and walks 16-row
//     weight tiles, its 16 warps split K and reduce through smem.  Weights are stored tile-major (wb_quant.cuh): a warp fetches
//     the MMA operands of one (tile, block) with three fully coalesced loads -- with the planar layout every 4-byte gather cost
//     16 sectors per 128 useful bytes and the phase was L1-sector bound;
//   * cross-attention (the HBM-heavy part: 2*n_keys*d f16 per row and layer) is cut into (row, head, half of the keys) units,
//     streamed through a cp.async ring; one lane owns one key-quarter and keeps its own online-softmax state.
// Lessons that shaped the code (profiles/): no pointers handed through structs (they turn shared-memory accesses into generic
// loads), no 64-bit divisions in loops, approximate reciprocals where they only feed a rounding, no register moves of
// in-flight loads, whole 32-byte sectors per load instruction, bulk copies instead of register staging for anything > 16 KB.
#include <cmath>
#include <type_traits>
#include "wb_decode_mk.cuh"
#include "wb_common.h"
#include "wb_dev.cuh"
#include "wb_ptx.cuh"

namespace wb {

constexpr int MK_THREADS = 512, MK_WARPS = 16, MK_XKEYS = 128, MK_MAXTOK = 64, MK_RED = 4 * 16 * 16 * 17, MK_PART = 68;
constexpr int MK_TP = 4;                // weight tiles a GEMV phase keeps in flight per iteration
static_assert(MK_TP * 16 * 16 * 17 <= MK_RED, "split-K reduction buffer");
constexpr int MK_ROWB = 1280;            // bytes of one staged activation row chunk (1280 int8 values or 640 halves)

// dynamic shared memory, fixed carve-up (every phase function addresses it directly: pointers handed through a struct in
// local memory turned every access into a generic load)
extern __shared__ __align__(16) uint8_t mk_smem[];
constexpr int MK_OFF_XQ   = 0;                                                   // staged activation rows [64][SW words]
constexpr int MK_OFF_XD   = MK_OFF_XQ + MK_MAXTOK * (MK_ROWB + 16);              // their Q8_0 block scales [64][chunk/32]
constexpr int MK_OFF_RED  = MK_OFF_XD + MK_MAXTOK * (MK_ROWB / 32) * 4;          // split-K partials [MK_TP tiles][16 warps][16 rows][17]
constexpr int MK_OFF_PART = MK_OFF_RED + MK_RED * 4;                             // attention warp partials [2][16][68]: m, l, -, -, o[64]
constexpr int MK_OFF_STAT = MK_OFF_PART + 2 * MK_WARPS * MK_PART * 4;            // LayerNorm partial sums [32]
constexpr int MK_OFF_FLAG = MK_OFF_STAT + 32 * 4;                                // [16] ints
constexpr int MK_SMEM     = MK_OFF_FLAG + 16 * 4;
#define SM_XQ   (reinterpret_cast<uint32_t *>(mk_smem + MK_OFF_XQ))
#define SM_XD   (reinterpret_cast<float *>(mk_smem + MK_OFF_XD))
#define SM_RED  (reinterpret_cast<float *>(mk_smem + MK_OFF_RED))
#define SM_PART (reinterpret_cast<float *>(mk_smem + MK_OFF_PART))
#define SM_STAT (reinterpret_cast<float *>(mk_smem + MK_OFF_STAT))
#define SM_FLAG (reinterpret_cast<int *>(mk_smem + MK_OFF_FLAG))

#define MK_STAMP() do { if (TRACE) { if (blockIdx.x == 0 && threadIdx.x == 0) a.trace[n_stamp] = clock64(); ++n_stamp; } } while (0)

__device__ __forceinline__ void l2_prefetch(const void * p, uint32_t bytes) {
    if (bytes >= 16) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(p), "r"(bytes & ~15u) : "memory");
}

// Grid barrier.  Arrival = one atomic on a counter; the LAST CTA to arrive releases everybody by writing one flag per CTA
// (each on its own 128-byte line), and every other CTA polls only its own flag.
__device__ __noinline__ void mk_grid_sync(const MkArgs & a, unsigned long long target) {
    __syncthreads();
    if (a.prefetch & 32) {
        // variant: arrival = a reduction without a return value, and EVERY CTA polls the counter itself -- no dependent chain
        // "atomic returns -> last arriver writes the flags -> the flags become visible" (one L2 round trip less per barrier)
        if (threadIdx.x == 0) {
            __threadfence();
            asm volatile("red.release.gpu.global.add.u64 [%0], 1;" :: "l"(a.bar) : "memory");
            const long long t0 = clock64();
            unsigned long long v;
            for (;;) {
                asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(a.bar) : "memory");
                if (v >= target) break;
                if (clock64() - t0 > (6LL << 30)) { *a.err = 1; __threadfence_system(); __trap(); }
            }
        }
        __syncthreads();
        return;
    }
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned long long old = atomicAdd(a.bar, 1ULL);
        SM_FLAG[2] = (old + 1 == target);
        if (old + 1 == target) __threadfence();                  // the last arriver acquires what the others released
    }
    __syncthreads();
    if (SM_FLAG[2]) {
        for (int i = threadIdx.x; i < (int) gridDim.x; i += MK_THREADS)
            asm volatile("st.release.gpu.global.u64 [%0], %1;" :: "l"(a.bar + 16 + 16 * i), "l"(target) : "memory");
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

// ---- activation rows ("actq" format) ----------------------------------------------------------------------------------------
// Rows handed from one phase to the next are quantised ONCE by their producer and live in global memory: block-32 weight types:
// int8 [64][K] followed by f32 block scales [64][K/32]; F16 weights: half [64][K].  A GEMV phase stages them chunk-wise in smem.
// Q8_0 rows follow quantize_row_q8_0 (d = amax/127 stored as f16, q = rint(x * 127/amax)); the two quotients are formed with a
// reciprocal multiply / __fdividef: <= 2 ulp from the IEEE quotient, which only matters within 1e-6 of a rounding boundary.

// one warp quantises the 32 values its lanes hold (one Q8_0 block of row t starting at element e0)
template <int WT>
__device__ __forceinline__ void mk_store_q(uint8_t * dst, int K, int t, int e0, int lane, float v) {
    if (WT == WT_F16) { reinterpret_cast<__half *>(dst)[(size_t) t * K + e0 + lane] = __float2half_rn(v); return; }
    const float amax = warp_max(fabsf(v));
    const float id = (amax != 0.0f) ? __fdividef(127.0f, amax) : 0.0f;
    reinterpret_cast<int8_t *>(dst)[(size_t) t * K + e0 + lane] = (int8_t) __float2int_rn(v * id);
    if (lane == 0) reinterpret_cast<float *>(dst + (size_t) MK_MAXTOK * K)[t * (K >> 5) + (e0 >> 5)] = __half2float(__float2half_rn(amax * (1.0f / 127.0f)));
}

// LayerNorm (ggml-cpu/ops.cpp:3698-3765, two passes, then mul/add whisper.cpp:2536-2543) of the f32 residual rows, quantised into
// `dst`.  Distributed: CTA r normalises row r (one warp per 128 values); followed by a grid barrier.
template <int WT>
__device__ __noinline__ void mk_lnq(const MkArgs & a, const float * src, int K, const float * __restrict__ ln_w, const float * __restrict__ ln_b, uint8_t * dst) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int row = blockIdx.x; row < a.n_tok; row += gridDim.x) {
        const bool act = warp < (K >> 7);
        const int e0 = warp * 128 + lane * 4;
        float4 v = act ? __ldcg(reinterpret_cast<const float4 *>(src + (size_t) row * K + e0)) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        float s = warp_sum((v.x + v.y) + (v.z + v.w));
        if (lane == 0) SM_STAT[warp] = s;
        __syncthreads();
        s = (lane < MK_WARPS) ? SM_STAT[lane] : 0.0f;
        const float mean = warp_sum(s) / K;
        float q = 0.0f;
        if (act) { v.x -= mean; v.y -= mean; v.z -= mean; v.w -= mean; q = (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w); }
        q = warp_sum(q);
        if (lane == 0) SM_STAT[16 + warp] = q;
        __syncthreads();
        q = (lane < MK_WARPS) ? SM_STAT[16 + lane] : 0.0f;
        const float rstd = 1.0f / sqrtf(warp_sum(q) / K + a.eps);
        if (act) {
            const float4 w = __ldg(reinterpret_cast<const float4 *>(ln_w + e0)), b = __ldg(reinterpret_cast<const float4 *>(ln_b + e0));
            float4 y;
            y.x = __fadd_rn(__fmul_rn(__fmul_rn(v.x, rstd), w.x), b.x);
            y.y = __fadd_rn(__fmul_rn(__fmul_rn(v.y, rstd), w.y), b.y);
            y.z = __fadd_rn(__fmul_rn(__fmul_rn(v.z, rstd), w.z), b.z);
            y.w = __fadd_rn(__fmul_rn(__fmul_rn(v.w, rstd), w.w), b.w);
            if (WT == WT_F16) {
                const __half2 h0 = __floats2half2_rn(y.x, y.y), h1 = __floats2half2_rn(y.z, y.w);
                *reinterpret_cast<uint2 *>(reinterpret_cast<__half *>(dst) + (size_t) row * K + e0) = make_uint2(*reinterpret_cast<const uint32_t *>(&h0), *reinterpret_cast<const uint32_t *>(&h1));
            } else {
                float amax = fmaxf(fmaxf(fabsf(y.x), fabsf(y.y)), fmaxf(fabsf(y.z), fabsf(y.w)));
                amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 1));
                amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 2));
                amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 4));
                const float id = (amax != 0.0f) ? __fdividef(127.0f, amax) : 0.0f;
                const uint32_t q0 = (uint32_t) __float2int_rn(y.x * id) & 0xffu, q1 = (uint32_t) __float2int_rn(y.y * id) & 0xffu;
                const uint32_t q2 = (uint32_t) __float2int_rn(y.z * id) & 0xffu, q3 = (uint32_t) __float2int_rn(y.w * id) & 0xffu;
                *reinterpret_cast<uint32_t *>(dst + (size_t) row * K + e0) = q0 | (q1 << 8) | (q2 << 16) | (q3 << 24);
                if ((lane & 7) == 0) reinterpret_cast<float *>(dst + (size_t) MK_MAXTOK * K)[row * (K >> 5) + (e0 >> 5)] = __half2float(__float2half_rn(amax * (1.0f / 127.0f)));
            }
        }
        __syncthreads();
    }
}

// quantise f32 rows (no LayerNorm) into the actq format: one warp per (row, 128 values); followed by a grid barrier
template <int WT>
__device__ __noinline__ void mk_q8_rows(const MkArgs & a, const float * src, int K, uint8_t * dst) {
    const int lane = threadIdx.x & 31, wg = blockIdx.x * MK_WARPS + (threadIdx.x >> 5), nwg = gridDim.x * MK_WARPS;
    const int cpr = K >> 7;
    int row = wg / cpr, ch = wg - row * cpr;
    const int drow = nwg / cpr, dch = nwg - drow * cpr;
    for (; row < a.n_tok; ) {
        const int e0 = ch * 128 + lane * 4;
        const float4 y = __ldcg(reinterpret_cast<const float4 *>(src + (size_t) row * K + e0));
        if (WT == WT_F16) {
            const __half2 h0 = __floats2half2_rn(y.x, y.y), h1 = __floats2half2_rn(y.z, y.w);
            *reinterpret_cast<uint2 *>(reinterpret_cast<__half *>(dst) + (size_t) row * K + e0) = make_uint2(*reinterpret_cast<const uint32_t *>(&h0), *reinterpret_cast<const uint32_t *>(&h1));
        } else {
            float amax = fmaxf(fmaxf(fabsf(y.x), fabsf(y.y)), fmaxf(fabsf(y.z), fabsf(y.w)));
            amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 1));
            amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 2));
            amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 4));
            const float id = (amax != 0.0f) ? __fdividef(127.0f, amax) : 0.0f;
            const uint32_t q0 = (uint32_t) __float2int_rn(y.x * id) & 0xffu, q1 = (uint32_t) __float2int_rn(y.y * id) & 0xffu;
            const uint32_t q2 = (uint32_t) __float2int_rn(y.z * id) & 0xffu, q3 = (uint32_t) __float2int_rn(y.w * id) & 0xffu;
            *reinterpret_cast<uint32_t *>(dst + (size_t) row * K + e0) = q0 | (q1 << 8) | (q2 << 16) | (q3 << 24);
            if ((lane & 7) == 0) reinterpret_cast<float *>(dst + (size_t) MK_MAXTOK * K)[row * (K >> 5) + (e0 >> 5)] = __half2float(__float2half_rn(amax * (1.0f / 127.0f)));
        }
        row += drow; ch += dch; if (ch >= cpr) { ch -= cpr; ++row; }
    }
}

// TMA bulk copy global -> shared, completion on an mbarrier
__device__ __forceinline__ void bulk_g2s(void * dst, const void * src, uint32_t bytes, uint64_t * bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
#define SM_MBAR (reinterpret_cast<uint64_t *>(mk_smem + MK_OFF_FLAG + 32))

struct MkEpi {
    const float * bias = nullptr, * scale = nullptr; int act = 0; const float * res = nullptr; float * out = nullptr;
    __half * kc = nullptr, * vc = nullptr; int kv_d = 0;
    uint8_t * qout = nullptr;            // PAIR epilogue: quantised rows [64][N] (+ block scales) for the next GEMV
};

// L2 prefetch of weight tiles for a later GEMV phase (tile-major: the records of a tile are contiguous); every tile once per grid
__device__ __noinline__ void mk_prefetch_w(const QMat & W) {
    if (threadIdx.x != MK_THREADS - 32) return;
    const int n_tiles = (W.N + 15) >> 4;
    const uint32_t tile_bytes = (uint32_t) (W.K / wt_tm_rec_k(W.type)) * wt_tm_rec_bytes(W.type);
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
        l2_prefetch(reinterpret_cast<const uint8_t *>(W.base) + (size_t) tile * tile_bytes, tile_bytes);
}

#define MK_FINE(j) do { if (fb >= 0 && blockIdx.x == 0 && threadIdx.x == 0) a.trace[fb + (j)] = clock64(); } while (0)

__device__ __forceinline__ void cp_async16(uint32_t saddr, const void * g) { asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(saddr), "l"(g) : "memory"); }
__device__ __forceinline__ void cp_async16_hint(uint32_t saddr, const void * g, uint64_t pol) { asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" :: "r"(saddr), "l"(g), "l"(pol) : "memory"); }
__device__ __forceinline__ uint64_t policy_evict_first() { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// k-loop of one iteration of mk_gemv for NV tile slots per warp (compile-time: predicated slots made the register allocator spill);
// leaves the split-K partials of slots j0 .. j0+NV-1 in SM_RED.  KS warps split K of a slot, UB records per slot are requested together
// (always 16 / 3: the order in which a warp adds up its blocks is part of the result, and a row must not depend on how many tiles its
// CTA happens to hold).  Four slots in one call hold 12 records + 32 accumulators per lane and spilled (7.4 us per iteration against
// 2.4 us for three slots): an iteration with four slots makes two calls of two slots, same partials, no spills.
template <int WT, int NV, int TU, int KS, int UB>
__device__ __forceinline__ void mk_gemv_kloop(const MkArgs & a, const uint8_t * wbase, int tile0, int tstep, int nrec, int nb, int SW, int nt, int NH, bool & staged, uint32_t stage_parity, int fb, int j0) {
    constexpr int REC = (WT == WT_Q4_0) ? 288 : (WT == WT_Q5_0 ? 352 : (WT == WT_Q8_0 ? 544 : 512));
    constexpr int QSB = (WT == WT_Q8_0) ? 512 : 256;
    constexpr int RLD = 17;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, c = lane & 3, ks = warp & (KS - 1);
    float acc[NV][2][4];
#pragma unroll
    for (int j = 0; j < NV; ++j)
#pragma unroll
        for (int h = 0; h < 2; ++h) { acc[j][h][0] = acc[j][h][1] = acc[j][h][2] = acc[j][h][3] = 0.0f; }
    for (int kb = ks; kb < nrec; kb += KS * UB) {
        uint4 wq[NV][UB]; uint2 wh[NV][UB]; uint32_t wd[NV][UB];
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            {
                const uint8_t * tb = wbase + (size_t) ((tile0 + ((j0 + j) / TU) * tstep) * TU + (j0 + j) % TU) * nrec * REC;
#pragma unroll
                for (int u = 0; u < UB; ++u) {
                    const uint8_t * rec = tb + (size_t) min(kb + u * KS, nrec - 1) * REC;
                    if (WT == WT_F16 || WT == WT_Q8_0) wq[j][u] = __ldg(reinterpret_cast<const uint4 *>(rec) + lane);
                    else { const uint2 q2 = __ldg(reinterpret_cast<const uint2 *>(rec) + lane); wq[j][u].x = q2.x; wq[j][u].y = q2.y; }
                    if (WT == WT_Q5_0) wh[j][u] = __ldg(reinterpret_cast<const uint2 *>(rec + QSB) + g);
                    if (WT != WT_F16)  wd[j][u] = __ldg(reinterpret_cast<const uint32_t *>(rec + QSB + (WT == WT_Q5_0 ? 64 : 0)) + g);
                }
            }
        }
        if (!staged) { mbar_wait(SM_MBAR, stage_parity); staged = true; MK_FINE(1); }     // the rows have landed (this warp's weights are in flight)
#pragma unroll
        for (int u = 0; u < UB; ++u) {
            const int b = kb + u * KS;
            if (b < nrec) {
                uint32_t bf[2][2]; float dx[2][2];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const bool ok = h * 8 + g < nt;
                    bf[h][0] = ok ? SM_XQ[(h * 8 + g) * SW + b * 8 + c] : 0u;
                    bf[h][1] = ok ? SM_XQ[(h * 8 + g) * SW + b * 8 + 4 + c] : 0u;
                    dx[h][0] = dx[h][1] = 0.0f;
                    if (WT != WT_F16) { dx[h][0] = SM_XD[min(h * 8 + 2 * c, nt - 1) * nb + b]; dx[h][1] = SM_XD[min(h * 8 + 2 * c + 1, nt - 1) * nb + b]; }
                }
#pragma unroll
                for (int j = 0; j < NV; ++j) {
                    {
                        uint32_t af[4];
                        float dw0 = 0.0f, dw1 = 0.0f;
                        if (WT == WT_F16 || WT == WT_Q8_0) { af[0] = wq[j][u].x; af[1] = wq[j][u].y; af[2] = wq[j][u].z; af[3] = wq[j][u].w; }
                        else {
                            uint32_t lo0 = wq[j][u].x & 0x0F0F0F0Fu, hi0 = (wq[j][u].x >> 4) & 0x0F0F0F0Fu, lo1 = wq[j][u].y & 0x0F0F0F0Fu, hi1 = (wq[j][u].y >> 4) & 0x0F0F0F0Fu;
                            if (WT == WT_Q5_0) {
                                lo0 |= spread4_to_bit4(wh[j][u].x >> (4 * c)); hi0 |= spread4_to_bit4(wh[j][u].x >> (16 + 4 * c));
                                lo1 |= spread4_to_bit4(wh[j][u].y >> (4 * c)); hi1 |= spread4_to_bit4(wh[j][u].y >> (16 + 4 * c));
                                af[0] = __vsub4(lo0, 0x10101010u); af[2] = __vsub4(hi0, 0x10101010u);
                                af[1] = __vsub4(lo1, 0x10101010u); af[3] = __vsub4(hi1, 0x10101010u);
                            } else {
                                af[0] = __vsub4(lo0, 0x08080808u); af[2] = __vsub4(hi0, 0x08080808u);
                                af[1] = __vsub4(lo1, 0x08080808u); af[3] = __vsub4(hi1, 0x08080808u);
                            }
                        }
                        if (WT != WT_F16) {
                            dw0 = __half2float(__ushort_as_half((unsigned short) (wd[j][u] & 0xffffu)));
                            dw1 = __half2float(__ushort_as_half((unsigned short) (wd[j][u] >> 16)));
                        }
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            if (h < NH) {
                                if (WT == WT_F16) mma_f16_16816(acc[j][h], af, bf[h][0], bf[h][1]);
                                else {
                                    int dd[4]; mma_s8_16832(dd, af, bf[h][0], bf[h][1]);
                                    acc[j][h][0] = fmaf(dw0 * dx[h][0], (float) dd[0], acc[j][h][0]);
                                    acc[j][h][1] = fmaf(dw0 * dx[h][1], (float) dd[1], acc[j][h][1]);
                                    acc[j][h][2] = fmaf(dw1 * dx[h][0], (float) dd[2], acc[j][h][2]);
                                    acc[j][h][3] = fmaf(dw1 * dx[h][1], (float) dd[3], acc[j][h][3]);
                                }
                            }
                        }
                    }
                }
            }
        }
    }
    if (!staged) { mbar_wait(SM_MBAR, stage_parity); staged = true; }
    // split-K partials -> smem: red[tile slot j][warp][row][batch row]
#pragma unroll
    for (int j = 0; j < NV; ++j)
        {
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int i = 0; i < 4; ++i) SM_RED[(((j0 + j) * MK_WARPS + warp) * 16 + g + (i >> 1) * 8) * RLD + h * 8 + 2 * c + (i & 1)] = acc[j][h][i];
        }
}

// y[t][n] = act((W[n,:] . x[t,:] + bias[n]) * scale[n]) + res[t][n].
// With up to 64 rows the activations (64 x K bytes) outweigh a weight tile (16 x K x ~0.7 bytes), so the work is cut by ROW GROUP
// first: CTA c serves the 16 rows of group c % NGc (8 rows for the widest F16 matrices), stages only those rows (one TMA bulk copy
// per row + one for its block scales) and walks the weight tiles its group-mates do not take, up to four tiles per iteration with their
// loads issued together.  The 16 warps split K of a tile; every weight block is decoded once and multiplied with both 8-row halves
// (mma.sync.m16n8k32.s8); partials are reduced through smem.  x: quantised rows in global memory (actq format).  A weight tile is
// read by the NGc CTAs of the different groups (from L2, prefetched a phase ahead).
// Latency hiding inside a phase (each phase is a chain barrier -> stage -> weights -> reduce -> epilogue of a few microseconds):
// the weight loads of the first iteration are issued BEFORE the staged rows are waited for, the epilogue's bias / scale / residual
// operands are fetched before the k-loop, tiles are dealt out evenly over the iterations and slots without a tile cost nothing.
// PAIR (FC1): tiles are dealt out in pairs = 32 consecutive output features, and the epilogue writes GELU(y) straight into the
// quantised-row format FC2 consumes (one warp = one Q8_0 block of one row) -- no f32 round trip, no extra phase.
template <int WT, bool PAIR>
__device__ __noinline__ void mk_gemv(const MkArgs & a, const QMat & W, const uint8_t * x, const MkEpi & e, int fb = -1) {
    constexpr int REC = (WT == WT_Q4_0) ? 288 : (WT == WT_Q5_0 ? 352 : (WT == WT_Q8_0 ? 544 : 512));
    constexpr int QSB = (WT == WT_Q8_0) ? 512 : 256;             // bytes of the qs part of a record
    constexpr int RK = (WT == WT_F16) ? 16 : 32;                 // K values per record
    constexpr int UB = 3;                                        // records whose loads are issued together
    constexpr int TP = MK_TP;                                    // tile slots per iteration
    constexpr int TU = PAIR ? 2 : 1;                             // tiles per unit of distribution
    constexpr int RLD = 17;                                      // row stride of the reduction buffer (16 batch rows + pad)
    const int N = W.N, K = W.K;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, c = lane & 3;
    const int n_tiles = (N + 15) >> 4, nrec = K / RK, nb = K >> 5;
    const int rowb = (WT == WT_F16) ? K * 2 : K, SW = (rowb >> 2) + 4;
    const int RG = (16 * (rowb + 16) <= MK_MAXTOK * (MK_ROWB + 16)) ? 16 : 8;   // rows per CTA (what fits the staging area)
    const int NGc = (a.n_tok + RG - 1) / RG;
    const int grp = blockIdx.x % NGc, ci = blockIdx.x / NGc, cg = ((int) gridDim.x - grp + NGc - 1) / NGc;   // row group, index / count of its CTAs
    const int t_base = grp * RG, nt = min(RG, a.n_tok - t_base), NH = (nt + 7) >> 3;
    const int n_units = n_tiles / TU;                            // PAIR: N is a multiple of 32
    const int upc = ci < n_units ? (n_units - ci + cg - 1) / cg : 0;            // units of this CTA: ci, ci + cg, ...
    const int n_it = (upc * TU + TP - 1) / TP, upi = n_it ? (upc + n_it - 1) / n_it : 0;   // iterations, units per iteration (even split)
    MK_FINE(0);
    uint32_t stage_parity;
    {   // stage the rows of the group: one TMA bulk copy per row (+ one for its block scales), completion on an mbarrier that every
        // thread waits for on its own, after it has issued its first weight loads
        const int ph = SM_FLAG[4];                               // staging round (mbarrier phase parity)
        __syncthreads();                                         // everybody has read `ph`; the previous users of the buffer are done
        stage_parity = (uint32_t) ph & 1u;
        if (tid == 0) {
            SM_FLAG[4] = ph + 1;
            mbar_arrive_expect_tx(SM_MBAR, (uint32_t) nt * (uint32_t) (rowb + (WT == WT_F16 ? 0 : nb * 4)));
        }
        if (tid < nt) {
            asm volatile("fence.proxy.async;" ::: "memory");     // rows were written with ordinary stores (by other CTAs, before the grid barrier)
            bulk_g2s(SM_XQ + tid * SW, x + (size_t) (t_base + tid) * rowb, (uint32_t) rowb, SM_MBAR);
            if (WT != WT_F16)
                bulk_g2s(SM_XD + tid * nb, reinterpret_cast<const float *>(x + (size_t) MK_MAXTOK * K) + (size_t) (t_base + tid) * nb, (uint32_t) nb * 4, SM_MBAR);
        }
    }
    bool staged = false;
    for (int it = 0; it < n_it; ++it) {
        const int u0 = it * upi, nv = TU * (min(upc, u0 + upi) - u0);           // valid tile slots of this iteration (CTA-uniform)
        {
            const uint8_t * wbase = reinterpret_cast<const uint8_t *>(W.base);
            const int unit0 = ci + u0 * cg;                      // slot j holds tile (unit0 + (j / TU) * cg) * TU + j % TU
            switch (nv) {
                case 1: if (!PAIR) { mk_gemv_kloop<WT, 1, TU, 16, 3>(a, wbase, unit0, cg, nrec, nb, SW, nt, NH, staged, stage_parity, fb, 0); break; }
                case 2: mk_gemv_kloop<WT, 2, TU, 16, 3>(a, wbase, unit0, cg, nrec, nb, SW, nt, NH, staged, stage_parity, fb, 0); break;
                case 3: if (!PAIR) { mk_gemv_kloop<WT, 3, TU, 16, 3>(a, wbase, unit0, cg, nrec, nb, SW, nt, NH, staged, stage_parity, fb, 0); break; }
                default:                                         // four slots: two passes of two (see mk_gemv_kloop)
                    mk_gemv_kloop<WT, 2, TU, 16, 3>(a, wbase, unit0, cg, nrec, nb, SW, nt, NH, staged, stage_parity, fb, 0);
                    mk_gemv_kloop<WT, 2, TU, 16, 3>(a, wbase, unit0, cg, nrec, nb, SW, nt, NH, staged, stage_parity, fb, 2);
                    break;
            }
        }
        if (it == 0) MK_FINE(2);
        float pb[2], ps[2], pr[2];
        // epilogue operands of this thread's (at most two) outputs, fetched before the partials go through smem (their latency hides behind the reduction)
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            int jj, tl, rl;
            if (PAIR) { const int idx = warp + MK_WARPS * i; jj = (idx >> 4) * 2 + (lane >> 4); tl = idx & 15; rl = lane & 15; }
            else      { const int o = tid + MK_THREADS * i; jj = o >> 8; tl = (o & 255) >> 4; rl = o & 15; }
            const int row = ((ci + (u0 + jj / TU) * cg) * TU + jj % TU) * 16 + rl;
            const bool ok = jj < nv && tl < nt && row < N;
            pb[i] = (ok && e.bias) ? __ldg(e.bias + row) : 0.0f;
            ps[i] = (ok && e.scale) ? __ldg(e.scale + row) : 1.0f;
            pr[i] = (ok && e.res) ? __ldcg(e.res + (size_t) (t_base + tl) * N + row) : 0.0f;
        }
        __syncthreads();
        if (it == 0) MK_FINE(3);
#pragma unroll
        for (int i = 0; i < 2; ++i) {                            // epilogue: up to TP tiles x 16 batch rows x 16 weight rows, two outputs per thread
            int jj, tl, rl;
            if (PAIR) { const int idx = warp + MK_WARPS * i; jj = (idx >> 4) * 2 + (lane >> 4); tl = idx & 15; rl = lane & 15; }
            else      { const int o = tid + MK_THREADS * i; jj = o >> 8; tl = (o & 255) >> 4; rl = o & 15; }
            const int row = ((ci + (u0 + jj / TU) * cg) * TU + jj % TU) * 16 + rl, t = t_base + tl;
            const bool ok = jj < nv && tl < nt && row < N;
            float v = 0.0f;
            if (ok) {
#pragma unroll
                for (int w = 0; w < MK_WARPS; ++w) v += SM_RED[((jj * MK_WARPS + w) * 16 + rl) * RLD + tl];
                v = (v + pb[i]) * ps[i];
                if (e.act == 1) v = gelu_ref_f16(v);
                v += pr[i];
            }
            if (PAIR) {
                // (row slot pair, token) is warp-uniform: the 32 lanes hold one Q8_0 block of the FC2 input row t (features row0 .. row0+31)
                const int idx = warp + MK_WARPS * i;
                if ((idx >> 4) * 2 < nv && tl < nt) mk_store_q<WT>(e.qout, N, t, row - lane, lane, v);
            } else if (ok) {
                if (e.out) e.out[(size_t) t * N + row] = v;
                if (e.kc && row >= e.kv_d) {
                    const size_t cell = a.cell[t];
                    if (row < 2 * e.kv_d) e.kc[cell * e.kv_d + (row - e.kv_d)] = __float2half_rn(v);
                    else                  e.vc[cell * e.kv_d + (row - 2 * e.kv_d)] = __float2half_rn(v);
                }
            }
        }
        __syncthreads();                                         // SM_RED is rewritten by the next iteration
        if (it == 0) MK_FINE(4);
    }
    if (!staged) mbar_wait(SM_MBAR, stage_parity);               // a CTA without tiles still waits for its copies before the area is reused
    MK_FINE(5);
}

// Logits (token-embedding matrix, N = n_vocab, K = d): one WARP per 16-row weight tile over the whole K -- no split-K, no reduction through
// shared memory, no CTA barrier per iteration (the split-K form needs 22 iterations of ~7 us for the 3242 tiles of large-v3: 220 us for a
// 45.6 MB stream).  A lane requests 8 records of its tile at once; the rows of the CTA's group are staged as in mk_gemv.  A row's result
// is one accumulator over the blocks in ascending order: independent of the batch.
template <int WT>
__device__ __noinline__ void mk_logits_warp(const MkArgs & a, const QMat & W, const uint8_t * x, float * out) {
    constexpr int REC = (WT == WT_Q4_0) ? 288 : (WT == WT_Q5_0 ? 352 : (WT == WT_Q8_0 ? 544 : 512));
    constexpr int QSB = (WT == WT_Q8_0) ? 512 : 256;
    constexpr int RK = (WT == WT_F16) ? 16 : 32;
    constexpr int UB = 8;
    const int N = W.N, K = W.K;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, c = lane & 3;
    const int n_tiles = (N + 15) >> 4, nrec = K / RK, nb = K >> 5;
    const int rowb = (WT == WT_F16) ? K * 2 : K, SW = (rowb >> 2) + 4;
    const int RG = (16 * (rowb + 16) <= MK_MAXTOK * (MK_ROWB + 16)) ? 16 : 8;
    const int NGc = (a.n_tok + RG - 1) / RG;
    const int grp = blockIdx.x % NGc, ci = blockIdx.x / NGc, cg = ((int) gridDim.x - grp + NGc - 1) / NGc;
    const int t_base = grp * RG, nt = min(RG, a.n_tok - t_base), NH = (nt + 7) >> 3;
    uint32_t stage_parity;
    {
        const int ph = SM_FLAG[4];
        __syncthreads();
        stage_parity = (uint32_t) ph & 1u;
        if (tid == 0) {
            SM_FLAG[4] = ph + 1;
            mbar_arrive_expect_tx(SM_MBAR, (uint32_t) nt * (uint32_t) (rowb + (WT == WT_F16 ? 0 : nb * 4)));
        }
        if (tid < nt) {
            asm volatile("fence.proxy.async;" ::: "memory");
            bulk_g2s(SM_XQ + tid * SW, x + (size_t) (t_base + tid) * rowb, (uint32_t) rowb, SM_MBAR);
            if (WT != WT_F16)
                bulk_g2s(SM_XD + tid * nb, reinterpret_cast<const float *>(x + (size_t) MK_MAXTOK * K) + (size_t) (t_base + tid) * nb, (uint32_t) nb * 4, SM_MBAR);
        }
    }
    bool staged = false;
    const uint8_t * wbase = reinterpret_cast<const uint8_t *>(W.base);
    for (int tile = ci * MK_WARPS + warp; tile < n_tiles; tile += cg * MK_WARPS) {
        float acc[2][4];
#pragma unroll
        for (int h = 0; h < 2; ++h) { acc[h][0] = acc[h][1] = acc[h][2] = acc[h][3] = 0.0f; }
        const uint8_t * tb = wbase + (size_t) tile * nrec * REC;
        for (int kb = 0; kb < nrec; kb += UB) {
            uint4 wq[UB]; uint2 wh[UB]; uint32_t wd[UB];
#pragma unroll
            for (int u = 0; u < UB; ++u) {
                const uint8_t * rec = tb + (size_t) min(kb + u, nrec - 1) * REC;
                if (WT == WT_F16 || WT == WT_Q8_0) wq[u] = __ldg(reinterpret_cast<const uint4 *>(rec) + lane);
                else { const uint2 q2 = __ldg(reinterpret_cast<const uint2 *>(rec) + lane); wq[u].x = q2.x; wq[u].y = q2.y; }
                if (WT == WT_Q5_0) wh[u] = __ldg(reinterpret_cast<const uint2 *>(rec + QSB) + g);
                if (WT != WT_F16)  wd[u] = __ldg(reinterpret_cast<const uint32_t *>(rec + QSB + (WT == WT_Q5_0 ? 64 : 0)) + g);
            }
            if (!staged) { mbar_wait(SM_MBAR, stage_parity); staged = true; }
#pragma unroll
            for (int u = 0; u < UB; ++u) {
                const int b = kb + u;
                if (b < nrec) {
                    uint32_t af[4];
                    float dw0 = 0.0f, dw1 = 0.0f;
                    if (WT == WT_F16 || WT == WT_Q8_0) { af[0] = wq[u].x; af[1] = wq[u].y; af[2] = wq[u].z; af[3] = wq[u].w; }
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
                    if (WT != WT_F16) {
                        dw0 = __half2float(__ushort_as_half((unsigned short) (wd[u] & 0xffffu)));
                        dw1 = __half2float(__ushort_as_half((unsigned short) (wd[u] >> 16)));
                    }
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        if (h < NH) {
                            const bool ok = h * 8 + g < nt;
                            const uint32_t b0 = ok ? SM_XQ[(h * 8 + g) * SW + b * 8 + c] : 0u, b1 = ok ? SM_XQ[(h * 8 + g) * SW + b * 8 + 4 + c] : 0u;
                            if (WT == WT_F16) mma_f16_16816(acc[h], af, b0, b1);
                            else {
                                const float dx0 = SM_XD[min(h * 8 + 2 * c, nt - 1) * nb + b], dx1 = SM_XD[min(h * 8 + 2 * c + 1, nt - 1) * nb + b];
                                int dd[4]; mma_s8_16832(dd, af, b0, b1);
                                acc[h][0] = fmaf(dw0 * dx0, (float) dd[0], acc[h][0]);
                                acc[h][1] = fmaf(dw0 * dx1, (float) dd[1], acc[h][1]);
                                acc[h][2] = fmaf(dw1 * dx0, (float) dd[2], acc[h][2]);
                                acc[h][3] = fmaf(dw1 * dx1, (float) dd[3], acc[h][3]);
                            }
                        }
                    }
                }
            }
        }
        // C fragment: acc[h][i] = weight row g + 8*(i>>1) of the tile, batch row h*8 + 2c + (i&1)
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int row = tile * 16 + g + (i >> 1) * 8, tl = h * 8 + 2 * c + (i & 1);
                if (tl < nt && row < N) out[(size_t) (t_base + tl) * N + row] = acc[h][i];
            }
    }
    if (!staged) mbar_wait(SM_MBAR, stage_parity);
}

// ---- attention -----------------------------------------------------------------------------------------------------------
// the 16 query values of this lane's quarter of head h of row t, rounded to f16 (ggml_flash_attn_ext converts Q to f16:
// ggml-cpu/ops.cpp:8560-8571)
// Lane quarter r of a key owns dims {8r..8r+7} and {32+8r..32+8r+7} of the head: the 4 lanes of a key then cover 64 contiguous
// bytes per load instruction (whole 32-byte sectors; 16 dims in a row per lane would use half of every sector it touches).
__device__ __forceinline__ void load_q16(const float * qh, int r, float (&qv)[16]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float4 v = __ldcg(reinterpret_cast<const float4 *>(qh + 8 * r + (i >> 1) * 32) + (i & 1));
        qv[4 * i]     = __half2float(__float2half_rn(v.x)); qv[4 * i + 1] = __half2float(__float2half_rn(v.y));
        qv[4 * i + 2] = __half2float(__float2half_rn(v.z)); qv[4 * i + 3] = __half2float(__float2half_rn(v.w));
    }
}

// One lane owns one key per step: it holds 16 of the 64 dims of that key's K and V rows (the 4 lanes of a key share the score) and
// keeps its OWN running max / sum / output over the keys it has seen -- no cross-lane traffic per key beyond the 2-step score
// reduction.  The 8 key slots of a warp are merged once at the end (warp_merge), the warps of a CTA through shared memory.
struct KV4 { uint4 k0, k1, v0, v1; };

__device__ __forceinline__ float dot16(const uint4 & k0, const uint4 & k1, const float (&q)[16]) {
    float s = 0.0f;
    const __half2 * h0 = reinterpret_cast<const __half2 *>(&k0), * h1 = reinterpret_cast<const __half2 *>(&k1);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 f = __half22float2(h0[i]), f2 = __half22float2(h1[i]);
        s = fmaf(f.x, q[2 * i], s); s = fmaf(f.y, q[2 * i + 1], s); s = fmaf(f2.x, q[8 + 2 * i], s); s = fmaf(f2.y, q[8 + 2 * i + 1], s);
    }
    return s;
}

__device__ __forceinline__ float dot16s(const uint4 & k0, const uint4 & k1, const float * q) {   // q: 16 floats in shared memory
    float s = 0.0f;
    const __half2 * h0 = reinterpret_cast<const __half2 *>(&k0), * h1 = reinterpret_cast<const __half2 *>(&k1);
#pragma unroll
    for (int i = 0; i < 4; i += 2) {
        const float4 qa = *reinterpret_cast<const float4 *>(q + 2 * i), qb = *reinterpret_cast<const float4 *>(q + 8 + 2 * i);
        const float2 f0 = __half22float2(h0[i]), f1 = __half22float2(h0[i + 1]), g0 = __half22float2(h1[i]), g1 = __half22float2(h1[i + 1]);
        s = fmaf(f0.x, qa.x, s); s = fmaf(f0.y, qa.y, s); s = fmaf(f1.x, qa.z, s); s = fmaf(f1.y, qa.w, s);
        s = fmaf(g0.x, qb.x, s); s = fmaf(g0.y, qb.y, s); s = fmaf(g1.x, qb.z, s); s = fmaf(g1.y, qb.w, s);
    }
    return s;
}

struct LaneAcc { float m, l, o[16]; };
__device__ __forceinline__ void lane_init(LaneAcc & A) { A.m = -INFINITY; A.l = 0.0f;
#pragma unroll
    for (int i = 0; i < 16; ++i) A.o[i] = 0.0f; }

// sc: score of this lane's key (already reduced over the 4 lanes of the key and scaled); masked keys must not call this
__device__ __forceinline__ void lane_update(LaneAcc & A, float sc, const uint4 & v0, const uint4 & v1) {
    const float mn = fmaxf(A.m, sc);
    const float resc = __expf(A.m - mn), p = __expf(sc - mn);
    A.l = fmaf(A.l, resc, p); A.m = mn;
    const __half2 * h0 = reinterpret_cast<const __half2 *>(&v0), * h1 = reinterpret_cast<const __half2 *>(&v1);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 f = __half22float2(h0[i]), f2 = __half22float2(h1[i]);
        A.o[2 * i]     = fmaf(A.o[2 * i],     resc, p * f.x);  A.o[2 * i + 1] = fmaf(A.o[2 * i + 1], resc, p * f.y);
        A.o[8 + 2 * i] = fmaf(A.o[8 + 2 * i], resc, p * f2.x); A.o[9 + 2 * i] = fmaf(A.o[9 + 2 * i], resc, p * f2.y);
    }
}
// merge the 8 key slots of the warp (lanes with equal lane & 3); afterwards every lane holds the warp totals of its quarter
__device__ __forceinline__ void warp_merge(LaneAcc & A) {
#pragma unroll
    for (int off = 4; off < 32; off <<= 1) {
        const float m2 = __shfl_xor_sync(0xffffffffu, A.m, off), l2 = __shfl_xor_sync(0xffffffffu, A.l, off);
        const float mn = fmaxf(A.m, m2);
        const float w1 = (A.m > -INFINITY) ? __expf(A.m - mn) : 0.0f, w2 = (m2 > -INFINITY) ? __expf(m2 - mn) : 0.0f;
        A.l = A.l * w1 + l2 * w2; A.m = mn;
#pragma unroll
        for (int i = 0; i < 16; ++i) { const float o2 = __shfl_xor_sync(0xffffffffu, A.o[i], off); A.o[i] = A.o[i] * w1 + o2 * w2; }
    }
}
__device__ __forceinline__ void part_store(float * part, const LaneAcc & A, int lane) {    // lanes 0..3 write the warp partial
    if (lane < 4) {
        if (lane == 0) { part[0] = A.m; part[1] = A.l; }
#pragma unroll
        for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4 *>(part + 4 + lane * 8 + (i >> 3) * 32 + (i & 7)) = make_float4(A.o[i], A.o[i + 1], A.o[i + 2], A.o[i + 3]);
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

// self-attention over the paged cache (whisper.cpp:2603-2625).  Item = (row, head); a pair of warps takes one (8 items per CTA at a time:
// with 64 rows x 20 heads every item has its warps in the first round).  A lane owns a key quarter; the cells and the K / V pieces of
// four key groups are requested together, so a row with 200 keys costs a handful of memory round trips instead of one per group.
template <int WT>
__device__ __noinline__ void mk_attn_self(const MkArgs & a, const MkLayer & L) {
    constexpr int WPI = 2, SB = 2, CV = 16;                      // warps per item, key groups in flight together (double-buffered), cell indices per lane requested at once
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, grp = warp / WPI, hw = warp % WPI;
    const int d = a.d, H = a.n_head, n_pairs = a.n_tok * H;
    const int kslot = lane >> 2, r = lane & 3;
    for (int p = blockIdx.x * (MK_WARPS / WPI) + grp; p < n_pairs; p += gridDim.x * (MK_WARPS / WPI)) {
        const int t = p / H, h = p - t * H;
        const int nk = a.nkv[t];
        const int * cells = a.idx + (size_t) t * a.ld_idx;
        float q[16];
        load_q16(a.qkv + (size_t) t * 3 * d + h * 64, r, q);
        LaneAcc A; lane_init(A);
        // Latency chain of an item: cell indices -> K / V rows -> arithmetic.  The indices of 256 keys (16 per lane) are requested at once,
        // then the K / V of two key groups at a time, the next two in flight while the current two are used: one round trip for the
        // indices plus one exposed K / V round trip per item instead of two per 64 keys.  Every lane still meets its keys in ascending order.
        for (int kb = 0; kb < nk; kb += WPI * 8 * CV) {
            int cv[CV];
#pragma unroll
            for (int i = 0; i < CV; ++i) { const int k = kb + hw * 8 + i * WPI * 8 + kslot; cv[i] = (k < nk) ? __ldg(cells + k) : -1; }
            KV4 f[2][SB];
            auto fetch = [&](KV4 (&dst)[SB], int c0, int c1) {
                const int cc[SB] = { c0, c1 };
#pragma unroll
                for (int u = 0; u < SB; ++u) {
                    dst[u].k0 = dst[u].k1 = dst[u].v0 = dst[u].v1 = make_uint4(0, 0, 0, 0);
                    if (cc[u] >= 0) {
                        const size_t off = (size_t) cc[u] * d + h * 64 + r * 8;
                        const uint4 * kp = reinterpret_cast<const uint4 *>(L.kc + off), * vp = reinterpret_cast<const uint4 *>(L.vc + off);
                        dst[u].k0 = __ldcg(kp); dst[u].k1 = __ldcg(kp + 4); dst[u].v0 = __ldcg(vp); dst[u].v1 = __ldcg(vp + 4);
                    }
                }
            };
            fetch(f[0], cv[0], cv[1]);
#pragma unroll
            for (int g = 0; g < CV / SB; ++g) {
                const bool more = (g + 1 < CV / SB) && (kb + (g + 1) * SB * WPI * 8 < nk);       // same for both warps of the item
                if (more) fetch(f[(g + 1) & 1], cv[(2 * g + 2) % CV], cv[(2 * g + 3) % CV]);
#pragma unroll
                for (int u = 0; u < SB; ++u) {
                    float sc = dot16(f[g & 1][u].k0, f[g & 1][u].k1, q);
                    sc += __shfl_xor_sync(0xffffffffu, sc, 1);
                    sc += __shfl_xor_sync(0xffffffffu, sc, 2);
                    if (cv[2 * g + u] >= 0) lane_update(A, sc, f[g & 1][u].v0, f[g & 1][u].v1);
                }
                if (!more) break;
            }
        }
        warp_merge(A);
        part_store(SM_PART + (grp * WPI + hw) * MK_PART, A, lane);
        bar_named(1 + grp, WPI * 32);
        {
            float M, Lsum;
            const float o = attn_merge(SM_PART + grp * WPI * MK_PART, WPI, hw * 32 + lane, M, Lsum);
            mk_store_q<WT>(a.actq, d, t, h * 64 + hw * 32, lane, (Lsum > 0.0f) ? __fdividef(o, Lsum) : 0.0f);
        }
        bar_named(1 + grp, WPI * 32);
    }
}

// cross-attention over the n_keys padded encoder positions, zero rows included (whisper.cpp:2688-2705).
// K and V of a window live HEAD-MAJOR in HBM ([layer][head][key][64], written that way by the cross GEMM's epilogue): the keys of one
// (row, head) pair are one contiguous 192 KB stream instead of 128-byte pieces 2.5 KB apart.
// Work unit = one (row, head) pair over ALL keys; CTA b takes the contiguous range of pairs [b*P/G, (b+1)*P/G).  Warp w, key slot s owns
// key 8w+s of every 128-key chunk; chunks are copied three ahead with cp.async into a 4-deep ring in shared memory (the GEMV staging
// area, idle here); every thread reads back only the 64 bytes it copied itself, so the ring needs no barrier -- it is an asynchronous
// extension of the register file (96 KB in flight per SM; a 6-deep ring was measured slower: the 40 KB of L1 it takes away cost the other
// phases more than the tail of this one gains).  The arithmetic of a pair (per-lane online softmax over its keys, 8 key slots
// merged per warp, 16 warps merged per CTA, all in fixed order) never depends on which other rows share the pass: a batch of 64 and a
// single row give bit-identical results.  (The phase is HBM-bound: a CTA with one pair more than its neighbour is not a cost, the
// others simply draw more bandwidth meanwhile.)
constexpr int MK_RING = 4, MK_RING_SLOT = MK_THREADS * 64, MK_OFF_QSM = MK_RING * MK_RING_SLOT;
static_assert(MK_OFF_QSM + 2 * 64 * 4 <= MK_OFF_PART, "the cp.async ring must fit below the attention partials");

__device__ __forceinline__ void cp_wait_ring() { asm volatile("cp.async.wait_group %0;" :: "n"(MK_RING - 1) : "memory"); }

template <int WT>
__device__ __noinline__ void mk_attn_cross(const MkArgs & a, const MkLayer & L) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int d = a.d, H = a.n_head, nchp = a.n_keys / MK_XKEYS;               // chunks per pair
    const unsigned P = (unsigned) (a.n_tok * H), G = min(gridDim.x, P);
    if (blockIdx.x >= G) return;
    const unsigned p0 = (blockIdx.x * P) / G, p1 = ((blockIdx.x + 1) * P) / G;
    // loader state: chunk being fetched
    unsigned pl = p0, il = 0;
    int jl = 0, tl_ = (int) p0 / H, hl = (int) p0 - tl_ * H;
    const size_t head_stride = (size_t) a.n_keys * 64;
    const size_t lane_off = (size_t) (warp * 8 + (lane >> 2)) * 64 + (lane & 3) * 8;       // key 8w+s of the chunk, dims {8r.., 32+8r..}
    const uint32_t ring = (uint32_t) __cvta_generic_to_shared(mk_smem) + tid * 16;
    const float kq_scale = a.kq_scale;
    // everything the copy path needs in registers (the asm statements are compiler memory barriers: values re-read from the argument
    // structs after each of them are dependent loads in front of every chunk); the slot of a row is fetched one row ahead
    const __half * const xk = L.xk, * const xv = L.xv;
    const int64_t slot_stride = a.slot_stride;
    const int * const slots = a.slot;
    const int n_tok = a.n_tok;
    const bool ef = (a.prefetch & 4) != 0;
    const uint64_t pol = policy_evict_first();
    int slot_cur = slots[tl_], slot_nxt = (tl_ + 1 < n_tok) ? slots[tl_ + 1] : 0;
    size_t pair_off = (size_t) slot_cur * slot_stride + (size_t) hl * head_stride + lane_off;
    auto issue = [&]() {                                           // copy the next chunk (if any) into ring slot il % 4; always commits a group
        if (pl < p1) {
            const size_t off = pair_off + (size_t) (jl * MK_XKEYS) * 64;
            const uint32_t sa = ring + (il % MK_RING) * MK_RING_SLOT;
            if (ef) {                                             // L2 evict-first: 500 MB of K / V per layer would otherwise flush the prefetched weights and the activations
                cp_async16_hint(sa, xk + off, pol); cp_async16_hint(sa + MK_THREADS * 16, xk + off + 32, pol);
                cp_async16_hint(sa + 2 * MK_THREADS * 16, xv + off, pol); cp_async16_hint(sa + 3 * MK_THREADS * 16, xv + off + 32, pol);
            } else {
                cp_async16(sa, xk + off); cp_async16(sa + MK_THREADS * 16, xk + off + 32);
                cp_async16(sa + 2 * MK_THREADS * 16, xv + off); cp_async16(sa + 3 * MK_THREADS * 16, xv + off + 32);
            }
            if (++jl == nchp) {
                jl = 0; ++pl;
                if (++hl == H) { hl = 0; ++tl_; slot_cur = slot_nxt; slot_nxt = (tl_ + 1 < n_tok) ? slots[tl_ + 1] : 0; }
                pair_off = (size_t) slot_cur * slot_stride + (size_t) hl * head_stride + lane_off;
            }
        }
        ++il;
        cp_commit();
    };
#pragma unroll
    for (int k = 0; k < MK_RING - 1; ++k) issue();
    int t = (int) p0 / H, h = (int) p0 - t * H;
    float * qsm = reinterpret_cast<float *>(mk_smem + MK_OFF_QSM);   // f16-rounded query of the pair: [2][64]
    // qsm is stored in lane order: position 16*r + 8*hi + i holds dim 32*hi + 8*r + i
    if (warp >= 2 && warp < 4) { const int dim = (warp - 2) * 32 + lane; qsm[((dim & 31) >> 3) * 16 + (dim >> 5) * 8 + (dim & 7)] = __half2float(__float2half_rn(__ldcg(a.q2 + (size_t) t * d + h * 64 + dim))); }
    __syncthreads();
    LaneAcc A; lane_init(A);
    int buf = 0, j = 0;
    unsigned ic = 0;
    for (unsigned p = p0; p < p1; ) {
        issue();
        cp_wait_ring();                                          // this thread's copies of chunk ic have landed
        const uint8_t * sl = mk_smem + (ic % MK_RING) * MK_RING_SLOT + tid * 16;
        ++ic;
        const uint4 k0 = *reinterpret_cast<const uint4 *>(sl), k1 = *reinterpret_cast<const uint4 *>(sl + MK_THREADS * 16);
        const uint4 v0 = *reinterpret_cast<const uint4 *>(sl + 2 * MK_THREADS * 16), v1 = *reinterpret_cast<const uint4 *>(sl + 3 * MK_THREADS * 16);
        float sc = dot16s(k0, k1, qsm + buf * 64 + (lane & 3) * 16);
        sc += __shfl_xor_sync(0xffffffffu, sc, 1);
        sc += __shfl_xor_sync(0xffffffffu, sc, 2);
        lane_update(A, sc * kq_scale, v0, v1);
        if (++j < nchp) continue;
        j = 0;                                                   // last chunk of the pair: merge the CTA and write the row's head
        warp_merge(A);
        part_store(SM_PART + (buf * MK_WARPS + warp) * MK_PART, A, lane);
        if (p + 1 < p1 && warp >= 2 && warp < 4) {               // query of the next pair
            int tn = t, hn = h;
            if (++hn == H) { hn = 0; ++tn; }
            const int dim = (warp - 2) * 32 + lane;
            qsm[(buf ^ 1) * 64 + ((dim & 31) >> 3) * 16 + (dim >> 5) * 8 + (dim & 7)] = __half2float(__float2half_rn(__ldcg(a.q2 + (size_t) tn * d + hn * 64 + dim)));
        }
        __syncthreads();
        if (warp < 2) {                                          // merge the 16 warp partials (one output dim per thread)
            const int dim = warp * 32 + lane;
            float M, Lsum;
            const float o = attn_merge(SM_PART + buf * MK_WARPS * MK_PART, MK_WARPS, dim, M, Lsum);
            mk_store_q<WT>(a.actq, d, t, h * 64 + warp * 32, lane, __fdividef(o, Lsum));
        }
        buf ^= 1;
        lane_init(A);
        ++p;
        if (++h == H) { h = 0; ++t; }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
}

#define MK_SYNC() do { MK_STAMP(); target += gridDim.x; mk_grid_sync(a, target); MK_STAMP(); } while (0)

template <int WT, bool TRACE>
__global__ void __launch_bounds__(MK_THREADS, 1)
k_decode_pass(const __grid_constant__ MkArgs a) {
    if (threadIdx.x == 0) { mbar_init(SM_MBAR, 1); SM_FLAG[4] = 0; mbar_fence_init(); }
    __syncthreads();
    unsigned long long target = a.bar_base;
    int n_stamp = 0;
    const int d = a.d;
    const bool pf_w = a.prefetch & 1;
    MK_STAMP();
    if (pf_w) mk_prefetch_w(a.layers[0].qkv);

    for (int l = 0; l < a.n_layer; ++l) {
        const MkLayer & L = a.layers[l];
        MkEpi e;
        // 1: LN -> quantised rows (whisper.cpp:2536-2543)
        mk_lnq<WT>(a, a.x, d, L.ln0_w, L.ln0_b, a.actq);
        if (pf_w) mk_prefetch_w(L.o);
        MK_SYNC();
        // 2: QKV + KV append (2545-2599)
        e = MkEpi(); e.bias = L.qkv_bias; e.scale = L.qkv_scale; e.out = a.qkv; e.kc = L.kc; e.vc = L.vc; e.kv_d = d;
        mk_gemv<WT, false>(a, L.qkv, a.actq, e, (TRACE && l == 1) ? 2048 + 0 : -1);
        MK_SYNC();
        // 3: self-attention (2603-2625) -> quantised rows for the O projection
        if (pf_w) mk_prefetch_w(L.cq);
        mk_attn_self<WT>(a, L);
        MK_SYNC();
        // 4: O + residual (2647-2659)
        if (pf_w) mk_prefetch_w(L.co);
        e = MkEpi(); e.bias = L.o_bias; e.res = a.x; e.out = a.x;
        mk_gemv<WT, false>(a, L.o, a.actq, e, (TRACE && l == 1) ? 2048 + 8 : -1);
        MK_SYNC();
        // 5: LN -> quantised rows
        mk_lnq<WT>(a, a.x, d, L.lnc_w, L.lnc_b, a.actq);
        if (pf_w) mk_prefetch_w(L.fc1);
        MK_SYNC();
        // 6: cross Q (2661-2681)
        e = MkEpi(); e.bias = L.cq_bias; e.out = a.q2;
        mk_gemv<WT, false>(a, L.cq, a.actq, e, (TRACE && l == 1) ? 2048 + 16 : -1);
        MK_SYNC();
        // 7: cross-attention (2688-2705)
        if (pf_w) mk_prefetch_w(L.fc2);
        mk_attn_cross<WT>(a, L);
        MK_SYNC();
        // 8: cross O + residual (2754-2766)
        e = MkEpi(); e.bias = L.co_bias; e.res = a.x; e.out = a.x;
        mk_gemv<WT, false>(a, L.co, a.actq, e, (TRACE && l == 1) ? 2048 + 24 : -1);
        MK_SYNC();
        // 9: LN -> quantised rows
        mk_lnq<WT>(a, a.x, d, L.lnm_w, L.lnm_b, a.actq);
        if (pf_w) { if (l + 1 < a.n_layer) mk_prefetch_w(a.layers[l + 1].qkv); else if (a.want_logits) mk_prefetch_w(a.te); }
        MK_SYNC();
        // 10: FC1 + GELU (2770-2794), then the rows are quantised for FC2
        e = MkEpi(); e.bias = L.fc1_bias; e.act = 1; e.qout = a.hq;
        mk_gemv<WT, true>(a, L.fc1, a.actq, e, (TRACE && l == 1) ? 2048 + 32 : -1);
        MK_STAMP(); MK_STAMP();                                  // (trace slot of the former FC1 -> Q8_0 phase)
        MK_SYNC();
        // 11: FC2 + residual (2797-2806)
        e = MkEpi(); e.bias = L.fc2_bias; e.res = a.x; e.out = a.x;
        mk_gemv<WT, false>(a, L.fc2, a.hq, e, (TRACE && l == 1) ? 2048 + 40 : -1);
        MK_SYNC();
    }
    if (a.want_logits) {                                         // final LN + logits (2811-2827)
        mk_lnq<WT>(a, a.x, d, a.lnf_w, a.lnf_b, a.actq);
        MK_SYNC();
        if (a.prefetch & 16) mk_logits_warp<WT>(a, a.te, a.actq, a.logits);
        else { MkEpi e; e.out = a.logits; mk_gemv<WT, false>(a, a.te, a.actq, e); }
        MK_STAMP();
    }
}

bool mk_cross_head_major() {
    return true;
}
int mk_barriers(int n_layer, bool want_logits) { return 11 * n_layer + (want_logits ? 1 : 0); }
bool mk_supported(int wtype) { return wtype == WT_F16 || wt_is_block32(wtype); }
size_t mk_smem_bytes(int, int) { return MK_SMEM; }
int mk_max_rows() { return MK_MAXTOK; }

template <int WT, bool TRACE>
static bool mk_launch_t(const MkArgs & a, int n_sm, cudaStream_t st) {
    const size_t smem = mk_smem_bytes(WT, a.d);
    if (ensure_dyn_smem(reinterpret_cast<const void *>(k_decode_pass<WT, TRACE>), 227 * 1024) != cudaSuccess) {
        set_error("decode megakernel: cannot raise the shared-memory limit"); return false;
    }
    if (smem > 227 * 1024) { set_error("decode megakernel: %zu bytes of shared memory needed", smem); return false; }
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
