// Implicit-GEMM convolution / GEMM on the 5th-generation tensor cores (tcgen05) for sm_100a.
//
//   y[n,oh,ow,co] = act( sum_{r,s,ci} x[n, oh*stride + r*dil - pad, ow*stride + s*dil - pad, ci] * w[co,r,s,ci] + bias[co] (+ residual) )
//
// This is the dense-contraction engine behind every conv / linear layer of the hot path (SURVEY.md §8a rows A2-A4, A10,
// B6, C5; Appendix B lists the shapes).  The reference runs them through cuDNN fp32 NCHW.
//
// Design (one CTA per SM, persistent over output tiles, warp-specialised):
//   * activations NHWC fp16/bf16; an output tile is bh x bw = 128 pixels of one image (or 128 consecutive pixels for 1x1).
//   * warp 0 (one lane): TMA producer.  For every filter tap (r,s) and every 64-channel chunk it issues ONE 4-D tiled
//     cp.async.bulk.tensor load of the *shifted* activation box {64 ch, bw, bh, 1} -- out-of-bounds coordinates are
//     zero-filled by the TMA unit, which is the convolution's zero padding; stride-2 convs use the tensor map's
//     elementStrides -- plus one 2-D load of the weight tile {64 k, BLOCK_N rows} of the K-major packed filter.
//     Both land in 128B-swizzled shared memory, i.e. directly in the canonical K-major UMMA operand layout.
//   * warp 1 (one lane): issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BLOCK_N<=256, K=16) with fp32 accumulators
//     in TMEM; tcgen05.commit releases the smem stage to the producer and, after the last k-block, hands the
//     accumulator to the epilogue.  Two accumulator stages (2 x 256 TMEM columns) overlap epilogue and MMA.
//   * warps 4-7: epilogue.  tcgen05.ld (32 lanes x 32 columns per warp) -> bias, residual, activation in fp32 ->
//     fp16/bf16 (or fp32) NHWC store, optionally into a channel slice of a wider tensor (concat fusion).
//   * mbarrier pipeline: full/empty per smem stage (TMA <-> MMA), tmem_full/tmem_empty per accumulator stage.
//
// Roofline: tensor pipe.  FLOPs = 2 * N*Hout*Wout * Cout * R*S*Cin.
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>

#include <cstdlib>
#include <map>
#include <mutex>
#include <tuple>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace {

constexpr int kBlockM = 128;
constexpr int kUmmaK = 16;
// warp 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, then EG groups of 4 epilogue warps (EG = 2: 384 threads x 168 registers; EG = 3: 512 threads compiled
// for 128 registers, where the first warpgroup shrinks to 40 with setmaxnreg.dec and the three epilogue warpgroups grow to 152).  The GELU / SiLU /
// residual epilogues of the short-K layers are bound by dependency latency (issue slots 48 % busy, ncu): a third warp per scheduler with the
// full register budget fills those gaps.
constexpr int kMaxEpiGroups = 3;
constexpr int kTmemCols = 512;
constexpr int kAccStages = 2;
constexpr int kMaxStages = 8;

struct ConvKernelParams {
    // geometry (output space)
    int N, H, W;                 // batch, output height/width (flat mode: N=1, H=1, W=pixels)
    int Cin, Cout;
    int R, S, stride, pad, dil;
    int bh, bw, tiles_h, tiles_w, tiles_m, tiles_n, block_n;
    int bk, kchunks, stages;     // bk: channels per k-block (64/32/16) -> swizzle 128/64/32 B
    int in_coff;
    int grouped;                 // 1: block-diagonal conv in 64-channel slices (ResNeXt grouped 3x3): n-tile j reads only input slice j
    // epilogue
    const float* bias;
    const void* residual;
    int res_ld, res_coff, res_mode;   // 0 none, 1 add before activation, 2 add after activation
    int act;
    int gelu_form;               // 1 (default): tanh form (one MUFU per element), 0: mixed polynomial / sigmoid-form exact GELU
    const float* act_param;      // per-channel PReLU slope
    void* out;
    float* out_f32;
    int out_ld, out_coff;
    int is_bf16;
    int tma_store;               // 0: per-lane 16 B global stores; 1 / 2: smem-staged TMA tile stores (64 B-swizzled / linear staging)
    uint32_t stage_off;          // byte offset of the epilogue staging area (8 warps x 2 KiB) from the 1 KiB-aligned smem base
    // LayerNorm folded into a 1x1 conv (csb_conv2d_ln_nhwc): x is UN-normalised, w = gamma (.) W, and the epilogue finishes
    //   y = rstd[row] * (acc - mean[row] * colsum[col]) + bias'[col]      with colsum[col] = sum_k w[col][k], bias' = bias + W beta
    // mean / rstd come from per-row partial (sum, sum of squares) pairs, one per 64-channel chunk, written by k_dwconv_tile<.., STATS>.
    const float* ln_stats;
    const float* ln_colsum;
    int ln_nchunk;
    float ln_inv_c, ln_eps;
};

template <int ACT>
__device__ __forceinline__ float apply_act(float x, float slope) {
    if constexpr (ACT == CSB_ACT_RELU) return fmaxf(x, 0.0f);
    else if constexpr (ACT == CSB_ACT_SILU) return __fdividef(x, 1.0f + __expf(-x));
    else if constexpr (ACT == CSB_ACT_GELU) {
        // exact-erf GELU, erf by Abramowitz-Stegun 7.1.26 (|error| < 1.5e-7 + MUFU approximation error ~1e-7: two orders below the fp16 output
        // rounding) with rcp.approx / ex2.approx: ~18 instructions and 2 MUFU per element instead of libdevice erff's ~35.  The C -> 4C layers of
        // ConvNeXt are epilogue-instruction-bound (K is only 128..1024), so this is on the critical path.
        const float ax = fabsf(x) * 0.70710678118654752f;
        float t, e;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, ax, 1.0f)));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(ax * ax * -1.4426950408889634f));
        const float poly = t * fmaf(t, fmaf(t, fmaf(t, fmaf(t, 1.061405429f, -1.453152027f), 1.421413741f), -0.284496736f), 0.254829592f);
        const float erf_ax = fmaf(-poly, e, 1.0f);                  // erf(|x|/sqrt2)
        return 0.5f * x + 0.5f * fabsf(x) * erf_ax;                  // 0.5 x (1 + sign(x) erf(|x|/sqrt2))
    }
    else if constexpr (ACT == CSB_ACT_PRELU) return x > 0.0f ? x : x * slope;
    else if constexpr (ACT == CSB_ACT_SIGMOID) return __fdividef(1.0f, 1.0f + __expf(-x));
    else if constexpr (ACT == CSB_ACT_SOFTPLUS) return x > 20.0f ? x : log1pf(__expf(x));
    else if constexpr (ACT == CSB_ACT_HARDSIGMOID) return fminf(fmaxf(x * (1.0f / 6.0f) + 0.5f, 0.0f), 1.0f);
    else return x;
}

// Exact-erf GELU on two values, MUFU-free: erf(x/sqrt2) = xc * P(t), xc = clamp(x, +-4.5), t = 2 xc^2 / 4.5^2 - 1, P = degree-9 least-squares fit on
// Chebyshev nodes (|erf error| < 9e-6 on the clamped range, 1 - erf(4.5/sqrt2) = 7e-6 beyond it; measured |GELU error| <= 1.4e-5 for |x| < 4.5 and
// <= 7e-6 |x| outside, i.e. 25x below the fp16 rounding of the output).  15 packed + 4 scalar instructions per PAIR, against ~36 + 4 MUFU.
__device__ __forceinline__ void gelu2(float& a, float& b) {
    const float ca = fminf(fmaxf(a, -4.5f), 4.5f), cb = fminf(fmaxf(b, -4.5f), 4.5f);
    const uint64_t x = pk2(a, b), xc = pk2(ca, cb);
#define CSB_C2(v) pk2(v, v)
    const uint64_t t = ffma2(fmul2(xc, xc), CSB_C2(2.0f / 20.25f), CSB_C2(-1.0f));
    uint64_t q = CSB_C2(-0.00369934f);
    q = ffma2(q, t, CSB_C2(0.00984567f));
    q = ffma2(q, t, CSB_C2(-0.01246448f));
    q = ffma2(q, t, CSB_C2(0.01944603f));
    q = ffma2(q, t, CSB_C2(-0.03675472f));
    q = ffma2(q, t, CSB_C2(0.05764049f));
    q = ffma2(q, t, CSB_C2(-0.08052489f));
    q = ffma2(q, t, CSB_C2(0.10928746f));
    q = ffma2(q, t, CSB_C2(-0.15436857f));
    q = ffma2(q, t, CSB_C2(0.31381178f));
    const uint64_t h = fmul2(x, CSB_C2(0.5f));
#undef CSB_C2
    upk2(ffma2(h, fmul2(xc, q), h), a, b);                          // 0.5 x (1 + erf(x / sqrt2))
}

// Exact-erf GELU on two values in sigmoid form, on the MUFU pipe: gelu(x) = x * Phi(x) = x / (1 + exp(-2 g(x))), g(x) = atanh(erf(x / sqrt2)) fitted
// by the odd polynomial x (c1 + c3 x^2 + c5 x^4) (x^2 clamped at 64, where Phi has saturated; weighted minimax fit, max |GELU error| 2.6e-5
// over all x -- the same class as gelu2 above and 20x below the fp16 rounding of the output).  The constants carry the factor -2 log2(e), so
// exp(-2 g) is one ex2.approx.  6 packed + 2 FMNMX + 4 MUFU per PAIR, against 15 packed + 4 FMNMX for gelu2: the epilogue of the C -> 4C
// layers is bound by instruction issue / dependency latency, and a chunk that uses both forms (kGeluMufuPairs of its 16 pairs here, the rest
// gelu2) spreads the work over the FMA and MUFU pipes (measured on B200, 512 -> 2048 @32x64x64 / 128 -> 512 @16x256x256: 0 pairs 348 / 548 us,
// 8 pairs 322 / 489 us, 16 pairs 333 / 503 us).  ex2 overflow (x << 0) gives rcp(inf) = 0 -> 0, the correct limit.
__device__ __forceinline__ void gelu2_mufu(float& a, float& b) {
    const uint64_t x = pk2(a, b);
    float s0, s1;
    upk2(fmul2(x, x), s0, s1);
    const uint64_t x2 = pk2(fminf(s0, 64.0f), fminf(s1, 64.0f));
#define CSB_C2(v) pk2(v, v)
    uint64_t q = ffma2(x2, CSB_C2(0.0010142643004655838f), CSB_C2(-0.10677573084831238f));
    q = ffma2(q, x2, CSB_C2(-2.301121234893799f));
    float u0, u1, e0, e1, r0, r1;
    upk2(fmul2(q, x), u0, u1);
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(u0));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(u1));
    float d0, d1;
    upk2(fadd2(pk2(e0, e1), CSB_C2(1.0f)), d0, d1);
#undef CSB_C2
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(d0));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(d1));
    upk2(fmul2(x, pk2(r0, r1)), a, b);
}
// Exact-erf GELU on two values in tanh form: gelu(x) = 0.5 x (1 + tanh(g(x))) with the same fitted g(x) = atanh(erf(x / sqrt2)) as gelu2_mufu, through ONE
// MUFU op per element (tanh.approx.f32) instead of ex2 + rcp: 6 packed + 2 FMNMX + 2 MUFU per PAIR.  The default form (GM < 0); measured against the
// correctly rounded fp16 GELU over x in [-12, 12] (tools/gelu_check.py): rms abs error 6.4e-5 (polynomial / sigmoid mix: 5.5e-5), outputs that differ from
// the correctly rounded value 15.5 % (mix: 28 %), same maximum.
__device__ __forceinline__ void gelu2_tanh(float& a, float& b) {
    const uint64_t x = pk2(a, b);
    float s0, s1;
    upk2(fmul2(x, x), s0, s1);
    const uint64_t x2 = pk2(fminf(s0, 64.0f), fminf(s1, 64.0f));
#define CSB_C2(v) pk2(v, v)
    uint64_t q = ffma2(x2, CSB_C2(-0.00035151723f), CSB_C2(0.037005644f));
    q = ffma2(q, x2, CSB_C2(0.79750788f));
    float u0, u1, t0, t1;
    upk2(fmul2(q, x), u0, u1);
    asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(u0));
    asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(u1));
    const uint64_t h = fmul2(x, CSB_C2(0.5f));
#undef CSB_C2
    upk2(ffma2(h, pk2(t0, t1), h), a, b);
}
#ifndef CSB_GELU_MUFU_PAIRS
#define CSB_GELU_MUFU_PAIRS 8
#endif
constexpr int kGeluMufuPairs = CSB_GELU_MUFU_PAIRS;      // of the 16 pairs of a 32-column chunk

// One 32-column chunk of one accumulator row.  The fast path (full chunk, 16 B-aligned slices) is straight-line code: 8 float4 bias
// loads (warp-uniform -> broadcast), 4 x 16 B residual loads, activation on 32 independent values, 4 x 16 B stores.
template <class T, int ACT, bool LN, int GM = kGeluMufuPairs>
__device__ __forceinline__ void epilogue_chunk(const ConvKernelParams& p, const uint32_t (&acc)[32], size_t pix, int n0, bool row_ok, bool fast, uint32_t stage,
                                               const CUtensorMap* tmC, int cw, int chh, int cimg, float neg_mean, float rstd, uint32_t stage_any) {
    // stage != 0: this full chunk leaves through shared memory and one TMA tile store per warp (rows outside the image are clipped by the TMA unit;
    // their arithmetic runs on in-bounds addresses because `pix` is clamped by the caller)
    if (!row_ok && !(stage && fast && !p.out_f32) && !p.out_f32) return;      // (the fp32 path shuffles across the warp: every lane takes part)
    if (fast && !p.out_f32) {
        float y[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) y[j] = __uint_as_float(acc[j]);
        if constexpr (LN) {                                 // rstd * acc + (-mean * rstd * colsum + bias'): 32 FFMA2 (the plain path spends 16 FADD2)
            const uint64_t nm = pk2(neg_mean, neg_mean), rs = pk2(rstd, rstd);      // neg_mean = -mean * rstd
#pragma unroll
            for (int g = 0; g < 8; ++g) {
                const float4 c = __ldg(reinterpret_cast<const float4*>(p.ln_colsum + n0) + g), b = __ldg(reinterpret_cast<const float4*>(p.bias + n0) + g);
                upk2(ffma2(rs, pk2(y[4 * g], y[4 * g + 1]), ffma2(nm, pk2(c.x, c.y), pk2(b.x, b.y))), y[4 * g], y[4 * g + 1]);
                upk2(ffma2(rs, pk2(y[4 * g + 2], y[4 * g + 3]), ffma2(nm, pk2(c.z, c.w), pk2(b.z, b.w))), y[4 * g + 2], y[4 * g + 3]);
            }
        } else if (p.bias) {                                // packed adds: 16 FADD2 instead of 32 FADD
#pragma unroll
            for (int g = 0; g < 8; ++g) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n0) + g);
                upk2(fadd2(pk2(y[4 * g], y[4 * g + 1]), pk2(b.x, b.y)), y[4 * g], y[4 * g + 1]);
                upk2(fadd2(pk2(y[4 * g + 2], y[4 * g + 3]), pk2(b.z, b.w)), y[4 * g + 2], y[4 * g + 3]);
            }
        }
        float r[32];
        if (p.res_mode) {
            const uint4* rp = reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(p.residual) + pix * p.res_ld + p.res_coff + n0);
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const uint4 u = rp[g];
                float2 f;
                f = unpack2<T>(u.x); r[8 * g] = f.x; r[8 * g + 1] = f.y;
                f = unpack2<T>(u.y); r[8 * g + 2] = f.x; r[8 * g + 3] = f.y;
                f = unpack2<T>(u.z); r[8 * g + 4] = f.x; r[8 * g + 5] = f.y;
                f = unpack2<T>(u.w); r[8 * g + 6] = f.x; r[8 * g + 7] = f.y;
            }
            if (p.res_mode == 1) {
#pragma unroll
                for (int j = 0; j < 32; j += 2) upk2(fadd2(pk2(y[j], y[j + 1]), pk2(r[j], r[j + 1])), y[j], y[j + 1]);
            }
        }
        if constexpr (ACT == CSB_ACT_PRELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) y[j] = apply_act<ACT>(y[j], p.act_param ? __ldg(p.act_param + n0 + j) : 0.25f);
        } else if constexpr (ACT == CSB_ACT_GELU) {
#pragma unroll
            for (int j = 0; j < 32; j += 2) {                // interleaved, so that neighbouring pairs run on different pipes
                if constexpr (GM < 0) gelu2_tanh(y[j], y[j + 1]);
                else if ((j >> 1) * GM / 16 != ((j >> 1) + 1) * GM / 16) gelu2_mufu(y[j], y[j + 1]);
                else gelu2(y[j], y[j + 1]);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) y[j] = apply_act<ACT>(y[j], 0.f);
        }
        if (p.res_mode == 2) {
#pragma unroll
            for (int j = 0; j < 32; j += 2) upk2(fadd2(pk2(y[j], y[j + 1]), pk2(r[j], r[j + 1])), y[j], y[j + 1]);
        }
        if (stage) {
            // Each lane owns one 64 B row of the warp's 32 x 32 staging tile.  A 16 B-per-lane global store touches 32 different lines per
            // instruction (the C -> 4C layers were bound by exactly that); the TMA engine writes whole rows and takes the work off the LSU.
            const int lane = threadIdx.x & 31;
            if (lane == 0) tma_store_wait_read();                 // the previous tile of this warp has left shared memory
            __syncwarp();
            const uint32_t row = stage + (uint32_t) lane * 64u;
            const uint32_t sw = p.tma_store == 1 ? (uint32_t) ((lane >> 1) & 3) : 0u;      // CU_TENSOR_MAP_SWIZZLE_64B: 16 B chunk ^= (row >> 1) & 3
#pragma unroll
            for (int g = 0; g < 4; ++g)
                st_shared_v4(row + (((uint32_t) g ^ sw) << 4), make_uint4(pack2<T>(y[8 * g], y[8 * g + 1]), pack2<T>(y[8 * g + 2], y[8 * g + 3]),
                                                                         pack2<T>(y[8 * g + 4], y[8 * g + 5]), pack2<T>(y[8 * g + 6], y[8 * g + 7])));
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) {
                tma_store_4d(tmC, stage, n0, cw, chh, cimg);
                tma_store_commit();
            }
            return;
        }
        uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<T*>(p.out) + pix * p.out_ld + p.out_coff + n0);
#pragma unroll
        for (int g = 0; g < 4; ++g)
            o[g] = make_uint4(pack2<T>(y[8 * g], y[8 * g + 1]), pack2<T>(y[8 * g + 2], y[8 * g + 3]), pack2<T>(y[8 * g + 4], y[8 * g + 5]),
                              pack2<T>(y[8 * g + 6], y[8 * g + 7]));
        return;
    }
    // fp32 outputs (head predictions, e.g. the 169 dynamic-kernel channels of RTMDet-Ins): a lane owns one accumulator ROW, so a per-lane store
    // touches 32 different lines per instruction (the 256 -> 169 head ran at 60 TFLOP/s on exactly that).  The chunk is transposed through the
    // warp's 2 KiB staging tile, 16 columns at a time, and leaves as 64 B row segments (two rows per store instruction).
    if (p.out_f32) {
        const int lane = threadIdx.x & 31;
        const T* res = p.res_mode ? reinterpret_cast<const T*>(p.residual) + pix * p.res_ld + p.res_coff + n0 : nullptr;
        float y[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const int c = n0 + j;
            float v = __uint_as_float(acc[j]);
            if (c < p.Cout) {
                if constexpr (LN) v = fmaf(rstd, v, neg_mean * __ldg(p.ln_colsum + c));
                if (p.bias) v += __ldg(p.bias + c);
                if (p.res_mode == 1) v += to_f<T>(res[j]);
                v = apply_act<ACT>(v, p.act_param ? __ldg(p.act_param + c) : 0.25f);
                if (p.res_mode == 2) v += to_f<T>(res[j]);
            }
            y[j] = v;
        }
        const unsigned long long obase = (unsigned long long) pix * (unsigned long long) p.out_ld + (unsigned long long) (p.out_coff + n0);
        const uint32_t my_row = stage_any + (uint32_t) lane * 64u, sw = (uint32_t) ((lane >> 1) & 3);
#pragma unroll
        for (int hb = 0; hb < 2; ++hb) {
            __syncwarp();
#pragma unroll
            for (int g = 0; g < 4; ++g)
                st_shared_v4(my_row + (((uint32_t) g ^ sw) << 4), make_uint4(__float_as_uint(y[16 * hb + 4 * g]), __float_as_uint(y[16 * hb + 4 * g + 1]),
                                                                         __float_as_uint(y[16 * hb + 4 * g + 2]), __float_as_uint(y[16 * hb + 4 * g + 3])));
            __syncwarp();
            const int c = lane & 15;
            const bool col_ok = n0 + 16 * hb + c < p.Cout;
#pragma unroll 4
            for (int i = 0; i < 16; ++i) {
                const int r = 2 * i + (lane >> 4);
                const unsigned long long ob = __shfl_sync(0xffffffffu, obase, r);
                const bool ok = __shfl_sync(0xffffffffu, (int) row_ok, r) != 0;
                float v;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(stage_any + (uint32_t) r * 64u + ((((uint32_t) c >> 2) ^ (uint32_t) ((r >> 1) & 3)) << 4) + (((uint32_t) c & 3u) << 2)));
                if (ok && col_ok) p.out_f32[ob + (unsigned long long) (16 * hb + c)] = v;
            }
        }
        return;
    }
    // generic path: channel tails and unaligned slices
    if (!row_ok) return;
    const T* res = p.res_mode ? reinterpret_cast<const T*>(p.residual) + pix * p.res_ld + p.res_coff + n0 : nullptr;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const int c = n0 + j;
        if (c >= p.Cout) continue;
        float v = __uint_as_float(acc[j]);
        if constexpr (LN) v = fmaf(rstd, v, neg_mean * __ldg(p.ln_colsum + c));
        if (p.bias) v += __ldg(p.bias + c);
        if (p.res_mode == 1) v += to_f<T>(res[j]);
        v = apply_act<ACT>(v, p.act_param ? __ldg(p.act_param + c) : 0.25f);
        if (p.res_mode == 2) v += to_f<T>(res[j]);
        reinterpret_cast<T*>(p.out)[pix * p.out_ld + p.out_coff + c] = from_f<T>(v);
    }
}

// Epilogue role: EG groups of 4 warps; warp w owns TMEM lane quarter (w & 3) and the 32-column chunks with chunk % EG == (w - 4) / 4.
// CG = 2: `total_tiles` counts PAIR tiles (two m-tiles that share an n-tile); this CTA owns m-tile 2 * pair + rank (an m-tile past the end is
// computed on zero-filled operands and never stored), and hands its accumulator stage back on the LEADER's tmem_empty barrier.
template <class T, int ACT, int CG, bool LN = false, int GM = kGeluMufuPairs>
__device__ __forceinline__ void epilogue_role(const ConvKernelParams& p, uint32_t tmem_base, uint32_t tfull0, uint32_t tempty0, int total_tiles, uint32_t stage_base,
                                              const CUtensorMap* tmC) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int EG = (int) (blockDim.x >> 7) - 1;
    const int q = warp & 3;
    const bool aligned = ((p.out_ld | p.out_coff) % 8 == 0) && (!p.res_mode || ((p.res_ld | p.res_coff) % 8 == 0)) && (!p.bias || ((uintptr_t) p.bias % 16 == 0));
    const int rank = CG == 2 ? (int) cluster_ctarank() : 0;
    const int cta = CG == 2 ? (int) (blockIdx.x >> 1) : (int) blockIdx.x, ncta = CG == 2 ? (int) (gridDim.x >> 1) : (int) gridDim.x;
    const uint32_t tempty_lead = CG == 2 ? mapa_u32(tempty0, 0) : tempty0;
    auto release = [&](int s) {
        if constexpr (CG == 2) mbar_arrive_cluster(tempty_lead + 8u * s);
        else mbar_arrive(tempty0 + 8u * s);
    };
    int as = 0, rot = (warp - 4) >> 2;
    uint32_t aphase = 0;
    for (int tile = cta; tile < total_tiles; tile += ncta) {
        // the warp groups take the tile's 32-column chunks round-robin, starting one group later every tile: with 8 chunks and 3 groups the
        // 3 / 3 / 2 split rotates, so over three tiles every group converts 8 chunks (the accumulator double buffer absorbs the skew)
        const int half = rot;
        if (++rot == EG) rot = 0;
        const int nt = tile % p.tiles_n, mt = (tile / p.tiles_n) * CG + rank;
        const int tw = mt % p.tiles_w, th = (mt / p.tiles_w) % p.tiles_h, img = mt / (p.tiles_w * p.tiles_h);
        const int m = q * 32 + lane;
        const int oh = th * p.bh + m / p.bw, ow = tw * p.bw + m % p.bw;
        const bool row_ok = oh < p.H && ow < p.W && img < p.N;
        // rows outside the image still run the arithmetic on the TMA-store path (the store clips them): keep their residual address in bounds
        const size_t pix = ((size_t) (img < p.N ? img : p.N - 1) * p.H + (oh < p.H ? oh : p.H - 1)) * p.W + (ow < p.W ? ow : p.W - 1);
        const uint32_t stage = p.tma_store ? stage_base + (uint32_t) (warp - 4) * 2048u : 0u;
        const int row0 = q * 32;
        const int cw = tw * p.bw + row0 % p.bw, chh = th * p.bh + row0 / p.bw;
        float neg_mean = 0.f, rstd = 1.f;
        if constexpr (LN) {                                 // this row's LayerNorm statistics from the per-chunk partials (loaded before the wait)
            const float2* sp = reinterpret_cast<const float2*>(p.ln_stats) + pix * p.ln_nchunk;
            float s1 = 0.f, s2 = 0.f;
            for (int c = 0; c < p.ln_nchunk; ++c) { const float2 t = __ldg(sp + c); s1 += t.x; s2 += t.y; }
            const float mean = s1 * p.ln_inv_c;
            rstd = rsqrtf(fmaxf(fmaf(s2, p.ln_inv_c, -mean * mean), 0.f) + p.ln_eps);
            neg_mean = -mean * rstd;
        }
        mbar_wait(tfull0 + 8u * as, aphase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t) (q * 32) << 16) + (uint32_t) as * 256u;
        const int ncols = min(p.block_n, p.Cout - nt * p.block_n);
        const int nchunks = (ncols + 31) / 32;
        int last = -1;                                     // last chunk this warp reads
        for (int ch = half; ch < nchunks; ch += EG) last = ch;
        if (last < 0) {                                    // nothing to read for this warp in this tile: release immediately
            tc_fence_before();
            __syncwarp();
            if (lane == 0) release(as);
        }
        for (int ch = half; ch < nchunks; ch += EG) {
            const int n0 = nt * p.block_n + ch * 32;
            const bool fast = aligned && n0 + 32 <= p.Cout;
            uint32_t acc[32];
            tmem_ld32(taddr + (uint32_t) ch * 32u, acc);
            if (ch == last) {                              // accumulator rows of this warp fully read: hand the TMEM stage back
                tc_fence_before();
                __syncwarp();
                if (lane == 0) release(as);
            }
            epilogue_chunk<T, ACT, LN, GM>(p, acc, pix, n0, row_ok, fast, stage, tmC, cw, chh, img, neg_mean, rstd, stage_base + (uint32_t) (warp - 4) * 2048u);
        }
        if (++as == kAccStages) { as = 0; aphase ^= 1u; }
    }
    if (p.tma_store && lane == 0) tma_store_wait_all();           // every tile store of this warp has completed before the CTA may exit
}

template <class T, int CG>
__device__ __forceinline__ void epilogue_dispatch(const ConvKernelParams& p, uint32_t tmem_base, uint32_t tfull0, uint32_t tempty0, int total_tiles, uint32_t stage_base,
                                                  const CUtensorMap* tmC) {
    if (p.act == CSB_ACT_GELU && p.gelu_form == 1) {    // tanh-form GELU (default)
        if (p.ln_stats) epilogue_role<T, CSB_ACT_GELU, CG, true, -1>(p, tmem_base, tfull0, tempty0, total_tiles, stage_base, tmC);
        else epilogue_role<T, CSB_ACT_GELU, CG, false, -1>(p, tmem_base, tfull0, tempty0, total_tiles, stage_base, tmC);
        return;
    }
    if (p.ln_stats) {     // LayerNorm-folded 1x1 conv: only the activations that follow a LayerNorm on this path
        if (p.act == CSB_ACT_GELU) epilogue_role<T, CSB_ACT_GELU, CG, true>(p, tmem_base, tfull0, tempty0, total_tiles, stage_base, tmC);
        else epilogue_role<T, CSB_ACT_NONE, CG, true>(p, tmem_base, tfull0, tempty0, total_tiles, stage_base, tmC);
        return;
    }
    switch (p.act) {      // hoisted out of every loop: each instantiation is straight-line code
        case CSB_ACT_RELU: epilogue_role<T, CSB_ACT_RELU, CG>(p, tmem_base, tfull0, tempty0, total_tiles, stage_base, tmC); break;
        case CSB_ACT_SILU: epilogue_role<T, CSB_ACT_SILU, CG>(p, tmem_base, tfull0, tempty0, total_tiles, stage_base, tmC); break;
        case CSB_ACT_GELU: epilogue_role<T, CSB_ACT_GELU, CG>(p, tmem_base, tfull0, tempty0, total_tiles, stage_base, tmC); break;
        case CSB_ACT_PRELU: epilogue_role<T, CSB_ACT_PRELU, CG>(p, tmem_base, tfull0, tempty0, total_tiles, stage_base, tmC); break;
        case CSB_ACT_SIGMOID: epilogue_role<T, CSB_ACT_SIGMOID, CG>(p, tmem_base, tfull0, tempty0, total_tiles, stage_base, tmC); break;
        case CSB_ACT_SOFTPLUS: epilogue_role<T, CSB_ACT_SOFTPLUS, CG>(p, tmem_base, tfull0, tempty0, total_tiles, stage_base, tmC); break;
        case CSB_ACT_HARDSIGMOID: epilogue_role<T, CSB_ACT_HARDSIGMOID, CG>(p, tmem_base, tfull0, tempty0, total_tiles, stage_base, tmC); break;
        default: epilogue_role<T, CSB_ACT_NONE, CG>(p, tmem_base, tfull0, tempty0, total_tiles, stage_base, tmC); break;
    }
}

// CG = 1: one CTA per SM, every CTA an independent pipeline.  CG = 2: clusters of two CTAs (one TPC) work on two m-tiles that share an n-tile; each
// CTA's producer loads its own activation tile and HALF of the weight tile, the leader's MMA warp issues tcgen05.mma.cta_group::2 (M = 256), so the
// weight tile crosses the L2 -> SM fabric once per pair instead of once per CTA (the 128 x 256 x 64 k-block drops from 48 KiB to 32 KiB per SM).
// Barriers: full[s] lives in the leader (one expect_tx arrival of the leader's producer + the bytes of both CTAs' loads), empty[s] and tmem_full[a]
// exist in both CTAs and are signalled by multicast commits, tmem_empty[a] lives in the leader and counts the epilogue warps of both CTAs.
template <int EG, int CG>
__global__ void __launch_bounds__(128 * (EG + 1), 1) k_conv_tc(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                               const __grid_constant__ CUtensorMap tmC, const ConvKernelParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t row_bytes = (uint32_t) p.bk * 2u;
    const uint32_t a_bytes = kBlockM * row_bytes, b_bytes = (uint32_t) (p.block_n / CG) * row_bytes;       // per CTA
    const uint32_t stage_bytes = (a_bytes + b_bytes + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + (uint32_t) p.stages * stage_bytes;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (kMaxStages + s); };
    auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * kMaxStages + s); };
    auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * kMaxStages + kAccStages + s); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * kMaxStages + 2 * kAccStages);
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = CG == 2 ? (int) cluster_ctarank() : 0;
    const int cta = CG == 2 ? (int) (blockIdx.x >> 1) : (int) blockIdx.x, ncta = CG == 2 ? (int) (gridDim.x >> 1) : (int) gridDim.x;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        if (p.tma_store) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmC) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int s = 0; s < kAccStages; ++s) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), 4 * EG * CG); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 2) {
        if constexpr (CG == 1) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t) kTmemCols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t) kTmemCols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if constexpr (CG == 2) cluster_sync_all();          // the peer's barriers are initialised before anything arrives on them remotely
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    const int total_tiles = ((p.tiles_m + CG - 1) / CG) * p.tiles_n;     // CG = 2: pair tiles
    const int kblocks = p.R * p.S * p.kchunks;

    if (warp < 4) {
        // 16 warps are compiled for 128 registers; the TMA / MMA / alloc warpgroup needs ~40, so it hands its share to the three epilogue
        // warpgroups (4 x 32 x 40 + 12 x 32 x 152 = 63488 <= 65536): 12 epilogue warps run with the register budget of the 8-warp build.
        if constexpr (EG == 3) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == 0) {
        // ===================================================== TMA producer (one per CTA; in a pair both signal the leader's full barrier)
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            const uint32_t full_lead0 = CG == 2 ? mapa_u32(full_bar(0), 0) : full_bar(0);
            for (int tile = cta; tile < total_tiles; tile += ncta) {
                const int nt = tile % p.tiles_n, mt = (tile / p.tiles_n) * CG + rank;
                const int tw = mt % p.tiles_w, th = (mt / p.tiles_w) % p.tiles_h, img = mt / (p.tiles_w * p.tiles_h);     // img >= N: zero-filled by TMA
                const int oh0 = th * p.bh, ow0 = tw * p.bw, n0 = nt * p.block_n;
                for (int r = 0; r < p.R; ++r)
                    for (int s = 0; s < p.S; ++s)
                        for (int kc = 0; kc < p.kchunks; ++kc) {
                            mbar_wait(empty_bar(stage), phase ^ 1u);
                            const uint32_t a_dst = smem_base + (uint32_t) stage * stage_bytes, b_dst = a_dst + a_bytes;
                            if (rank == 0) mbar_expect_tx(full_bar(stage), CG * (a_bytes + b_bytes));
                            const uint32_t fb = full_lead0 + 8u * stage;
                            tma_load_4d<CG>(a_dst, &tmA, fb, p.in_coff + (p.grouped ? n0 : kc * p.bk), ow0 * p.stride + s * p.dil - p.pad,
                                            oh0 * p.stride + r * p.dil - p.pad, img);
                            tma_load_2d<CG>(b_dst, &tmB, fb, ((r * p.S + s) * p.kchunks + kc) * p.bk, n0 + rank * (p.block_n / CG));
                            if (++stage == p.stages) { stage = 0; phase ^= 1u; }
                        }
            }
        }
    } else if (warp == 1 && rank == 0) {
        // ===================================================== MMA issuer (the leader CTA of a pair)
        // instruction descriptor (cute::UMMA::InstrDescriptor): fp32 accumulate, A/B fp16 or bf16, both K-major, N>>3 @17, M>>4 @24
        const uint32_t fmt = p.is_bf16 ? 1u : 0u;
        const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t) (p.block_n >> 3) << 17) | ((uint32_t) ((kBlockM * CG) >> 4) << 24);
        int stage = 0, as = 0;
        uint32_t phase = 0, aphase = 0;
        for (int tile = cta; tile < total_tiles; tile += ncta) {
            mbar_wait(tempty_bar(as), aphase ^ 1u);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + (uint32_t) as * 256u;
            for (int kb = 0; kb < kblocks; ++kb) {
                mbar_wait(full_bar(stage), phase);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t a_addr = smem_base + (uint32_t) stage * stage_bytes, b_addr = a_addr + a_bytes;
                    const uint64_t adesc = make_desc(a_addr, row_bytes), bdesc = make_desc(b_addr, row_bytes);
                    for (int k = 0; k < p.bk / kUmmaK; ++k)   // advance 16 elements = 32 B = 2 descriptor units inside the swizzle atom
                        umma_f16<CG>(tmem_d, adesc + 2u * k, bdesc + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
                    umma_commit<CG>(empty_bar(stage));
                    if (kb == kblocks - 1) umma_commit<CG>(tfull_bar(as));
                }
                __syncwarp();
                if (++stage == p.stages) { stage = 0; phase ^= 1u; }
            }
            if (++as == kAccStages) { as = 0; aphase ^= 1u; }
        }
    }
    } else {
        // ===================================================== epilogue (TMEM -> registers -> global), 4 * EG warps
        if constexpr (EG == 3) asm volatile("setmaxnreg.inc.sync.aligned.u32 152;");
        if (p.is_bf16) epilogue_dispatch<__nv_bfloat16, CG>(p, tmem_base, tfull_bar(0), tempty_bar(0), total_tiles, smem_base + p.stage_off, &tmC);
        else epilogue_dispatch<__half, CG>(p, tmem_base, tfull_bar(0), tempty_bar(0), total_tiles, smem_base + p.stage_off, &tmC);
    }
    tc_fence_before();
    __syncthreads();
    if constexpr (CG == 2) cluster_sync_all();          // neither CTA leaves (or frees TMEM) while its peer may still signal it / read its shared memory
    if (warp == 2) {
        tc_fence_after();
        if constexpr (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t) kTmemCols) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t) kTmemCols) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn) f;
    });
    return fn;
}

CUtensorMapSwizzle swizzle_of(int bk) { return bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (bk == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B); }

std::atomic<int> g_pair_mode{[] { const char* e = getenv("CSB_CTA_PAIR"); return e ? atoi(e) : 1; }()};

// tanh form by default: measured on B200 (profiles/r2_gelu_form_ab.md) it is at least as accurate after the fp16 rounding of the output and 2-12 % faster
// per C -> 4C layer, +3 % on the whole step (fewer FMA-pipe instructions -> less power -> higher clocks under the power cap); CSB_GELU_FORM=poly restores
// the polynomial / sigmoid mix
std::atomic<int> g_gelu_form{[] { const char* e = getenv("CSB_GELU_FORM"); return e && (e[0] == 'p' || e[0] == '0') ? 0 : 1; }()};

}  // namespace

// A/B switch of the GELU epilogue form (0 mixed polynomial / sigmoid, 1 tanh = default); returns the previous form.
extern "C" int csb_conv_set_gelu_form(int form) { return g_gelu_form.exchange(form); }

// A/B switch of the CTA-pair path for tests and benchmarks (same values as the CSB_CTA_PAIR environment variable); returns the previous mode.
extern "C" int csb_conv_set_pair_mode(int mode) { return g_pair_mode.exchange(mode); }

static int conv_impl(const csb_conv_desc* d, const void* x, const void* w, const float* bias, const float* act_param, const void* residual, void* y,
                     float* y_f32, const float* ln_stats, const float* ln_colsum, float ln_eps, void* stream) {
    CSB_REQUIRE(d && x && w && (y || y_f32), "null pointer");
    CSB_REQUIRE(d->N > 0 && d->Hin > 0 && d->Win > 0 && d->Cin > 0 && d->Cout > 0 && d->R > 0 && d->S > 0, "bad shape");
    CSB_REQUIRE(d->stride == 1 || d->stride == 2 || d->stride == 4, "stride must be 1, 2 or 4");
    CSB_REQUIRE(d->Cin % 16 == 0, "Cin must be a multiple of 16 (pad the channel dimension)");
    CSB_REQUIRE(d->in_ld % 8 == 0 && d->in_coff % 8 == 0 && d->in_ld >= d->in_coff + d->Cin, "input channel stride/offset must be multiples of 8");
    CSB_REQUIRE(((uintptr_t) x & 15) == 0 && ((uintptr_t) w & 15) == 0, "x and w must be 16-byte aligned");
    CSB_REQUIRE(d->res_mode == 0 || residual, "residual pointer missing");
    // thin stride-1 RxS layers (Cout <= 128): one halo tile per 64-channel chunk instead of one activation box per tap (tc_halo.cu)
    if (!ln_stats && d->groups <= 1 && d->Cout <= 128 && (d->act == CSB_ACT_NONE || d->act == CSB_ACT_RELU || d->act == CSB_ACT_SILU || d->act == CSB_ACT_PRELU) &&
        csb_conv_halo_supported(d))
        return csb::conv_halo_launch(d, x, w, bias, act_param, residual, y, y_f32, nullptr, stream);
    EncodeTiledFn enc = encode_fn();
    if (!enc) return csb::fail(CSB_ERR_CUDA, "%s: %s", "csb_conv2d_nhwc", "cuTensorMapEncodeTiled unavailable");
    const int Hout = (d->Hin + 2 * d->pad - d->dil * (d->R - 1) - 1) / d->stride + 1;
    const int Wout = (d->Win + 2 * d->pad - d->dil * (d->S - 1) - 1) / d->stride + 1;
    CSB_REQUIRE(Hout > 0 && Wout > 0, "empty output");

    ConvKernelParams p{};
    p.Cin = d->Cin; p.Cout = d->Cout; p.R = d->R; p.S = d->S; p.stride = d->stride; p.pad = d->pad; p.dil = d->dil;
    p.bk = d->Cin % 64 == 0 ? 64 : (d->Cin % 32 == 0 ? 32 : 16);
    p.kchunks = d->Cin / p.bk;
    p.grouped = d->groups > 1;
    if (p.grouped) {
        CSB_REQUIRE(d->Cin == d->Cout && d->Cin % 64 == 0 && d->Cin % d->groups == 0 && 64 % (d->Cin / d->groups) == 0,
                    "grouped conv needs Cin == Cout, a multiple of 64, with 64 % (Cin/groups) == 0");
        p.kchunks = 1;      // w is [Cout][R][S][64]: the 64x64 diagonal block of each output slice (zeros outside the groups)
    }
    p.in_coff = d->in_coff;
    const bool flat = d->R == 1 && d->S == 1 && d->stride == 1 && d->pad == 0 && !p.grouped;
    if (ln_stats) {
        CSB_REQUIRE(flat && ln_colsum && d->Cin % 64 == 0 && ((uintptr_t) ln_colsum & 15) == 0 && ((uintptr_t) ln_stats & 7) == 0,
                    "LayerNorm folding needs a dense 1x1 stride-1 conv with Cin a multiple of 64");
        CSB_REQUIRE(d->act == CSB_ACT_GELU || d->act == CSB_ACT_NONE, "LayerNorm folding supports act none / gelu");
        CSB_REQUIRE(bias && ((uintptr_t) bias & 15) == 0, "LayerNorm folding needs the folded bias (bias + W beta), 16-byte aligned");
        p.ln_stats = ln_stats; p.ln_colsum = ln_colsum; p.ln_nchunk = d->Cin / 64; p.ln_inv_c = 1.0f / (float) d->Cin; p.ln_eps = ln_eps;
    }
    cuuint64_t gdim[4], gstr[3];
    cuuint32_t box[4], estr[4];
    const cuuint64_t esz = 2;
    if (flat) {   // 1x1: the NHWC tensor is a [pixels, C] matrix -> tiles of 128 consecutive pixels, no spatial waste
        p.N = 1; p.H = 1; p.W = d->N * Hout * Wout; p.bh = 1; p.bw = kBlockM;
        gdim[0] = (cuuint64_t) d->in_ld; gdim[1] = (cuuint64_t) p.W; gdim[2] = 1; gdim[3] = 1;
        gstr[0] = (cuuint64_t) d->in_ld * esz; gstr[1] = gstr[0] * gdim[1]; gstr[2] = gstr[1];
    } else {
        p.N = d->N; p.H = Hout; p.W = Wout;
        long long best = -1;
        for (int bw = 8; bw <= 128; bw *= 2) {   // tile shape minimising padded area; ties -> squarer tile
            const int bh = kBlockM / bw;
            if (bw * d->stride > 256 || bh * d->stride > 256) continue;
            long long area = (long long) ((Wout + bw - 1) / bw) * bw * ((Hout + bh - 1) / bh) * bh;
            long long score = area * 1024 + (bw > bh ? bw / bh : bh / bw);
            if (best < 0 || score < best) { best = score; p.bw = bw; p.bh = bh; }
        }
        gdim[0] = (cuuint64_t) d->in_ld; gdim[1] = (cuuint64_t) d->Win; gdim[2] = (cuuint64_t) d->Hin; gdim[3] = (cuuint64_t) d->N;
        gstr[0] = (cuuint64_t) d->in_ld * esz; gstr[1] = gstr[0] * d->Win; gstr[2] = gstr[1] * d->Hin;
    }
    p.tiles_h = (p.H + p.bh - 1) / p.bh; p.tiles_w = (p.W + p.bw - 1) / p.bw; p.tiles_m = p.N * p.tiles_h * p.tiles_w;
    p.block_n = p.grouped ? 64 : (d->Cout >= 256 ? 256 : ((d->Cout + 15) / 16) * 16);
    p.tiles_n = (d->Cout + p.block_n - 1) / p.block_n;
    // CTA pairs (tcgen05 cta_group::2, see k_conv_tc): for the wide-N layers, whose k-blocks are bound by the L2 -> shared-memory fabric.
    // CSB_CTA_PAIR: 0 never, 1 (default) when block_n >= 128 and there are at least two m-tiles, 2 whenever the shape allows it.
    const int pair_mode = g_pair_mode.load(std::memory_order_relaxed);
    static const int eg = [] { const char* e = getenv("CSB_EPI_GROUPS"); return e && atoi(e) == 2 ? 2 : 3; }();      // 12 (default) or 8 epilogue warps
    const int cg = (eg == 3 && pair_mode > 0 && !p.grouped && p.block_n % 32 == 0 && p.tiles_m >= 2 && (pair_mode == 2 || p.block_n >= 128)) ? 2 : 1;
    box[0] = (cuuint32_t) p.bk; box[1] = (cuuint32_t) (p.bw * d->stride); box[2] = (cuuint32_t) (p.bh * d->stride); box[3] = 1;
    estr[0] = 1; estr[1] = (cuuint32_t) d->stride; estr[2] = (cuuint32_t) d->stride; estr[3] = 1;
    const CUtensorMapDataType dt = d->dtype == 1 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    CUtensorMap tmA, tmB;
    CUresult r = enc(&tmA, dt, 4, const_cast<void*>(x), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_of(p.bk),
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return csb::fail(CSB_ERR_INVALID, "%s: %s", "csb_conv2d_nhwc", "cuTensorMapEncodeTiled(A) failed");
    const cuuint64_t ktot = (cuuint64_t) d->R * d->S * (p.grouped ? 64 : d->Cin);
    cuuint64_t wdim[2] = {ktot, (cuuint64_t) d->Cout}, wstr[1] = {ktot * esz};
    cuuint32_t wbox[2] = {(cuuint32_t) p.bk, (cuuint32_t) (p.block_n / cg)}, westr[2] = {1, 1};      // a pair's CTAs load half of the weight tile each
    r = enc(&tmB, dt, 2, const_cast<void*>(w), wdim, wstr, wbox, westr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_of(p.bk),
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return csb::fail(CSB_ERR_INVALID, "%s: %s", "csb_conv2d_nhwc", "cuTensorMapEncodeTiled(B) failed");

    const uint32_t row_bytes = p.bk * 2, stage_bytes = ((kBlockM + p.block_n / cg) * row_bytes + 1023u) & ~1023u;
    const int kblocks = p.R * p.S * p.kchunks;
    int stages = (int) ((200u * 1024u) / stage_bytes);
    stages = stages > kMaxStages ? kMaxStages : stages;
    stages = stages > kblocks * 2 ? (kblocks * 2 < 2 ? 2 : kblocks * 2) : stages;
    p.stages = stages < 2 ? 2 : stages;
    // [stages | barriers + tmem slot | pad to 1 KiB | epilogue staging 8 x 2 KiB]
    p.stage_off = (uint32_t) (((size_t) p.stages * stage_bytes + 8 * (2 * kMaxStages + 2 * kAccStages) + 16 + 1023) & ~(size_t) 1023);
    const size_t smem = (size_t) p.stage_off + 4 * kMaxEpiGroups * 2048 + 1024 /*align*/;
    p.bias = bias; p.act = d->act; p.act_param = act_param;
    p.gelu_form = g_gelu_form.load(std::memory_order_relaxed);
    p.residual = residual; p.res_ld = d->res_ld; p.res_coff = d->res_coff; p.res_mode = d->res_mode;
    p.out = y; p.out_f32 = y_f32; p.out_ld = d->out_ld; p.out_coff = d->out_coff; p.is_bf16 = d->dtype == 1;
    // Output tensor map for the epilogue's TMA tile stores: the channel slice [out_coff, out_coff + Cout) of the NHWC output, box = 32 channels x
    // the 32 rows one epilogue warp owns ({32, 32, 1, 1} or {32, bw, 32 / bw, 1}); channels >= Cout and rows outside the image are clipped.
    static const int store_mode = [] { const char* e = getenv("CSB_TMA_STORE"); return e ? atoi(e) : 1; }();
    CUtensorMap tmC = tmA;
    p.tma_store = 0;
    if (store_mode && y && !y_f32 && d->out_ld % 8 == 0 && d->out_coff % 8 == 0 && ((uintptr_t) y & 15) == 0 && d->Cout >= 32) {
        const int cbw = p.bw < 32 ? p.bw : 32, cbh = 32 / cbw;
        cuuint64_t cdim[4] = {(cuuint64_t) d->Cout, (cuuint64_t) p.W, (cuuint64_t) p.H, (cuuint64_t) p.N};
        cuuint64_t cstr[3] = {(cuuint64_t) d->out_ld * esz, (cuuint64_t) d->out_ld * esz * p.W, (cuuint64_t) d->out_ld * esz * p.W * p.H};
        cuuint32_t cbox[4] = {32, (cuuint32_t) cbw, (cuuint32_t) cbh, 1}, cestr[4] = {1, 1, 1, 1};
        char* cbase = reinterpret_cast<char*>(y) + (size_t) d->out_coff * esz;
        r = enc(&tmC, dt, 4, cbase, cdim, cstr, cbox, cestr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                store_mode == 1 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r == CUDA_SUCCESS) p.tma_store = store_mode == 1 ? 1 : 2;
    }
    // the opt-in to > 48 KiB of dynamic shared memory is per device: set it once per (device, kernel)
    static unsigned char attr_done[64] = {};
    if (csb::first_use_on_device(attr_done)) {
        cudaFuncSetAttribute(k_conv_tc<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        cudaFuncSetAttribute(k_conv_tc<3, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        cudaFuncSetAttribute(k_conv_tc<3, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    }
    // 12 epilogue warps (EG = 3, registers re-balanced with setmaxnreg) by default: -1.1 ms on the detector's convs and -0.7 ms on LeReS' per 32
    // frames against the 8-warp build, which stays selectable (CSB_EPI_GROUPS=2) for A/B runs
    const int sms = csb::num_sms();
    if (cg == 2) {
        const int pairs = ((p.tiles_m + 1) / 2) * p.tiles_n;
        const int grid = 2 * (pairs < sms / 2 ? pairs : sms / 2);
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned) grid); cfg.blockDim = dim3(512); cfg.dynamicSmemBytes = smem; cfg.stream = (cudaStream_t) stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        cudaLaunchKernelEx(&cfg, k_conv_tc<3, 2>, tmA, tmB, tmC, p);
    } else {
        const int total = p.tiles_m * p.tiles_n;
        const int grid = total < sms ? total : sms;
        if (eg == 3) k_conv_tc<3, 1><<<grid, 512, smem, (cudaStream_t) stream>>>(tmA, tmB, tmC, p);
        else k_conv_tc<2, 1><<<grid, 384, smem, (cudaStream_t) stream>>>(tmA, tmB, tmC, p);
    }
    if (csb::g_profiling.load(std::memory_order_relaxed) == 2) {           // detailed profile: one key per layer shape
        char label[160];
        snprintf(label, sizeof label, "k_conv_tc[%dx%dx%dx%d->%d k%dx%d s%d d%d g%d act%d res%d]", d->N, d->Hin, d->Win, d->Cin, d->Cout, d->R, d->S, d->stride,
                 d->dil, d->groups, d->act, d->res_mode);
        return csb::launched(csb::profile_intern(label), (cudaStream_t) stream);
    }
    return csb::launched("k_conv_tc", (cudaStream_t) stream);
}

extern "C" int csb_conv2d_nhwc(const csb_conv_desc* d, const void* x, const void* w, const float* bias, const float* act_param,
                               const void* residual, void* y, float* y_f32, void* stream) {
    return conv_impl(d, x, w, bias, act_param, residual, y, y_f32, nullptr, nullptr, 0.f, stream);
}

extern "C" int csb_conv2d_ln_nhwc(const csb_conv_desc* d, const void* x, const void* w, const float* bias, const float* colsum, const float* stats, float eps,
                                  const void* residual, void* y, void* stream) {
    CSB_REQUIRE(stats && colsum, "null pointer");
    return conv_impl(d, x, w, bias, nullptr, residual, y, nullptr, stats, colsum, eps, stream);
}

// ================================================================================================ fused ConvNeXt MLP
// LayerNorm -> Linear C -> 4C -> GELU -> Linear 4C -> C (layer scale folded) -> + residual in ONE kernel (mmpretrain ConvNeXtBlock,
// SURVEY.md Appendix A.4), for the wide-image stages (C = 128 / 256) where the 4C intermediate dominates the HBM traffic of the block (2.1 GB per
// launch at 32 x 256^2 x 512, written by fc1 and read back by fc2): here it never leaves the SM.
//
// Per 128-pixel tile: the activation tile X [128 x C] is loaded once (C / 64 swizzled k-blocks).  The hidden dimension is walked in chunks of HC
// units (128 for C = 128, 64 for C = 256):
//     GEMM1  Hacc[b] (TMEM, HC columns, two buffers) = X . W1[chunk]^T                         M128 x N=HC x K=C
//     EPI1   12 epilogue warps: tcgen05.ld -> folded LayerNorm + bias + GELU -> fp16 -> shared memory, written directly in the canonical K-major
//            128B-swizzled operand layout (16 B chunk index ^ (row & 7)), two buffers
//     GEMM2  Y (TMEM, C columns) += Hs[b] . W2[:, chunk]^T                                      M128 x N=C x K=HC
// and after the last chunk EPI2 = the ordinary epilogue (bias, residual, TMA tile store).  The MMA warp issues GEMM1(j+1) before GEMM2(j), so the
// tensor pipe works on the next chunk while the epilogue warps convert the current one.  Weight chunks stream through a 3-stage ring of 32 KiB
// stages in exactly the order the MMA warp consumes them: W1(0), W1(1), W2(0), W1(2), W2(1), ..., W2(last).
// The arithmetic (k-order of both accumulations, fp16 rounding of the hidden activations) is that of csb_conv2d_ln_nhwc followed by
// csb_conv2d_nhwc, so the results are bit-identical to the two-launch path.
namespace {

struct MlpParams {
    int C, Hd, HC, NJ, k1b, k2b, tiles_m;
    const float* b1;
    const float* colsum;
    const float* stats;
    int ln_nchunk;
    float ln_inv_c, ln_eps;
    uint32_t x_bytes, ring_off, h_off, h_bytes, bar_off;      // shared-memory layout (byte offsets from the 1 KiB-aligned base)
    int sub16;                                                // EPI1 task width: 16 accumulator columns (1) or 32 (0)
    int diag;                                                 // CSB_MLP_DIAG (profiling only, wrong results): 1 skip the GELU, 2 skip the LayerNorm fold, 4 skip the smem store
    int stages;                                               // weight ring depth
    uint32_t stage_bytes;                                     // bytes of one weight chunk over the CTA group (a pair's CTAs hold half each)
    int res_pf;                                               // prefetch the residual rows into L2 at tile start
    uint32_t vec_off;                                         // shared-memory copy of colsum[Hd] | b1[Hd] (fp32), 0 = read them from global memory
    int NB;                                                   // hidden-chunk buffers in flight (Hacc in TMEM, Hs in shared memory): 2 or 3; GEMM1 runs NB - 1 chunks ahead of GEMM2
};

constexpr int kMlpMaxStages = 6;
constexpr int kMlpMaxBuf = 3;
// barriers (8 B each): full[6] empty[6] xfull xempty haccf[3] hacce[3] hsf[3] hse[3] yfull yempty
enum { MB_FULL = 0, MB_EMPTY = 6, MB_XFULL = 12, MB_XEMPTY = 13, MB_HACCF = 14, MB_HACCE = 17, MB_HSF = 20, MB_HSE = 23, MB_YFULL = 26, MB_YEMPTY = 27, MB_COUNT = 28 };

__device__ __forceinline__ void mbar_arrive_cluster_release(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// wait with cluster-scope acquire: the leader's MMA warp consumes shared memory written by the PEER CTA's epilogue warps (they arrive with
// release.cluster after fence.proxy.async)
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    long long t0 = 0;
    for (uint32_t spin = 1;; ++spin) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2, %3;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(bar), "r"(parity), "r"(4000u) : "memory");
        if (ok) return;
        if ((spin & 255u) == 0) {
            if (t0 == 0) t0 = clock64();
            else if (clock64() - t0 > 8000000000ll) __trap();
        }
    }
}
__device__ __forceinline__ float4 ld_shared_f4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
// arrival of an epilogue warp on a barrier that lives in the pair's leader CTA (CG = 2) or in this CTA (CG = 1)
template <int CG>
__device__ __forceinline__ void arrive_lead(uint32_t local_bar) {
    if constexpr (CG == 2) mbar_arrive_cluster_release(mapa_u32(local_bar, 0));
    else mbar_arrive(local_bar);
}

template <int NCOL>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, uint32_t (&v)[NCOL]) {
    if constexpr (NCOL == 32) tmem_ld32(taddr, v);
    else tmem_ld16(taddr, v);
}

// EPI1 of one hidden chunk for one warp: its tasks are the NCOL-column slices first, first + EG, ... of the chunk's HC accumulator columns (TMEM lane
// quarter of the warp).  Per task: tcgen05.ld -> folded LayerNorm + bias + GELU -> fp16 -> shared memory in the K-major 128B-swizzled operand layout
// (k-block = column / 64, row m, 16 B chunk (column % 64) / 8 ^ (m & 7)).  NCOL = 16 balances 8 tasks per quarter over 3 warps as 3 / 3 / 2.
template <class T, int NCOL, int EG, int CG>
__device__ __forceinline__ void mlp_epi1(const ConvKernelParams& p, const MlpParams& q, int first, int j, int b, int m, int lane, uint32_t tm_lane, uint32_t hrow,
                                         uint32_t hacce_bar, uint64_t nm, uint64_t rs, uint32_t vec) {
    const int ntask = q.HC / NCOL;
    int last = -1;
    for (int t = first; t < ntask; t += EG) last = t;
    if (last < 0) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) arrive_lead<CG>(hacce_bar);
    }
    for (int t = first; t < ntask; t += EG) {
        uint32_t acc[NCOL];
        tmem_ld_cols<NCOL>(tm_lane + (uint32_t) (b * q.HC + t * NCOL), acc);
        if (t == last) {                                             // this warp has read all it needs from the accumulator buffer
            tc_fence_before();
            __syncwarp();
            if (lane == 0) arrive_lead<CG>(hacce_bar);
        }
        const int c0 = t * NCOL, n0 = j * q.HC + c0;                 // column inside the chunk, hidden unit
        float y[NCOL];
        if (q.diag & 2) {
#pragma unroll
            for (int i = 0; i < NCOL; ++i) y[i] = __uint_as_float(acc[i]);
        } else
#pragma unroll
        for (int g = 0; g < NCOL / 4; ++g) {                         // folded LayerNorm: rstd * acc + (-mean * rstd * colsum + bias')
            float4 c4, b4;
            if (vec) {                                               // staged once per CTA: with ~218 KiB of shared memory in use the L1 keeps ~10 KiB, and the
                c4 = ld_shared_f4(vec + (uint32_t) (n0 + 4 * g) * 4u);    // per-task __ldg of these vectors was the top stall of the kernel (ncu: long_scoreboard)
                b4 = ld_shared_f4(vec + (uint32_t) (q.Hd + n0 + 4 * g) * 4u);
            } else {
                c4 = __ldg(reinterpret_cast<const float4*>(q.colsum + n0) + g);
                b4 = __ldg(reinterpret_cast<const float4*>(q.b1 + n0) + g);
            }
            upk2(ffma2(rs, pk2(__uint_as_float(acc[4 * g]), __uint_as_float(acc[4 * g + 1])), ffma2(nm, pk2(c4.x, c4.y), pk2(b4.x, b4.y))), y[4 * g], y[4 * g + 1]);
            upk2(ffma2(rs, pk2(__uint_as_float(acc[4 * g + 2]), __uint_as_float(acc[4 * g + 3])), ffma2(nm, pk2(c4.z, c4.w), pk2(b4.z, b4.w))), y[4 * g + 2], y[4 * g + 3]);
        }
        if (q.diag & 1) {
        } else if (p.gelu_form == 1) {
#pragma unroll
            for (int i = 0; i < NCOL; i += 2) gelu2_tanh(y[i], y[i + 1]);
        } else {                                                     // the polynomial / sigmoid mix, pair for pair as epilogue_chunk selects it (position inside the 32-column chunk)
#pragma unroll
            for (int i = 0; i < NCOL; i += 2) {
                const int pi = ((c0 & 31) + i) >> 1;                 // NCOL = 16: c0 & 31 is 0 or 16, warp-uniform
                if (pi * kGeluMufuPairs / 16 != (pi + 1) * kGeluMufuPairs / 16) gelu2_mufu(y[i], y[i + 1]);
                else gelu2(y[i], y[i + 1]);
            }
        }
        const uint32_t rowaddr = hrow + (uint32_t) (c0 >> 6) * 16384u;
        const uint32_t ck0 = (uint32_t) ((c0 & 63) >> 3);
        if (!(q.diag & 4))
#pragma unroll
        for (int g = 0; g < NCOL / 8; ++g)
            st_shared_v4(rowaddr + (((ck0 + (uint32_t) g) ^ (uint32_t) (m & 7)) << 4),
                         make_uint4(pack2<T>(y[8 * g], y[8 * g + 1]), pack2<T>(y[8 * g + 2], y[8 * g + 3]), pack2<T>(y[8 * g + 4], y[8 * g + 5]), pack2<T>(y[8 * g + 6], y[8 * g + 7])));
    }
}

// CG = 2: a CTA pair (one TPC) works on two m-tiles with tcgen05 cta_group::2 (M = 256 per MMA): each CTA loads its own activation tile and HALF of
// every weight chunk, so the weight stream (4.3 GB per launch from L2) crosses the L2 -> SM fabric once per pair.  Kept as an option: correct, not faster.  Barriers as in k_conv_tc<., 2>: TMA "full" barriers live in the leader and receive both CTAs' bytes; MMA completions are multicast commits;
// the epilogue warps of both CTAs arrive on the leader's hacce / hsf / yempty barriers (cluster-scope release).
template <class T, int EG, int CG>
__global__ void __launch_bounds__(128 * (EG + 1), 1) k_mlp_tc(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
                                                   const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmC,
                                                   const ConvKernelParams p, const MlpParams q) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t ring_base = smem_base + q.ring_off, h_base = smem_base + q.h_off, bar_base = smem_base + q.bar_off;
    auto bar = [&](int i) { return bar_base + 8u * (uint32_t) i; };
    const uint32_t tmem_slot = bar(MB_COUNT);
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int NEPI = 4 * EG;
    const int rank = CG == 2 ? (int) cluster_ctarank() : 0;
    const int nstage = q.stages;
    const uint32_t stage_bytes = q.stage_bytes / CG;
    const int NB = q.NB, LEAD = q.NB - 1;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW1) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW2) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmC) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kMlpMaxStages; ++s) { mbar_init(bar(MB_FULL + s), 1); mbar_init(bar(MB_EMPTY + s), 1); }
        mbar_init(bar(MB_XFULL), 1); mbar_init(bar(MB_XEMPTY), 1);
        for (int b = 0; b < kMlpMaxBuf; ++b) {
            mbar_init(bar(MB_HACCF + b), 1); mbar_init(bar(MB_HACCE + b), NEPI * CG);
            mbar_init(bar(MB_HSF + b), NEPI * CG); mbar_init(bar(MB_HSE + b), 1);
        }
        mbar_init(bar(MB_YFULL), 1); mbar_init(bar(MB_YEMPTY), NEPI * CG);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 2) {
        if constexpr (CG == 1) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t) kTmemCols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t) kTmemCols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if constexpr (CG == 2) cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    const uint32_t tm_y = tmem_base, tm_h0 = tmem_base + (uint32_t) q.C;          // Y: columns [0, C), Hacc[b]: [C + b * HC, ...)
    const int cta = CG == 2 ? (int) (blockIdx.x >> 1) : (int) blockIdx.x, ncta = CG == 2 ? (int) (gridDim.x >> 1) : (int) gridDim.x;
    const int ngroups = (q.tiles_m + CG - 1) / CG;                                // tiles of 128 x CG rows
    const uint32_t w1_kb_bytes = (uint32_t) (q.HC / CG) * 128u, w2_kb_bytes = (uint32_t) (q.C / CG) * 128u;      // per CTA

    if (warp < 4) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        if (warp == 0) {
            // ===================================================== TMA producer (one per CTA; in a pair both signal the leader's full barriers)
            if (elect_one()) {
                int stage = 0;
                uint32_t phase = 0, xphase = 0;
                const uint32_t full_lead0 = CG == 2 ? mapa_u32(bar(MB_FULL), 0) : bar(MB_FULL);
                const uint32_t xfull_lead = CG == 2 ? mapa_u32(bar(MB_XFULL), 0) : bar(MB_XFULL);
                auto load_w1 = [&](int j) {
                    mbar_wait(bar(MB_EMPTY + stage), phase ^ 1u);
                    if (rank == 0) mbar_expect_tx(bar(MB_FULL + stage), (uint32_t) (CG * q.k1b) * w1_kb_bytes);
                    for (int kb = 0; kb < q.k1b; ++kb)
                        tma_load_2d<CG>(ring_base + (uint32_t) stage * stage_bytes + (uint32_t) kb * w1_kb_bytes, &tmW1, full_lead0 + 8u * stage, kb * 64,
                                        j * q.HC + rank * (q.HC / CG));
                    if (++stage == nstage) { stage = 0; phase ^= 1u; }
                };
                auto load_w2 = [&](int j) {
                    mbar_wait(bar(MB_EMPTY + stage), phase ^ 1u);
                    if (rank == 0) mbar_expect_tx(bar(MB_FULL + stage), (uint32_t) (CG * q.k2b) * w2_kb_bytes);
                    for (int kb = 0; kb < q.k2b; ++kb)
                        tma_load_2d<CG>(ring_base + (uint32_t) stage * stage_bytes + (uint32_t) kb * w2_kb_bytes, &tmW2, full_lead0 + 8u * stage, j * q.HC + kb * 64,
                                        rank * (q.C / CG));
                    if (++stage == nstage) { stage = 0; phase ^= 1u; }
                };
                for (int grp = cta; grp < ngroups; grp += ncta) {
                    const int tile = grp * CG + rank;                         // a tile past the end (odd tile count) loads zero rows and stores nothing
                    mbar_wait(bar(MB_XEMPTY), xphase ^ 1u);
                    if (rank == 0) mbar_expect_tx(bar(MB_XFULL), CG * q.x_bytes);
                    for (int kb = 0; kb < q.k1b; ++kb)
                        tma_load_4d<CG>(smem_base + (uint32_t) kb * 16384u, &tmX, xfull_lead, p.in_coff + kb * 64, tile * kBlockM, 0, 0);
                    xphase ^= 1u;
                    for (int j = 0; j < q.NJ + LEAD; ++j) {               // the MMA warp's consumption order
                        if (j < q.NJ) load_w1(j);
                        if (j >= LEAD) load_w2(j - LEAD);
                    }
                }
            }
        } else if (warp == 1 && rank == 0) {
            // ===================================================== MMA issuer (the leader CTA of a pair)
            const uint32_t fmt = p.is_bf16 ? 1u : 0u;
            const uint32_t idesc1 = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t) (q.HC >> 3) << 17) | ((uint32_t) ((kBlockM * CG) >> 4) << 24);
            const uint32_t idesc2 = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t) (q.C >> 3) << 17) | ((uint32_t) ((kBlockM * CG) >> 4) << 24);
            int stage = 0;
            uint32_t phase = 0, xphase = 0, yphase = 0;
            // chunks use the buffers round-robin over the whole run of the CTA: the g-th chunk uses buffer g % NB for the (g / NB)-th time
            int b1 = 0, b2 = 0;
            uint32_t u1 = 0, u2 = 0;
            auto gemm1 = [&](int j, bool last) {
                const int b = b1;
                const uint32_t u = u1;
                if (++b1 == NB) { b1 = 0; ++u1; }
                mbar_wait(bar(MB_HACCE + b), (u & 1u) ^ 1u);                     // the epilogue warps (of both CTAs) have read this accumulator buffer
                mbar_wait(bar(MB_FULL + stage), phase);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t wb = ring_base + (uint32_t) stage * stage_bytes;
                    for (int kb = 0; kb < q.k1b; ++kb) {
                        const uint64_t adesc = make_desc(smem_base + (uint32_t) kb * 16384u, 128), bdesc = make_desc(wb + (uint32_t) kb * w1_kb_bytes, 128);
                        for (int k = 0; k < 4; ++k) umma_f16<CG>(tm_h0 + (uint32_t) (b * q.HC), adesc + 2u * k, bdesc + 2u * k, idesc1, (kb | k) != 0 ? 1u : 0u);
                    }
                    umma_commit<CG>(bar(MB_EMPTY + stage));
                    umma_commit<CG>(bar(MB_HACCF + b));
                    if (last) umma_commit<CG>(bar(MB_XEMPTY));                       // X tile free: the producers may load the next tile's activations
                }
                __syncwarp();
                if (++stage == nstage) { stage = 0; phase ^= 1u; }
            };
            auto gemm2 = [&](int j, bool last) {
                const int b = b2;
                const uint32_t u = u2;
                if (++b2 == NB) { b2 = 0; ++u2; }
                // the epilogue warps (of both CTAs) have written this chunk's fp16 activations
                if constexpr (CG == 2) mbar_wait_cluster(bar(MB_HSF + b), u & 1u);
                else mbar_wait(bar(MB_HSF + b), u & 1u);
                mbar_wait(bar(MB_FULL + stage), phase);
                if (j == 0) { mbar_wait(bar(MB_YEMPTY), yphase ^ 1u); yphase ^= 1u; } // the previous tile's output accumulator has been read
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t wb = ring_base + (uint32_t) stage * stage_bytes, hb = h_base + (uint32_t) b * q.h_bytes;
                    for (int kb = 0; kb < q.k2b; ++kb) {
                        const uint64_t adesc = make_desc(hb + (uint32_t) kb * 16384u, 128), bdesc = make_desc(wb + (uint32_t) kb * w2_kb_bytes, 128);
                        for (int k = 0; k < 4; ++k) umma_f16<CG>(tm_y, adesc + 2u * k, bdesc + 2u * k, idesc2, (j | kb | k) != 0 ? 1u : 0u);
                    }
                    umma_commit<CG>(bar(MB_EMPTY + stage));
                    umma_commit<CG>(bar(MB_HSE + b));
                    if (last) umma_commit<CG>(bar(MB_YFULL));
                }
                __syncwarp();
                if (++stage == nstage) { stage = 0; phase ^= 1u; }
            };
            for (int grp = cta; grp < ngroups; grp += ncta) {
                mbar_wait(bar(MB_XFULL), xphase);
                xphase ^= 1u;
                tc_fence_after();
                for (int j = 0; j < q.NJ + LEAD; ++j) {                       // GEMM1 runs LEAD chunks ahead of GEMM2
                    if (j < q.NJ) gemm1(j, j == q.NJ - 1);
                    if (j >= LEAD) gemm2(j - LEAD, j - LEAD == q.NJ - 1);
                }
            }
        }
    } else {
        // ===================================================== epilogue warps: EPI1 per hidden chunk, EPI2 per tile
        if constexpr (EG == 3) asm volatile("setmaxnreg.inc.sync.aligned.u32 152;");
        else asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");      // 16 epilogue warps: the pool is the CTA's launch allocation (640 x 96 = 61440): 4 x 32 x 40 + 16 x 32 x 104 = 58368
        const int qd = warp & 3, grpw = (warp - 4) >> 2;
        const bool aligned = ((p.out_ld | p.out_coff) % 8 == 0) && (!p.res_mode || ((p.res_ld | p.res_coff) % 8 == 0)) && (!p.bias || ((uintptr_t) p.bias % 16 == 0));
        const uint32_t stage_any = smem_base + p.stage_off + (uint32_t) (warp - 4) * 2048u;
        const uint32_t lane_base = ((uint32_t) (qd * 32) << 16);
        const int nc2 = q.C / 32;
        uint32_t ue = 0, yphase = 0;
        int be = 0, rot = grpw;
        const uint32_t vec = q.vec_off ? smem_base + q.vec_off : 0u;
        if (vec) {                                                   // colsum | b1 -> shared memory, once per CTA (epilogue warps only: named barrier 1)
            for (int i = (int) threadIdx.x - 128; i < q.Hd; i += NEPI * 32) {
                asm volatile("st.shared.f32 [%0], %1;" ::"r"(vec + (uint32_t) i * 4u), "f"(__ldg(q.colsum + i)) : "memory");
                asm volatile("st.shared.f32 [%0], %1;" ::"r"(vec + (uint32_t) (q.Hd + i) * 4u), "f"(__ldg(q.b1 + i)) : "memory");
            }
            asm volatile("bar.sync 1, %0;" ::"r"(NEPI * 32) : "memory");
        }
        for (int grp = cta; grp < ngroups; grp += ncta) {
            const int tile = grp * CG + rank;
            const int m = qd * 32 + lane;
            const long long row = (long long) tile * kBlockM + m;
            const bool row_ok = row < p.W;
            const size_t pix = (size_t) (row_ok ? row : (long long) p.W - 1);
            float neg_mean, rstd;
            {
                const float2* sp = reinterpret_cast<const float2*>(q.stats) + pix * q.ln_nchunk;
                float s1 = 0.f, s2 = 0.f;
                for (int c = 0; c < q.ln_nchunk; ++c) { const float2 t = __ldg(sp + c); s1 += t.x; s2 += t.y; }
                const float mean = s1 * q.ln_inv_c;
                rstd = rsqrtf(fmaxf(fmaf(s2, q.ln_inv_c, -mean * mean), 0.f) + q.ln_eps);
                neg_mean = -mean * rstd;
            }
            const uint64_t nm = pk2(neg_mean, neg_mean), rs = pk2(rstd, rstd);
            if (q.res_pf && grpw == 0 && p.res_mode && row_ok) {     // the residual row EPI2 will add: pull it into L2 now (EPI2's loads were exposed HBM latency)
                const char* rrow = reinterpret_cast<const char*>(reinterpret_cast<const T*>(p.residual) + pix * p.res_ld + p.res_coff);
                for (int o = 0; o < q.C * 2; o += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(rrow + o));
            }
            for (int j = 0; j < q.NJ; ++j) {
                const int b = be;
                const uint32_t u = ue;
                if (++be == NB) { be = 0; ++ue; }
                const int first = rot;                                       // this warp group's chunks of the hidden chunk: first, first + EG, ...
                if (++rot == EG) rot = 0;
                mbar_wait(bar(MB_HACCF + b), u & 1u);
                tc_fence_after();
                mbar_wait(bar(MB_HSE + b), (u & 1u) ^ 1u);                    // GEMM2 of the chunk that used this shared-memory buffer before has completed
                if (q.sub16) mlp_epi1<T, 16, EG, CG>(p, q, first, j, b, m, lane, tm_h0 + lane_base, h_base + (uint32_t) b * q.h_bytes + (uint32_t) m * 128u, bar(MB_HACCE + b), nm, rs, vec);
                else mlp_epi1<T, 32, EG, CG>(p, q, first, j, b, m, lane, tm_h0 + lane_base, h_base + (uint32_t) b * q.h_bytes + (uint32_t) m * 128u, bar(MB_HACCE + b), nm, rs, vec);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core's async-proxy reads
                __syncwarp();
                if (lane == 0) arrive_lead<CG>(bar(MB_HSF + b));
            }
            // ---- EPI2: the block's output tile (bias, residual, TMA tile store), as k_conv_tc's epilogue
            mbar_wait(bar(MB_YFULL), yphase);
            yphase ^= 1u;
            tc_fence_after();
            const uint32_t stage = p.tma_store ? stage_any : 0u;
            const int cw = tile * kBlockM + qd * 32;
            int last2 = -1;
            for (int ch = grpw; ch < nc2; ch += EG) last2 = ch;
            if (last2 < 0) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) arrive_lead<CG>(bar(MB_YEMPTY));
            }
            for (int ch = grpw; ch < nc2; ch += EG) {
                uint32_t acc[32];
                tmem_ld32(tm_y + lane_base + (uint32_t) (ch * 32), acc);
                if (ch == last2) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) arrive_lead<CG>(bar(MB_YEMPTY));
                }
                epilogue_chunk<T, CSB_ACT_NONE, false>(p, acc, pix, ch * 32, row_ok, aligned && ch * 32 + 32 <= p.Cout, stage, &tmC, cw, 0, 0, 0.f, 1.f, stage_any);
            }
        }
        if (p.tma_store && lane == 0) tma_store_wait_all();
    }
    tc_fence_before();
    __syncthreads();
    if constexpr (CG == 2) cluster_sync_all();          // neither CTA leaves (or frees TMEM) while its peer may still signal it / read its shared memory
    if (warp == 2) {
        tc_fence_after();
        if constexpr (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t) kTmemCols) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t) kTmemCols) : "memory");
    }
}

std::atomic<int> g_mlp_mode{[] { const char* e = getenv("CSB_FUSE_MLP"); return e ? atoi(e) : 1; }()};

}  // namespace

// 1 if csb_convnext_mlp_nhwc can run this block shape (C = 128 or 256, hidden = 4C) and the fused path is enabled (CSB_FUSE_MLP, default 1).
extern "C" int csb_convnext_mlp_supported(int C, int hidden) { return g_mlp_mode.load(std::memory_order_relaxed) != 0 && (C == 128 || C == 256) && hidden == 4 * C; }
extern "C" int csb_convnext_mlp_set_mode(int mode) { return g_mlp_mode.exchange(mode); }

extern "C" int csb_convnext_mlp_nhwc(const void* x, int x_ld, int x_coff, long long pixels, int C, int hidden, const void* w1, const float* b1, const float* colsum,
                                     const float* stats, float eps, const void* w2, const float* b2, const void* residual, int res_ld, int res_coff, void* y,
                                     int y_ld, int y_coff, int dtype, void* stream) {
    CSB_REQUIRE(x && w1 && b1 && colsum && stats && w2 && y, "null pointer");
    CSB_REQUIRE((C == 128 || C == 256) && hidden == 4 * C && pixels > 0 && pixels < (1ll << 31), "fused MLP: C must be 128 or 256 with hidden = 4 C");
    CSB_REQUIRE(x_ld % 8 == 0 && x_coff % 8 == 0 && y_ld % 8 == 0 && y_coff % 8 == 0 && (!residual || (res_ld % 8 == 0 && res_coff % 8 == 0)), "channel strides / offsets must be multiples of 8");
    CSB_REQUIRE((((uintptr_t) x | (uintptr_t) w1 | (uintptr_t) w2 | (uintptr_t) y | (uintptr_t) b1 | (uintptr_t) colsum | (uintptr_t) b2) & 15) == 0 && ((uintptr_t) stats & 7) == 0,
                "pointers must be 16-byte aligned");
    EncodeTiledFn enc = encode_fn();
    if (!enc) return csb::fail(CSB_ERR_CUDA, "%s: %s", "csb_convnext_mlp_nhwc", "cuTensorMapEncodeTiled unavailable");
    MlpParams q{};
    // C = 128: two chunk buffers of 128 hidden units; CSB_MLP_NB=3 = three buffers of 64 units with GEMM1 two chunks ahead of GEMM2 (measured slower,
    // gpurun r2c39: 1366 against 1282 us -- twice the hand-offs per tile).  C = 256: two buffers of 64 units (shared memory: the X tile is 64 KiB).
    static const int nb_env = [] { const char* e = getenv("CSB_MLP_NB"); return e ? atoi(e) : 2; }();
    q.NB = (C == 128 && nb_env == 3) ? 3 : 2;
    q.C = C; q.Hd = hidden; q.HC = (C == 128 && q.NB == 2) ? 128 : 64; q.NJ = hidden / q.HC; q.k1b = C / 64; q.k2b = q.HC / 64;
    q.tiles_m = (int) ((pixels + kBlockM - 1) / kBlockM);
    q.b1 = b1; q.colsum = colsum; q.stats = stats; q.ln_nchunk = C / 64; q.ln_inv_c = 1.0f / (float) C; q.ln_eps = eps;
    q.x_bytes = (uint32_t) q.k1b * 16384u;
    q.ring_off = q.x_bytes;
    // CTA pairs (CSB_MLP_PAIR=1; off by default): each CTA holds half of every weight chunk (16 KiB stages, 6 deep).  Bit-identical, but measured
    // slower (gpurun r2c38: 1706 / 1119 us against 1284 / 801 us): the kernel is bound by the per-chunk GEMM1 -> epilogue -> GEMM2 hand-offs, which the
    // cross-CTA arrivals lengthen, not by the weight stream it halves.
    static const int pair_env = [] { const char* e = getenv("CSB_MLP_PAIR"); return e ? atoi(e) : 0; }();
    static const int eg = [] { const char* e = getenv("CSB_MLP_EG"); return e && atoi(e) == 4 ? 4 : 3; }();      // 12 (default) or 16 epilogue warps
    const int cg = (pair_env && eg == 3 && dtype == 0 && q.tiles_m >= 2) ? 2 : 1;      // (the experimental forms are built for fp16 only)
    q.stage_bytes = (uint32_t) q.HC * (uint32_t) C * 2u;             // W1 chunk [HC x C] = W2 chunk [C x HC]: 16 or 32 KiB
    q.stages = (int) (98304u / q.stage_bytes) * cg;                    // 96 KiB of ring per CTA
    if (q.stages > kMlpMaxStages) q.stages = kMlpMaxStages;
    q.h_off = q.ring_off + (uint32_t) q.stages * (q.stage_bytes / cg);
    q.h_bytes = (uint32_t) q.k2b * 16384u;
    q.bar_off = q.h_off + (uint32_t) q.NB * q.h_bytes;
    { const char* e = getenv("CSB_MLP_SUB16"); q.sub16 = e ? atoi(e) : 0; }     // measured (gpurun r2c28): 32-column tasks 1237 / 781 us, 16-column 1261 / 878 us
    // CSB_MLP_EG=4: 16 epilogue warps (640 threads, 104 registers each) on 16-column tasks: 8 / 4 tasks per TMEM lane quarter split evenly over 4 warps.
    // Measured (gpurun r2c36): 1252 / 857 us against 1238 / 836 us with 12 warps -- neither more warps nor the even split moves the GELU epilogue.
    const int eg_run = (eg == 4 && dtype == 0) ? 4 : 3;
    if (eg_run == 4) q.sub16 = 1;
    { const char* e = getenv("CSB_MLP_DIAG"); q.diag = e ? atoi(e) : 0; }
    ConvKernelParams p{};
    p.N = 1; p.H = 1; p.W = (int) pixels; p.Cin = hidden; p.Cout = C; p.R = p.S = 1; p.stride = 1; p.bh = 1; p.bw = kBlockM;
    p.tiles_h = 1; p.tiles_w = q.tiles_m; p.tiles_m = q.tiles_m; p.tiles_n = 1; p.block_n = C; p.in_coff = x_coff;
    p.bias = b2; p.residual = residual; p.res_ld = res_ld; p.res_coff = res_coff; p.res_mode = residual ? 2 : 0; p.act = CSB_ACT_NONE;
    p.out = y; p.out_ld = y_ld; p.out_coff = y_coff; p.is_bf16 = dtype == 1;
    p.gelu_form = g_gelu_form.load(std::memory_order_relaxed);
    p.stage_off = (q.bar_off + 8u * (MB_COUNT + 2) + 1023u) & ~1023u;
    // [X | ring | Hs | barriers | pad | epilogue staging (2 KiB per epilogue warp) | colsum, b1 copies]
    static const int vec_env = [] { const char* e = getenv("CSB_MLP_VEC_SMEM"); return e ? atoi(e) : 1; }();
    { const char* e = getenv("CSB_MLP_RES_PF"); q.res_pf = e ? atoi(e) : 1; }
    const uint32_t staging = (uint32_t) (4 * eg_run) * 2048u, vec_bytes = 2u * (uint32_t) hidden * 4u;
    q.vec_off = (vec_env && (size_t) p.stage_off + staging + vec_bytes + 1024 <= 227 * 1024) ? p.stage_off + staging : 0u;
    const size_t smem = (size_t) p.stage_off + staging + (q.vec_off ? vec_bytes : 0u) + 1024;
    CSB_REQUIRE(smem <= 227 * 1024, "fused MLP: shared-memory budget exceeded");
    const CUtensorMapDataType dt = dtype == 1 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    const cuuint64_t esz = 2;
    CUtensorMap tmX, tmW1, tmW2, tmC;
    {
        cuuint64_t gdim[4] = {(cuuint64_t) x_ld, (cuuint64_t) pixels, 1, 1}, gstr[3] = {(cuuint64_t) x_ld * esz, (cuuint64_t) x_ld * esz * pixels, (cuuint64_t) x_ld * esz * pixels};
        cuuint32_t box[4] = {64, (cuuint32_t) kBlockM, 1, 1}, estr[4] = {1, 1, 1, 1};
        if (enc(&tmX, dt, 4, const_cast<void*>(x), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return csb::fail(CSB_ERR_INVALID, "%s: %s", "csb_convnext_mlp_nhwc", "cuTensorMapEncodeTiled(X) failed");
    }
    {
        cuuint64_t wdim[2] = {(cuuint64_t) C, (cuuint64_t) hidden}, wstr[1] = {(cuuint64_t) C * esz};
        cuuint32_t wbox[2] = {64, (cuuint32_t) (q.HC / cg)}, westr[2] = {1, 1};          // a pair's CTAs load half of the chunk's rows each
        if (enc(&tmW1, dt, 2, const_cast<void*>(w1), wdim, wstr, wbox, westr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return csb::fail(CSB_ERR_INVALID, "%s: %s", "csb_convnext_mlp_nhwc", "cuTensorMapEncodeTiled(W1) failed");
        cuuint64_t vdim[2] = {(cuuint64_t) hidden, (cuuint64_t) C}, vstr[1] = {(cuuint64_t) hidden * esz};
        cuuint32_t vbox[2] = {64, (cuuint32_t) (C / cg)};
        if (enc(&tmW2, dt, 2, const_cast<void*>(w2), vdim, vstr, vbox, westr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return csb::fail(CSB_ERR_INVALID, "%s: %s", "csb_convnext_mlp_nhwc", "cuTensorMapEncodeTiled(W2) failed");
    }
    {
        cuuint64_t cdim[4] = {(cuuint64_t) C, (cuuint64_t) pixels, 1, 1};
        cuuint64_t cstr[3] = {(cuuint64_t) y_ld * esz, (cuuint64_t) y_ld * esz * pixels, (cuuint64_t) y_ld * esz * pixels};
        cuuint32_t cbox[4] = {32, 32, 1, 1}, cestr[4] = {1, 1, 1, 1};
        char* cbase = reinterpret_cast<char*>(y) + (size_t) y_coff * esz;
        if (enc(&tmC, dt, 4, cbase, cdim, cstr, cbox, cestr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return csb::fail(CSB_ERR_INVALID, "%s: %s", "csb_convnext_mlp_nhwc", "cuTensorMapEncodeTiled(C) failed");
        p.tma_store = 1;
    }
    static unsigned char attr_done[64] = {};
    if (csb::first_use_on_device(attr_done)) {
        cudaFuncSetAttribute(k_mlp_tc<__half, 3, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        cudaFuncSetAttribute(k_mlp_tc<__nv_bfloat16, 3, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        cudaFuncSetAttribute(k_mlp_tc<__half, 4, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        cudaFuncSetAttribute(k_mlp_tc<__half, 3, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    }
    const int sms = csb::num_sms();
    if (cg == 2) {
        const int pairs = (q.tiles_m + 1) / 2;
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned) (2 * (pairs < sms / 2 ? pairs : sms / 2))); cfg.blockDim = dim3(512); cfg.dynamicSmemBytes = smem; cfg.stream = (cudaStream_t) stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        cudaLaunchKernelEx(&cfg, k_mlp_tc<__half, 3, 2>, tmX, tmW1, tmW2, tmC, p, q);
    } else {
        const int grid = q.tiles_m < sms ? q.tiles_m : sms;
        if (eg_run == 4) k_mlp_tc<__half, 4, 1><<<grid, 640, smem, (cudaStream_t) stream>>>(tmX, tmW1, tmW2, tmC, p, q);
        else if (dtype == 1) k_mlp_tc<__nv_bfloat16, 3, 1><<<grid, 512, smem, (cudaStream_t) stream>>>(tmX, tmW1, tmW2, tmC, p, q);
        else k_mlp_tc<__half, 3, 1><<<grid, 512, smem, (cudaStream_t) stream>>>(tmX, tmW1, tmW2, tmC, p, q);
    }
    if (csb::g_profiling.load(std::memory_order_relaxed) == 2) {
        char label[160];
        snprintf(label, sizeof label, "k_mlp_tc[%lldx%d->%d->%d]", pixels, C, hidden, C);
        return csb::launched(csb::profile_intern(label), (cudaStream_t) stream);
    }
    return csb::launched("k_mlp_tc", (cudaStream_t) stream);
}
