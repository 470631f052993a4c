// Halo-tile implicit-GEMM convolution on tcgen05 (sm_100a): stride-1 RxS convolutions whose k-blocks are too thin for the per-tap pipeline of
// tc_conv.cu -- the narrow 3x3 layers of ISNet / the Inpaint and Refine nets, the grouped 3x3 convolutions of
// ResNeXt-101 32x8d (row B6) and the depthwise 7x7 / 5x5 convolutions of ConvNeXt / CSPNeXt (row A2), which become block-diagonal MMAs.
//
// k_conv_tc loads, for every filter tap, a fresh shifted 128-pixel activation box: R*S x the activation bytes cross the L2 -> shared-memory
// fabric, and with N <= 128 output channels that traffic -- not the tensor pipe -- bounds the layer (measured: 0.25-0.5 PFLOP/s, 11 TB/s L2 -> SM).
// Here an output tile is 16 rows x 8 columns of one image and the producer loads ONE halo box {64 channels, 16 pixels, 16 + (R-1) dil rows} per
// 64-channel chunk.  Every tap (r, s) is then the same shared-memory tile read through a shifted UMMA descriptor:
//     start address += ((r dil) * 16 + s dil) * 128 B,   stride between 8-row groups (SBO) = one halo row = 2 KiB,   matrix base offset = 0
// so the 8 consecutive pixels of an output row are 8 consecutive 128 B rows of shared memory and consecutive output rows are SBO apart.  The start
// address is then NOT aligned to the 1 KiB swizzle pattern.  Measured on the B200 (tests/test_halo_gpu.py, both settings): the tensor core applies
// the 128B-swizzle XOR to the ABSOLUTE shared-memory address bits [7:9] -> [4:6], exactly as TMA wrote the tile, so a shifted start needs no
// correction; setting the descriptor's "matrix base offset" field to (start >> 7) & 7 gives wrong results (kept as diagnostic mode 2).
// Weights: one tile per tap, either streamed through a ring or -- when all taps fit -- resident in shared memory for as long as the CTA stays on
// the same n-tile (tiles are handed out as contiguous ranges in n-major order, so a CTA reloads them at most twice).
// Grouped convolutions (groups > 1, w given compactly as [Cout][R][S][gk], gk = max(16, Cin / groups)) issue, per tap and 64-channel slice,
// 64 / gk MMAs of N = K = gk on the diagonal blocks: no multiply by the zeros of the block-diagonal form.  Depthwise = groups == C, gk = 16.
//
// Roles as in tc_conv.cu: warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator, 12 epilogue warps (setmaxnreg 40 / 152).
// Roofline: tensor pipe for the dense layers (2 * pixels * Cout * R * S * Cin FLOP); the grouped / depthwise layers execute gk / cpg times
// their useful FLOPs on N = 16 / 32 MMAs and sit between the tensor roof and the HBM floor (4 C B per pixel).
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>

#include <cstdlib>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace {

constexpr int kTileW = 8, kTileH = 16, kPitch = 16;          // output tile 16 x 8 = 128 pixels; halo rows are 16 pixels = 2 KiB apart
constexpr uint32_t kRowB = 128, kSbo = kPitch * kRowB;
constexpr int kMaxA = 4, kMaxB = 64, kMaxAcc = 8;
constexpr int kEpiGroups = 3;

struct HaloParams {
    int N, H, W;                 // output geometry
    int Cout;
    int R, S, dil, pad, taps;
    int kchunks;                 // 64-channel chunks of the contraction (grouped: 1, the slice of the n-tile)
    int tiles_h, tiles_w, tiles_m, tiles_n, block_n;
    int in_coff, grouped;
    int gk, nsub, n_mma;         // per tap and chunk: nsub MMAs series of N = n_mma columns and K = kq channels each
    int kq;
    uint32_t b_row_bytes, b_tile_bytes;
    int b_slots, b_resident;
    int a_stages;
    int n_iss;                   // MMA-issuing warps (resident weights only)
    uint32_t halo_bytes;
    int acc_stages, acc_stride;
    int base_off;                // 0 (correct on sm_100a): descriptor base offset 0; 1: (start >> 7) & 7 (diagnostic mode, see the header comment)
    uint32_t b_off, bar_off;
    // epilogue
    const float* bias;
    const void* residual;
    int res_ld, res_coff, res_mode;
    int act;
    const float* act_param;
    void* out;
    float* out_f32;
    int out_ld, out_coff;
    int is_bf16;
    float* stats;                // depthwise + LayerNorm statistics: [pixels][stats_nchunk][2] (sum, sum of squares) of the rounded outputs
    int stats_nchunk;
};

__device__ __forceinline__ uint64_t make_desc_ex(uint32_t saddr, uint32_t sbo_bytes, uint32_t base_off, uint64_t layout) {
    return (uint64_t) ((saddr & 0x3ffff) >> 4) | (1ull << 16) | ((uint64_t) (sbo_bytes >> 4) << 32) | (1ull << 46) | ((uint64_t) (base_off & 7u) << 49) |
           (layout << 61);
}

template <int ACT>
__device__ __forceinline__ float halo_act(float t, float slope) {
    if constexpr (ACT == CSB_ACT_RELU) return fmaxf(t, 0.f);
    else if constexpr (ACT == CSB_ACT_SILU) return __fdividef(t, 1.0f + __expf(-t));
    else if constexpr (ACT == CSB_ACT_PRELU) return t > 0.f ? t : t * slope;
    else return t;
}

// Epilogue of the halo kernel.  The layers on this path are thin (<= 128 output channels, often 9..36 MMAs per tile), so the conversion of a
// tile must be as cheap as the MMAs that produced it: activation and residual mode are template parameters, bias / residual are 16 B loads,
// a full 32-column chunk is straight-line code (the tail path handles Cout % 32 != 0, fp32 outputs and unaligned slices).
template <class T, int ACT>
__device__ __forceinline__ void halo_epilogue(const HaloParams& p, uint32_t tmem_base, uint32_t tfull0, uint32_t tempty0, int nt, int m0, int m1) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = warp & 3, grp = (warp - 4) >> 2;
    const bool aligned = ((p.out_ld | p.out_coff) % 8 == 0) && (!p.res_mode || ((p.res_ld | p.res_coff) % 8 == 0)) && !p.out_f32 &&
                         (!p.bias || ((uintptr_t) p.bias % 16 == 0));
    int as = 0, rot = grp;
    uint32_t aphase = 0;
    for (int mt = m0; mt < m1; ++mt) {
        const int turn = rot;                               // see tc_conv.cu: the groups' starting chunk rotates from tile to tile
        if (++rot == kEpiGroups) rot = 0;
        const int tw = mt % p.tiles_w, th = (mt / p.tiles_w) % p.tiles_h, img = mt / (p.tiles_w * p.tiles_h);
        const int m = q * 32 + lane;
        const int oh = th * kTileH + (m >> 3), ow = tw * kTileW + (m & 7);
        const bool row_ok = oh < p.H && ow < p.W;
        const size_t pix = ((size_t) img * p.H + (oh < p.H ? oh : p.H - 1)) * p.W + (ow < p.W ? ow : p.W - 1);
        const int ncols = min(p.block_n, p.Cout - nt * p.block_n);
        const int nchunks = (ncols + 31) / 32;
        // statistics mode: ONE group converts the whole tile (it owns complete rows of the 64-channel slice); otherwise chunks go round-robin
        const int first = p.stats ? (turn == 0 ? 0 : nchunks) : turn, step = p.stats ? 1 : kEpiGroups;
        int last = -1;
        for (int ch = first; ch < nchunks; ch += step) last = ch;
        mbar_wait(tfull0 + 8u * as, aphase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t) (q * 32) << 16) + (uint32_t) (as * p.acc_stride);
        if (last < 0) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty0 + 8u * as);
        }
        float s1 = 0.f, s2 = 0.f;
        for (int ch = first; ch < nchunks; ch += step) {
            const int n0 = nt * p.block_n + ch * 32;
            const int nv = min(32, p.Cout - n0);
            uint32_t acc[32];
            tmem_ld32(taddr + (uint32_t) ch * 32u, acc);
            if (ch == last) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tempty0 + 8u * as);
            }
            if (!row_ok) continue;
            if (aligned && nv % 8 == 0) {
                const T* res = p.res_mode ? reinterpret_cast<const T*>(p.residual) + pix * p.res_ld + p.res_coff + n0 : nullptr;
                T* o = reinterpret_cast<T*>(p.out) + pix * p.out_ld + p.out_coff + n0;
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    if (g * 8 >= nv) break;
                    float y[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) y[e] = __uint_as_float(acc[g * 8 + e]);
                    if (p.bias) {
                        const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + n0) + 2 * g), b1 = __ldg(reinterpret_cast<const float4*>(p.bias + n0) + 2 * g + 1);
                        y[0] += b0.x; y[1] += b0.y; y[2] += b0.z; y[3] += b0.w; y[4] += b1.x; y[5] += b1.y; y[6] += b1.z; y[7] += b1.w;
                    }
                    float r[8];
                    if (p.res_mode) {
                        const uint4 u = *reinterpret_cast<const uint4*>(res + g * 8);
                        float2 f;
                        f = unpack2<T>(u.x); r[0] = f.x; r[1] = f.y;
                        f = unpack2<T>(u.y); r[2] = f.x; r[3] = f.y;
                        f = unpack2<T>(u.z); r[4] = f.x; r[5] = f.y;
                        f = unpack2<T>(u.w); r[6] = f.x; r[7] = f.y;
                        if (p.res_mode == 1) {
#pragma unroll
                            for (int e = 0; e < 8; ++e) y[e] += r[e];
                        }
                    }
                    if constexpr (ACT == CSB_ACT_PRELU) {
                        float sl[8];
                        if (p.act_param) {
                            const float4 a0 = __ldg(reinterpret_cast<const float4*>(p.act_param + n0) + 2 * g), a1 = __ldg(reinterpret_cast<const float4*>(p.act_param + n0) + 2 * g + 1);
                            sl[0] = a0.x; sl[1] = a0.y; sl[2] = a0.z; sl[3] = a0.w; sl[4] = a1.x; sl[5] = a1.y; sl[6] = a1.z; sl[7] = a1.w;
                        } else {
#pragma unroll
                            for (int e = 0; e < 8; ++e) sl[e] = 0.25f;
                        }
#pragma unroll
                        for (int e = 0; e < 8; ++e) y[e] = halo_act<ACT>(y[e], sl[e]);
                    } else {
#pragma unroll
                        for (int e = 0; e < 8; ++e) y[e] = halo_act<ACT>(y[e], 0.f);
                    }
                    if (p.res_mode == 2) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) y[e] += r[e];
                    }
                    const uint4 pk = make_uint4(pack2<T>(y[0], y[1]), pack2<T>(y[2], y[3]), pack2<T>(y[4], y[5]), pack2<T>(y[6], y[7]));
                    if (p.stats) {                          // statistics of the ROUNDED values, i.e. of what the consumer will read
                        float2 f;
                        f = unpack2<T>(pk.x); s1 += f.x + f.y; s2 = fmaf(f.x, f.x, fmaf(f.y, f.y, s2));
                        f = unpack2<T>(pk.y); s1 += f.x + f.y; s2 = fmaf(f.x, f.x, fmaf(f.y, f.y, s2));
                        f = unpack2<T>(pk.z); s1 += f.x + f.y; s2 = fmaf(f.x, f.x, fmaf(f.y, f.y, s2));
                        f = unpack2<T>(pk.w); s1 += f.x + f.y; s2 = fmaf(f.x, f.x, fmaf(f.y, f.y, s2));
                    }
                    *reinterpret_cast<uint4*>(o + g * 8) = pk;
                }
            } else {
                const T* res = p.res_mode ? reinterpret_cast<const T*>(p.residual) + pix * p.res_ld + p.res_coff + n0 : nullptr;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    if (j >= nv) break;
                    const int c = n0 + j;
                    float t = __uint_as_float(acc[j]) + (p.bias ? __ldg(p.bias + c) : 0.f);
                    if (p.res_mode == 1) t += to_f<T>(res[j]);
                    t = halo_act<ACT>(t, ACT == CSB_ACT_PRELU && p.act_param ? __ldg(p.act_param + c) : 0.25f);
                    if (p.res_mode == 2) t += to_f<T>(res[j]);
                    if (p.out_f32) p.out_f32[pix * p.out_ld + p.out_coff + c] = t;
                    else {
                        const T h = from_f<T>(t);
                        reinterpret_cast<T*>(p.out)[pix * p.out_ld + p.out_coff + c] = h;
                        if (p.stats) { const float f = to_f<T>(h); s1 += f; s2 = fmaf(f, f, s2); }
                    }
                }
            }
        }
        if (p.stats && last >= 0 && row_ok) reinterpret_cast<float2*>(p.stats)[pix * p.stats_nchunk + nt] = make_float2(s1, s2);
        if (++as == p.acc_stages) { as = 0; aphase ^= 1u; }
    }
}

template <class T>
__device__ __forceinline__ void halo_epilogue_dispatch(const HaloParams& p, uint32_t tmem_base, uint32_t tfull0, uint32_t tempty0, int nt, int m0, int m1) {
    switch (p.act) {
        case CSB_ACT_RELU: halo_epilogue<T, CSB_ACT_RELU>(p, tmem_base, tfull0, tempty0, nt, m0, m1); break;
        case CSB_ACT_SILU: halo_epilogue<T, CSB_ACT_SILU>(p, tmem_base, tfull0, tempty0, nt, m0, m1); break;
        case CSB_ACT_PRELU: halo_epilogue<T, CSB_ACT_PRELU>(p, tmem_base, tfull0, tempty0, nt, m0, m1); break;
        default: halo_epilogue<T, CSB_ACT_NONE>(p, tmem_base, tfull0, tempty0, nt, m0, m1); break;
    }
}

// (The issuers' barrier waits are keyed by index parity; the host limits n_iss so that they stay unambiguous, see conv_halo_launch.)
// MMA issue loop.  With N <= 128 an MMA occupies the tensor pipe for only 8..64 cycles, so the ISSUE rate is the critical path of this kernel
// (measured: ~350 cycles per tap for one issuing warp).  Two measures:
//   * the loop is specialised (RESIDENT, NSUB x KSTEPS) and advances descriptor start addresses by additions -- no divisions, no table loads;
//   * when the weights are resident, `n_iss` (up to 3) warps issue for DIFFERENT tiles (tile i -> warp i % n_iss, accumulator stage i % 4): their
//     MMA series are independent (distinct TMEM columns), the shared-memory operands are read-only, and the ring slots of the halos are owned by
//     one tile at a time, so no ordering between the issuers is needed.
// NSUB x KSTEPS MMAs per tap: (1, 4) dense, N = block_n, K = 64; (2, 2) / (4, 1) grouped with 32 / 16-channel diagonal blocks.
// a_stages and acc_stages are powers of two: slot = index & (stages - 1), phase = (index >> log2 stages) & 1.
template <int NSUB, int KSTEPS, bool RESIDENT>
__device__ __forceinline__ void mma_role(const HaloParams& p, uint32_t tmem_base, uint32_t a_base, uint32_t b_base, uint32_t bar_base, int ntiles, int idx, int n_iss) {
    const bool leader = elect_one();
    const uint32_t afull0 = bar_base, aempty0 = bar_base + 8u * kMaxA, bfull0 = bar_base + 8u * (2 * kMaxA), bempty0 = bar_base + 8u * (2 * kMaxA + kMaxB);
    const uint32_t tfull0 = bar_base + 8u * (2 * kMaxA + 2 * kMaxB), tempty0 = tfull0 + 8u * kMaxAcc;
    const uint32_t fmt = p.is_bf16 ? 1u : 0u;
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t) (p.n_mma >> 3) << 17) | ((uint32_t) (128 >> 4) << 24);
    const uint64_t b_layout = p.b_row_bytes == 128 ? 2ull : (p.b_row_bytes == 64 ? 4ull : 6ull);
    // descriptors with a zero start address: the per-MMA start (16 B units, < 2^14) is added into the low word
    const uint64_t a_hi = make_desc_ex(0u, kSbo, 0u, 2ull), b_hi = make_desc_ex(0u, 8u * p.b_row_bytes, 0u, b_layout);
    const uint32_t a_lo0 = (a_base & 0x3ffffu) >> 4, b_lo0 = (b_base & 0x3ffffu) >> 4;
    const uint32_t a_step = p.halo_bytes >> 4, b_step = p.b_tile_bytes >> 4;
    const uint32_t q_rows = ((uint32_t) p.gk * p.b_row_bytes) >> 4;      // start-address step between the diagonal blocks of a weight tile
    const uint32_t col_step = ((uint32_t) p.dil * kRowB) >> 4;           // next tap in the row: dil pixels
    const uint32_t row_step = ((uint32_t) (p.dil * kPitch) * kRowB >> 4) - (uint32_t) (p.S - 1) * col_step;   // first tap of the next filter row
    const int taps = p.taps, S = p.S, kchunks = p.kchunks;
    const int a_mask = p.a_stages - 1, a_log = p.a_stages == 4 ? 2 : 1, t_mask = p.acc_stages - 1, t_log = p.acc_stages == 4 ? 2 : 1;
    int b_slot = 0;
    uint32_t b_ph = 0;
    bool wait_b = true;                                             // resident weights: awaited once, at the issuer's first tile
    for (int i = idx; i < ntiles; i += n_iss) {
        const int as = i & t_mask;
        mbar_wait(tempty0 + 8u * as, (((uint32_t) i >> t_log) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t) (as * p.acc_stride);
        uint32_t first = 0u;                                        // accumulate flag: 0 for the first MMA series of the tile
        for (int kc = 0; kc < kchunks; ++kc) {
            const int g = i * kchunks + kc, a_slot = g & a_mask;
            mbar_wait(afull0 + 8u * a_slot, ((uint32_t) g >> a_log) & 1u);
            tc_fence_after();
            uint64_t ad = a_hi | (uint64_t) (a_lo0 + (uint32_t) a_slot * a_step);
            uint32_t slot = RESIDENT ? (uint32_t) kc : (uint32_t) b_slot;
            uint64_t bd = b_hi | (uint64_t) (b_lo0 + slot * b_step);
            int sx = 0;
            for (int tap = 0; tap < taps; ++tap) {
                if (!RESIDENT || wait_b) { mbar_wait(bfull0 + 8u * slot, RESIDENT ? 0u : b_ph); tc_fence_after(); }
                if (leader) {
#pragma unroll
                    for (int q = 0; q < NSUB; ++q)
#pragma unroll
                        for (int k = 0; k < KSTEPS; ++k)
                            umma_f16<1>(tmem_d + (uint32_t) (q * (64 / NSUB)), ad + (uint64_t) (2 * (q * KSTEPS + k)), bd + (uint64_t) (q * q_rows + 2 * k), idesc,
                                        k == 0 ? first : 1u);
                    if (!RESIDENT) umma_commit<1>(bempty0 + 8u * slot);
                }
                first = 1u;
                if (++sx == S) { sx = 0; ad += row_step; } else ad += col_step;
                if (RESIDENT) { slot += (uint32_t) kchunks; bd += (uint64_t) kchunks * b_step; }
                else if (++b_slot == p.b_slots) { b_slot = 0; b_ph ^= 1u; slot = 0u; bd = b_hi | (uint64_t) b_lo0; }
                else { ++slot; bd += b_step; }
            }
            if (leader) umma_commit<1>(aempty0 + 8u * a_slot);
        }
        wait_b = false;
        if (leader) umma_commit<1>(tfull0 + 8u * as);
        __syncwarp();
    }
}

template <int NSUB, int KSTEPS>
__device__ __forceinline__ void mma_dispatch(const HaloParams& p, uint32_t tmem_base, uint32_t a_base, uint32_t b_base, uint32_t bar_base, int ntiles, int idx, int n_iss) {
    if (p.b_resident) mma_role<NSUB, KSTEPS, true>(p, tmem_base, a_base, b_base, bar_base, ntiles, idx, n_iss);
    else mma_role<NSUB, KSTEPS, false>(p, tmem_base, a_base, b_base, bar_base, ntiles, idx, n_iss);
}

__global__ void __launch_bounds__(128 * (kEpiGroups + 1), 1) k_conv_halo(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                                        const HaloParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_base = smem_base, b_base = smem_base + p.b_off, bar_base = smem_base + p.bar_off;
    auto afull = [&](int s) { return bar_base + 8u * s; };
    auto aempty = [&](int s) { return bar_base + 8u * (kMaxA + s); };
    auto bfull = [&](int s) { return bar_base + 8u * (2 * kMaxA + s); };
    auto bempty = [&](int s) { return bar_base + 8u * (2 * kMaxA + kMaxB + s); };
    auto tfull = [&](int s) { return bar_base + 8u * (2 * kMaxA + 2 * kMaxB + s); };
    auto tempty = [&](int s) { return bar_base + 8u * (2 * kMaxA + 2 * kMaxB + kMaxAcc + s); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * kMaxA + 2 * kMaxB + 2 * kMaxAcc);
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kMaxA; ++s) { mbar_init(afull(s), 1); mbar_init(aempty(s), 1); }
        for (int s = 0; s < kMaxB; ++s) { mbar_init(bfull(s), 1); mbar_init(bempty(s), 1); }
        for (int s = 0; s < kMaxAcc; ++s) { mbar_init(tfull(s), 1); mbar_init(tempty(s), 4 * kEpiGroups); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    // every CTA works on ONE n-tile (grid = tiles_n * ctas_per_n): its weight set is loaded once; the m-tiles of the n-tile are split evenly
    const int ctas_per_n = (int) gridDim.x / p.tiles_n, nt = (int) blockIdx.x % p.tiles_n, part = (int) blockIdx.x / p.tiles_n;
    const int m0 = (int) ((long long) p.tiles_m * part / ctas_per_n), m1 = (int) ((long long) p.tiles_m * (part + 1) / ctas_per_n);
    const int ntiles = m1 - m0;
    const int n_iss = p.n_iss;                                   // issuing warps (1, 2, 3): see mma_role

    if (warp < 4) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        if (warp == 0) {
            // ===================================================== TMA producer (all ring positions are counters with wrap: no divisions)
            if (elect_one()) {
                const int n0 = nt * p.block_n;
                if (p.b_resident)                                    // the whole weight set of this n-tile, once
                    for (int kc = 0; kc < p.kchunks; ++kc)
                        for (int tap = 0; tap < p.taps; ++tap) {
                            const int slot = tap * p.kchunks + kc;
                            mbar_expect_tx(bfull(slot), p.b_tile_bytes);
                            tma_load_2d<1>(b_base + (uint32_t) slot * p.b_tile_bytes, &tmB, bfull(slot), p.grouped ? tap * p.gk : (tap * p.kchunks + kc) * 64, n0);
                        }
                int a_slot = 0, b_slot = 0;
                uint32_t a_ph = 0, b_ph = 0;
                // halo of (tile, kc): one 4-D box; the halo of the NEXT chunk is issued before this chunk's weight tiles
                int at = m0, akc = 0;                                // cursor of the next halo to issue
                auto issue_a = [&]() {
                    const int tw = at % p.tiles_w, th = (at / p.tiles_w) % p.tiles_h, img = at / (p.tiles_w * p.tiles_h);
                    mbar_wait(aempty(a_slot), a_ph ^ 1u);
                    mbar_expect_tx(afull(a_slot), p.halo_bytes);
                    tma_load_4d<1>(a_base + (uint32_t) a_slot * p.halo_bytes, &tmA, afull(a_slot), p.in_coff + (p.grouped ? nt * 64 : akc * 64),
                                   tw * kTileW - p.pad, th * kTileH - p.pad, img);
                    if (++a_slot == p.a_stages) { a_slot = 0; a_ph ^= 1u; }
                    if (++akc == p.kchunks) { akc = 0; ++at; }
                };
                if (at < m1) issue_a();
                for (int mt = m0; mt < m1; ++mt)
                    for (int kc = 0; kc < p.kchunks; ++kc) {
                        if (at < m1) issue_a();
                        if (p.b_resident) continue;
                        for (int tap = 0; tap < p.taps; ++tap) {
                            mbar_wait(bempty(b_slot), b_ph ^ 1u);
                            mbar_expect_tx(bfull(b_slot), p.b_tile_bytes);
                            tma_load_2d<1>(b_base + (uint32_t) b_slot * p.b_tile_bytes, &tmB, bfull(b_slot), p.grouped ? tap * p.gk : (tap * p.kchunks + kc) * 64, n0);
                            if (++b_slot == p.b_slots) { b_slot = 0; b_ph ^= 1u; }
                        }
                    }
            }
        } else if (warp - 1 < n_iss) {
            // ===================================================== MMA issuers
            if (p.nsub == 1) mma_dispatch<1, 4>(p, tmem_base, a_base, b_base, bar_base, ntiles, warp - 1, n_iss);
            else if (p.nsub == 2) mma_dispatch<2, 2>(p, tmem_base, a_base, b_base, bar_base, ntiles, warp - 1, n_iss);
            else mma_dispatch<4, 1>(p, tmem_base, a_base, b_base, bar_base, ntiles, warp - 1, n_iss);
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 152;");
        if (p.is_bf16) halo_epilogue_dispatch<__nv_bfloat16>(p, tmem_base, tfull(0), tempty(0), nt, m0, m1);
        else halo_epilogue_dispatch<__half>(p, tmem_base, tfull(0), tempty(0), nt, m0, m1);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) return (EncodeTiledFn) f;
        return (EncodeTiledFn) nullptr;
    }();
    return fn;
}

// 0 off, 1 halo path (default), 2 halo path with the descriptor base offset set to (start >> 7) & 7 (diagnostic: wrong results on sm_100a)
std::atomic<int> g_halo_mode{[] { const char* e = getenv("CSB_CONV_HALO"); return e ? atoi(e) : 1; }()};

int groups_gk(const csb_conv_desc* d) { const int cpg = d->Cin / d->groups; return cpg < 16 ? 16 : cpg; }

}  // namespace

extern "C" int csb_conv_set_halo_mode(int mode) { return g_halo_mode.exchange(mode); }

// The shapes the halo kernel takes: stride 1, (S - 1) * dil <= 8 (the shifted 8-pixel rows stay inside the 16-pixel halo pitch), whole 64-channel
// chunks; dense layers with Cout <= 256 (one n-tile), grouped layers with Cin == Cout and 16 / 32 / 64-channel diagonal blocks.
extern "C" int csb_conv_halo_supported(const csb_conv_desc* d) {
    if (!d || g_halo_mode.load(std::memory_order_relaxed) == 0) return 0;
    if (d->stride != 1 || d->R < 2 || d->S < 2 || d->R * d->S > kMaxB || (d->S - 1) * d->dil > 8 || (d->R - 1) * d->dil > 40) return 0;
    if (d->Cin % 64 != 0 || d->in_ld % 8 != 0 || d->in_coff % 8 != 0) return 0;
    if (d->groups > 1) {
        if (d->Cin != d->Cout || d->Cin % d->groups != 0) return 0;
        const int cpg = d->Cin / d->groups;
        if (cpg > 64 || 64 % cpg != 0) return 0;
    } else if (d->Cout > 256) return 0;
    return 1;
}

namespace csb {

// Launch of the halo kernel; `w` is [Cout][R][S][Cin] (dense) or [Cout][R][S][gk] (groups > 1).  stats != nullptr: (sum, sum of squares) per
// pixel and 64-channel slice of the rounded outputs (depthwise conv feeding a folded LayerNorm, see csb_conv2d_ln_nhwc).
int conv_halo_launch(const csb_conv_desc* d, const void* x, const void* w, const float* bias, const float* act_param, const void* residual, void* y, float* y_f32,
                     float* stats, void* stream) {
    CSB_REQUIRE(d && x && w && (y || y_f32), "null pointer");
    CSB_REQUIRE(csb_conv_halo_supported(d), "shape not supported by the halo-tile kernel");
    CSB_REQUIRE(((uintptr_t) x & 15) == 0 && ((uintptr_t) w & 15) == 0, "x and w must be 16-byte aligned");
    CSB_REQUIRE(d->res_mode == 0 || residual, "residual pointer missing");
    CSB_REQUIRE(d->act == CSB_ACT_NONE || d->act == CSB_ACT_RELU || d->act == CSB_ACT_SILU || d->act == CSB_ACT_PRELU, "activation not supported by the halo-tile kernel");
    EncodeTiledFn enc = encode_fn();
    if (!enc) return csb::fail(CSB_ERR_CUDA, "%s: %s", "csb_conv2d_halo_nhwc", "cuTensorMapEncodeTiled unavailable");
    const int Hout = d->Hin + 2 * d->pad - d->dil * (d->R - 1), Wout = d->Win + 2 * d->pad - d->dil * (d->S - 1);
    CSB_REQUIRE(Hout > 0 && Wout > 0, "empty output");
    HaloParams p{};
    p.N = d->N; p.H = Hout; p.W = Wout; p.Cout = d->Cout;
    p.R = d->R; p.S = d->S; p.dil = d->dil; p.pad = d->pad; p.taps = d->R * d->S;
    p.grouped = d->groups > 1;
    p.in_coff = d->in_coff;
    p.tiles_h = (Hout + kTileH - 1) / kTileH; p.tiles_w = (Wout + kTileW - 1) / kTileW; p.tiles_m = d->N * p.tiles_h * p.tiles_w;
    if (p.grouped) {
        p.gk = groups_gk(d); p.nsub = 64 / p.gk; p.n_mma = p.gk; p.kq = p.gk;
        p.kchunks = 1; p.block_n = 64; p.tiles_n = d->Cout / 64;
        CSB_REQUIRE(d->Cout % 64 == 0, "grouped conv needs channels in multiples of 64");
        p.b_row_bytes = (uint32_t) p.gk * 2u; p.b_tile_bytes = 64u * p.b_row_bytes;
    } else {
        p.gk = 0; p.nsub = 1; p.kq = 64;
        p.kchunks = d->Cin / 64; p.block_n = ((d->Cout + 15) / 16) * 16; p.n_mma = p.block_n; p.tiles_n = 1;
        p.b_row_bytes = 128u; p.b_tile_bytes = (uint32_t) p.block_n * 128u;
    }
    const int ph = kTileH + (d->R - 1) * d->dil;
    p.halo_bytes = (uint32_t) ph * kPitch * kRowB;
    p.acc_stride = p.block_n <= 32 ? 32 : (p.block_n <= 64 ? 64 : (p.block_n <= 128 ? 128 : 256));
    p.acc_stages = 512 / p.acc_stride > 4 ? 4 : 512 / p.acc_stride;
    const int mode = g_halo_mode.load(std::memory_order_relaxed);
    p.base_off = mode == 2 ? 1 : 0;
    // shared memory: [halo ring | weight tiles | barriers]; weights stay resident when all taps (x chunks) fit next to >= 2 halo stages
    const uint32_t budget = 224u * 1024u, bars = 2048u;
    const uint32_t all_b = (uint32_t) (p.taps * p.kchunks) * p.b_tile_bytes;
    if (p.taps * p.kchunks <= kMaxB && all_b + 2u * p.halo_bytes + bars <= budget) {
        p.b_resident = 1; p.b_slots = p.taps * p.kchunks;
    } else {
        p.b_resident = 0;
        int slots = (int) ((budget - bars - 2u * p.halo_bytes) / p.b_tile_bytes);
        slots = slots > 8 ? 8 : slots;
        CSB_REQUIRE(slots >= 2, "weight tile too large for the halo-tile kernel");
        p.b_slots = slots;
    }
    const uint32_t b_bytes = (uint32_t) p.b_slots * p.b_tile_bytes;
    const int a_stages = (int) ((budget - bars - b_bytes) / p.halo_bytes);
    p.a_stages = a_stages >= 4 ? 4 : 2;                          // a power of two (slot = index & mask in the issue loop)
    CSB_REQUIRE(a_stages >= 2, "halo tile too large");
    // Issuers work on different tiles concurrently and find their ring slots by index parity: tile i may only start once tile i - n_iss is
    // complete, so the halo ring must hold the chunks of n_iss tiles (a_stages >= n_iss * kchunks) and there must be n_iss accumulator stages --
    // otherwise a wait could be satisfied by the slot's PREVIOUS phase (mbarrier parity is ambiguous two phases apart).
    p.n_iss = 1;
    if (p.b_resident) {
        int n = p.a_stages / p.kchunks;
        n = n > 3 ? 3 : n;
        n = n > p.acc_stages ? p.acc_stages : n;
        p.n_iss = n < 1 ? 1 : n;
    }
    p.b_off = (uint32_t) p.a_stages * p.halo_bytes;              // halo_bytes is a multiple of 2 KiB: every region stays 1 KiB aligned
    p.bar_off = p.b_off + ((b_bytes + 1023u) & ~1023u);
    const size_t smem = (size_t) p.bar_off + bars + 1024 /*align*/;

    const cuuint64_t esz = 2;
    const CUtensorMapDataType dt = d->dtype == 1 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    cuuint64_t gdim[4] = {(cuuint64_t) d->in_ld, (cuuint64_t) d->Win, (cuuint64_t) d->Hin, (cuuint64_t) d->N};
    cuuint64_t gstr[3] = {(cuuint64_t) d->in_ld * esz, (cuuint64_t) d->in_ld * esz * d->Win, (cuuint64_t) d->in_ld * esz * d->Win * d->Hin};
    cuuint32_t box[4] = {64, (cuuint32_t) kPitch, (cuuint32_t) ph, 1}, estr[4] = {1, 1, 1, 1};
    CUtensorMap tmA, tmB;
    CUresult r = enc(&tmA, dt, 4, const_cast<void*>(x), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return csb::fail(CSB_ERR_INVALID, "%s: %s", "csb_conv2d_halo_nhwc", "cuTensorMapEncodeTiled(A) failed");
    const cuuint64_t ktot = (cuuint64_t) p.taps * (p.grouped ? p.gk : d->Cin);
    cuuint64_t wdim[2] = {ktot, (cuuint64_t) d->Cout}, wstr[1] = {ktot * esz};
    cuuint32_t wbox[2] = {(cuuint32_t) (p.grouped ? p.gk : 64), (cuuint32_t) (p.grouped ? 64 : p.block_n)}, westr[2] = {1, 1};
    const CUtensorMapSwizzle wsw = p.b_row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (p.b_row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    r = enc(&tmB, dt, 2, const_cast<void*>(w), wdim, wstr, wbox, westr, CU_TENSOR_MAP_INTERLEAVE_NONE, wsw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return csb::fail(CSB_ERR_INVALID, "%s: %s", "csb_conv2d_halo_nhwc", "cuTensorMapEncodeTiled(B) failed");

    p.bias = bias; p.act = d->act; p.act_param = act_param;
    p.residual = residual; p.res_ld = d->res_ld; p.res_coff = d->res_coff; p.res_mode = d->res_mode;
    p.out = y; p.out_f32 = y_f32; p.out_ld = d->out_ld; p.out_coff = d->out_coff; p.is_bf16 = d->dtype == 1;
    p.stats = stats; p.stats_nchunk = d->Cout / 64;
    CSB_REQUIRE(!stats || (p.grouped && d->Cout % 64 == 0 && !y_f32), "statistics need a grouped / depthwise layer with 64-channel slices");

    static unsigned char attr_done[64] = {};
    if (csb::first_use_on_device(attr_done)) cudaFuncSetAttribute(k_conv_halo, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    // one n-tile per CTA: grid = tiles_n * (CTAs per n-tile), at most one CTA per SM
    const int sms = csb::num_sms();
    CSB_REQUIRE(p.tiles_n <= sms, "too many 64-channel slices for the halo-tile kernel");
    int per_n = sms / p.tiles_n;
    per_n = per_n > p.tiles_m ? p.tiles_m : per_n;
    const int grid = p.tiles_n * per_n;
    k_conv_halo<<<grid, 128 * (kEpiGroups + 1), smem, (cudaStream_t) stream>>>(tmA, tmB, p);
    if (csb::g_profiling.load(std::memory_order_relaxed) == 2) {
        char label[160];
        snprintf(label, sizeof label, "k_conv_halo[%dx%dx%dx%d->%d k%dx%d s%d d%d g%d act%d res%d]", d->N, d->Hin, d->Win, d->Cin, d->Cout, d->R, d->S, d->stride, d->dil,
                 d->groups, d->act, d->res_mode);
        return csb::launched(csb::profile_intern(label), (cudaStream_t) stream);
    }
    return csb::launched("k_conv_halo", (cudaStream_t) stream);
}

}  // namespace csb

extern "C" int csb_conv2d_halo_nhwc(const csb_conv_desc* d, const void* x, const void* w, const float* bias, const float* act_param, const void* residual, void* y,
                                    float* y_f32, float* stats, void* stream) {
    return csb::conv_halo_launch(d, x, w, bias, act_param, residual, y, y_f32, stats, stream);
}
