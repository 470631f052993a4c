// Multi-head self-attention with an additive relative-position bias on the 5th-generation tensor cores (tcgen05) -- the BEiT-L blocks of the
// ZoeDepth / MiDaS DPT encoder at the resolution the reference's Ken-Burns pipeline uses (img_size [672, 672] -> T = 1765 tokens,
// anime_3dkenburns/kenburns_effect.py:543).  Same contract as k_attention (zoe_attn.cu, the mma.sync kernel kept for
// short sequences and as the A/B reference):
//
//   out[b, q, h*64 + :] = softmax_k( Q[b,h,q,:] . K[b,h,k,:] * scale + bias[h, q, k] ) . V[b,h,k,:]
//
// One CTA = 128 queries of one (batch, head): warps 0-3 do the softmax (thread t <-> query row t <-> TMEM lane t, so it needs no cross-thread
// reduction), warp 4 is the TMA + MMA issuer; the two sides talk through mbarriers only (no CTA barrier inside the key loop -- with
// __syncthreads the softmax warps spent 27 % of their time waiting for the warp that also issued, ncu).
// Per 64-key block:
//   S   = Q K^T        4 x tcgen05.mma M128 N64 K16, operands TMA-staged in 128B-swizzled shared memory (Q once, K double-buffered), S
//                      double-buffered in TMEM: S_{j+2} is issued right behind P V_j
//   P   = exp2(S * scale * log2e + bias * log2e - m_ref)   S is read from TMEM once (64 fp32 per thread), fp32 -> fp16 into a 128B-swizzled
//                      shared tile that is the A operand of the second GEMM.  The 128 x 64 bias tile is TMA-loaded INTO that tile (same shape,
//                      same swizzle, ring of three) and each thread overwrites its bias chunk with its P chunk in place: a thread reading
//                      its own bias row straight from global memory touches 32 different lines per warp instruction, which made the first
//                      version of this kernel L1-tag bound at the speed of the mma.sync kernel (2.19 ms per launch at B = 32, T = 1765)
//   O  += P V          4 x tcgen05.mma M128 N64 K16; V arrives TRANSPOSED ([b, h, d, key], written by k_transpose_v) so that both operands are
//                      K-major; O accumulates in TMEM over all key blocks (online softmax with lazy rescaling, see below)
// Two CTAs per SM (105 KB shared memory, 256 TMEM columns each).  Work per (b, h): 4 T^2 64 FLOP; 0.94 ms per launch at B = 32, T = 1765
// (434 TFLOP/s); the kernel is bound by the softmax's instruction issue (654 warp instructions per block and warp, one MUFU.EX2 per score),
// not by the tensor pipe.
#include <cuda.h>
#include <cuda_fp16.h>

#include <mutex>

#include "common.cuh"

namespace {

constexpr int kD = 64;
constexpr int kBQ = 128;
constexpr int kBK = 64;             // keys per block

struct AttnParams {
    int B, T, heads, Tp, nqb, nkb;
    float scale_log2;            // scale * log2(e)
    const __half* bias;          // [heads][Tp][Tp]
    __half* out;                 // [B][T][heads*64]
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded wait: a broken pipeline traps (the launch reports an error) instead of hanging the GPU.  The try_wait carries a suspend-time hint: the
// hardware parks the warp until the phase completes or the hint elapses, instead of returning after a few tens of nanoseconds -- ncu of the round-1
// loop (profiles/r2_ncu_fc1.md) showed about half of all issued warp instructions of a GEMM launch in these polling loops (ISETP / IMAD / BRA of the
// producer, MMA and epilogue warps), i.e. issue slots and power taken from the epilogue.  CSB_MBAR_HINT_NS=0 compiles the old loop.
#ifndef CSB_MBAR_HINT_NS
#define CSB_MBAR_HINT_NS 4000
#endif
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
#if CSB_MBAR_HINT_NS > 0
    uint32_t ok = 0;
    long long t0 = 0;
    for (uint32_t spin = 1;; ++spin) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(bar), "r"(parity), "r"((uint32_t) CSB_MBAR_HINT_NS) : "memory");
        if (ok) return;
        if ((spin & 255u) == 0) {                                       // every 256 returns without completion (about 1 ms when the hint is honoured)
            if (t0 == 0) t0 = clock64();
            else if (clock64() - t0 > 8000000000ll) __trap();
        }
    }
#else
    uint32_t ok = 0;
    long long t0 = 0;
    for (uint32_t spin = 0;; ++spin) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return;
        if (spin == 64) t0 = clock64();
        if (spin > 64 && (spin & 1023) == 0 && clock64() - t0 > 4000000000ll) __trap();
    }
#endif
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, "
        "%21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
          "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
          "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
          "=r"(v[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, "
        "%22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]),
          "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
          "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
// K-major operand descriptor, 128 B rows, SWIZZLE_128B (as in tc_conv.cu): start address, LBO ignored (1), SBO = 8 rows x 128 B, version 1.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    return (uint64_t) ((saddr & 0x3ffff) >> 4) | (1ull << 16) | ((uint64_t) (1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

// shared memory map (offsets from the 1 KiB-aligned base): K double-buffered (consumed early, by the S GEMM two blocks ahead), V^T and bias / P in a
// ring of three (their loads are issued a whole block before use, once P V of the block that last used the stage has completed)
constexpr uint32_t kOffQ = 0;                      // 128 q x 128 B
constexpr uint32_t kOffK = 16384;                  // 2 stages x (64 keys x 128 B)
constexpr uint32_t kOffV = 32768;                  // 3 stages x (64 d x 128 B of keys)
constexpr uint32_t kOffP = 57344;                  // 3 stages x (128 q x 128 B of keys): bias tile, overwritten in place by P
constexpr uint32_t kOffBar = 106496;               // barriers q, k[2], v[3], b[3], s[2], o, p[3], done + TMEM slot
constexpr uint32_t kSmemBytes = kOffBar + 192 + 1024 /*alignment slack*/;
constexpr int kThreadsAttn = 160;                  // warps 0-3: softmax (thread t <-> query row t <-> TMEM lane t); warp 4: TMA + MMA issuer

__global__ void __launch_bounds__(kThreadsAttn, 2) k_attention_tc(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmK,
                                                         const __grid_constant__ CUtensorMap tmVT, const __grid_constant__ CUtensorMap tmBias, const AttnParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar_q = base + kOffBar, bar_k = bar_q + 8, bar_v = bar_q + 24, bar_b = bar_q + 48, bar_s = bar_q + 72, bar_o = bar_q + 88, bar_p = bar_q + 96, bar_done = bar_q + 120,
                   tmem_slot = bar_q + 128;
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
    const int warp = threadIdx.x >> 5, tid = threadIdx.x;
    // batch fastest: the CTAs that share one (head, query block) slice of the bias run together and hit it in L2
    const int b = blockIdx.x % p.B, qb = (blockIdx.x / p.B) % p.nqb, h = blockIdx.x / (p.B * p.nqb);
    const int q0 = qb * kBQ;
    const int cq = h * kD, ck = (p.heads + h) * kD;           // channel offsets of Q and K inside the fused qkv row
    const int n = p.nkb;

    if (tid == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmQKV) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmK) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmVT) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmBias) : "memory");
        for (int i = 0; i < 16; ++i) mbar_init(bar_q + 8u * i, (i >= 12 && i <= 14) ? 128u : 1u);      // bar_p: one arrival per softmax thread
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;                 // columns [0,64) S0, [64,128) S1, [128,192) O
    const uint32_t lane_addr = (uint32_t) (warp * 32) << 16;

    // instruction descriptor: fp32 accumulate, fp16 A/B, both K-major, N = 64 (>>3 @17), M = 128 (>>4 @24) -- the same for both GEMMs
    const uint32_t idesc = (1u << 4) | ((uint32_t) (kBK >> 3) << 17) | ((uint32_t) (kBQ >> 4) << 24);

    auto load_k = [&](int j) {          // K rows [j*64, +64) of (b, h) -> stage j & 1 (rows >= T are zero-filled by the TMA unit)
        const uint32_t bar = bar_k + 8u * (j & 1);
        mbar_expect_tx(bar, kBK * 128);
        tma_load_3d(base + kOffK + (uint32_t) (j & 1) * 8192u, &tmK, bar, ck, j * kBK, b);
    };
    auto load_v = [&](int j) {          // V^T [64 d][keys j*64 .. +64) -> stage j % 3
        const uint32_t bar = bar_v + 8u * (uint32_t) (j % 3);
        mbar_expect_tx(bar, kD * 128);
        tma_load_3d(base + kOffV + (uint32_t) (j % 3) * 8192u, &tmVT, bar, j * kBK, 0, b * p.heads + h);
    };
    auto load_bias = [&](int j) {       // bias[h][q0 .. +128)[j*64 .. +64) INTO P stage j % 3
        const uint32_t bar = bar_b + 8u * (uint32_t) (j % 3);
        mbar_expect_tx(bar, kBQ * 128);
        tma_load_3d(base + kOffP + (uint32_t) (j % 3) * 16384u, &tmBias, bar, j * kBK, q0, h);
    };
    auto issue_s = [&](int j) {         // S_j = Q K_j^T -> S[j & 1]
        const uint64_t adesc = make_desc(base + kOffQ), bdesc = make_desc(base + kOffK + (uint32_t) (j & 1) * 8192u);
#pragma unroll
        for (int k = 0; k < kD / 16; ++k) umma_f16(tmem_base + (uint32_t) (j & 1) * 64u, adesc + 2u * k, bdesc + 2u * k, idesc, k != 0);
        umma_commit(bar_s + 8u * (j & 1));
    };

    if (warp == 4) {
        // ===================================================== TMA + MMA issuer (one lane): the softmax warps never wait for each other or for a GEMM
        if (threadIdx.x == 128) {
            mbar_expect_tx(bar_q, kBQ * 128);
            tma_load_3d(base + kOffQ, &tmQKV, bar_q, cq, q0, b);
            load_k(0);
            load_bias(0);
            if (n > 1) load_k(1);
            load_v(0);
            if (n > 1) { load_bias(1); load_v(1); }
            mbar_wait(bar_q, 0);
            mbar_wait(bar_k, 0);
            tc_fence_after();
            issue_s(0);
            if (n > 1) {
                mbar_wait(bar_k + 8, 0);
                tc_fence_after();
                issue_s(1);
            }
            for (int j = 0; j < n; ++j) {
                const int st = j & 1, s3 = j % 3;
                const uint32_t ph = (uint32_t) ((j >> 1) & 1), ph3 = (uint32_t) ((j / 3) & 1);
                if (j + 2 < n) {                                                            // S_j complete -> K stage st is free: prefetch K_{j+2}
                    mbar_wait(bar_s + 8u * st, ph);
                    load_k(j + 2);
                }
                if (j > 0) mbar_wait(bar_o, (uint32_t) ((j - 1) & 1));                      // P V_{j-1} done: V / P stage (j + 2) % 3 is free
                if (j + 2 < n) { load_bias(j + 2); load_v(j + 2); }
                mbar_wait(bar_p + 8u * (uint32_t) s3, ph3);                                 // P_j written by all 128 rows, S_j fully read
                mbar_wait(bar_v + 8u * (uint32_t) s3, ph3);
                tc_fence_after();
                const uint64_t adesc = make_desc(base + kOffP + (uint32_t) s3 * 16384u), bdesc = make_desc(base + kOffV + (uint32_t) s3 * 8192u);
#pragma unroll
                for (int k = 0; k < kBK / 16; ++k) umma_f16(tmem_base + 128u, adesc + 2u * k, bdesc + 2u * k, idesc, (j | k) != 0);      // O += P V_j
                umma_commit(bar_o);
                if (j + 2 < n) {                                                            // S GEMM of block j + 2 into the S buffer just released
                    mbar_wait(bar_k + 8u * st, ph ^ 1u);
                    tc_fence_after();
                    issue_s(j + 2);
                }
            }
            mbar_wait(bar_o, (uint32_t) ((n - 1) & 1));                                     // the last P V: O is final
            mbar_arrive(bar_done);
        }
    } else {
    // Online softmax with LAZY rescaling: O accumulates in TMEM across the key blocks (tcgen05.mma accumulate) relative to a reference maximum
    // m_ref per row that is only moved -- and O / l only rescaled, by the warp that owns the rows -- when some row's maximum has grown by more
    // than 2^8 since (P then stays <= 256, exact enough in fp16 and fp32).  The per-block read-modify of O through registers that the textbook
    // form needs is gone: TMEM reads (~64 B/clk per SM) are the scarce resource of this kernel.
    float m_ref = -INFINITY, l_run = 0.f;
    const int qrow = q0 + tid;                                                            // < Tp: the bias rows are padded to a multiple of 128
    const uint32_t psw = (uint32_t) (tid & 7);                                             // SWIZZLE_128B: 16 B chunk ^= row & 7
    const uint32_t tmem_o = tmem_base + 128u + lane_addr;
    constexpr float kLog2e = 1.4426950408889634f;

    for (int j = 0; j < n; ++j) {
        const int st = j & 1, s3 = j % 3;                                                  // S / K stage, V / bias / P stage
        const uint32_t ph = (uint32_t) ((j >> 1) & 1), ph3 = (uint32_t) ((j / 3) & 1);
        // ---------------- 1. softmax of block j (S_j in TMEM, bias tile in shared memory); nothing of the previous block is waited for
        mbar_wait(bar_b + 8u * (uint32_t) s3, ph3);
        mbar_wait(bar_s + 8u * st, ph);
        tc_fence_after();
        const uint32_t tmem_s = tmem_base + (uint32_t) st * 64u + lane_addr;
        const uint32_t prow = base + kOffP + (uint32_t) s3 * 16384u + (uint32_t) tid * 128u;
        // S_j is read from TMEM ONCE (64 fp32 per thread stay in registers between the maximum and the exponentials)
        float t[kBK];
        float mxp[8];                                                                       // 8 independent maxima: no 64-deep FMNMX chain
#pragma unroll
        for (int i = 0; i < 8; ++i) mxp[i] = -INFINITY;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            uint32_t s[32];
            tmem_ld32(tmem_s + (uint32_t) c * 32u, s);
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const uint4 u = ld_shared_v4(prow + ((((uint32_t) (c * 4 + g)) ^ psw) << 4));
                const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float2 bf = __half22float2(*reinterpret_cast<const __half2*>(&w[e]));
                    const float t0 = fmaf(__uint_as_float(s[8 * g + 2 * e]), p.scale_log2, bf.x * kLog2e);
                    const float t1 = fmaf(__uint_as_float(s[8 * g + 2 * e + 1]), p.scale_log2, bf.y * kLog2e);
                    t[c * 32 + 8 * g + 2 * e] = t0;
                    t[c * 32 + 8 * g + 2 * e + 1] = t1;
                    mxp[2 * e] = fmaxf(mxp[2 * e], t0);
                    mxp[2 * e + 1] = fmaxf(mxp[2 * e + 1], t1);
                }
            }
        }
        const float mx = fmaxf(fmaxf(fmaxf(mxp[0], mxp[1]), fmaxf(mxp[2], mxp[3])), fmaxf(fmaxf(mxp[4], mxp[5]), fmaxf(mxp[6], mxp[7])));
        if (__any_sync(0xffffffffu, mx - m_ref > 8.0f)) {                                   // warp-uniform; always true for the first block
            const float m_new = fmaxf(m_ref, mx);
            const float alpha = ex2(m_ref - m_new);                                        // first block: exp2(-inf) = 0
            m_ref = m_new;
            l_run *= alpha;
            if (j > 0) {                                                                    // O rows of this warp *= alpha, once P V_{j-1} is complete
                // (no parity aliasing: S_j is complete, hence P V_{j-2} is; bar_o has seen j - 1 or j completions)
                mbar_wait(bar_o, (uint32_t) ((j - 1) & 1));
                tc_fence_after();
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    uint32_t v[32];
                    tmem_ld32(tmem_o + (uint32_t) c * 32u, v);
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
                    tmem_st32(tmem_o + (uint32_t) c * 32u, v);
                }
            }
        }
        // ---------------- P = exp2(t - m_ref) -> fp16 over the bias chunk it came from (A operand of P V), row sum
        float sump[4] = {0.f, 0.f, 0.f, 0.f};                                               // independent partial sums
#pragma unroll
        for (int c8 = 0; c8 < 8; ++c8) {
            uint32_t pk[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float p0 = ex2(t[8 * c8 + 2 * e] - m_ref), p1 = ex2(t[8 * c8 + 2 * e + 1] - m_ref);
                sump[e] += p0 + p1;
                pk[e] = pack_h2(p0, p1);
            }
            st_shared_v4(prow + ((((uint32_t) c8) ^ psw) << 4), pk[0], pk[1], pk[2], pk[3]);
        }
        l_run += (sump[0] + sump[1]) + (sump[2] + sump[3]);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");                        // P (generic proxy) -> tcgen05.mma (async proxy)
        tc_fence_before();
        mbar_arrive(bar_p + 8u * (uint32_t) s3);                                            // this row of P is written, this row of S_j is read
    }
    // ---------------- normalise and store this query row (128 B).  (A parity wait can only tell the current phase from the one before it, and the
    // softmax threads do not follow bar_o phase by phase: the issuer signals the final P V on its own barrier.)
    mbar_wait(bar_done, 0);
    tc_fence_after();
    {
        const float inv = 1.0f / l_run;
        __half* orow = p.out + ((size_t) b * p.T + (qrow < p.T ? qrow : 0)) * (p.heads * kD) + h * kD;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            uint32_t v[32];
            tmem_ld32(tmem_o + (uint32_t) c * 32u, v);
            if (qrow < p.T) {
                uint4* dst = reinterpret_cast<uint4*>(orow + c * 32);
#pragma unroll
                for (int g = 0; g < 4; ++g)
                    dst[g] = make_uint4(pack_h2(__uint_as_float(v[8 * g]) * inv, __uint_as_float(v[8 * g + 1]) * inv),
                                        pack_h2(__uint_as_float(v[8 * g + 2]) * inv, __uint_as_float(v[8 * g + 3]) * inv),
                                        pack_h2(__uint_as_float(v[8 * g + 4]) * inv, __uint_as_float(v[8 * g + 5]) * inv),
                                        pack_h2(__uint_as_float(v[8 * g + 6]) * inv, __uint_as_float(v[8 * g + 7]) * inv));
            }
        }
    }
    }   // softmax warps
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
    }
}

// V of the fused qkv tensor -> vt [B][heads][64][Tkp] (keys contiguous, zero for keys >= T): the K-major B operand of O = P V.
__global__ void __launch_bounds__(256) k_transpose_v(const __half* __restrict__ qkv, int T, int heads, int Tkp, __half* __restrict__ vt) {
    __shared__ __half tile[64][kD + 2];
    const int t0 = blockIdx.x * 64, h = blockIdx.y, b = blockIdx.z;
    const size_t ld = (size_t) 3 * heads * kD;
    const __half* src = qkv + (size_t) b * T * ld + (size_t) (2 * heads + h) * kD;
    for (int i = threadIdx.x; i < 64 * 32; i += blockDim.x) {                               // 64 keys x 32 half2
        const int r = i >> 5, c2 = i & 31;
        const __half2 v = t0 + r < T ? *reinterpret_cast<const __half2*>(src + (size_t) (t0 + r) * ld + 2 * c2) : __floats2half2_rn(0.f, 0.f);
        tile[r][2 * c2] = __low2half(v);
        tile[r][2 * c2 + 1] = __high2half(v);
    }
    __syncthreads();
    __half* dst = vt + ((size_t) (b * heads + h) * kD) * Tkp + t0;
    for (int i = threadIdx.x; i < 64 * 32; i += blockDim.x) {                               // 64 d x 32 half2 of keys
        const int d = i >> 5, k2 = i & 31;
        *reinterpret_cast<__half2*>(dst + (size_t) d * Tkp + 2 * k2) = __halves2half2(tile[2 * k2][d], tile[2 * k2 + 1][d]);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = (EncodeTiledFn) f;
    });
    return fn;
}

}  // namespace

extern "C" long long csb_attention_tc_scratch_bytes(int B, int T, int heads) {
    const long long Tkp = (T + kBK - 1) / kBK * kBK;
    return (long long) B * heads * kD * Tkp * 2;
}

extern "C" int csb_attention_bias_tc(const void* qkv, int B, int T, int heads, int head_dim, const void* bias, int Tp, float scale, void* vt_scratch, void* out,
                                     void* stream) {
    CSB_REQUIRE(qkv && bias && out && vt_scratch, "null pointer");
    CSB_REQUIRE(B > 0 && T > 0 && heads > 0 && head_dim == kD, "head_dim must be 64");
    CSB_REQUIRE(Tp % kBK == 0 && Tp >= T, "the bias must be padded to a multiple of 64 keys (and as many rows)");
    CSB_REQUIRE((((uintptr_t) qkv | (uintptr_t) out | (uintptr_t) bias | (uintptr_t) vt_scratch) & 15) == 0, "pointers must be 16-byte aligned");
    EncodeTiledFn enc = encode_fn();
    if (!enc) return csb::fail(CSB_ERR_CUDA, "%s: %s", "csb_attention_bias_tc", "cuTensorMapEncodeTiled unavailable");
    cudaStream_t st = (cudaStream_t) stream;
    const int Tkp = (T + kBK - 1) / kBK * kBK;
    k_transpose_v<<<dim3(Tkp / 64, heads, B), 256, 0, st>>>((const __half*) qkv, T, heads, Tkp, (__half*) vt_scratch);
    CSB_TRY(csb::launched("k_transpose_v", st));

    CUtensorMap tmQKV, tmK, tmVT;
    const cuuint64_t ld = (cuuint64_t) 3 * heads * kD;
    for (int which = 0; which < 2; ++which) {   // qkv [B][T][3*heads*64]: box {64 channels, 128 tokens (Q) or 64 tokens (K), 1}
        cuuint64_t gdim[3] = {ld, (cuuint64_t) T, (cuuint64_t) B}, gstr[2] = {ld * 2, ld * 2 * (cuuint64_t) T};
        cuuint32_t box[3] = {(cuuint32_t) kD, (cuuint32_t) (which ? kBK : kBQ), 1}, estr[3] = {1, 1, 1};
        if (enc(which ? &tmK : &tmQKV, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(qkv), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return csb::fail(CSB_ERR_INVALID, "%s: %s", "csb_attention_bias_tc", "cuTensorMapEncodeTiled(qkv) failed");
    }
    {   // vt [B*heads][64][Tkp]: box {64 keys, 64 d, 1}
        cuuint64_t gdim[3] = {(cuuint64_t) Tkp, (cuuint64_t) kD, (cuuint64_t) B * heads}, gstr[2] = {(cuuint64_t) Tkp * 2, (cuuint64_t) Tkp * 2 * kD};
        cuuint32_t box[3] = {64, (cuuint32_t) kD, 1}, estr[3] = {1, 1, 1};
        if (enc(&tmVT, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, vt_scratch, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return csb::fail(CSB_ERR_INVALID, "%s: %s", "csb_attention_bias_tc", "cuTensorMapEncodeTiled(vt) failed");
    }
    CUtensorMap tmBias;
    {   // bias [heads][Tp][Tp]: box {64 keys, 128 query rows, 1}
        cuuint64_t gdim[3] = {(cuuint64_t) Tp, (cuuint64_t) Tp, (cuuint64_t) heads}, gstr[2] = {(cuuint64_t) Tp * 2, (cuuint64_t) Tp * 2 * (cuuint64_t) Tp};
        cuuint32_t box[3] = {64, (cuuint32_t) kBQ, 1}, estr[3] = {1, 1, 1};
        if (enc(&tmBias, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(bias), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return csb::fail(CSB_ERR_INVALID, "%s: %s", "csb_attention_bias_tc", "cuTensorMapEncodeTiled(bias) failed");
    }
    AttnParams p;
    p.B = B; p.T = T; p.heads = heads; p.Tp = Tp; p.nqb = (T + kBQ - 1) / kBQ; p.nkb = Tkp / kBK;
    p.scale_log2 = scale * 1.4426950408889634f;
    p.bias = (const __half*) bias; p.out = (__half*) out;
    static unsigned char attr_done[64] = {};
    if (csb::first_use_on_device(attr_done)) cudaFuncSetAttribute(k_attention_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) kSmemBytes);
    const long long grid = (long long) B * p.nqb * heads;
    k_attention_tc<<<(unsigned) grid, kThreadsAttn, kSmemBytes, st>>>(tmQKV, tmK, tmVT, tmBias, p);
    return csb::launched("k_attention_tc", st);
}
