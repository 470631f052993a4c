// Depthwise KxK convolution (K = 5 / 7, stride 1, zero padding K/2) on TMA-staged halo tiles -- the first kernel of every ConvNeXt block
// (dw 7x7) and the dw 5x5 of the CSPNeXt blocks of the neck (row A3).
//
// Round 1's k_dwconv_tile loaded its (2 + K - 1) x (8 + K - 1) input window per 2 x 8 output pixels straight from global memory into registers:
// 7x redundant loads, all of their latency exposed to 12 warps per SM (ncu: issue slots 29 % busy, long_scoreboard the top stall, 0.31 of the
// packed-FFMA peak, 16.4 ms per 32 frames).  Here the memory side is asynchronous:
//   * a CTA walks a contiguous range of (64-channel chunk, 32 x 16 pixel tile) tasks; for every task ONE 4-D TMA box {64 ch, 32 + K - 1, 16 + K - 1, 1}
//     lands in shared memory (zero padding = the TMA unit's out-of-bounds fill: no boundary predicates in the inner loop), double buffered, the box of
//     task t + 1 in flight while task t is computed (one mbarrier per buffer);
//   * 16 warps each own an 8-pixel-wide column strip x 4 rows of the tile and slide the round-1 register micro-kernel (2 x 8 outputs x 2 channels per
//     lane, fp32 packed FFMA2, taps in shared memory) down it, reading inputs with immediate-offset LDS.32 (a warp reads one 128 B pixel: no conflicts);
//   * halo re-reads drop from 7x to 1.63x and come from L2; DRAM traffic is the compulsory 4 C B/px.
// Arithmetic order per output is the one of k_dwconv_tile (bias, then taps row-major), so results are bit-identical to round 1.
// Roofline: fp32 FMA issue (49 or 25 FMA per output element); HBM floor 4 C B/px.
#include <cuda.h>
#include <cuda_fp16.h>

#include <cstdlib>

#include "common.cuh"

namespace {

constexpr int kWarps = 16, kThreads = kWarps * 32;    // 4 warps per scheduler: ncu of the 8-warp build showed the FMA pipe 59 % busy, `wait` / `math_pipe_throttle` the top stalls
constexpr int BW = 32, BH = 16;                    // output pixels of one CTA tile (per 64-channel chunk)

struct DwHaloParams {
    const float* w;          // [K][K][C] fp32
    const float* bias;       // [C] or null
    __half* y;
    float* stats;            // STATS: [pixel][C/64][2] (sum, sum of squares) of the fp16-rounded outputs
    int N, H, W, C, ldy, yoff, xoff, act;
    int tiles_x, tiles_y, chunks;
    long long ntiles;        // N * tiles_y * tiles_x
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
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
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)),
        "l"(*reinterpret_cast<unsigned long long*>(&c)));
    return *reinterpret_cast<float2*>(&d);
}
__device__ __forceinline__ float act_f(float x, int act) {
    switch (act) {
        case CSB_ACT_RELU: return fmaxf(x, 0.0f);
        case CSB_ACT_SILU: return x / (1.0f + __expf(-x));
        case CSB_ACT_GELU: return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f));
        case CSB_ACT_SIGMOID: return 1.0f / (1.0f + __expf(-x));
        default: return x;
    }
}

template <int K>
struct Geo {
    static constexpr int R = K / 2, IW = BW + K - 1, IH = BH + K - 1;
    static constexpr uint32_t kStageBytes = (uint32_t) IW * IH * 128u;
    static constexpr uint32_t kWeightBytes = (uint32_t) K * K * 32u * 8u;
    static constexpr uint32_t kSmem = 2u * kStageBytes + kWeightBytes + 64u + 128u /*alignment slack*/;
};

template <int K, bool STATS>
__global__ void __launch_bounds__(kThreads, 1) k_dwconv_halo(const __grid_constant__ CUtensorMap tmX, const DwHaloParams p) {
    using G = Geo<K>;
    constexpr int TSX = 8, INX = TSX + K - 1, IW = G::IW;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);
    const uint32_t sbase = smem_u32(smem);
    float2* wsm = reinterpret_cast<float2*>(smem + 2u * G::kStageBytes);
    const uint32_t bar0 = sbase + 2u * G::kStageBytes + G::kWeightBytes;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cx = warp & 3, ry = warp >> 2;                       // 4 column strips of 8 px x kWarps/4 row groups
    constexpr int kRowsPerWarp = BH / (kWarps / 4);
    const long long total = p.ntiles * p.chunks;                   // chunk-major task order: a CTA's range touches at most a few chunks
    const long long t0 = total * blockIdx.x / gridDim.x, t1 = total * (blockIdx.x + 1) / gridDim.x;

    if (threadIdx.x == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
        mbar_init(bar0, 1);
        mbar_init(bar0 + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    auto issue = [&](long long t, int s) {                         // one thread: the halo box of task t -> buffer s
        const int chunk = (int) (t / p.ntiles);
        const long long tile = t % p.ntiles;
        const int tx = (int) (tile % p.tiles_x), ty = (int) ((tile / p.tiles_x) % p.tiles_y), img = (int) (tile / ((long long) p.tiles_x * p.tiles_y));
        mbar_expect_tx(bar0 + 8u * s, G::kStageBytes);
        tma_load_4d(sbase + (uint32_t) s * G::kStageBytes, &tmX, bar0 + 8u * s, p.xoff + chunk * 64, tx * BW - G::R, ty * BH - G::R, img);
    };
    if (threadIdx.x == 0 && t0 < t1) issue(t0, 0);

    int cur_chunk = -1;
    float2 b2 = make_float2(0.f, 0.f);
    for (long long t = t0; t < t1; ++t) {
        const int it = (int) (t - t0), s = it & 1;
        // buffer s^1 was read during iteration it-1, which ended with __syncthreads: it is free for the next box
        if (threadIdx.x == 0 && t + 1 < t1) issue(t + 1, s ^ 1);
        const int chunk = (int) (t / p.ntiles);
        if (chunk != cur_chunk) {                                  // (re)load this chunk's K*K x 64 taps: [tap][lane] float2
            cur_chunk = chunk;
            for (int i = threadIdx.x; i < K * K * 32; i += kThreads)
                wsm[i] = __ldg(reinterpret_cast<const float2*>(p.w + (size_t) (i >> 5) * p.C + chunk * 64 + (i & 31) * 2));
            b2 = p.bias ? __ldg(reinterpret_cast<const float2*>(p.bias + chunk * 64 + lane * 2)) : make_float2(0.f, 0.f);
            __syncthreads();
        }
        const long long tile = t % p.ntiles;
        const int tx = (int) (tile % p.tiles_x), ty = (int) ((tile / p.tiles_x) % p.tiles_y);
        const long long img = tile / ((long long) p.tiles_x * p.tiles_y);
        mbar_wait(bar0 + 8u * s, (uint32_t) (it >> 1) & 1u);
        const int ox0 = tx * BW + cx * TSX;
        const int c0 = chunk * 64 + lane * 2;
        // this warp's strip inside the halo tile: pixel (row, col) of the box sits at ((row * IW + col) * 128) bytes, lane l reads bytes [4l, 4l+4)
        const __half2* strip = reinterpret_cast<const __half2*>(smem + (size_t) s * G::kStageBytes) + (size_t) (cx * TSX) * 32 + lane;
        if (ox0 < p.W) {
#pragma unroll 1
            for (int rp = 0; rp < kRowsPerWarp / 2; ++rp) {        // row pairs of this warp
                const int ly = ry * kRowsPerWarp + rp * 2, oy0 = ty * BH + ly;
                if (oy0 >= p.H) break;
                const __half2* win = strip + (size_t) ly * IW * 32;           // input row iy of this row pair = win + iy * IW * 32
                auto load_row = [&](int iy, __half2 (&row)[INX]) {
#pragma unroll
                    for (int ix = 0; ix < INX; ++ix) row[ix] = win[(iy * IW + ix) * 32];
                };
                float2 acc0[TSX], acc1[TSX], xa[INX], xc[INX];
                __half2 nxt[INX];
#pragma unroll
                for (int j = 0; j < TSX; ++j) acc0[j] = acc1[j] = b2;
                load_row(0, nxt);
#pragma unroll
                for (int ix = 0; ix < INX; ++ix) xa[ix] = __half22float2(nxt[ix]);
                load_row(1, nxt);
#pragma unroll
                for (int ix = 0; ix < INX; ++ix) xc[ix] = __half22float2(nxt[ix]);
                load_row(2, nxt);
                auto taps = [&](int r, const float2 (&top)[INX], const float2 (&bot)[INX]) {
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        const float2 wv = wsm[(r * K + k) * 32 + lane];
#pragma unroll
                        for (int tt = 0; tt < TSX; ++tt) {
                            acc0[tt] = ffma2(top[tt + k], wv, acc0[tt]);
                            acc1[tt] = ffma2(bot[tt + k], wv, acc1[tt]);
                        }
                    }
                };
#pragma unroll
                for (int r = 0; r + 1 < K; r += 2) {               // fully unrolled: every LDS has an immediate offset
                    taps(r, xa, xc);
#pragma unroll
                    for (int ix = 0; ix < INX; ++ix) xa[ix] = __half22float2(nxt[ix]);
                    load_row(r + 3, nxt);
                    taps(r + 1, xc, xa);
#pragma unroll
                    for (int ix = 0; ix < INX; ++ix) xc[ix] = __half22float2(nxt[ix]);
                    if (r + 4 <= K) load_row(r + 4, nxt);
                }
                taps(K - 1, xa, xc);
                if (p.act != CSB_ACT_NONE) {
#pragma unroll
                    for (int j = 0; j < TSX; ++j) {
                        acc0[j].x = act_f(acc0[j].x, p.act), acc0[j].y = act_f(acc0[j].y, p.act);
                        acc1[j].x = act_f(acc1[j].x, p.act), acc1[j].y = act_f(acc1[j].y, p.act);
                    }
                }
                __half* yb = p.y + ((img * p.H + oy0) * p.W + ox0) * p.ldy + p.yoff + c0;
                const bool row1 = oy0 + 1 < p.H;
                if constexpr (STATS) {
                    // per output row: 16 values per lane ([0,8): per-pixel sum over this lane's 2 channels, [8,16): sum of squares of the fp16-rounded
                    // outputs) -> transposed butterfly: lane l ends with the warp total of value (l & 15) (16 shuffles instead of 80)
                    const bool full_x = ox0 + TSX <= p.W;              // warp-uniform: interior strips store without per-pixel predicates
                    auto row_stats = [&](const float2 (&acc)[TSX], int oy, bool store, int writer_half) {
                        float v[16];
#pragma unroll
                        for (int j = 0; j < TSX; ++j) {
                            const __half2 h = __floats2half2_rn(acc[j].x, acc[j].y);
                            const float2 f = __half22float2(h);
                            v[j] = f.x + f.y; v[8 + j] = fmaf(f.x, f.x, f.y * f.y);
                            if (store && (full_x || ox0 + j < p.W)) *reinterpret_cast<__half2*>(yb + ((size_t) (oy - oy0) * p.W + j) * p.ldy) = h;
                        }
#pragma unroll
                        for (int sft = 0; sft < 4; ++sft) {
                            const int m = 8 >> sft;
                            const bool up = (lane & m) != 0;
#pragma unroll
                            for (int i = 0; i < m; ++i) {
                                const float send = up ? v[i] : v[i + m], keep = up ? v[i + m] : v[i];
                                v[i] = keep + __shfl_xor_sync(0xffffffffu, send, m);
                            }
                        }
                        v[0] += __shfl_xor_sync(0xffffffffu, v[0], 16);
                        const int px = ox0 + (lane & 7);
                        if (store && (lane >> 4) == writer_half && px < p.W)
                            p.stats[(((img * p.H + oy) * p.W + px) * (p.C >> 6) + chunk) * 2 + ((lane >> 3) & 1)] = v[0];
                    };
                    row_stats(acc0, oy0, true, 0);
                    row_stats(acc1, oy0 + 1, row1, 1);
                } else {
#pragma unroll
                    for (int j = 0; j < TSX; ++j) {
                        if (ox0 + j >= p.W) break;
                        *reinterpret_cast<__half2*>(yb + (size_t) j * p.ldy) = __floats2half2_rn(acc0[j].x, acc0[j].y);
                        if (row1) *reinterpret_cast<__half2*>(yb + ((size_t) p.W + j) * p.ldy) = __floats2half2_rn(acc1[j].x, acc1[j].y);
                    }
                }
            }
        }
        __syncthreads();                                           // everybody is done with buffer s (and with the taps, should the chunk change next)
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

template <int K, bool STATS>
int launch(const CUtensorMap& tm, const DwHaloParams& p, cudaStream_t st) {
    // the attribute belongs to the device that is current when it is set: once per (instantiation, device), outside any stream capture
    static bool attr_done[64] = {};
    static int sm_count[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    dev = dev < 0 || dev >= 64 ? 0 : dev;
    if (!attr_done[dev]) {
        cudaFuncSetAttribute(k_dwconv_halo<K, STATS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) Geo<K>::kSmem);
        cudaDeviceGetAttribute(&sm_count[dev], cudaDevAttrMultiProcessorCount, dev);
        attr_done[dev] = true;
    }
    const long long total = p.ntiles * p.chunks;
    const int sms = sm_count[dev] > 0 ? sm_count[dev] : 148;
    const int grid = (int) (total < sms ? total : sms);
    k_dwconv_halo<K, STATS><<<grid, kThreads, Geo<K>::kSmem, st>>>(tm, p);
    return csb::launched("k_dwconv_halo", st);
}

}  // namespace

// Returns CSB_OK after enqueueing, or a negative "not applicable" marker (-1000) when the shape is outside this kernel's domain so that the caller
// (nn_elem.cu) takes the register-tiled path.
int csb_dwconv_halo_try(const void* x, int ldx, int xoff, const float* w, const float* bias, int act, int N, int H, int W, int C, int K, void* y, int ldy, int yoff,
                        float* stats, cudaStream_t st) {
    const char* env = getenv("CSB_DW_HALO");                         // read per call: tests A/B the two kernels inside one process
    if ((env && atoi(env) == 0) || !(K == 5 || K == 7) || C % 64 != 0 || ldx % 8 != 0 || xoff % 8 != 0 || ldy % 2 != 0 || yoff % 2 != 0 || ((uintptr_t) x & 15) != 0) return -1000;
    if (stats && K != 7) return -1000;
    EncodeTiledFn enc = encode_fn();
    if (!enc) return -1000;
    CUtensorMap tm;
    const cuuint64_t gdim[4] = {(cuuint64_t) ldx, (cuuint64_t) W, (cuuint64_t) H, (cuuint64_t) N};
    const cuuint64_t gstr[3] = {(cuuint64_t) ldx * 2, (cuuint64_t) ldx * 2 * W, (cuuint64_t) ldx * 2 * W * H};
    const cuuint32_t box[4] = {64, (cuuint32_t) (BW + K - 1), (cuuint32_t) (BH + K - 1), 1}, estr[4] = {1, 1, 1, 1};
    if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(x), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return -1000;
    DwHaloParams p{};
    p.w = w; p.bias = bias; p.y = (__half*) y; p.stats = stats;
    p.N = N; p.H = H; p.W = W; p.C = C; p.ldy = ldy; p.yoff = yoff; p.xoff = xoff; p.act = act;
    p.tiles_x = (W + BW - 1) / BW; p.tiles_y = (H + BH - 1) / BH; p.chunks = C / 64;
    p.ntiles = (long long) N * p.tiles_x * p.tiles_y;
    if (K == 7) return stats ? launch<7, true>(tm, p, st) : launch<7, false>(tm, p, st);
    return launch<5, false>(tm, p, st);
}
