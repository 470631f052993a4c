// depth_adjustment_animesseg for sm_100a (reference anime_3dkenburns/kenburns_effect.py:39-91, non-median branch):
// for every instance mask IN ORDER (each step sees the disparity already flattened by the previous instances):
//     plane  = disparity * mask;              if plane.sum() == 0: skip                                   (:57, :68)
//     top, bottom = first / last row with plane.sum(row) > 0                                              (:75, :77)
//     v      = max(plane[rows >= round(top + 0.97 * (bottom - top))])     (Python round: half to even)     (:78)
//     disparity = (1 - mask) * disparity + mask * v                                                       (:78)
// The reference does this with ~8 full-frame ATen kernels and 5 `.item()` host syncs per instance.  Here: three launches per instance, no
// host sync (skip / row range / max stay in a small device state), each touching the mask (1 B/px) and the disparity only under the mask.
//   k_adj_rows   per row: does the row contain plane > 0 ?  -> first/last row via atomicMin/Max, any-flag
//   k_adj_max    max of plane over rows >= cut (float atomicMax on non-negative values via int bits)
//   k_adj_apply  disparity = (1-m)*d + m*v  (exact reference arithmetic incl. its fp32 rounding)
#include "common.cuh"

namespace {

struct AdjState {
    int top, bottom, any;
    unsigned vbits;
};

__global__ void k_adj_init(AdjState* st) {
    if (threadIdx.x == 0 && blockIdx.x == 0) *st = AdjState{0x7fffffff, -1, 0, 0u};
}

// one warp per row
__global__ void __launch_bounds__(256) k_adj_rows(const float* __restrict__ disp, const uint8_t* __restrict__ mask, int H, int W, AdjState* st) {
    const int lane = threadIdx.x & 31;
    for (int y = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); y < H; y += gridDim.x * (blockDim.x >> 5)) {
        // plane.sum(row) > 0  with non-negative terms  <=>  some term > 0 (fp32 sums of non-negatives cannot cancel)
        bool pos = false;
        bool neg = false;
        for (int x = lane; x < W; x += 32) {
            if (mask[(size_t) y * W + x]) {
                const float d = disp[(size_t) y * W + x];
                pos |= d > 0.f;
                neg |= d < 0.f;
            }
        }
        pos = __any_sync(0xffffffffu, pos);
        neg = __any_sync(0xffffffffu, neg);
        if (neg) {   // negative disparities present: fall back to the literal row sum (sequential fp32 order of one thread per row)
            float s = 0.f;
            if (lane == 0) for (int x = 0; x < W; ++x) s += mask[(size_t) y * W + x] ? disp[(size_t) y * W + x] : 0.f;
            pos = __shfl_sync(0xffffffffu, s > 0.f ? 1 : 0, 0);
        }
        if (lane == 0 && pos) {
            atomicMin(&st->top, y);
            atomicMax(&st->bottom, y);
            st->any = 1;
        }
    }
}

__device__ __forceinline__ int py_round_to_int(double v) { return (int) rint(v); }      // Python round(): half to even

__global__ void __launch_bounds__(256) k_adj_max(const float* __restrict__ disp, const uint8_t* __restrict__ mask, int H, int W, AdjState* st) {
    if (!st->any) return;
    const int top = st->top, bottom = st->bottom;
    const int cut = py_round_to_int((double) top + (0.97 * (double) (bottom - top)));
    const long long n = (long long) (H - cut) * W;
    float m = 0.f;      // the reference takes the max over the whole row slice of plane = d * mask: masked-out pixels contribute 0
    for (long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x; i < n; i += (long long) gridDim.x * blockDim.x) {
        const size_t o = (size_t) cut * W + i;
        const float p = mask[o] ? disp[o] : 0.f;
        m = fmaxf(m, p);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(&st->vbits, __float_as_uint(m));
}

__global__ void __launch_bounds__(256) k_adj_apply(float* __restrict__ disp, const uint8_t* __restrict__ mask, int H, int W, AdjState* st) {
    const int any = st->any;
    const float v = __uint_as_float(st->vbits);
    if (any) {
        const long long n = (long long) H * W;
        for (long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x; i < n; i += (long long) gridDim.x * blockDim.x) {
            // ((1.0 - m) * d) + (m * v) with m in {0,1}: exact for both cases, evaluated like the reference
            const float mk = mask[i] ? 1.0f : 0.0f;
            disp[i] = __fadd_rn(__fmul_rn(__fsub_rn(1.0f, mk), disp[i]), __fmul_rn(mk, v));
        }
    }
}

// ---- batched, single launch: G co-resident CTAs per image walk that image's instances in order, separated by per-image software barriers.
// state (per image, per instance slot): tops (init 0x7f7f7f7f), bottoms (init -1), vbits (init 0); bar: one monotonically increasing counter per image.
__device__ __forceinline__ void group_barrier(unsigned* counter, unsigned target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        unsigned v;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
            if (v < target) __nanosleep(64);
        } while (v < target);
    }
    __syncthreads();
}

__global__ void __launch_bounds__(512) k_adj_batch(float* __restrict__ disp, const uint8_t* __restrict__ masks, const int* __restrict__ num, int Kmax, int H, int W,
                                                   int G, int* __restrict__ tops, int* __restrict__ bottoms, int* __restrict__ mtops,
                                                   int* __restrict__ mbottoms, unsigned* __restrict__ vbits, unsigned* __restrict__ bars) {
    const int img = blockIdx.x / G, g = blockIdx.x % G;
    const int K = min(num[img], Kmax);
    float* D = disp + (size_t) img * H * W;
    unsigned* bar = bars + img;
    unsigned target = 0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    for (int k = 0; k < K; ++k) {
        const uint8_t* M = masks + ((size_t) img * Kmax + k) * H * W;
        const int slot = img * Kmax + k;
        // phase A: rows with plane > 0 (top/bottom of the reference) and rows touched by the mask at all (extent of phase C).  The only full
        // pass over the 1 B/px mask; phases B and C are confined to the instance's rows.
        for (int y = g * nwarp + warp; y < H; y += G * nwarp) {
            bool pos = false, any = false;
            for (int x = lane; x < W; x += 32)
                if (M[(size_t) y * W + x]) { any = true; pos |= D[(size_t) y * W + x] > 0.f; }
            pos = __any_sync(0xffffffffu, pos);
            any = __any_sync(0xffffffffu, any);
            if (lane == 0 && pos) { atomicMin(tops + slot, y); atomicMax(bottoms + slot, y); }
            if (lane == 0 && any) { atomicMin(mtops + slot, y); atomicMax(mbottoms + slot, y); }
        }
        target += G;
        group_barrier(bar, target);
        const int top = ((volatile int*) tops)[slot], bottom = ((volatile int*) bottoms)[slot];
        if (bottom < 0) continue;                                   // plane.sum() == 0: skipped by every CTA of the image alike
        const int cut = py_round_to_int((double) top + (0.97 * (double) (bottom - top)));
        // phase B: max of plane over rows >= cut (rows below `bottom` hold plane <= 0 and cannot raise a maximum that starts at 0)
        float m = 0.f;
        const long long n = (long long) (bottom - cut + 1) * W;
        for (long long i = (long long) g * blockDim.x + threadIdx.x; i < n; i += (long long) G * blockDim.x) {
            const size_t o = (size_t) cut * W + i;
            m = fmaxf(m, M[o] ? D[o] : 0.f);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (lane == 0 && m > 0.f) atomicMax(vbits + slot, __float_as_uint(m));
        target += G;
        group_barrier(bar, target);
        const float v = __uint_as_float(((volatile unsigned*) vbits)[slot]);
        // phase C: flatten -- ((1 - m) * d) + (m * v); with m == 0 the value is unchanged, so only the mask's rows are visited
        const int mt = ((volatile int*) mtops)[slot], mb = ((volatile int*) mbottoms)[slot];
        const long long hw = (long long) (mb - mt + 1) * W, base = (long long) mt * W;
        for (long long i = (long long) g * blockDim.x + threadIdx.x; i < hw; i += (long long) G * blockDim.x) {
            const float mk = M[base + i] ? 1.0f : 0.0f;
            D[base + i] = __fadd_rn(__fmul_rn(__fsub_rn(1.0f, mk), D[base + i]), __fmul_rn(mk, v));
        }
        target += G;
        group_barrier(bar, target);
    }
}

}  // namespace

extern "C" int csb_depth_adjust_batch(float* disparity, const uint8_t* masks, const int* num, int N, int Kmax, int H, int W, int32_t* state, void* stream) {
    CSB_REQUIRE(disparity && masks && num && state, "null pointer");
    CSB_REQUIRE(N > 0 && Kmax > 0 && H > 0 && W > 0, "bad shape");
    cudaStream_t st = (cudaStream_t) stream;
    int G = csb::num_sms() / N;
    CSB_REQUIRE(G >= 1, "at most one image per SM (N <= SM count): split the batch");
    G = G > 16 ? 16 : G;
    const size_t slots = (size_t) N * Kmax;
    int* tops = state;                       // [tops | mtops] initialised to 0x7f7f7f7f, [bottoms | mbottoms] to -1, [vbits | bars] to 0
    int* mtops = state + slots;
    int* bottoms = state + 2 * slots;
    int* mbottoms = state + 3 * slots;
    unsigned* vbits = reinterpret_cast<unsigned*>(state + 4 * slots);
    unsigned* bars = reinterpret_cast<unsigned*>(state + 5 * slots);
    CSB_TRY(csb::cuda_ok(cudaMemsetAsync(tops, 0x7f, sizeof(int) * 2 * slots, st), "memset"));
    CSB_TRY(csb::cuda_ok(cudaMemsetAsync(bottoms, 0xff, sizeof(int) * 2 * slots, st), "memset"));
    CSB_TRY(csb::cuda_ok(cudaMemsetAsync(vbits, 0, sizeof(int) * (slots + N), st), "memset"));
    csb::memset_done(st);
    void* args[] = {&disparity, &masks, &num, &Kmax, &H, &W, &G, &tops, &bottoms, &mtops, &mbottoms, &vbits, &bars};
    // cooperative launch: the runtime verifies that all N*G CTAs are co-resident, which the software barriers rely on
    CSB_TRY(csb::cuda_ok(cudaLaunchCooperativeKernel((void*) k_adj_batch, dim3(N * G), dim3(512), args, 0, st), "cudaLaunchCooperativeKernel(k_adj_batch)"));
    return csb::launched("k_adj_batch", st);
}

extern "C" int csb_depth_adjust_instances(float* disparity, const uint8_t* masks, int K, int H, int W, void* state, void* stream) {
    CSB_REQUIRE(disparity && state && (masks || K == 0), "null pointer");
    CSB_REQUIRE(K >= 0 && H > 0 && W > 0, "bad shape");
    cudaStream_t st = (cudaStream_t) stream;
    AdjState* s = reinterpret_cast<AdjState*>(state);
    const int g = csb::wave_grid((long long) H * W, 256, 4);
    for (int k = 0; k < K; ++k) {
        const uint8_t* m = masks + (size_t) k * H * W;
        k_adj_init<<<1, 32, 0, st>>>(s);
        CSB_TRY(csb::launched("k_adj_init", st));
        k_adj_rows<<<csb::wave_grid((long long) H * 32, 256, 4), 256, 0, st>>>(disparity, m, H, W, s);
        CSB_TRY(csb::launched("k_adj_rows", st));
        k_adj_max<<<g, 256, 0, st>>>(disparity, m, H, W, s);
        CSB_TRY(csb::launched("k_adj_max", st));
        k_adj_apply<<<g, 256, 0, st>>>(disparity, m, H, W, s);
        CSB_TRY(csb::launched("k_adj_apply", st));
    }
    return CSB_OK;
}
