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
                                                   int* __restrict__ mbottoms, unsigned* __restrict__ vbits, unsigned* __restrict__ bars,
                                                   const int* __restrict__ need) {
    if (need && *need == 0) return;                                 // the order-free path (below) already produced the result
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

// ---- order-free formulation (the common case: every disparity under a mask is > 0, which the depth estimators guarantee) -------------------
// With positive disparities the sequential recurrence has a closed dependency structure:
//   * top_k / bottom_k / cut_k depend on the mask only (plane.sum(row) > 0  <=>  the row intersects the mask);
//   * after the loop a pixel holds v_last, `last` = the highest instance index covering it;
//   * v_k = max over the pixels q of mask_k in rows >= cut_k of the value q holds just before step k, i.e. of v_pred(q,k) (pred = highest covering
//     index below k) or of the original disparity if there is none (masked-out pixels contribute plane = 0 < v_k).
// So: one pass packs the K masks into a K-bit cover word per pixel and finds top/bottom; one pass derives base_k = max of original disparities
// and the relation R[k] = {pred(q,k)}; K scalar steps resolve v_k = max(base_k, max_{j in R[k]} v_j); one pass writes v_last.  Every step is a
// max or a select -- bit-identical to the reference's arithmetic ((1-m)*d + m*v is exactly v or d for finite d).  The first pass also raises a
// flag if any masked disparity is not > 0; then nothing is written and the sequential kernel above runs instead.
constexpr int kParMaxK = 128;

__global__ void __launch_bounds__(256) k_adjp_cover(const float* __restrict__ disp, const uint8_t* __restrict__ masks, const int* __restrict__ num, int Kmax, int H,
                                                    int W, int Kw, int* __restrict__ tops, int* __restrict__ bottoms, unsigned* __restrict__ cover,
                                                    int* __restrict__ need) {
    __shared__ int s_top[kParMaxK], s_bot[kParMaxK];
    __shared__ int s_fail;
    const int img = blockIdx.y, K = min(num[img], Kmax);
    for (int i = threadIdx.x; i < kParMaxK; i += blockDim.x) { s_top[i] = 0x7fffffff; s_bot[i] = -1; }
    if (threadIdx.x == 0) s_fail = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int segs = (W + 127) / 128;
    const long long HW = (long long) H * W;
    const float* D = disp + (long long) img * HW;
    const uint8_t* M = masks + (long long) img * Kmax * HW;
    unsigned* C = cover + (long long) img * HW * Kw;                      // word-planar: C[w * HW + pixel]
    bool fail = false;
    for (long long sgm = (long long) blockIdx.x * nwarp + warp; sgm < (long long) H * segs; sgm += (long long) gridDim.x * nwarp) {
        const int y = (int) (sgm / segs), x = (int) (sgm % segs) * 128 + lane * 4;
        const bool valid = x < W;                                      // W % 4 == 0
        const long long p = (long long) y * W + x;
        unsigned c[4][4];
#pragma unroll
        for (int j = 0; j < 4; ++j) c[j][0] = c[j][1] = c[j][2] = c[j][3] = 0u;
        for (int k = 0; k < K; ++k) {
            uchar4 m = make_uchar4(0, 0, 0, 0);
            if (valid) m = *reinterpret_cast<const uchar4*>(M + (long long) k * HW + p);
            const unsigned bit = 1u << (k & 31);
            const int wi = k >> 5;
#pragma unroll
            for (int w = 0; w < 4; ++w) {                                  // static register indexing
                if (w == wi) {
                    if (m.x) c[0][w] |= bit;
                    if (m.y) c[1][w] |= bit;
                    if (m.z) c[2][w] |= bit;
                    if (m.w) c[3][w] |= bit;
                }
            }
            const unsigned b = __ballot_sync(0xffffffffu, (m.x | m.y | m.z | m.w) != 0);
            if (lane == 0 && b) {
                if (y < s_top[k]) atomicMin(&s_top[k], y);
                if (y > s_bot[k]) atomicMax(&s_bot[k], y);
            }
        }
        if (valid) {
            const float4 d = *reinterpret_cast<const float4*>(D + p);
            const float dv[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const unsigned any = c[j][0] | c[j][1] | c[j][2] | c[j][3];
                if (any && !(dv[j] > 0.f)) fail = true;
            }
#pragma unroll
            for (int w = 0; w < 4; ++w)
                if (w < Kw) *reinterpret_cast<uint4*>(C + (long long) w * HW + p) = make_uint4(c[0][w], c[1][w], c[2][w], c[3][w]);
        }
    }
    if (fail) s_fail = 1;
    __syncthreads();
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        if (s_bot[k] >= 0) {
            atomicMin(tops + img * Kmax + k, s_top[k]);
            atomicMax(bottoms + img * Kmax + k, s_bot[k]);
        }
    }
    if (threadIdx.x == 0 && s_fail) atomicOr(need, 1);
}

// W % 16 == 0: 16 pixels per thread.  The version above is latency-bound (one 4 B mask load in flight per thread: 0.8 TB/s of the 1 B/px x K mask
// stream); here every thread issues eight independent 16 B loads (eight instances) before it touches them, and the byte -> bit transposition is
// word-parallel: masks are 0/1 bytes, so `acc |= m << i` (i < 8) gathers eight instances of four pixels in one instruction without crossing bytes.
__global__ void __launch_bounds__(256) k_adjp_cover16(const float* __restrict__ disp, const uint8_t* __restrict__ masks, const int* __restrict__ num, int Kmax, int H,
                                                      int W, int Kw, int* __restrict__ tops, int* __restrict__ bottoms, unsigned* __restrict__ cover,
                                                      int* __restrict__ need) {
    __shared__ int s_top[kParMaxK], s_bot[kParMaxK];
    __shared__ int s_fail;
    const int img = blockIdx.y, K = min(num[img], Kmax);
    for (int i = threadIdx.x; i < kParMaxK; i += blockDim.x) { s_top[i] = 0x7fffffff; s_bot[i] = -1; }
    if (threadIdx.x == 0) s_fail = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int segs = (W + 511) / 512;
    const long long HW = (long long) H * W;
    const float* D = disp + (long long) img * HW;
    const uint8_t* M = masks + (long long) img * Kmax * HW;
    unsigned* C = cover + (long long) img * HW * Kw;
    bool fail = false;
    for (long long sgm = (long long) blockIdx.x * nwarp + warp; sgm < (long long) H * segs; sgm += (long long) gridDim.x * nwarp) {
        const int y = (int) (sgm / segs), x = (int) (sgm % segs) * 512 + lane * 16;
        const bool valid = x < W;                                      // W % 16 == 0
        const long long p = (long long) y * W + x;
        unsigned covered = 0u;                                         // bit j: pixel j is under some mask
#pragma unroll
        for (int wi = 0; wi < 4; ++wi) {
            if (wi * 32 >= K) break;                                   // warp-uniform
            unsigned cw[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) cw[j] = 0u;
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const int k0 = wi * 32 + g * 8;
                if (k0 >= K) break;                                    // warp-uniform
                uint4 m[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    m[i] = make_uint4(0u, 0u, 0u, 0u);
                    if (valid && k0 + i < K) m[i] = __ldg(reinterpret_cast<const uint4*>(M + (long long) (k0 + i) * HW + p));
                }
                unsigned a[4] = {0u, 0u, 0u, 0u};
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const unsigned b = __ballot_sync(0xffffffffu, (m[i].x | m[i].y | m[i].z | m[i].w) != 0u);
                    if (lane == 0 && b) {
                        if (y < s_top[k0 + i]) atomicMin(&s_top[k0 + i], y);
                        if (y > s_bot[k0 + i]) atomicMax(&s_bot[k0 + i], y);
                    }
                    a[0] |= m[i].x << i; a[1] |= m[i].y << i; a[2] |= m[i].z << i; a[3] |= m[i].w << i;
                }
#pragma unroll
                for (int q = 0; q < 4; ++q)
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) cw[4 * q + jj] |= ((a[q] >> (8 * jj)) & 0xffu) << (8 * g);
            }
            if (valid) {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    *reinterpret_cast<uint4*>(C + (long long) wi * HW + p + 4 * q) = make_uint4(cw[4 * q], cw[4 * q + 1], cw[4 * q + 2], cw[4 * q + 3]);
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) covered |= (cw[j] != 0u ? 1u : 0u) << j;
        }
        if (valid) {
            for (int wi = (K + 31) / 32; wi < Kw; ++wi)                // words beyond this image's instance count
#pragma unroll
                for (int q = 0; q < 4; ++q) *reinterpret_cast<uint4*>(C + (long long) wi * HW + p + 4 * q) = make_uint4(0u, 0u, 0u, 0u);
            if (covered) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 d = *reinterpret_cast<const float4*>(D + p + 4 * q);
                    if (((covered >> (4 * q)) & 1u) && !(d.x > 0.f)) fail = true;
                    if (((covered >> (4 * q + 1)) & 1u) && !(d.y > 0.f)) fail = true;
                    if (((covered >> (4 * q + 2)) & 1u) && !(d.z > 0.f)) fail = true;
                    if (((covered >> (4 * q + 3)) & 1u) && !(d.w > 0.f)) fail = true;
                }
            }
        }
    }
    if (fail) s_fail = 1;
    __syncthreads();
    for (int k = threadIdx.x; k < K; k += blockDim.x) {
        if (s_bot[k] >= 0) {
            atomicMin(tops + img * Kmax + k, s_top[k]);
            atomicMax(bottoms + img * Kmax + k, s_bot[k]);
        }
    }
    if (threadIdx.x == 0 && s_fail) atomicOr(need, 1);
}


// One CTA walks whole rows.  Only instances whose bottom band (rows >= cut_k) contains the row matter, so the row's K-bit band mask is built first and
// ANDed with the cover word: the vast majority of pixels leave after four coalesced loads (the previous version walked every set bit of every
// pixel).  pred(q, k) = the highest covering index below k, from the FULL cover word.
__global__ void __launch_bounds__(256) k_adjp_rel(const float* __restrict__ disp, const int* __restrict__ num, int Kmax, int H, int W, int Kw,
                                                  const int* __restrict__ tops, const int* __restrict__ bottoms, const unsigned* __restrict__ cover,
                                                  unsigned* __restrict__ base, unsigned* __restrict__ rel, const int* __restrict__ need) {
    if (*need) return;
    __shared__ int s_cut[kParMaxK];
    __shared__ unsigned s_base[kParMaxK];
    __shared__ unsigned s_rel[kParMaxK * 4];
    __shared__ unsigned s_band[4];
    const int img = blockIdx.y, K = min(num[img], Kmax);
    for (int k = threadIdx.x; k < kParMaxK; k += blockDim.x) {
        int cut = 0x7fffffff;
        if (k < K) {
            const int top = tops[img * Kmax + k], bottom = bottoms[img * Kmax + k];
            if (bottom >= 0) cut = py_round_to_int((double) top + (0.97 * (double) (bottom - top)));
        }
        s_cut[k] = cut;
        s_base[k] = 0u;
    }
    for (int i = threadIdx.x; i < kParMaxK * 4; i += blockDim.x) s_rel[i] = 0u;
    __syncthreads();
    const long long HW = (long long) H * W;
    const float* D = disp + (long long) img * HW;
    const unsigned* C = cover + (long long) img * HW * Kw;                // word-planar
    for (int y = blockIdx.x; y < H; y += gridDim.x) {
        if (threadIdx.x < kParMaxK) {                                     // warps 0..3: band bit k = (y >= cut_k)
            const unsigned b = __ballot_sync(0xffffffffu, y >= s_cut[threadIdx.x]);
            if ((threadIdx.x & 31) == 0) s_band[threadIdx.x >> 5] = b;
        }
        __syncthreads();
        const unsigned band[4] = {s_band[0], s_band[1], s_band[2], s_band[3]};
        if (band[0] | band[1] | band[2] | band[3]) {
            for (int x = threadIdx.x; x < W; x += blockDim.x) {
                const long long p = (long long) y * W + x;
                unsigned cw[4], lowprev[4];
                int lp = -1;
#pragma unroll
                for (int wi = 0; wi < 4; ++wi) {
                    cw[wi] = wi < Kw ? C[(long long) wi * HW + p] : 0u;
                    lowprev[wi] = (unsigned) lp;                          // highest covering index in the words below wi (or -1)
                    if (cw[wi]) lp = wi * 32 + 31 - __clz(cw[wi]);
                }
#pragma unroll
                for (int wi = 0; wi < 4; ++wi) {
                    unsigned r = cw[wi] & band[wi];
                    while (r) {
                        const int b = __ffs(r) - 1;
                        r &= r - 1;
                        const int k = wi * 32 + b;
                        const unsigned lower = cw[wi] & ((1u << b) - 1u);
                        const int prev = lower ? wi * 32 + 31 - __clz(lower) : (int) lowprev[wi];
                        if (prev < 0) {
                            const unsigned u = __float_as_uint(D[p]);             // > 0: bit patterns order like the values
                            if (u > s_base[k]) atomicMax(&s_base[k], u);
                        } else {
                            const unsigned bit = 1u << (prev & 31);
                            const int idx = k * 4 + (prev >> 5);
                            if (!(s_rel[idx] & bit)) atomicOr(&s_rel[idx], bit);
                        }
                    }
                }
            }
        }
        __syncthreads();                                                  // s_band is rewritten for the next row
    }
    for (int k = threadIdx.x; k < K; k += blockDim.x)
        if (s_base[k]) atomicMax(base + img * Kmax + k, s_base[k]);
    for (int i = threadIdx.x; i < K * 4; i += blockDim.x)
        if (s_rel[i]) atomicOr(rel + (long long) img * Kmax * 4 + i, s_rel[i]);
}

// one warp per image: v_k = max(base_k, max_{j in R[k]} v_j), k ascending (R[k] only holds j < k)
__global__ void k_adjp_resolve(const int* __restrict__ num, int Kmax, const unsigned* __restrict__ base, const unsigned* __restrict__ rel,
                               unsigned* __restrict__ vbits, const int* __restrict__ need) {
    if (*need) return;
    __shared__ unsigned v[kParMaxK];
    const int img = blockIdx.x, K = min(num[img], Kmax), lane = threadIdx.x;
    for (int k = 0; k < K; ++k) {
        unsigned m = 0u;
        if (lane < 4) {
            unsigned w = rel[((long long) img * Kmax + k) * 4 + lane];
            while (w) {
                const int j = lane * 32 + (__ffs(w) - 1);
                w &= w - 1;
                m = max(m, v[j]);
            }
        }
        for (int o = 2; o; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (lane == 0) v[k] = max(m, base[img * Kmax + k]);
        __syncwarp();
    }
    for (int k = lane; k < K; k += 32) vbits[img * Kmax + k] = v[k];
}

__global__ void __launch_bounds__(256) k_adjp_apply(float* __restrict__ disp, int Kmax, int H, int W, int Kw, const unsigned* __restrict__ cover,
                                                    const unsigned* __restrict__ vbits, const int* __restrict__ need) {
    if (*need) return;
    const int img = blockIdx.y;
    const long long HW = (long long) H * W;
    float* D = disp + (long long) img * HW;
    const unsigned* C = cover + (long long) img * HW * Kw;
    for (long long p = (long long) blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += (long long) gridDim.x * blockDim.x) {
        for (int wi = Kw - 1; wi >= 0; --wi) {
            const unsigned w = C[(long long) wi * HW + p];
            if (w) {
                D[p] = __uint_as_float(vbits[img * Kmax + wi * 32 + 31 - __clz(w)]);
                break;
            }
        }
    }
}

}  // namespace

// state (int32 words): [tops | mtops | bottoms | mbottoms | vbits] x (N*Kmax), bars [N], need [4], base [N*Kmax], rel [N*Kmax*4], (pad to 16 B) cover [N][Kw][H*W]
extern "C" long long csb_depth_adjust_state_words(int N, int Kmax, int H, int W) {
    const long long slots = (long long) N * Kmax;
    const long long Kw = (Kmax + 31) / 32;
    long long words = 5 * slots + N + 4 + slots + slots * 4;
    if (Kmax <= kParMaxK && W % 4 == 0) words += 3 + (long long) N * H * W * Kw;      // + alignment of the cover planes
    return words;
}

extern "C" int csb_depth_adjust_batch(float* disparity, const uint8_t* masks, const int* num, int N, int Kmax, int H, int W, int32_t* state, void* stream) {
    CSB_REQUIRE(disparity && masks && num && state, "null pointer");
    CSB_REQUIRE(N > 0 && Kmax > 0 && H > 0 && W > 0, "bad shape");
    cudaStream_t st = (cudaStream_t) stream;
    int G = csb::num_sms() / N;
    CSB_REQUIRE(G >= 1, "at most one image per SM (N <= SM count): split the batch");
    G = G > 16 ? 16 : G;
    const size_t slots = (size_t) N * Kmax;
    int* tops = state;                       // [tops | mtops] initialised to 0x7f7f7f7f, [bottoms | mbottoms] to -1, everything after to 0
    int* mtops = state + slots;
    int* bottoms = state + 2 * slots;
    int* mbottoms = state + 3 * slots;
    unsigned* vbits = reinterpret_cast<unsigned*>(state + 4 * slots);
    unsigned* bars = reinterpret_cast<unsigned*>(state + 5 * slots);
    int* need = state + 5 * slots + N;
    unsigned* base = reinterpret_cast<unsigned*>(need + 4);
    unsigned* rel = base + slots;
    unsigned* cover = reinterpret_cast<unsigned*>(state) + ((10 * slots + N + 4 + 3) & ~(size_t) 3);       // 16 B aligned (uint4 plane stores)
    CSB_TRY(csb::cuda_ok(cudaMemsetAsync(tops, 0x7f, sizeof(int) * 2 * slots, st), "memset"));
    CSB_TRY(csb::cuda_ok(cudaMemsetAsync(bottoms, 0xff, sizeof(int) * 2 * slots, st), "memset"));
    CSB_TRY(csb::cuda_ok(cudaMemsetAsync(vbits, 0, sizeof(int) * (slots + N + 4 + slots + slots * 4), st), "memset"));
    csb::memset_done(st);
    const int* need_arg = nullptr;
    if (Kmax <= kParMaxK && W % 4 == 0) {                            // order-free path; falls through to the sequential kernel if `need` gets raised
        const int Kw = (Kmax + 31) / 32;
        int gx = (4 * csb::num_sms() + N - 1) / N;
        gx = gx < 1 ? 1 : gx;
        // mask-row extents: the sequential kernel computes the same values into the same slots, so a fallback run is not disturbed
        if (W % 16 == 0 && ((uintptr_t) masks & 15) == 0 && ((long long) H * W) % 16 == 0)
            k_adjp_cover16<<<dim3(gx, N), 256, 0, st>>>(disparity, masks, num, Kmax, H, W, Kw, mtops, mbottoms, cover, need);
        else
            k_adjp_cover<<<dim3(gx, N), 256, 0, st>>>(disparity, masks, num, Kmax, H, W, Kw, mtops, mbottoms, cover, need);
        CSB_TRY(csb::launched("k_adjp_cover", st));
        k_adjp_rel<<<dim3(gx, N), 256, 0, st>>>(disparity, num, Kmax, H, W, Kw, mtops, mbottoms, cover, base, rel, need);
        CSB_TRY(csb::launched("k_adjp_rel", st));
        k_adjp_resolve<<<N, 32, 0, st>>>(num, Kmax, base, rel, vbits, need);
        CSB_TRY(csb::launched("k_adjp_resolve", st));
        k_adjp_apply<<<dim3(gx, N), 256, 0, st>>>(disparity, Kmax, H, W, Kw, cover, vbits, need);
        CSB_TRY(csb::launched("k_adjp_apply", st));
        need_arg = need;
    }
    void* args[] = {&disparity, &masks, &num, &Kmax, &H, &W, &G, &tops, &bottoms, &mtops, &mbottoms, &vbits, &bars, &need_arg};
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

// ------------------------------------------------------------------------------------------------ median variant (use_medium=True)
// reference kenburns_effect.py:80: `tenAdjusted[tenPlane > 0] = tenAdjusted[tenPlane > 0].median()` per instance, in order (torch.median = the LOWER
// median: sorted[(n - 1) / 2]).  Exact selection by a 4-pass byte radix select on the float bit patterns (the selected disparities are > 0, and
// positive floats order like their bits); per instance 4 x (histogram, pick) + 1 apply launch, no host read.  An instance with no positive masked
// disparity is skipped (the reference skips when the masked SUM is zero, which for the positive disparities of this path is the same condition).
namespace {
struct MedState { uint32_t hist[256]; uint32_t prefix, rank, count, pad; };

__global__ void __launch_bounds__(256) k_med_hist(const float* __restrict__ d, const uint8_t* __restrict__ m, long long P, MedState* st, int shift) {
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t pre = st->prefix;
    for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < P; i += (long long) gridDim.x * blockDim.x) {
        const float f = d[i];
        if (!m[i] || !(f > 0.f)) continue;
        const uint32_t k = __float_as_uint(f);
        if (shift == 24 || (k >> (shift + 8)) == (pre >> (shift + 8))) atomicAdd(&h[(k >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (h[threadIdx.x]) atomicAdd(&st->hist[threadIdx.x], h[threadIdx.x]);
}

__global__ void k_med_pick(MedState* st, int shift) {                 // one warp
    const int lane = threadIdx.x;
    uint32_t c[8], s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) { c[j] = st->hist[lane * 8 + j]; s += c[j]; }
    uint32_t incl = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    uint32_t rank = st->rank, pre = shift == 24 ? 0u : st->prefix;
    if (shift == 24) rank = total ? (total - 1) / 2 : 0;
    __syncwarp();
    const uint32_t excl = incl - s;
    if (total && rank >= excl && rank < incl) {
        uint32_t run = excl;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (rank >= run && rank < run + c[j]) { st->prefix = pre | ((uint32_t) (lane * 8 + j) << shift); st->rank = rank - run; }
            run += c[j];
        }
    }
    if (shift == 24 && lane == 0) st->count = total;
#pragma unroll
    for (int j = 0; j < 8; ++j) st->hist[lane * 8 + j] = 0;
}

__global__ void __launch_bounds__(256) k_med_apply(float* __restrict__ d, const uint8_t* __restrict__ m, long long P, const MedState* __restrict__ st) {
    if (st->count == 0) return;
    const float v = __uint_as_float(st->prefix);
    for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < P; i += (long long) gridDim.x * blockDim.x)
        if (m[i] && d[i] > 0.f) d[i] = v;
}
}  // namespace

extern "C" int csb_depth_adjust_median(float* disparity, const uint8_t* masks, int K, int H, int W, void* state, void* stream) {
    CSB_REQUIRE(disparity && state && (masks || K == 0), "null pointer");
    CSB_REQUIRE(K >= 0 && H > 0 && W > 0, "bad shape");
    cudaStream_t st = (cudaStream_t) stream;
    MedState* s = reinterpret_cast<MedState*>(state);
    const long long P = (long long) H * W;
    const int g = csb::wave_grid(P, 256, 4);
    CSB_TRY(csb::cuda_ok(cudaMemsetAsync(s, 0, sizeof(MedState), st), "memset"));
    csb::memset_done(st);
    for (int k = 0; k < K; ++k) {
        const uint8_t* m = masks + (long long) k * P;
        for (int shift = 24; shift >= 0; shift -= 8) {
            k_med_hist<<<g, 256, 0, st>>>(disparity, m, P, s, shift);
            CSB_TRY(csb::launched("k_med_hist", st));
            k_med_pick<<<1, 32, 0, st>>>(s, shift);
            CSB_TRY(csb::launched("k_med_pick", st));
        }
        k_med_apply<<<g, 256, 0, st>>>(disparity, m, P, s);
        CSB_TRY(csb::launched("k_med_apply", st));
    }
    return CSB_OK;
}
