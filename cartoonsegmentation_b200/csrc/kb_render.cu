// Point-cloud render for sm_100a: z-pass -> degrid -> splat -> normalise, plus the batched autozoom coverage
// variant.  Replaces the three cupy kernels of the reference's render_pointcloud
// (anime_3dkenburns/models/utils.py:56-315) and folds process_shift's tensor part
// (anime_3dkenburns/common.py:76-81) into the point loaders.
//
// Numerics follow the SASS of the reference kernels (built unmodified by oracle/build_ref_kernels.py):
// double-precision evaluation wherever the kernel strings use bare literals, one FFMA for the ray/plane
// intersection, IEEE division.  Every fp32 operation whose rounding matters is written with an explicit
// intrinsic so that nvcc's -fmad contraction cannot change it.
//
// Data layout in HBM
//   points [B,3,N] / data [B,C,N]  channel-planar as in the reference -> consecutive lanes read consecutive
//                                  points: every load is a fully coalesced 128 B line per warp.
//   zkey   [B,H,W] int32           z-buffer as order-preserving integer keys of fltError, so the z-pass is one
//                                  native RED.MIN.S32 per point (the reference spins on atomicCAS,
//                                  utils/cupy_utils.py:21-29).  Initialised by a byte memset (0x7f7f7f7f > any
//                                  key that can occur); the 1e6 pre-fill of the reference (:59) is applied on read.
//   acc    [B,H,W,CP] fp32         CP = roundup(C+1, 4): channel-INTERLEAVED accumulator, weight in slot C, so
//                                  one corner of one point is CP/4 vector reductions (RED.ADD.F32x4) instead of
//                                  C+1 scalar atomics (:268-298: 4*(C+1) atomicAdds per point).
#include <cuda_fp16.h>

#include "common.cuh"

namespace {

constexpr float kZeeInit = 1000000.0f;

struct Shift {
    float sx, sy, sz;
    int on;             // 0: none, 1: host values above, 2: read 3 floats from `dev` (sync-free pipelines)
    const float* dev;
};
__device__ __forceinline__ Shift resolve(Shift s) {
    if (s.on == 2) { s.sx = __ldg(s.dev); s.sy = __ldg(s.dev + 1); s.sz = __ldg(s.dev + 2); }
    return s;
}
struct ShiftTable {
    float v[256][3];
};

struct Proj {
    float ox, oy, err;
    int x0, y0;
    float w[4];  // NW, NE, SW, SE
    bool ok;
};

// common.py:78-81 (torch fp32: mul, then add; the ratio is 1 whenever z + 1e-7f rounds back to z).
__device__ __forceinline__ void apply_shift(float& x, float& y, float& z, float sx, float sy, float sz) {
    float zp = __fadd_rn(z, 0.0000001f);
    float r = (zp == z && fabsf(z) <= 3.402823466e38f) ? 1.0f : __fdiv_rn(z, zp);
    x = __fadd_rn(__fmul_rn(x, r), sx);
    y = __fadd_rn(__fmul_rn(y, r), sy);
    z = __fadd_rn(z, sz);
}

// models/utils.py:76-113 == :229-266
__device__ __forceinline__ Proj project(float x, float y, float z, int H, int W, double focal, double focal_baseline) {
    Proj p;
    p.ok = false;
    if (z < 0.001f) return p;                                                  // :82  (double) z < 0.001: float(0.001) is the smallest float >= 0.001, so the float compare decides identically
    float nx = 0.0f - x, ny = 0.0f - y, nz = 0.0f - z;                         // fltLineVector :80
    float s = __fmaf_rn(0.0f, nx, __fmul_rn(0.0f, ny));                        // x/y terms of both dot products
    float num = __fadd_rn(__fsub_rn((float) focal, z), s);                     // :86
    float den = __fadd_rn(nz, s);                                              // :87
    float dist = __fdiv_rn(num, den);                                          // :88
    if (fabsf(den) < 0.001f) return p;                                         // :90  (same argument)
    float ix = __fmaf_rn(nx, dist, x);                                         // :94
    float iy = __fmaf_rn(ny, dist, y);
    p.ox = __fadd_rn(ix, (float) (0.5 * W - 0.5));                             // :96  the double sum is exact and rounded once: one float add of the exact constant
    p.oy = __fadd_rn(iy, (float) (0.5 * H - 0.5));                             // :97
    p.err = (float) (1000000.0 - (focal_baseline / ((double) z + 0.0000001))); // :99
    p.x0 = (int) floorf(p.ox);
    p.y0 = (int) floorf(p.oy);
    float ex = __fsub_rn((float) (p.x0 + 1), p.ox), ey = __fsub_rn((float) (p.y0 + 1), p.oy);
    float fx = __fsub_rn(p.ox, (float) p.x0), fy = __fsub_rn(p.oy, (float) p.y0);
    p.w[0] = __fmul_rn(ex, ey);                                                // :110-113
    p.w[1] = __fmul_rn(fx, ey);
    p.w[2] = __fmul_rn(ex, fy);
    p.w[3] = __fmul_rn(fx, fy);
    p.ok = true;
    return p;
}

// Order-preserving float <-> int32 key (handles negative fltError too).
__device__ __forceinline__ int fkey(float f) {
    int k = __float_as_int(f);
    return k ^ ((k >> 31) & 0x7fffffff);
}
__device__ __forceinline__ float fdec(int k) { return __int_as_float(k ^ ((k >> 31) & 0x7fffffff)); }

__device__ __forceinline__ void zpass_point(float x, float y, float z, int H, int W, double focal, double fb, int* zk) {
    Proj p = project(x, y, z, H, W, focal, fb);
    if (!p.ok) return;
    const float a = p.w[0], b = p.w[1], c = p.w[2], d = p.w[3];
    int tx, ty;                                                                 // :115-135
    if (a >= b && a >= c && a >= d) { tx = p.x0; ty = p.y0; }
    else if (b >= a && b >= c && b >= d) { tx = p.x0 + 1; ty = p.y0; }
    else if (c >= a && c >= b && c >= d) { tx = p.x0; ty = p.y0 + 1; }
    else if (d >= a && d >= b && d >= c) { tx = p.x0 + 1; ty = p.y0 + 1; }
    else return;
    if (tx < 0 || tx >= W || ty < 0 || ty >= H) return;
    if (p.err == p.err) atomicMin(zk + (size_t) ty * W + tx, fkey(p.err));     // float atomicMin, cupy_utils.py:21-29
}

__global__ void __launch_bounds__(256) k_zpass(const float* __restrict__ pts, int B, int N, int H, int W, double focal,
                                               double fb, Shift sh_, int* __restrict__ zkey) {
    const Shift sh = resolve(sh_);
    const int b = blockIdx.y;                                   // batch on grid.y: no 64-bit division per point
    const float* P = pts + (size_t) b * 3 * N;
    int* zk = zkey + (size_t) b * H * W;
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
        float x = __ldg(P + n), y = __ldg(P + N + n), z = __ldg(P + 2 * (size_t) N + n);
        if (sh.on) apply_shift(x, y, z, sh.sx, sh.sy, sh.sz);
        zpass_point(x, y, z, H, W, focal, fb, zk);
    }
}

// Batched over candidate shifts (blockIdx.y): autozoom, common.py:99-131.
__global__ void __launch_bounds__(256) k_zpass_batched(const float* __restrict__ pts, int N, int H, int W, double focal, double fb,
                                                       const __grid_constant__ ShiftTable tab, int s0, int* __restrict__ zkey) {
    const int s = blockIdx.y;
    const float sx = tab.v[s][0], sy = tab.v[s][1], sz = tab.v[s][2];
    int* zk = zkey + (size_t) (s0 + s) * H * W;
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
        float x = __ldg(pts + n), y = __ldg(pts + N + n), z = __ldg(pts + 2 * (size_t) N + n);
        apply_shift(x, y, z, sx, sy, sz);
        zpass_point(x, y, z, H, W, focal, fb, zk);
    }
}

// kernel_pointrender_updateDegrid :152-212, out of place (the reference's in-place update is a race; reading the
// pre-update value everywhere is the outcome of the schedule "all loads before all stores").
__global__ void __launch_bounds__(256) k_degrid(const int* __restrict__ zkey, int planes, int H, int W, float* __restrict__ zee) {
    const int HW = H * W;
    for (int plane = blockIdx.y; plane < planes; plane += gridDim.y)
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
        const int x = i % W, y = i / W;
        const int* Z = zkey + (size_t) plane * HW;
        auto zv = [&](int yy, int xx) { return fminf(fdec(__ldg(Z + (size_t) yy * W + xx)), kZeeInit); };
        const float c = zv(y, x);
        int cnt = 0;
        float sum = 0.0f;
        const int ox[4] = {1, 0, 1, 1}, oy[4] = {0, 1, 1, -1};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            int x1 = x + ox[k], y1 = y + oy[k], x2 = x - ox[k], y2 = y - oy[k];
            if (x1 < 0 || x1 >= W || y1 < 0 || y1 >= H || x2 < 0 || x2 >= W || y2 < 0 || y2 >= H) continue;
            float a = zv(y1, x1), d = zv(y2, x2);
            // :187-188 compares in double: c >= a + 1.0 with the sum exact.  For floats c, a that is c >= (smallest float >= a + 1) = add.ru
            if (c >= __fadd_ru(a, 1.0f) && c >= __fadd_ru(d, 1.0f)) {
                cnt += 2;
                sum = __fadd_rn(sum, a);
                sum = __fadd_rn(sum, d);
            }
        }
        float o = c;
        if (cnt > 0) o = fminf(c, __fdiv_rn(sum, (float) cnt));                        // :197
        zee[(size_t) plane * HW + i] = o;
    }
}

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// kernel_pointrender_updateOutput :215-313 on the interleaved accumulator.
__global__ void __launch_bounds__(256) k_splat(const float* __restrict__ pts, const float* __restrict__ data, const float* __restrict__ zee,
                                               int B, int N, int C, int CP, int H, int W, double focal, double fb, Shift sh_,
                                               float* __restrict__ acc) {
    const Shift sh = resolve(sh_);
    const int b = blockIdx.y;
    const float* P = pts + (size_t) b * 3 * N;
    const float* Z = zee + (size_t) b * H * W;
    float* A = acc + (size_t) b * H * W * CP;
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
        float x = __ldg(P + n), y = __ldg(P + N + n), z = __ldg(P + 2 * (size_t) N + n);
        if (sh.on) apply_shift(x, y, z, sh.sx, sh.sy, sh.sz);
        Proj p = project(x, y, z, H, W, focal, fb);
        if (!p.ok) continue;
        size_t off[4];
        unsigned pass = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            int cx = p.x0 + (k & 1), cy = p.y0 + (k >> 1);
            off[k] = 0;
            if (cx >= 0 && cx < W && cy >= 0 && cy < H) {
                size_t o = (size_t) cy * W + cx;
                if (p.err <= __fadd_rd(__ldg(Z + o), 1.0f)) {                             // :269 in double; err <= z + 1 <=> err <= (largest float <= z + 1) = add.rd
                    pass |= 1u << k;
                    off[k] = o * CP;
                }
            }
        }
        if (!pass) continue;
        const float* D = data + (size_t) b * C * N + n;
        for (int g = 0; g < CP; g += 4) {
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                int c = g + j;
                v[j] = c < C ? __ldg(D + (size_t) c * N) : (c == C ? 1.0f : 0.0f);        // ones channel :57
            }
            const bool single = (g == C);    // this group holds only the weight
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (!(pass & (1u << k))) continue;
                const float w = p.w[k];
                if (single) atomicAdd(A + off[k] + g, __fmul_rn(v[0], w));
                else red_add_v4(A + off[k] + g, __fmul_rn(v[0], w), __fmul_rn(v[1], w), __fmul_rn(v[2], w), __fmul_rn(v[3], w));
            }
        }
    }
}

// Splat with an fp16 channel-INTERLEAVED payload [N, ld] (the 68-channel context render of the Inpaint net, pointcloud_inpainting.py:135: the
// payload comes straight from the conv engine in NHWC fp16, each point reads one contiguous run of C halves).
__global__ void __launch_bounds__(256) k_splat_h16(const float* __restrict__ pts, const __half* __restrict__ data, int ld, const float* __restrict__ zee, int N,
                                                   int C, int CP, int H, int W, double focal, double fb, Shift sh_, float* __restrict__ acc) {
    const Shift sh = resolve(sh_);
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
        float x = __ldg(pts + n), y = __ldg(pts + N + n), z = __ldg(pts + 2 * (size_t) N + n);
        if (sh.on) apply_shift(x, y, z, sh.sx, sh.sy, sh.sz);
        Proj p = project(x, y, z, H, W, focal, fb);
        if (!p.ok) continue;
        size_t off[4];
        unsigned pass = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            int cx = p.x0 + (k & 1), cy = p.y0 + (k >> 1);
            off[k] = 0;
            if (cx >= 0 && cx < W && cy >= 0 && cy < H) {
                size_t o = (size_t) cy * W + cx;
                if (p.err <= __fadd_rd(__ldg(zee + o), 1.0f)) { pass |= 1u << k; off[k] = o * CP; }
            }
        }
        if (!pass) continue;
        const __half* D = data + (size_t) n * ld;
        for (int g = 0; g < CP; g += 4) {
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                int c = g + j;
                v[j] = c < C ? __half2float(D[c]) : (c == C ? 1.0f : 0.0f);
            }
            const bool single = (g == C);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (!(pass & (1u << k))) continue;
                const float w = p.w[k];
                if (single) atomicAdd(acc + off[k] + g, __fmul_rn(v[0], w));
                else red_add_v4(acc + off[k] + g, __fmul_rn(v[0], w), __fmul_rn(v[1], w), __fmul_rn(v[2], w), __fmul_rn(v[3], w));
            }
        }
    }
}

// Inpaint-net input assembly (pointcloud_inpainting.py:140-146): tenExisting = (existing > 0) * median5(existing > 0); render *= tenExisting;
// netInput sees cat([render, existing]).  flags: pass 1 writes (w > 0) per pixel, pass 2 takes the 5x5 reflect-padded majority (median of 25
// zeros/ones = 1 iff at least 13 ones), normalises and writes NHWC fp16 with the existing flag in channel C.
__global__ void __launch_bounds__(256) k_cov_flags(const float* __restrict__ acc, int CP, int C, int HW, uint8_t* __restrict__ flags) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) flags[i] = acc[(size_t) i * CP + C] > 0.0f ? 1 : 0;
}
__device__ __forceinline__ int reflect_idx(int v, int n) {
    if (v < 0) v = -v;
    if (v >= n) v = 2 * (n - 1) - v;
    return v;
}
__global__ void __launch_bounds__(256) k_inpaint_input(const float* __restrict__ acc, const uint8_t* __restrict__ flags, int CP, int C, int H, int W, int ldo,
                                                       __half* __restrict__ out, float* __restrict__ existing) {
    const int HW = H * W;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
        const int x = i % W, y = i / W;
        int cnt = 0;
#pragma unroll
        for (int dy = -2; dy <= 2; ++dy)
#pragma unroll
            for (int dx = -2; dx <= 2; ++dx) cnt += flags[(size_t) reflect_idx(y + dy, H) * W + reflect_idx(x + dx, W)];
        const float e = (flags[i] && cnt >= 13) ? 1.0f : 0.0f;
        existing[i] = e;
        const float* A = acc + (size_t) i * CP;
        const float d = __fadd_rn(A[C], 0.0000001f);
        __half* o = out + (size_t) i * ldo;
        for (int c = 0; c < ldo; ++c) o[c] = __float2half_rn(c < C ? __fmul_rn(__fdiv_rn(A[c], d), e) : (c == C ? e : 0.0f));
    }
}

// host tail :315 -- render = acc[:C] / (acc[C] + 1e-7), existing = acc[C]; interleaved -> planar.
__global__ void __launch_bounds__(256) k_normalise(const float* __restrict__ acc, int B, int C, int CP, int H, int W,
                                                   float* __restrict__ render, float* __restrict__ existing) {
    const int HW = H * W, b = blockIdx.y;
    for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < HW; o += gridDim.x * blockDim.x) {
        const size_t i = (size_t) b * HW + o;
        const float* A = acc + i * CP;
        const float w = A[C];
        if (existing) existing[i] = w;
        if (render) {
            const float d = __fadd_rn(w, 0.0000001f);
            for (int g = 0; g < C; g += 4) {
                float4 q = *reinterpret_cast<const float4*>(A + g);
                float v[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (g + j < C) render[((size_t) b * C + g + j) * HW + o] = __fdiv_rn(v[j], d);
            }
        }
    }
}

// Coverage pass for autozoom: tenExisting > 0  <=>  some point passes the z-test at this pixel with a weight that
// survives the FTZ float reduction.  Idempotent byte stores, no atomics.
__global__ void __launch_bounds__(256) k_cover_batched(const float* __restrict__ pts, const float* __restrict__ zee, int N, int H, int W,
                                                       double focal, double fb, const __grid_constant__ ShiftTable tab, int s0,
                                                       uint8_t* __restrict__ cover) {
    const int s = blockIdx.y;
    const float sx = tab.v[s][0], sy = tab.v[s][1], sz = tab.v[s][2];
    const float* Z = zee + (size_t) (s0 + s) * H * W;
    uint8_t* Cv = cover + (size_t) (s0 + s) * H * W;
    for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += gridDim.x * blockDim.x) {
        float x = __ldg(pts + n), y = __ldg(pts + N + n), z = __ldg(pts + 2 * (size_t) N + n);
        apply_shift(x, y, z, sx, sy, sz);
        Proj p = project(x, y, z, H, W, focal, fb);
        if (!p.ok) continue;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            int cx = p.x0 + (k & 1), cy = p.y0 + (k >> 1);
            if (cx < 0 || cx >= W || cy < 0 || cy >= H) continue;
            if (!(p.w[k] >= 1.17549435e-38f)) continue;
            size_t o = (size_t) cy * W + cx;
            if (p.err <= __fadd_rd(__ldg(Z + o), 1.0f)) Cv[o] = 1;
        }
    }
}

__global__ void __launch_bounds__(256) k_count_cover(const uint8_t* __restrict__ cover, long long HW, int s0, int* __restrict__ counts) {
    const int s = blockIdx.y;
    const uint8_t* Cv = cover + (size_t) (s0 + s) * HW;
    unsigned local = 0;
    const long long nvec = HW / 16;
    const uint4* V = reinterpret_cast<const uint4*>(Cv);
    for (long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x; i < nvec; i += (long long) gridDim.x * blockDim.x) {
        uint4 q = __ldg(V + i);
        local = __dp4a(q.x, 0x01010101u, local);
        local = __dp4a(q.y, 0x01010101u, local);
        local = __dp4a(q.z, 0x01010101u, local);
        local = __dp4a(q.w, 0x01010101u, local);
    }
    if (blockIdx.x == 0)
        for (long long i = nvec * 16 + threadIdx.x; i < HW; i += blockDim.x) local += Cv[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    __shared__ unsigned part[8];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = 0;
        for (int i = 0; i < (int) (blockDim.x >> 5); ++i) t += part[i];
        if (t) atomicAdd(counts + s0 + s, (int) t);
    }
}

// Scalar part of process_shift (common.py:60-72) on the device, in double like the Python original, so a pipeline can
// go from disparity_to_cloud's device scalars to the render without a host read.
__global__ void k_shift_scalars(const float* __restrict__ scalars, int W, int H, double focal, double shiftU, double shiftV, double ratio,
                                double scale, float* __restrict__ out) {
    if (threadIdx.x || blockIdx.x) return;
    const double dmin = (double) scalars[2], fu = (double) scalars[4], fv = (double) scalars[5];
    const double depthFrom = dmin, depthTo = dmin * ratio;
    const double closest = dmin + (depthTo - depthFrom);
    const double tu = fu + shiftU, tv = fv + shiftV;
    const double fx = ((fu - (W / 2.0)) * closest) / focal, fy = ((fv - (H / 2.0)) * closest) / focal;
    const double tx = ((tu - (W / 2.0)) * closest) / focal, ty = ((tv - (H / 2.0)) * closest) / focal;
    out[0] = (float) ((fx - tx) * scale);
    out[1] = (float) ((fy - ty) * scale);
    out[2] = (float) ((depthTo - depthFrom) * scale);
}

Shift make_shift(const float* shift, const float* shift_dev) {
    Shift s{0.f, 0.f, 0.f, 0, nullptr};
    if (shift_dev) s = Shift{0.f, 0.f, 0.f, 2, shift_dev};
    else if (shift) s = Shift{shift[0], shift[1], shift[2], 1, nullptr};
    return s;
}

}  // namespace

extern "C" int csb_render_acc_channels(int C) { return ((C + 1) + 3) / 4 * 4; }

static int zpass_impl(const float* points, int B, int N, int H, int W, double focal, double baseline, const float* shift,
                      const float* shift_dev, int32_t* zkey, cudaStream_t stream) {
    CSB_REQUIRE((points || N == 0) && zkey, "null pointer");
    CSB_REQUIRE(B > 0 && N >= 0 && H > 0 && W > 0, "bad shape");
    cudaStream_t st = (cudaStream_t) stream;
    CSB_TRY(csb::cuda_ok(cudaMemsetAsync(zkey, 0x7f, sizeof(int32_t) * (size_t) B * H * W, st), "memset zkey"));
    csb::memset_done(st);
    if ((long long) B * N == 0) return CSB_OK;
    k_zpass<<<dim3(csb::wave_grid(N, 256, 8), B), 256, 0, st>>>(points, B, N, H, W, focal, focal * baseline, make_shift(shift, shift_dev), zkey);
    return csb::launched("k_zpass", st);
}

extern "C" int csb_pointcloud_zpass(const float* points, int B, int N, int H, int W, double focal, double baseline,
                                    const float* shift, int32_t* zkey, void* stream) {
    return zpass_impl(points, B, N, H, W, focal, baseline, shift, nullptr, zkey, (cudaStream_t) stream);
}

extern "C" int csb_pointcloud_degrid(const int32_t* zkey, int B, int H, int W, float* zee, void* stream) {
    CSB_REQUIRE(zkey && zee, "null pointer");
    CSB_REQUIRE(B > 0 && H > 0 && W > 0, "bad shape");
    k_degrid<<<dim3(csb::wave_grid((long long) H * W, 256, 8), B < 65535 ? B : 65535), 256, 0, (cudaStream_t) stream>>>(zkey, B, H, W, zee);
    return csb::launched("k_degrid", (cudaStream_t) stream);
}

// Shared by csb_pointcloud_render and the fused frame (kb_frame.cu).
int csb_render_accumulate(const float* points, const float* data, int B, int N, int C, int H, int W, double focal, double baseline,
                          const float* shift, const float* shift_dev, int32_t* zkey, float* zee, float* acc, cudaStream_t st) {
    const int CP = csb_render_acc_channels(C);
    CSB_TRY(zpass_impl(points, B, N, H, W, focal, baseline, shift, shift_dev, zkey, st));
    CSB_TRY(csb_pointcloud_degrid(zkey, B, H, W, zee, st));
    CSB_TRY(csb::cuda_ok(cudaMemsetAsync(acc, 0, sizeof(float) * (size_t) B * H * W * CP, st), "memset acc"));
    csb::memset_done(st);
    if ((long long) B * N == 0) return CSB_OK;
    k_splat<<<dim3(csb::wave_grid(N, 256, 8), B), 256, 0, st>>>(points, data, zee, B, N, C, CP, H, W, focal, focal * baseline,
                                                                      make_shift(shift, shift_dev), acc);
    return csb::launched("k_splat", st);
}

extern "C" int csb_pointcloud_render(const float* points, const float* data, int B, int N, int C, int H, int W, double focal,
                                     double baseline, const float* shift, const float* shift_dev, int32_t* zkey, float* zee, float* acc,
                                     float* render, float* existing, void* stream) {
    CSB_REQUIRE(((points && data) || N == 0) && zkey && zee && acc, "null pointer");
    CSB_REQUIRE(B > 0 && N >= 0 && C > 0 && H > 0 && W > 0, "bad shape");
    CSB_REQUIRE(((uintptr_t) acc & 15) == 0, "acc must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t) stream;
    CSB_TRY(csb_render_accumulate(points, data, B, N, C, H, W, focal, baseline, shift, shift_dev, zkey, zee, acc, st));
    if (!render && !existing) return CSB_OK;
    k_normalise<<<dim3(csb::wave_grid((long long) H * W, 256, 8), B), 256, 0, st>>>(acc, B, C, csb_render_acc_channels(C), H, W, render, existing);
    return csb::launched("k_normalise", st);
}

extern "C" int csb_autozoom_coverage(const float* points, int N, int H, int W, double focal, double baseline, const float* shifts,
                                     int S, int32_t* zkey, float* zee, uint8_t* cover, int32_t* counts, void* stream) {
    CSB_REQUIRE(points && shifts && zkey && zee && cover && counts, "null pointer");
    CSB_REQUIRE(N > 0 && H > 0 && W > 0 && S > 0, "bad shape");
    CSB_REQUIRE(((uintptr_t) cover & 15) == 0 && ((long long) H * W) % 16 == 0, "cover must be 16-byte aligned and H*W a multiple of 16");
    cudaStream_t st = (cudaStream_t) stream;
    const size_t HW = (size_t) H * W;
    CSB_TRY(csb::cuda_ok(cudaMemsetAsync(zkey, 0x7f, sizeof(int32_t) * S * HW, st), "memset zkey"));
    csb::memset_done(st);
    CSB_TRY(csb::cuda_ok(cudaMemsetAsync(cover, 0, S * HW, st), "memset cover"));
    csb::memset_done(st);
    CSB_TRY(csb::cuda_ok(cudaMemsetAsync(counts, 0, sizeof(int32_t) * S, st), "memset counts"));
    csb::memset_done(st);
    const double fb = focal * baseline;
    const int gx = csb::wave_grid(N, 256, 2);
    for (int s0 = 0; s0 < S; s0 += 256) {
        const int sc = S - s0 < 256 ? S - s0 : 256;
        ShiftTable tab;
        for (int i = 0; i < sc; ++i)
            for (int j = 0; j < 3; ++j) tab.v[i][j] = shifts[(size_t) (s0 + i) * 3 + j];
        k_zpass_batched<<<dim3(gx, sc), 256, 0, st>>>(points, N, H, W, focal, fb, tab, s0, zkey);
        CSB_TRY(csb::launched("k_zpass_batched", st));
        k_degrid<<<dim3(csb::wave_grid((long long) HW, 256, 2), sc), 256, 0, st>>>(zkey + s0 * HW, sc, H, W, zee + s0 * HW);
        CSB_TRY(csb::launched("k_degrid", st));
        k_cover_batched<<<dim3(gx, sc), 256, 0, st>>>(points, zee, N, H, W, focal, fb, tab, s0, cover);
        CSB_TRY(csb::launched("k_cover_batched", st));
        k_count_cover<<<dim3(16, sc), 256, 0, st>>>(cover, (long long) HW, s0, counts);
        CSB_TRY(csb::launched("k_count_cover", st));
    }
    return CSB_OK;
}

extern "C" int csb_shift_from_scalars(const float* scalars, int W, int H, double focal, double shiftU, double shiftV, double depth_ratio,
                                      float* shift_dev, void* stream) {
    CSB_REQUIRE(scalars && shift_dev, "null pointer");
    k_shift_scalars<<<1, 32, 0, (cudaStream_t) stream>>>(scalars, W, H, focal, shiftU, shiftV, depth_ratio, 1.0, shift_dev);
    return csb::launched("k_shift_scalars", (cudaStream_t) stream);
}

extern "C" int csb_inpaint_context_render(const float* points, const void* data16, int ld, int N, int C, int H, int W, double focal, double baseline,
                                          const float* shift, int32_t* zkey, float* zee, float* acc, uint8_t* flags, void* out16, int ldo, float* existing,
                                          void* stream) {
    CSB_REQUIRE(points && data16 && shift && zkey && zee && acc && flags && out16 && existing, "null pointer");
    CSB_REQUIRE(N > 0 && C > 0 && C <= ld && H > 0 && W > 0 && ldo >= C + 1 && ldo % 8 == 0, "bad shape");
    cudaStream_t st = (cudaStream_t) stream;
    const int CP = csb_render_acc_channels(C);
    CSB_TRY(zpass_impl(points, 1, N, H, W, focal, baseline, shift, nullptr, zkey, st));
    CSB_TRY(csb_pointcloud_degrid(zkey, 1, H, W, zee, st));
    CSB_TRY(csb::cuda_ok(cudaMemsetAsync(acc, 0, sizeof(float) * (size_t) H * W * CP, st), "memset acc"));
    csb::memset_done(st);
    k_splat_h16<<<csb::wave_grid(N, 256, 8), 256, 0, st>>>(points, (const __half*) data16, ld, zee, N, C, CP, H, W, focal, focal * baseline,
                                                            make_shift(shift, nullptr), acc);
    CSB_TRY(csb::launched("k_splat_h16", st));
    k_cov_flags<<<csb::wave_grid((long long) H * W, 256, 8), 256, 0, st>>>(acc, CP, C, H * W, flags);
    CSB_TRY(csb::launched("k_cov_flags", st));
    k_inpaint_input<<<csb::wave_grid((long long) H * W, 256, 8), 256, 0, st>>>(acc, flags, CP, C, H, W, ldo, (__half*) out16, existing);
    return csb::launched("k_inpaint_input", st);
}
