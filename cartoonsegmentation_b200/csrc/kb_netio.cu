// Input / output glue of the two per-image Ken-Burns networks on the device: the reference wraps its Inpaint
// (anime_3dkenburns/models/pointcloud_inpainting.py:117-131, 190-200) and Refine (disparity_refinement.py:99-100, 128-135) nets in eager
// per-tensor ops -- mean / std reductions, (x - mean) / (std + 1e-7), NCHW <-> NHWC permutes, concatenations, de-normalisation, clip / threshold,
// and the valid-masked point cloud of the raw frame -- about 30 small ATen launches per pass.  Here: one statistics pass per tensor and one fused
// elementwise kernel per packing step, all HBM-bound (bytes next to each kernel).
#include <cuda_fp16.h>

#include "common.cuh"

namespace {

__device__ __forceinline__ unsigned okey(float f) {
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float okey_inv(unsigned k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); }

// scratch: [0] sum, [1] sum of squares (double), [2] max key (low 32 bits).  4 B/element.
__global__ void __launch_bounds__(256) k_stats_accum(const float* __restrict__ x, long long n, double* __restrict__ scratch) {
    double s = 0.0, ss = 0.0;
    unsigned mx = 0u;
    for (long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x; i < n; i += (long long) gridDim.x * blockDim.x) {
        const float v = x[i];
        s += (double) v;
        ss += (double) v * (double) v;
        mx = max(mx, okey(v));
    }
    for (int o = 16; o; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        ss += __shfl_xor_sync(0xffffffffu, ss, o);
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(scratch, s);
        atomicAdd(scratch + 1, ss);
        atomicMax(reinterpret_cast<unsigned*>(scratch + 2), mx);
    }
}
// out: [mean, population std (torch.std(unbiased=False)), max]
__global__ void k_stats_finish(const double* __restrict__ scratch, long long n, float* __restrict__ out) {
    if (threadIdx.x || blockIdx.x) return;
    const double mean = scratch[0] / (double) n;
    double var = scratch[1] / (double) n - mean * mean;
    var = var < 0.0 ? 0.0 : var;
    out[0] = (float) mean;
    out[1] = (float) sqrt(var);
    out[2] = okey_inv(*reinterpret_cast<const unsigned*>(scratch + 2));
}

// out [HW][16] fp16 = [ (a - mean_a) / (std_a + eps) (ca planes) | (b - mean_b) / (std_b + eps) (cb planes) | zeros ].  4 (ca + cb) B read, 32 B written per pixel.
__global__ void __launch_bounds__(256) k_pack_norm16(const float* __restrict__ a, int ca, const float* __restrict__ sa, const float* __restrict__ b, int cb,
                                                     const float* __restrict__ sb, long long HW, float eps, __half* __restrict__ out) {
    const float ma = sa ? sa[0] : 0.f, da = sa ? __fadd_rn(sa[1], eps) : 1.f, mb = sb ? sb[0] : 0.f, db = sb ? __fadd_rn(sb[1], eps) : 1.f;
    for (long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x; i < HW; i += (long long) gridDim.x * blockDim.x) {
        __align__(16) __half v[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            float f = 0.f;
            if (c < ca) f = __fdiv_rn(__fsub_rn(a[(size_t) c * HW + i], ma), da);
            else if (c < ca + cb) f = __fdiv_rn(__fsub_rn(b[(size_t) (c - ca) * HW + i], mb), db);
            v[c] = __float2half_rn(f);
        }
        uint4* o = reinterpret_cast<uint4*>(out + (size_t) i * 16);
        o[0] = *reinterpret_cast<const uint4*>(v);
        o[1] = *reinterpret_cast<const uint4*>(v + 8);
    }
}

// payload [HW][72] fp16 = [x16[0:4] | ctx[0:64] | zeros(4)], in 8-byte units (18 per pixel).  136 B read, 144 B written per pixel.
__global__ void __launch_bounds__(256) k_inpaint_payload(const uint2* __restrict__ x16, const uint2* __restrict__ ctx, long long HW, uint2* __restrict__ payload) {
    const long long total = HW * 18;
    for (long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x; i < total; i += (long long) gridDim.x * blockDim.x) {
        const long long pix = i / 18;
        const int u = (int) (i - pix * 18);
        uint2 v = make_uint2(0u, 0u);
        if (u == 0) v = x16[pix * 4];
        else if (u < 17) v = ctx[pix * 16 + (u - 1)];
        payload[i] = v;
    }
}

// out [C][HW] fp32 = post( (a [+ b]) [HW][C] * (std + eps) + mean ), post: 0 none, 1 clip to [0, 1], 2 threshold(0, 0).  torch's operation order (mul, then add).
__global__ void __launch_bounds__(256) k_net_output(const float* __restrict__ a, const float* __restrict__ b, int C, long long HW, const float* __restrict__ st,
                                                    float eps, int post, float* __restrict__ out) {
    const float m = st[0], s = __fadd_rn(st[1], eps);
    const long long total = HW * C;
    for (long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x; i < total; i += (long long) gridDim.x * blockDim.x) {
        const int c = (int) (i / HW);
        const long long pix = i - (long long) c * HW;
        float v = a[pix * C + c];
        if (b) v = __fadd_rn(v, b[pix * C + c]);
        v = __fadd_rn(__fmul_rn(v, s), m);
        if (post == 1) v = fminf(fmaxf(v, 0.0f), 1.0f);
        else if (post == 2) v = v <= 0.0f ? 0.0f : v;
        out[i] = v;
    }
}

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// pointcloud_inpainting.py:117-120: tenDepth = (focal * baseline) / (disp + 1e-7); tenValid = |laplacian(disp / disp.max())| < 0.03;
// tenPoints = depth_to_points(tenDepth * tenValid, focal).  Stencil and point formulas as in kb_points.cu (models/utils.py:16-20, 43-50).
__global__ void __launch_bounds__(256) k_inpaint_points(const float* __restrict__ disp, int H, int W, float fb, float inv, const float* __restrict__ st,
                                                        float* __restrict__ pts) {
    const long long HW = (long long) H * W;
    const float mx = st[2];
    for (long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x; i < HW; i += (long long) gridDim.x * blockDim.x) {
        const int x = (int) (i % W), y = (int) (i / W);
        auto px = [&](int dy, int dx) { return __fdiv_rn(__ldg(disp + (size_t) clampi(y + dy, 0, H - 1) * W + clampi(x + dx, 0, W - 1)), mx); };
        float acc = 0.0f;                                  // the reference's stencil, taps in row-major order (kb_points.cu: laplace5)
        acc = __fadd_rn(acc, __fmul_rn(-1.0f, px(-1, 0)));
        acc = __fadd_rn(acc, __fmul_rn(-1.0f, px(-1, 1)));
        acc = __fadd_rn(acc, __fmul_rn(-1.0f, px(0, -1)));
        acc = __fadd_rn(acc, __fmul_rn(4.0f, px(0, 0)));
        acc = __fadd_rn(acc, __fmul_rn(-1.0f, px(1, -1)));
        const float valid = fabsf(acc) < 0.03f ? 1.0f : 0.0f;
        // torch: `scalar / tensor` is Tensor.__rtruediv__ = tensor.reciprocal() * scalar (two roundings), not one division
        const float d = __fmul_rn(__fmul_rn(__frcp_rn(__fadd_rn(disp[i], 0.0000001f)), fb), valid);
        const float hx = __fmul_rn(__fadd_rn((float) x, __fadd_rn(-0.5f * (float) W, 0.5f)), inv);
        const float vy = __fmul_rn(__fadd_rn((float) y, __fadd_rn(-0.5f * (float) H, 0.5f)), inv);
        pts[i] = __fmul_rn(d, hx);
        pts[HW + i] = __fmul_rn(d, vy);
        pts[2 * HW + i] = d;
    }
}

}  // namespace

extern "C" int csb_tensor_stats(const float* x, long long n, double* scratch, float* out, void* stream) {
    CSB_REQUIRE(x && scratch && out && n > 0, "bad arguments");
    cudaStream_t st = (cudaStream_t) stream;
    CSB_TRY(csb::cuda_ok(cudaMemsetAsync(scratch, 0, 3 * sizeof(double), st), "memset"));
    csb::memset_done(st);
    k_stats_accum<<<csb::wave_grid(n, 256, 4), 256, 0, st>>>(x, n, scratch);
    CSB_TRY(csb::launched("k_stats_accum", st));
    k_stats_finish<<<1, 32, 0, st>>>(scratch, n, out);
    return csb::launched("k_stats_finish", st);
}

extern "C" int csb_pack_norm16(const float* a, int ca, const float* stats_a, const float* b, int cb, const float* stats_b, long long HW, float eps, void* out16,
                               void* stream) {
    CSB_REQUIRE(out16 && HW > 0 && ca >= 0 && cb >= 0 && ca + cb <= 16 && (ca == 0 || a) && (cb == 0 || b), "bad arguments");
    CSB_REQUIRE(((uintptr_t) out16 & 15) == 0, "out must be 16-byte aligned");
    k_pack_norm16<<<csb::wave_grid(HW, 256, 8), 256, 0, (cudaStream_t) stream>>>(a, ca, stats_a, b, cb, stats_b, HW, eps, (__half*) out16);
    return csb::launched("k_pack_norm16", (cudaStream_t) stream);
}

extern "C" int csb_inpaint_payload(const void* x16, const void* ctx64, long long HW, void* payload72, void* stream) {
    CSB_REQUIRE(x16 && ctx64 && payload72 && HW > 0, "bad arguments");
    CSB_REQUIRE((((uintptr_t) x16 | (uintptr_t) ctx64 | (uintptr_t) payload72) & 7) == 0, "pointers must be 8-byte aligned");
    k_inpaint_payload<<<csb::wave_grid(HW * 18, 256, 8), 256, 0, (cudaStream_t) stream>>>((const uint2*) x16, (const uint2*) ctx64, HW, (uint2*) payload72);
    return csb::launched("k_inpaint_payload", (cudaStream_t) stream);
}

extern "C" int csb_net_output(const float* a, const float* b, int C, long long HW, const float* stats, float eps, int post, float* out_nchw, void* stream) {
    CSB_REQUIRE(a && stats && out_nchw && C > 0 && HW > 0 && post >= 0 && post <= 2, "bad arguments");
    k_net_output<<<csb::wave_grid(HW * C, 256, 8), 256, 0, (cudaStream_t) stream>>>(a, b, C, HW, stats, eps, post, out_nchw);
    return csb::launched("k_net_output", (cudaStream_t) stream);
}

extern "C" int csb_inpaint_points(const float* disp, int H, int W, double focal, double baseline, const float* stats, float* points, void* stream) {
    CSB_REQUIRE(disp && stats && points && H > 0 && W > 0, "bad arguments");
    k_inpaint_points<<<csb::wave_grid((long long) H * W, 256, 8), 256, 0, (cudaStream_t) stream>>>(disp, H, W, (float) (focal * baseline), (float) (1.0 / focal), stats, points);
    return csb::launched("k_inpaint_points", (cudaStream_t) stream);
}
