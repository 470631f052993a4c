// Elementwise / stencil neighbours of the point-cloud render for sm_100a:
//   process_shift tensor part   anime_3dkenburns/common.py:76-81
//   depth_to_points             anime_3dkenburns/models/utils.py:43-50
//   spatial_filter              anime_3dkenburns/models/utils.py:9-40  (laplacian / median-3 / median-5)
//   disparity -> cloud          anime_3dkenburns/kenburns_effect.py:928-937 (three global reductions + stencil, 3 launches,
//                               results stay on the device: the reference does 3 .item()/D2H syncs here)
// All HBM-bound: algorithmic bytes are stated next to each kernel.
#include "common.cuh"

namespace {

__device__ __forceinline__ int fkey(float f) {
    int k = __float_as_int(f);
    return k ^ ((k >> 31) & 0x7fffffff);
}
__device__ __forceinline__ unsigned ukey(float f) { return (unsigned) fkey(f) ^ 0x80000000u; }
__device__ __forceinline__ float udec(unsigned u) {
    int k = (int) (u ^ 0x80000000u);
    return __int_as_float(k ^ ((k >> 31) & 0x7fffffff));
}

// 24 B/point
__global__ void __launch_bounds__(256) k_shift(const float* __restrict__ pts, int B, int N, float sx, float sy, float sz, float* __restrict__ out) {
    const long long total = (long long) B * N;
    for (long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x; i < total; i += (long long) gridDim.x * blockDim.x) {
        int b = (int) (i / N), n = (int) (i - (long long) b * N);
        const float* P = pts + (size_t) b * 3 * N;
        float* O = out + (size_t) b * 3 * N;
        float x = P[n], y = P[N + n], z = P[2 * (size_t) N + n];
        float r = __fdiv_rn(z, __fadd_rn(z, 0.0000001f));
        O[n] = __fadd_rn(__fmul_rn(x, r), sx);
        O[N + n] = __fadd_rn(__fmul_rn(y, r), sy);
        O[2 * (size_t) N + n] = __fadd_rn(z, sz);
    }
}

__device__ __forceinline__ void d2p(float d, int x, int y, int H, int W, float inv, float& px, float& py) {
    float hx = __fmul_rn(__fadd_rn((float) x, __fadd_rn(-0.5f * (float) W, 0.5f)), inv);    // linspace entry (exact) * float32(1/focal)
    float vy = __fmul_rn(__fadd_rn((float) y, __fadd_rn(-0.5f * (float) H, 0.5f)), inv);
    px = __fmul_rn(d, hx);
    py = __fmul_rn(d, vy);
}

// 16 B/px
__global__ void __launch_bounds__(256) k_depth_to_points(const float* __restrict__ depth, int B, int H, int W, float inv, float* __restrict__ pts) {
    const long long HW = (long long) H * W, total = (long long) B * HW;
    for (long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x; i < total; i += (long long) gridDim.x * blockDim.x) {
        int b = (int) (i / HW);
        long long o = i - b * HW;
        int x = (int) (o % W), y = (int) (o / W);
        float d = depth[i], px, py;
        d2p(d, x, y, H, W, inv, px, py);
        float* O = pts + (size_t) b * 3 * HW;
        O[o] = px;
        O[HW + o] = py;
        O[2 * HW + o] = d;
    }
}

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
__device__ __forceinline__ int reflecti(int v, int n) {
    if (v < 0) v = -v;
    if (v >= n) v = 2 * (n - 1) - v;
    return v;
}

// Reference stencil (models/utils.py:16-20): w[0][1] = w[0][2] = w[1][0] = w[2][0] = -1, w[1][1] = 4, replicate padding,
// taps accumulated in row-major order.
template <class F>
__device__ __forceinline__ float laplace5(F px) {
    float acc = 0.0f;
    acc = __fadd_rn(acc, __fmul_rn(-1.0f, px(-1, 0)));
    acc = __fadd_rn(acc, __fmul_rn(-1.0f, px(-1, 1)));
    acc = __fadd_rn(acc, __fmul_rn(-1.0f, px(0, -1)));
    acc = __fadd_rn(acc, __fmul_rn(4.0f, px(0, 0)));
    acc = __fadd_rn(acc, __fmul_rn(-1.0f, px(1, -1)));
    return acc;
}

// 8 B/px (stencil neighbours come from L1)
__global__ void __launch_bounds__(256) k_laplacian(const float* __restrict__ in, long long planes, int H, int W, float* __restrict__ out) {
    const long long HW = (long long) H * W, total = planes * HW;
    for (long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x; i < total; i += (long long) gridDim.x * blockDim.x) {
        long long o = i % HW;
        const float* I = in + (i - o);
        int x = (int) (o % W), y = (int) (o / W);
        out[i] = laplace5([&](int dy, int dx) { return __ldg(I + (size_t) clampi(y + dy, 0, H - 1) * W + clampi(x + dx, 0, W - 1)); });
    }
}

// torch.median over the K*K reflect-padded window = element (K*K-1)/2 of the sorted window (lower median).
template <int K>
__global__ void __launch_bounds__(256) k_median(const float* __restrict__ in, long long planes, int H, int W, float* __restrict__ out) {
    constexpr int M = K * K, R = K / 2;
    const long long HW = (long long) H * W, total = planes * HW;
    for (long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x; i < total; i += (long long) gridDim.x * blockDim.x) {
        long long o = i % HW;
        const float* I = in + (i - o);
        int x = (int) (o % W), y = (int) (o / W);
        float v[M];
#pragma unroll
        for (int dy = -R; dy <= R; ++dy)
#pragma unroll
            for (int dx = -R; dx <= R; ++dx)
                v[(dy + R) * K + dx + R] = __ldg(I + (size_t) reflecti(y + dy, H) * W + reflecti(x + dx, W));
        // partial selection: after pass j, v[j] holds the (j+1)-th smallest
#pragma unroll
        for (int j = 0; j <= (M - 1) / 2; ++j) {
#pragma unroll
            for (int k = j + 1; k < M; ++k) {
                float lo = fminf(v[j], v[k]), hi = fmaxf(v[j], v[k]);
                v[j] = lo;
                v[k] = hi;
            }
        }
        out[i] = v[(M - 1) / 2];
    }
}

// ---- disparity -> cloud (kenburns_effect.py:928-937) -------------------------------------------------------
// scratch (uint64[8]): [0] max(raw) key, [1] ~min(disp) key, [2] max(disp) key, [3] ~(min depth key, index), [4] (max depth key, ~index)
__global__ void k_d2c_init(unsigned long long* scratch) {
    if (threadIdx.x < 8) scratch[threadIdx.x] = 0ull;
}

__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long t = __shfl_xor_sync(0xffffffffu, v, o);
        v = t > v ? t : v;
    }
    return v;
}

template <int NV>
__device__ __forceinline__ void block_max_to_global(unsigned long long (&v)[NV], unsigned long long* dst) {
    __shared__ unsigned long long sm[NV][8];
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        unsigned long long w = warp_max_u64(v[j]);
        if ((threadIdx.x & 31) == 0) sm[j][threadIdx.x >> 5] = w;
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        unsigned long long m = 0;
        for (int i = 0; i < (int) (blockDim.x >> 5); ++i) m = sm[threadIdx.x][i] > m ? sm[threadIdx.x][i] : m;
        atomicMax(dst + threadIdx.x, m);
    }
}

// pass 1: max(raw).  4 B/px
__global__ void __launch_bounds__(256) k_d2c_max(const float* __restrict__ raw, long long HW, unsigned long long* scratch) {
    unsigned long long v[1] = {0ull};
    for (long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x; i < HW; i += (long long) gridDim.x * blockDim.x) {
        unsigned long long k = ukey(raw[i]);
        v[0] = k > v[0] ? k : v[0];
    }
    block_max_to_global<1>(v, scratch);
}

// pass 2: disparity, depth and their reductions.  12 B/px
__global__ void __launch_bounds__(256) k_d2c_scale(const float* __restrict__ raw, int H, int W, float fb, float ffb,
                                                   float* __restrict__ disp, float* __restrict__ depth, unsigned long long* scratch) {
    const long long HW = (long long) H * W;
    const float mx = udec((unsigned) scratch[0]);
    unsigned long long v[4] = {0ull, 0ull, 0ull, 0ull};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < (int) HW; i += gridDim.x * blockDim.x) {
        float d = __fmul_rn(__fdiv_rn(raw[i], mx), fb);                                  // :928
        float z = __fdiv_rn(ffb, __fadd_rn(d, 0.00001f));                                // :929
        disp[i] = d;
        depth[i] = z;
        unsigned long long kmin = (unsigned long long) (~ukey(d)), kmax = (unsigned long long) ukey(d);
        v[0] = kmin > v[0] ? kmin : v[0];
        v[1] = kmax > v[1] ? kmax : v[1];
        const int x = i % W, y = i / W;
        if (H > 256 && W > 256 && x >= 128 && x < W - 128 && y >= 128 && y < H - 128) {  // :937 crop [128:-128]
            unsigned idx = (unsigned) ((y - 128) * (W - 256) + (x - 128));
            unsigned long long pmin = ((unsigned long long) (~ukey(z)) << 32) | (unsigned) (~idx);   // smallest value, then first index
            unsigned long long pmax = ((unsigned long long) ukey(z) << 32) | (unsigned) (~idx);      // largest value, then first index
            v[2] = pmin > v[2] ? pmin : v[2];
            v[3] = pmax > v[3] ? pmax : v[3];
        }
    }
    block_max_to_global<4>(v, scratch + 1);
}

// pass 3: valid mask + both point clouds (+ scalar decode).  36 B/px
__global__ void __launch_bounds__(256) k_d2c_points(const float* __restrict__ disp, const float* __restrict__ depth, int H, int W, float inv,
                                                    float* __restrict__ valid, float* __restrict__ pts, float* __restrict__ unalt,
                                                    const unsigned long long* __restrict__ scratch, float* __restrict__ scalars,
                                                    const uint8_t* __restrict__ image, float* __restrict__ data4) {
    const long long HW = (long long) H * W;
    const float mx2 = udec((unsigned) scratch[2]);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        scalars[0] = udec(~(unsigned) scratch[1]);
        scalars[1] = mx2;
        const int cw = W - 256;
        unsigned long long pmin = scratch[3], pmax = scratch[4];
        unsigned imin = ~(unsigned) pmin, imax = ~(unsigned) pmax;
        scalars[2] = udec(~(unsigned) (pmin >> 32));
        scalars[3] = udec((unsigned) (pmax >> 32));
        scalars[4] = cw > 0 ? (float) (imin % cw) : 0.f;
        scalars[5] = cw > 0 ? (float) (imin / cw) : 0.f;
        scalars[6] = cw > 0 ? (float) (imax % cw) : 0.f;
        scalars[7] = cw > 0 ? (float) (imax / cw) : 0.f;
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < (int) HW; i += gridDim.x * blockDim.x) {
        const int x = i % W, y = i / W;
        float lap = laplace5([&](int dy, int dx) {
            return __fdiv_rn(__ldg(disp + (size_t) clampi(y + dy, 0, H - 1) * W + clampi(x + dx, 0, W - 1)), mx2);   // :931
        });
        float v = fabsf(lap) < 0.03f ? 1.0f : 0.0f;
        float d = depth[i], px, py;
        valid[i] = v;
        float dv = __fmul_rn(d, v);
        d2p(dv, x, y, H, W, inv, px, py);                                                // :932
        pts[i] = px; pts[HW + i] = py; pts[2 * HW + i] = dv;
        d2p(d, x, y, H, W, inv, px, py);                                                 // :933
        unalt[i] = px; unalt[HW + i] = py; unalt[2 * HW + i] = d;
        if (data4) {   // render payload: tenRawImage = BGR * float32(1/255) planar (kenburns_effect.py:921) ++ tenRawDepth (:1036)
            const float k = (float) (1.0 / 255.0);
            data4[i] = __fmul_rn((float) image[(size_t) i * 3 + 0], k);
            data4[HW + i] = __fmul_rn((float) image[(size_t) i * 3 + 1], k);
            data4[2 * HW + i] = __fmul_rn((float) image[(size_t) i * 3 + 2], k);
            data4[3 * HW + i] = d;
        }
    }
}

}  // namespace

extern "C" int csb_points_shift(const float* points, int B, int N, const float* shift, float* out, void* stream) {
    CSB_REQUIRE(points && shift && out, "null pointer");
    CSB_REQUIRE(B > 0 && N >= 0, "bad shape");
    if ((long long) B * N == 0) return CSB_OK;
    k_shift<<<csb::wave_grid((long long) B * N, 256, 8), 256, 0, (cudaStream_t) stream>>>(points, B, N, shift[0], shift[1], shift[2], out);
    return csb::launched("k_shift", (cudaStream_t) stream);
}

extern "C" int csb_depth_to_points(const float* depth, int B, int H, int W, double focal, float* points, void* stream) {
    CSB_REQUIRE(depth && points, "null pointer");
    CSB_REQUIRE(B > 0 && H > 0 && W > 0, "bad shape");
    k_depth_to_points<<<csb::wave_grid((long long) B * H * W, 256, 8), 256, 0, (cudaStream_t) stream>>>(depth, B, H, W, (float) (1.0 / focal), points);
    return csb::launched("k_depth_to_points", (cudaStream_t) stream);
}

extern "C" int csb_spatial_filter(const float* input, int B, int C, int H, int W, int kind, float* output, void* stream) {
    CSB_REQUIRE(input && output && input != output, "null or aliased pointer");
    CSB_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, "bad shape");
    CSB_REQUIRE(kind == 0 || ((kind == 3 || kind == 5) && H > kind / 2 && W > kind / 2), "kind must be 0 (laplacian), 3 or 5 (median)");
    const long long planes = (long long) B * C;
    const int grid = csb::wave_grid(planes * H * W, 256, 8);
    cudaStream_t st = (cudaStream_t) stream;
    if (kind == 0) k_laplacian<<<grid, 256, 0, st>>>(input, planes, H, W, output);
    else if (kind == 3) k_median<3><<<grid, 256, 0, st>>>(input, planes, H, W, output);
    else k_median<5><<<grid, 256, 0, st>>>(input, planes, H, W, output);
    return csb::launched("k_spatial_filter", st);
}

extern "C" int csb_disparity_to_cloud(const float* raw, int H, int W, double focal, double baseline, float* disparity, float* depth,
                                      float* valid, float* points, float* unaltered, float* scalars, uint64_t* scratch,
                                      const uint8_t* image_hwc, float* data4, void* stream) {
    CSB_REQUIRE((image_hwc == nullptr) == (data4 == nullptr), "image_hwc and data4 go together");
    CSB_REQUIRE(raw && disparity && depth && valid && points && unaltered && scalars && scratch, "null pointer");
    CSB_REQUIRE(H > 0 && W > 0, "bad shape");
    cudaStream_t st = (cudaStream_t) stream;
    const long long HW = (long long) H * W;
    const int grid = csb::wave_grid(HW, 256, 4);
    unsigned long long* sc = reinterpret_cast<unsigned long long*>(scratch);
    k_d2c_init<<<1, 32, 0, st>>>(sc);
    CSB_TRY(csb::launched("k_d2c_init", st));
    k_d2c_max<<<grid, 256, 0, st>>>(raw, HW, sc);
    CSB_TRY(csb::launched("k_d2c_max", st));
    k_d2c_scale<<<grid, 256, 0, st>>>(raw, H, W, (float) baseline, (float) (focal * baseline), disparity, depth, sc);
    CSB_TRY(csb::launched("k_d2c_scale", st));
    k_d2c_points<<<grid, 256, 0, st>>>(disparity, depth, H, W, (float) (1.0 / focal), valid, points, unaltered, sc, scalars, image_hwc, data4);
    return csb::launched("k_d2c_points", st);
}
