// LeReS post-processing on the device (reference depth_modules/leres/__init__.py:117-140 `apply_leres` tail and
// anime_3dkenburns/kenburns_effect.py:572-577): the reference downloads the network output and runs numpy + OpenCV per image on the host
//     out   = 65535 * (depth - min) / (max - min)            (float32; zeros if max - min <= eps)
//     d8    = bitwise_not(convertScaleAbs(out.astype(uint16), alpha = 255 / 65535))
//     depth = cv2.resize(d8, (W, H), INTER_AREA if upscaling else INTER_LANCZOS4).astype(float32)
// which stalls the GPU pipeline behind the host.  Here: per-image min/max (order-preserving keys), quantisation with the same float32 operation
// order and OpenCV's rounding (cvRound = round-half-even of src * (float) alpha), and OpenCV's INTER_AREA-upscale = INTER_LINEAR fixed-point
// kernel with area-mode source positions (imgproc/src/resize.cpp), all bit-exact against cv2 (oracle: orc_resize_area_up_u8c1, pinned to cv2
// in the CPU suite).  Downscaling (kenburns_effect.py:574-575: INTER_LANCZOS4 when the estimator output has more rows than the frame, i.e. inputs smaller
// than the estimator size rounded up to a multiple of 32) is OpenCV's 8-tap fixed-point resize (resize.cpp: HResizeLanczos4<uchar,int,short> +
// VResizeLanczos4 + FixedPtCast<int,uchar,22>): the coefficient tables (interpolateLanczos4, double sin/cos of the host libm as OpenCV uses, rounded
// to short x 2048) are built once per geometry on the host and cached on the device; k_lt_lanczos evaluates the 8 x 8 taps in OpenCV's int32 arithmetic.
#include <math.h>

#include <map>
#include <mutex>
#include <tuple>
#include <vector>

#include "common.cuh"

namespace {

__device__ __forceinline__ unsigned okey(float f) {
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float okey_inv(unsigned k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); }

// mm [N][2]: min key (init 0xffffffff), max key (init 0)
__global__ void __launch_bounds__(256) k_lt_minmax(const float* __restrict__ x, long long per, unsigned* __restrict__ mm) {
    const int img = blockIdx.y;
    const float* X = x + (long long) img * per;
    unsigned lo = 0xffffffffu, hi = 0u;
    for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < per; i += (long long) gridDim.x * blockDim.x) {
        const unsigned k = okey(X[i]);
        lo = min(lo, k);
        hi = max(hi, k);
    }
    for (int o = 16; o; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(mm + 2 * img, lo);
        atomicMax(mm + 2 * img + 1, hi);
    }
}

__global__ void __launch_bounds__(256) k_lt_quant(const float* __restrict__ x, long long per, const unsigned* __restrict__ mm, uint8_t* __restrict__ q) {
    const int img = blockIdx.y;
    const float mn = okey_inv(mm[2 * img]), mx = okey_inv(mm[2 * img + 1]);
    const float range = __fsub_rn(mx, mn);
    const bool ok = (double) range > 2.220446049250313e-16;                 // np.finfo("float").eps
    const float alpha = (float) (255.0 / 65535.0);
    const float* X = x + (long long) img * per;
    uint8_t* Q = q + (long long) img * per;
    for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < per; i += (long long) gridDim.x * blockDim.x) {
        float out = 0.f;
        if (ok) out = __fdiv_rn(__fmul_rn(65535.0f, __fsub_rn(X[i], mn)), range);
        const unsigned u16 = (unsigned) out & 0xffffu;                        // astype("uint16"): truncation (values are in [0, 65535])
        int v = __float2int_rn(__fmul_rn((float) u16, alpha));                // convertScaleAbs: saturate_cast<uchar>(|src * a|), cvRound
        v = v < 0 ? 0 : (v > 255 ? 255 : v);
        Q[i] = (uint8_t) (255 - v);                                           // bitwise_not
    }
}

__device__ __forceinline__ void area_coef(int d, double scale, double inv, int n, int& s, short& a0, short& a1) {
    s = (int) floor(d * scale);
    float f = (float) ((d + 1) - (s + 1) * inv);
    f = f <= 0 ? 0.f : f - floorf(f);
    if (s < 0) { f = 0.f; s = 0; }
    if (s >= n - 1) { f = 0.f; s = n - 1; }
    a0 = (short) __float2int_rn((1.f - f) * 2048.f);
    a1 = (short) __float2int_rn(f * 2048.f);
}

__global__ void __launch_bounds__(256) k_lt_resize(const uint8_t* __restrict__ q, int h, int w, int H, int W, float* __restrict__ out) {
    const int img = blockIdx.y;
    const uint8_t* S = q + (long long) img * h * w;
    float* O = out + (long long) img * H * W;
    const double inv_x = (double) W / w, inv_y = (double) H / h, scale_x = 1.0 / inv_x, scale_y = 1.0 / inv_y;
    for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < (long long) H * W; i += (long long) gridDim.x * blockDim.x) {
        const int dx = (int) (i % W), dy = (int) (i / W);
        int sx, sy;
        short a0, a1, b0, b1;
        area_coef(dx, scale_x, inv_x, w, sx, a0, a1);
        sy = (int) floor(dy * scale_y);
        float fy = (float) ((dy + 1) - (sy + 1) * inv_y);
        fy = fy <= 0 ? 0.f : fy - floorf(fy);
        b0 = (short) __float2int_rn((1.f - fy) * 2048.f);
        b1 = (short) __float2int_rn(fy * 2048.f);
        const int y0 = min(max(sy, 0), h - 1), y1 = min(max(sy + 1, 0), h - 1);
        const int x1 = sx + 1 < w ? sx + 1 : sx;
        const int s0 = S[(long long) y0 * w + sx] * a0 + S[(long long) y0 * w + x1] * a1;
        const int s1 = S[(long long) y1 * w + sx] * a0 + S[(long long) y1 * w + x1] * a1;
        int v = (((b0 * (s0 >> 4)) >> 16) + ((b1 * (s1 >> 4)) >> 16) + 2) >> 2;
        v = v < 0 ? 0 : (v > 255 ? 255 : v);
        O[i] = (float) v;
    }
}

// cv2.resize(u8 C1, INTER_LANCZOS4): tab = [xofs (W ints) | yofs (H ints)] followed by shorts [alpha (W x 8) | beta (H x 8)]
__global__ void __launch_bounds__(256) k_lt_lanczos(const uint8_t* __restrict__ q, int h, int w, int H, int W, const int* __restrict__ ofs,
                                                    const short* __restrict__ coef, float* __restrict__ out) {
    const int img = blockIdx.y;
    const uint8_t* S = q + (long long) img * h * w;
    float* O = out + (long long) img * H * W;
    const int *xofs = ofs, *yofs = ofs + W;
    const short *alpha = coef, *beta = coef + (long long) W * 8;
    for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < (long long) H * W; i += (long long) gridDim.x * blockDim.x) {
        const int dx = (int) (i % W), dy = (int) (i / W);
        const int sx = xofs[dx], sy = yofs[dy];
        int xs[8], a[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int x = sx - 3 + k;
            xs[k] = x < 0 ? 0 : (x >= w ? w - 1 : x);               // HResizeLanczos4 border loop (cn = 1): clamp to the row
            a[k] = alpha[dx * 8 + k];
        }
        int v = 0;                                                  // WT = int: OpenCV's own 32-bit arithmetic (wraps identically)
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            int y = sy - 3 + r;
            y = y < 0 ? 0 : (y >= h ? h - 1 : y);                   // clip(sy - ksize2 + 1 + k, 0, ssize.height)
            const uint8_t* row = S + (long long) y * w;
            int hsum = 0;
#pragma unroll
            for (int k = 0; k < 8; ++k) hsum += (int) row[xs[k]] * a[k];
            v += hsum * (int) beta[dy * 8 + r];
        }
        v = (v + (1 << 21)) >> 22;                                  // FixedPtCast<int, uchar, INTER_RESIZE_COEF_BITS * 2>
        O[i] = (float) (v < 0 ? 0 : (v > 255 ? 255 : v));
    }
}

// imgproc/src/imgwarp.cpp interpolateLanczos4 (OpenCV 4.x), operation for operation (float / double mix as in the source)
void lanczos4_coeffs(float x, float* coeffs) {
    static const double s45 = 0.70710678118654752440084436210485;
    static const double cs[][2] = {{1, 0}, {-s45, -s45}, {0, 1}, {s45, -s45}, {-1, 0}, {s45, s45}, {0, -1}, {-s45, s45}};
    float sum = 0;
    const double y0 = -(x + 3) * 3.1415926535897932384626433832795 * 0.25, s0 = sin(y0), c0 = cos(y0);
    for (int i = 0; i < 8; i++) {
        const float y0_ = (x + 3 - i);
        if (fabs(y0_) >= 1e-6f) {
            const double y = -y0_ * 3.1415926535897932384626433832795 * 0.25;
            coeffs[i] = (float) ((cs[i][0] * s0 + cs[i][1] * c0) / (y * y));
        } else {
            coeffs[i] = 1e30f;
        }
        sum += coeffs[i];
    }
    sum = 1.f / sum;
    for (int i = 0; i < 8; i++) coeffs[i] *= sum;
}

short sat_short(float v) {                                           // saturate_cast<short>(float): cvRound (round half to even) + clamp
    const long r = lrintf(v);
    return (short) (r < -32768 ? -32768 : (r > 32767 ? 32767 : r));
}

struct LanczosTab { int* ofs = nullptr; short* coef = nullptr; };

// resize.cpp cv::resize, INTER_LANCZOS4 branch: fx = (float) ((dx + 0.5) * scale - 0.5), sx = cvFloor(fx), fx -= sx (no index clamp for Lanczos)
int lanczos_tables(int h, int w, int H, int W, LanczosTab& out) {
    static std::mutex mu;
    static std::map<std::tuple<int, int, int, int, int>, LanczosTab> cache;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lk(mu);
    auto key = std::make_tuple(dev, h, w, H, W);
    auto it = cache.find(key);
    if (it != cache.end()) { out = it->second; return CSB_OK; }
    std::vector<int> ofs((size_t) W + H);
    std::vector<short> coef(((size_t) W + H) * 8);
    auto axis = [&](int dn, int sn, int* o, short* c) {
        const double scale = 1.0 / ((double) dn / sn);               // scale_x = 1. / inv_scale_x, inv_scale_x = (double) dsize.width / ssize.width
        for (int d = 0; d < dn; ++d) {
            float f = (float) ((d + 0.5) * scale - 0.5);
            const int s0 = (int) floorf(f);
            f -= s0;
            float cb[8];
            lanczos4_coeffs(f, cb);
            o[d] = s0;
            for (int k = 0; k < 8; ++k) c[d * 8 + k] = sat_short(cb[k] * 2048.f);
        }
    };
    axis(W, w, ofs.data(), coef.data());
    axis(H, h, ofs.data() + W, coef.data() + (size_t) W * 8);
    LanczosTab t;
    if (cudaMalloc(&t.ofs, ofs.size() * sizeof(int)) != cudaSuccess || cudaMalloc(&t.coef, coef.size() * sizeof(short)) != cudaSuccess)
        return csb::fail(CSB_ERR_CUDA, "%s: %s", "csb_leres_depth_tail", "cudaMalloc of the Lanczos tables failed");
    if (cudaMemcpy(t.ofs, ofs.data(), ofs.size() * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(t.coef, coef.data(), coef.size() * sizeof(short), cudaMemcpyHostToDevice) != cudaSuccess)
        return csb::fail(CSB_ERR_CUDA, "%s: %s", "csb_leres_depth_tail", "upload of the Lanczos tables failed");
    cache[key] = t;
    out = t;
    return CSB_OK;
}

__global__ void __launch_bounds__(256) k_lt_tofloat(const uint8_t* __restrict__ q, long long n, float* __restrict__ out) {
    for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long) gridDim.x * blockDim.x) out[i] = (float) q[i];
}

// depth[depth == 0] = depth[depth > 0].min()  (kenburns_effect.py:577), per image: positive floats order like their bit patterns
__global__ void __launch_bounds__(256) k_min_positive(const float* __restrict__ x, long long per, unsigned* __restrict__ slot) {
    const int img = blockIdx.y;
    const float* X = x + (long long) img * per;
    unsigned lo = 0x7f800000u;
    for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < per; i += (long long) gridDim.x * blockDim.x) {
        const float v = X[i];
        if (v > 0.f) lo = min(lo, __float_as_uint(v));
    }
    for (int o = 16; o; o >>= 1) lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    if ((threadIdx.x & 31) == 0) atomicMin(slot + img, lo);
}
__global__ void __launch_bounds__(256) k_zero_fill(float* __restrict__ x, long long per, const unsigned* __restrict__ slot) {
    const int img = blockIdx.y;
    float* X = x + (long long) img * per;
    const float m = __uint_as_float(slot[img]);
    for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < per; i += (long long) gridDim.x * blockDim.x)
        if (X[i] == 0.f) X[i] = m;
}

}  // namespace

extern "C" int csb_zero_to_min_positive(float* x, int N, long long per, unsigned* scratch, void* stream) {
    CSB_REQUIRE(x && scratch && N > 0 && per > 0, "bad arguments");
    cudaStream_t st = (cudaStream_t) stream;
    CSB_TRY(csb::cuda_ok(cudaMemsetAsync(scratch, 0x7f, sizeof(unsigned) * N, st), "memset"));      // 0x7f7f7f7f = 3.39e38 (+inf stand-in: no positive pixel)
    csb::memset_done(st);
    int gx = (4 * csb::num_sms() + N - 1) / N;
    gx = gx < 1 ? 1 : gx;
    k_min_positive<<<dim3(gx, N), 256, 0, st>>>(x, per, scratch);
    CSB_TRY(csb::launched("k_min_positive", st));
    k_zero_fill<<<dim3(gx, N), 256, 0, st>>>(x, per, scratch);
    return csb::launched("k_zero_fill", st);
}

extern "C" int csb_leres_depth_tail(const float* logits, int N, int h, int w, int H, int W, unsigned* minmax, uint8_t* q8, float* out, void* stream) {
    CSB_REQUIRE(logits && minmax && q8 && out && N > 0 && h > 0 && w > 0, "bad arguments");
    const bool lanczos = h > H;                                     // kenburns_effect.py:573-574: k = depth.shape[0] / ori_h > 1 -> INTER_LANCZOS4 (both axes)
    CSB_REQUIRE(lanczos || (H >= h && W >= w), "INTER_AREA with a shrinking width but growing height is not built (host path)");
    cudaStream_t st = (cudaStream_t) stream;
    LanczosTab tab;
    if (lanczos) CSB_TRY(lanczos_tables(h, w, H, W, tab));          // built once per geometry (synchronous upload on first use), then cached
    const long long per = (long long) h * w;
    CSB_TRY(csb::cuda_ok(cudaMemsetAsync(minmax, 0, sizeof(unsigned) * 2 * N, st), "memset"));
    csb::memset_done(st);
    // min slots start at 0xffffffff
    CSB_TRY(csb::cuda_ok(cudaMemset2DAsync(minmax, 2 * sizeof(unsigned), 0xff, sizeof(unsigned), N, st), "memset2d"));
    csb::memset_done(st);
    int gx = (2 * csb::num_sms() + N - 1) / N;
    gx = gx < 1 ? 1 : gx;
    k_lt_minmax<<<dim3(gx, N), 256, 0, st>>>(logits, per, minmax);
    CSB_TRY(csb::launched("k_lt_minmax", st));
    k_lt_quant<<<dim3(gx, N), 256, 0, st>>>(logits, per, minmax, q8);
    CSB_TRY(csb::launched("k_lt_quant", st));
    if (lanczos) {
        int gl = (4 * csb::num_sms() + N - 1) / N;
        k_lt_lanczos<<<dim3(gl < 1 ? 1 : gl, N), 256, 0, st>>>(q8, h, w, H, W, tab.ofs, tab.coef, out);
        return csb::launched("k_lt_lanczos", st);
    }
    if (H == h && W == w) {
        k_lt_tofloat<<<csb::wave_grid(per * N, 256, 4), 256, 0, st>>>(q8, per * N, out);
        return csb::launched("k_lt_tofloat", st);
    }
    int gy = (4 * csb::num_sms() + N - 1) / N;
    k_lt_resize<<<dim3(gy < 1 ? 1 : gy, N), 256, 0, st>>>(q8, h, w, H, W, out);
    return csb::launched("k_lt_resize", st);
}
