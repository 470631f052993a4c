// Multi-head self-attention with an additive relative-position bias -- the BEiT-L blocks of the ZoeDepth / MiDaS DPT encoder (SURVEY.md §8a row B3;
// restated from timm `beit.py` Attention as patched by MiDaS v3.1 `backbones/beit.py`, which torch.hub loads at the reference's
// `base_models/midas.py:341`; cross-checked against transformers' BeitSelfAttention).
//
//   out[b, q, h*64 + :] = softmax_k( Q[b,h,q,:] . K[b,h,k,:] / 8 + bias[h, q, k] ) . V[b,h,k,:]
//
// qkv is the fused projection output [B, T, 3*heads*64] fp16 (q | k | v, head-major inside each).  T = 577 for the 384 x 384 input: far too short
// for a tcgen05 pipeline to pay (per (b, h) the two GEMMs are 577 x 577 x 64), and the attention is ~6% of the encoder's FLOPs, so this is a
// flash-style warp-level kernel on mma.sync.m16n8k16 (fp16 operands, fp32 accumulate, online softmax in fp32) -- the linear layers around it
// (94% of the FLOPs) run on the tcgen05 conv engine.  One CTA = 64 queries of one (batch, head); 4 warps x 16 rows; K/V stream through a
// double-buffered cp.async ring in 64-key tiles; the bias arrives with rows padded to a multiple of 64 keys, the padding holding -60000 so that
// out-of-range keys vanish in the softmax without a branch.
#include <cuda_fp16.h>

#include "common.cuh"

namespace {

constexpr int kD = 64;            // head dim
constexpr int kBQ = 64;           // queries per CTA
constexpr int kBK = 64;           // keys per tile
constexpr int kLd = 72;           // padded smem row (halves): 144 B rows keep ldmatrix conflict-free

__device__ __forceinline__ void cp16(void* smem, const void* gmem, bool valid) {
    const unsigned s = (unsigned) __cvta_generic_to_shared(smem);
    const int sz = valid ? 16 : 0;                                   // src-size 0 -> zero fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldsm4(unsigned (&r)[4], const __half* p) {
    const unsigned s = (unsigned) __cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(s));
}
__device__ __forceinline__ void ldsm4t(unsigned (&r)[4], const __half* p) {
    const unsigned s = (unsigned) __cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(s));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ unsigned pack_h2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<unsigned*>(&h);
}

// one 64 x 64 tile of K or V rows [t0, t0+64) of (b, head) into smem (rows >= T zero-filled)
__device__ __forceinline__ void load_tile(__half (*dst)[kLd], const __half* __restrict__ base, long long row_stride, int t0, int T) {
    for (int i = threadIdx.x; i < kBK * 8; i += blockDim.x) {
        const int r = i >> 3, c = (i & 7) * 8;
        const bool ok = t0 + r < T;
        cp16(&dst[r][c], base + (long long) (ok ? t0 + r : 0) * row_stride + c, ok);
    }
}

__global__ void __launch_bounds__(128) k_attention(const __half* __restrict__ qkv, int T, int heads, const __half* __restrict__ bias, int Tp,
                                                   float scale, __half* __restrict__ out) {
    __shared__ __align__(16) __half sQ[kBQ][kLd];
    __shared__ __align__(16) __half sK[2][kBK][kLd];
    __shared__ __align__(16) __half sV[2][kBK][kLd];
    const int qb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const long long ld = 3LL * heads * kD;
    const __half* q_base = qkv + (long long) b * T * ld + (long long) h * kD;
    const __half* k_base = q_base + (long long) heads * kD;
    const __half* v_base = k_base + (long long) heads * kD;
    const int q0 = qb * kBQ;
    const int nkb = (T + kBK - 1) / kBK;

    load_tile(sQ, q_base, ld, q0, T);
    load_tile(sK[0], k_base, ld, 0, T);
    load_tile(sV[0], v_base, ld, 0, T);
    cp_commit();
    if (nkb > 1) {
        load_tile(sK[1], k_base, ld, kBK, T);
        load_tile(sV[1], v_base, ld, kBK, T);
    }
    cp_commit();
    cp_wait<1>();
    __syncthreads();

    // Q fragments of this warp's 16 rows: 4 k-steps over d
    unsigned qf[4][4];
    {
        const int m = lane >> 3, rin = lane & 7;
        const int row = warp * 16 + rin + (m & 1) * 8, col = (m >> 1) * 8;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) ldsm4(qf[kk], &sQ[row][kk * 16 + col]);
    }
    float o[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f;
    float mrow[2] = {-INFINITY, -INFINITY}, lrow[2] = {0.f, 0.f};
    const int qr0 = q0 + warp * 16 + g, qr1 = qr0 + 8;                      // the two query rows this thread owns
    const __half* brow0 = bias + ((long long) h * Tp + (qr0 < Tp ? qr0 : 0)) * Tp;
    const __half* brow1 = bias + ((long long) h * Tp + (qr1 < Tp ? qr1 : 0)) * Tp;
    const float sl2 = scale * 1.4426950408889634f;                          // scores are kept in log2 units: exp2 instead of exp

    for (int kb = 0; kb < nkb; ++kb) {
        const int buf = kb & 1;
        float s[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j) s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
        {   // S = Q K^T
            const int m = lane >> 3, rin = lane & 7;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
                for (int jp = 0; jp < 4; ++jp) {                              // two 8-key n-tiles per ldmatrix.x4
                    unsigned bf[4];
                    ldsm4(bf, &sK[buf][jp * 16 + (m >> 1) * 8 + rin][kk * 16 + (m & 1) * 8]);
                    mma16816(s[2 * jp], qf[kk], bf[0], bf[1]);
                    mma16816(s[2 * jp + 1], qf[kk], bf[2], bf[3]);
                }
            }
        }
        // + bias, running max
        float mx0 = mrow[0], mx1 = mrow[1];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int kc = kb * kBK + j * 8 + 2 * t;
            const float2 b0 = __half22float2(*reinterpret_cast<const __half2*>(brow0 + kc));
            const float2 b1 = __half22float2(*reinterpret_cast<const __half2*>(brow1 + kc));
            s[j][0] = fmaf(s[j][0], sl2, b0.x * 1.4426950408889634f);
            s[j][1] = fmaf(s[j][1], sl2, b0.y * 1.4426950408889634f);
            s[j][2] = fmaf(s[j][2], sl2, b1.x * 1.4426950408889634f);
            s[j][3] = fmaf(s[j][3], sl2, b1.y * 1.4426950408889634f);
            mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
            mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float c0 = exp2f(mrow[0] - mx0), c1 = exp2f(mrow[1] - mx1);    // first tile: exp2(-inf) = 0
        mrow[0] = mx0;
        mrow[1] = mx1;
        float sum0 = 0.f, sum1 = 0.f;
        unsigned pf[4][4];                                                   // P as A fragments: 4 k-steps of 16 keys
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float p0 = exp2f(s[j][0] - mx0), p1 = exp2f(s[j][1] - mx0), p2 = exp2f(s[j][2] - mx1), p3 = exp2f(s[j][3] - mx1);
            sum0 += p0 + p1;
            sum1 += p2 + p3;
            pf[j >> 1][(j & 1) * 2] = pack_h2(p0, p1);                        // a0 / a2: row g
            pf[j >> 1][(j & 1) * 2 + 1] = pack_h2(p2, p3);                    // a1 / a3: row g + 8
        }
        lrow[0] = lrow[0] * c0 + sum0;
        lrow[1] = lrow[1] * c1 + sum1;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            o[j][0] *= c0; o[j][1] *= c0; o[j][2] *= c1; o[j][3] *= c1;
        }
        {   // O += P V
            const int m = lane >> 3, rin = lane & 7;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
                for (int jp = 0; jp < 4; ++jp) {                              // two 8-wide d n-tiles per ldmatrix.x4.trans
                    unsigned bf[4];
                    ldsm4t(bf, &sV[buf][kk * 16 + (m & 1) * 8 + rin][jp * 16 + (m >> 1) * 8]);
                    mma16816(o[2 * jp], pf[kk], bf[0], bf[1]);
                    mma16816(o[2 * jp + 1], pf[kk], bf[2], bf[3]);
                }
            }
        }
        __syncthreads();                                                      // everyone is done with buffer `buf`
        if (kb + 2 < nkb) {
            load_tile(sK[buf], k_base, ld, (kb + 2) * kBK, T);
            load_tile(sV[buf], v_base, ld, (kb + 2) * kBK, T);
        }
        cp_commit();
        cp_wait<1>();                                                         // tile kb+1 has landed
        __syncthreads();
    }
    // row sums live in the 4 lanes of a quad
    lrow[0] += __shfl_xor_sync(0xffffffffu, lrow[0], 1);
    lrow[0] += __shfl_xor_sync(0xffffffffu, lrow[0], 2);
    lrow[1] += __shfl_xor_sync(0xffffffffu, lrow[1], 1);
    lrow[1] += __shfl_xor_sync(0xffffffffu, lrow[1], 2);
    const float i0 = 1.0f / lrow[0], i1 = 1.0f / lrow[1];
    const long long old = (long long) heads * kD;
    __half* o0 = out + ((long long) b * T + qr0) * old + h * kD + 2 * t;
    __half* o1 = out + ((long long) b * T + qr1) * old + h * kD + 2 * t;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        if (qr0 < T) *reinterpret_cast<__half2*>(o0 + j * 8) = __floats2half2_rn(o[j][0] * i0, o[j][1] * i0);
        if (qr1 < T) *reinterpret_cast<__half2*>(o1 + j * 8) = __floats2half2_rn(o[j][2] * i1, o[j][3] * i1);
    }
}

// ---- token plumbing of the DPT encoder -------------------------------------------------------------------------------------------------------
// tokens[b, 0] = cls; tokens[b, 1 + p] = patches[b, p]     (BeitEmbeddings: torch.cat((cls_tokens, embeddings), 1))
__global__ void k_tokens_assemble(const __half* __restrict__ patches, const __half* __restrict__ cls, int B, int P, int C, __half* __restrict__ tokens) {
    const long long n8 = (long long) B * (P + 1) * (C / 8);
    for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long) gridDim.x * blockDim.x) {
        const int c8 = (int) (i % (C / 8));
        const long long row = i / (C / 8);
        const int tk = (int) (row % (P + 1));
        const long long b = row / (P + 1);
        const uint4 v = tk == 0 ? reinterpret_cast<const uint4*>(cls)[c8] : reinterpret_cast<const uint4*>(patches)[(b * P + tk - 1) * (C / 8) + c8];
        reinterpret_cast<uint4*>(tokens)[i] = v;
    }
}

// DPT 'project' readout input: out[b, p] = [tokens[b, 1 + p] | tokens[b, 0]]   (ZoeDepthReassembleStage / MiDaS ProjectReadout)
__global__ void k_readout_concat(const __half* __restrict__ tokens, int B, int P, int C, __half* __restrict__ out) {
    const long long n8 = (long long) B * P * (2 * C / 8);
    for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long) gridDim.x * blockDim.x) {
        const int c8 = (int) (i % (2 * C / 8));
        const long long row = i / (2 * C / 8);
        const int p = (int) (row % P);
        const long long b = row / P;
        const long long src_row = b * (P + 1) + (c8 < C / 8 ? 1 + p : 0);
        reinterpret_cast<uint4*>(out)[i] = reinterpret_cast<const uint4*>(tokens)[src_row * (C / 8) + (c8 < C / 8 ? c8 : c8 - C / 8)];
    }
}

// ConvTranspose2d(C, C, kernel=k, stride=k) tail: x [B,h,w,k*k*C] (channel = (i*k + j)*C + c) -> y [B,h*k,w*k,C]
__global__ void k_pixel_shuffle(const __half* __restrict__ x, int B, int h, int w, int k, int C, __half* __restrict__ y) {
    const long long n8 = (long long) B * h * k * w * k * (C / 8);
    for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long) gridDim.x * blockDim.x) {
        const int c8 = (int) (i % (C / 8));
        long long r = i / (C / 8);
        const int X = (int) (r % (w * k)); r /= (w * k);
        const int Y = (int) (r % (h * k));
        const long long b = r / (h * k);
        const int yy = Y / k, ii = Y % k, xx = X / k, jj = X % k;
        reinterpret_cast<uint4*>(y)[i] = reinterpret_cast<const uint4*>(x)[(((b * h + yy) * w + xx) * (k * k) + ii * k + jj) * (C / 8) + c8];
    }
}

}  // namespace

extern "C" int csb_attention_bias(const void* qkv, int B, int T, int heads, int head_dim, const void* bias, int Tp, float scale, void* out, void* stream) {
    CSB_REQUIRE(qkv && bias && out, "null pointer");
    CSB_REQUIRE(B > 0 && T > 0 && heads > 0 && head_dim == kD, "head_dim must be 64");
    CSB_REQUIRE(Tp % kBK == 0 && Tp >= T, "the bias must be padded to a multiple of 64 keys (and as many rows)");
    CSB_REQUIRE((((uintptr_t) qkv | (uintptr_t) out | (uintptr_t) bias) & 15) == 0, "pointers must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t) stream;
    const dim3 grid((T + kBQ - 1) / kBQ, heads, B);
    k_attention<<<grid, 128, 0, st>>>((const __half*) qkv, T, heads, (const __half*) bias, Tp, scale, (__half*) out);
    return csb::launched("k_attention", st);
}

extern "C" int csb_tokens_assemble(const void* patches, const void* cls, int B, int P, int C, void* tokens, void* stream) {
    CSB_REQUIRE(patches && cls && tokens && B > 0 && P > 0 && C % 8 == 0, "bad arguments");
    cudaStream_t st = (cudaStream_t) stream;
    k_tokens_assemble<<<csb::wave_grid((long long) B * (P + 1) * (C / 8), 256, 4), 256, 0, st>>>((const __half*) patches, (const __half*) cls, B, P, C, (__half*) tokens);
    return csb::launched("k_tokens_assemble", st);
}

extern "C" int csb_readout_concat(const void* tokens, int B, int P, int C, void* out, void* stream) {
    CSB_REQUIRE(tokens && out && B > 0 && P > 0 && C % 8 == 0, "bad arguments");
    cudaStream_t st = (cudaStream_t) stream;
    k_readout_concat<<<csb::wave_grid((long long) B * P * (2 * C / 8), 256, 4), 256, 0, st>>>((const __half*) tokens, B, P, C, (__half*) out);
    return csb::launched("k_readout_concat", st);
}

extern "C" int csb_pixel_shuffle_nhwc(const void* x, int B, int h, int w, int k, int C, void* y, void* stream) {
    CSB_REQUIRE(x && y && B > 0 && h > 0 && w > 0 && k > 0 && C % 8 == 0, "bad arguments");
    cudaStream_t st = (cudaStream_t) stream;
    k_pixel_shuffle<<<csb::wave_grid((long long) B * h * k * w * k * (C / 8), 256, 4), 256, 0, st>>>((const __half*) x, B, h, w, k, C, (__half*) y);
    return csb::launched("k_pixel_shuffle", st);
}
