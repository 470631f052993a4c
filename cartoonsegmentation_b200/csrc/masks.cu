// Instance-mask utilities of `AnimeInstances` on the device (reference animeinsseg/anime_instances.py:268-298).
//
//   resize(h, w):     masks.float() -> F.interpolate(mode='area') -> > 0.3 ; bboxes[:, ::2] *= h/oh, bboxes[:, 1::2] *= w/ow (the reference's x/y swap,
//                     harmless for square scaling, kept), torch.round -> int                                              (:268-280)
//   compose_masks():  logical OR over the instances                                                                          (:282-298)
//
// 'area' interpolation is adaptive average pooling: output pixel (y, x) averages the source window rows [floor(y*H0/H), ceil((y+1)*H0/H)) x the same in
// x.  On {0,1} inputs the average is count / size, so the decision `avg > 0.3f` is an integer count compared with a float quotient; both are formed in
// fp32 exactly as torch does (sum in fp32, one division by the window size), so the result is bit-identical to the reference's torch ops.
// HBM-bound: reads K*H0*W0 + writes K*H*W bytes.
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256) k_masks_area_resize(const uint8_t* __restrict__ src, int K, int H0, int W0, uint8_t* __restrict__ dst, int H, int W, float thr) {
    const long long total = (long long) K * H * W;
    for (long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x; i < total; i += (long long) gridDim.x * blockDim.x) {
        const int x = (int) (i % W), y = (int) ((i / W) % H);
        const long long k = i / ((long long) W * H);
        // torch start_index / end_index of adaptive pooling: floor(a * in / out), ceil((a + 1) * in / out) in integer arithmetic
        const int y0 = (int) (((long long) y * H0) / H), y1 = (int) ((((long long) y + 1) * H0 + H - 1) / H);
        const int x0 = (int) (((long long) x * W0) / W), x1 = (int) ((((long long) x + 1) * W0 + W - 1) / W);
        const uint8_t* s = src + k * (long long) H0 * W0;
        int cnt = 0;
        for (int yy = y0; yy < y1; ++yy)
            for (int xx = x0; xx < x1; ++xx) cnt += s[(long long) yy * W0 + xx] != 0;
        const float avg = __fdiv_rn((float) cnt, (float) ((y1 - y0) * (x1 - x0)));
        dst[i] = avg > thr ? 1 : 0;
    }
}

// xywh int boxes: columns 0, 2 (x, w) scale by hs = h / oh, columns 1, 3 (y, h) by ws = w / ow -- as the reference writes it
__global__ void k_boxes_scale_round(const int* __restrict__ in, int K, float hs, float ws, int* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < K * 4) out[i] = (int) rintf(__fmul_rn((float) in[i], (i & 1) ? ws : hs));
}

// 16 pixels per thread (one uint4 per instance): out = OR over k
__global__ void __launch_bounds__(256) k_masks_compose(const uint8_t* __restrict__ masks, int K, long long P, uint8_t* __restrict__ out) {
    const long long nvec = P / 16;
    for (long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x; i < nvec; i += (long long) gridDim.x * blockDim.x) {
        uint4 acc = make_uint4(0, 0, 0, 0);
        for (int k = 0; k < K; ++k) {
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(masks + k * P) + i);
            acc.x |= v.x; acc.y |= v.y; acc.z |= v.z; acc.w |= v.w;
        }
        // bool bytes may hold any non-zero value in principle: normalise every byte to 0 / 1
        auto norm = [](unsigned w) { const unsigned nz = (w | (w >> 1) | (w >> 2) | (w >> 3) | (w >> 4) | (w >> 5) | (w >> 6) | (w >> 7)) & 0x01010101u; return nz; };
        reinterpret_cast<uint4*>(out)[i] = make_uint4(norm(acc.x), norm(acc.y), norm(acc.z), norm(acc.w));
    }
    if (blockIdx.x == 0) {                                                     // tail pixels (P % 16)
        for (long long p = nvec * 16 + threadIdx.x; p < P; p += blockDim.x) {
            uint8_t a = 0;
            for (int k = 0; k < K; ++k) a |= masks[k * P + p];
            out[p] = a ? 1 : 0;
        }
    }
}

// any H*W (instance planes not 16-byte aligned): one pixel per thread
__global__ void __launch_bounds__(256) k_masks_compose_bytes(const uint8_t* __restrict__ masks, int K, long long P, uint8_t* __restrict__ out) {
    for (long long p = blockIdx.x * (long long) blockDim.x + threadIdx.x; p < P; p += (long long) gridDim.x * blockDim.x) {
        uint8_t a = 0;
        for (int k = 0; k < K; ++k) a |= masks[k * P + p];
        out[p] = a ? 1 : 0;
    }
}

}  // namespace

extern "C" int csb_masks_area_resize(const uint8_t* masks, int K, int H0, int W0, uint8_t* out, int H, int W, float thr, const int* boxes_in, int* boxes_out,
                                     void* stream) {
    CSB_REQUIRE(masks && out && K > 0 && H0 > 0 && W0 > 0 && H > 0 && W > 0, "bad arguments");
    cudaStream_t st = (cudaStream_t) stream;
    k_masks_area_resize<<<csb::wave_grid((long long) K * H * W, 256, 8), 256, 0, st>>>(masks, K, H0, W0, out, H, W, thr);
    CSB_TRY(csb::launched("k_masks_area_resize", st));
    if (boxes_in && boxes_out) {
        k_boxes_scale_round<<<(K * 4 + 127) / 128, 128, 0, st>>>(boxes_in, K, (float) ((double) H / (double) H0), (float) ((double) W / (double) W0), boxes_out);
        CSB_TRY(csb::launched("k_boxes_scale_round", st));
    }
    return CSB_OK;
}

extern "C" int csb_masks_compose(const uint8_t* masks, int K, long long P, uint8_t* out, void* stream) {
    CSB_REQUIRE(masks && out && K > 0 && P > 0, "bad arguments");
    if (((uintptr_t) masks & 15) != 0 || ((uintptr_t) out & 15) != 0 || P % 16 != 0) {
        k_masks_compose_bytes<<<csb::wave_grid(P, 256, 8), 256, 0, (cudaStream_t) stream>>>(masks, K, P, out);
        return csb::launched("k_masks_compose", (cudaStream_t) stream);
    }
    k_masks_compose<<<csb::wave_grid(P / 16 + 1, 256, 8), 256, 0, (cudaStream_t) stream>>>(masks, K, P, out);
    return csb::launched("k_masks_compose", (cudaStream_t) stream);
}
