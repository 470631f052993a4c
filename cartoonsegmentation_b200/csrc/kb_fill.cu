// fill_disocclusion for sm_100a -- replaces kernel_discfill_updateOutput (anime_3dkenburns/common.py:149-245).
//
// One thread per pixel.  Valid pixels copy their C channels; hole pixels (depth <= 0) march along 16 fixed directions
// both ways until a valid pixel or the border, choose the direction with the shortest from-to span and copy from the
// farther (larger depth) endpoint.  The direction table is normalised exactly as :174-179 does (fp32 sqrt + division,
// evaluated once per thread in registers/constant folding), coordinates are rounded with roundf (half away from zero,
// :189-190) and the span uses powf(int, 2) like the reference (:210) so ties break identically.
//
// Roofline: HBM-bound for the bulk (read (C+1) planes, write C planes = (8C+4) B/px), plus a divergent ray-march tail
// whose depth reads hit L1/L2 (holes are spatially clustered).
#include "kb_fill.cuh"

namespace {

using csbfill::find_fill;

__global__ void __launch_bounds__(256) k_discfill(const float* __restrict__ in, const float* __restrict__ depth, int B, int C, int H, int W,
                                                  float* __restrict__ out) {
    const long long HW = (long long) H * W, total = (long long) B * HW;
    for (long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x; i < total; i += (long long) gridDim.x * blockDim.x) {
        const int b = (int) (i / HW);
        const long long o = i - b * HW;
        const float* Dp = depth + (size_t) b * HW;
        const float* I = in + (size_t) b * C * HW;
        float* O = out + (size_t) b * C * HW;
        long long src = o;
        if (!(__ldg(Dp + o) > 0.0f)) {                                          // :163
            const int x = (int) (o % W), y = (int) (o / W);
            long long f = find_fill(x, y, H, W, [&](int yy, int xx) { return __ldg(Dp + (size_t) yy * W + xx) > 0.0f; },
                                    [&](int yy, int xx) { return __ldg(Dp + (size_t) yy * W + xx); });
            if (f >= 0) src = f;
        }
        for (int c = 0; c < C; ++c) O[(size_t) c * HW + o] = __ldg(I + (size_t) c * HW + src);   // clone :146 + copy :237-239
    }
}

}  // namespace

extern "C" int csb_disocclusion_fill(const float* input, const float* depth, int B, int C, int H, int W, float* output, void* stream) {
    CSB_REQUIRE(input && depth && output, "null pointer");
    CSB_REQUIRE(input != output, "output may not alias input");
    CSB_REQUIRE(B > 0 && C > 0 && H > 0 && W > 0, "bad shape");
    k_discfill<<<csb::wave_grid((long long) B * H * W, 256, 8), 256, 0, (cudaStream_t) stream>>>(input, depth, B, C, H, W, output);
    return csb::launched("k_discfill", (cudaStream_t) stream);
}
