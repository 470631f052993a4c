// Point-cloud growth after an inpaint pass (reference anime_3dkenburns/kenburns_effect.py:462-512): the pixels that were holes in
// the shifted view (tenExisting == 0) are appended to the cloud's channel-planar arrays
//     inpainted_img [1,3,N], tenInpaDisparity [1,1,N], tenInpaDepth [1,1,N], tenInpaPoints [1,3,N]      N -> N + holes
// The reference does this with boolean-mask gathers + torch.cat (about 20 ATen launches and a hidden host sync per array).  Here: an order-preserving
// stream compaction -- per-block hole counts, one single-block scan, then ONE kernel that copies the old planes and scatters the selected pixels
// (in-block ranks by warp ballots) for all 8 planes.  Element order equals the boolean-mask gather's (ascending pixel index), so results are
// bit-identical.  HBM-bound: reads 4 P (flags) + 4 * planes * (N + P), writes 4 * planes * (N + holes).
#include "common.cuh"

namespace {

constexpr int kBlock = 1024;          // pixels per block (one thread per pixel)
constexpr int kMaxPlanes = 8;

struct PlaneSet { const float* src[kMaxPlanes]; const float* old_[kMaxPlanes]; float* dst[kMaxPlanes]; };

__global__ void __launch_bounds__(kBlock) k_compact_count(const float* __restrict__ existing, long long P, int* __restrict__ block_counts) {
    const long long i = (long long) blockIdx.x * kBlock + threadIdx.x;
    const int flag = (i < P && existing[i] == 0.0f) ? 1 : 0;
    const int c = __syncthreads_count(flag);
    if (threadIdx.x == 0) block_counts[blockIdx.x] = c;
}

// exclusive scan of nb block counts in one block; total -> offsets[nb]
__global__ void __launch_bounds__(1024) k_compact_scan(const int* __restrict__ counts, int nb, int* __restrict__ offsets) {
    __shared__ int warp_sums[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < nb; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < nb ? counts[i] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_sums[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int w = warp_sums[lane], wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += t;
            }
            warp_sums[lane] = wi - w;                                  // exclusive prefix of the warp sums
        }
        __syncthreads();
        const int excl = carry + warp_sums[warp] + incl - v;
        if (i < nb) offsets[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) offsets[nb] = carry;
}

__global__ void __launch_bounds__(kBlock) k_compact_append(const float* __restrict__ existing, long long P, const int* __restrict__ offsets, PlaneSet ps, int planes,
                                                           long long n_old, int copy_blocks) {
    if ((int) blockIdx.x >= (int) gridDim.x - copy_blocks) {           // the trailing blocks copy the old planes (grid-stride)
        const long long start = ((long long) (blockIdx.x - (gridDim.x - copy_blocks)) * kBlock + threadIdx.x);
        for (int c = 0; c < planes; ++c)
            for (long long i = start; i < n_old; i += (long long) copy_blocks * kBlock) ps.dst[c][i] = ps.old_[c][i];
        return;
    }
    __shared__ int warp_off[32];
    const long long i = (long long) blockIdx.x * kBlock + threadIdx.x;
    const bool flag = i < P && existing[i] == 0.0f;
    const unsigned ballot = __ballot_sync(0xffffffffu, flag);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) warp_off[warp] = __popc(ballot);
    __syncthreads();
    if (warp == 0) {
        const int w = warp_off[lane];
        int incl = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        warp_off[lane] = incl - w;
    }
    __syncthreads();
    if (flag) {
        const long long pos = n_old + offsets[blockIdx.x] + warp_off[warp] + __popc(ballot & ((1u << lane) - 1u));
        for (int c = 0; c < planes; ++c) ps.dst[c][pos] = ps.src[c][i];
    }
}

}  // namespace

extern "C" long long csb_cloud_append_scratch_ints(long long P) { return 2 * ((P + kBlock - 1) / kBlock) + 2; }

// Phase 1: count.  scratch[0 .. nb) block counts, scratch[nb .. 2 nb] exclusive offsets and, last, the total; `total_out` (device int) also receives it.
extern "C" int csb_cloud_append_count(const float* existing, long long P, int* scratch, int* total_out, void* stream) {
    CSB_REQUIRE(existing && scratch && total_out && P > 0, "bad arguments");
    cudaStream_t st = (cudaStream_t) stream;
    const int nb = (int) ((P + kBlock - 1) / kBlock);
    k_compact_count<<<nb, kBlock, 0, st>>>(existing, P, scratch);
    CSB_TRY(csb::launched("k_compact_count", st));
    k_compact_scan<<<1, 1024, 0, st>>>(scratch, nb, scratch + nb);
    CSB_TRY(csb::launched("k_compact_scan", st));
    return csb::cuda_ok(cudaMemcpyAsync(total_out, scratch + 2 * nb, sizeof(int), cudaMemcpyDeviceToDevice, st), "copy total");
}

// Phase 2: dst[c] = concat(old[c][0 .. n_old), src[c][existing == 0]) for `planes` planes (<= 8); dst planes hold n_old + total floats each.
extern "C" int csb_cloud_append(const float* existing, long long P, const int* scratch, const float* const* src, const float* const* old_, float* const* dst,
                                int planes, long long n_old, void* stream) {
    CSB_REQUIRE(existing && scratch && src && old_ && dst && planes > 0 && planes <= kMaxPlanes && P > 0 && n_old >= 0, "bad arguments");
    PlaneSet ps{};
    for (int c = 0; c < planes; ++c) { ps.src[c] = src[c]; ps.old_[c] = old_[c]; ps.dst[c] = dst[c]; }
    const int nb = (int) ((P + kBlock - 1) / kBlock);
    const int copy_blocks = n_old > 0 ? 2 * csb::num_sms() : 0;
    k_compact_append<<<nb + copy_blocks, kBlock, 0, (cudaStream_t) stream>>>(existing, P, scratch + nb, ps, planes, n_old, copy_blocks);
    return csb::launched("k_compact_append", (cudaStream_t) stream);
}
