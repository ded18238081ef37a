// densify.cu -- densification statistics of one rendered view, fused into one kernel.
//
// Replaces the torch ops of
//   /root/reference/sings/rec/models/sings_hybrid.py:1013-1015 (add_densification_stats):
//       xyz_gradient_accum[vis] += ||viewspace_points.grad[vis, :2]||_2 ;  denom[vis] += 1
//   /root/reference/sings/rec/trainer/gs_trainer.py:487-490:
//       max_radii2D[vis] = max(max_radii2D[vis], radii[vis])          with vis = radii > 0
// These three per-Gaussian arrays are the data-parallel all-reduce payload besides the
// parameter gradients (SUM for the first two, MAX for the third).  The norm is
// taken per view BEFORE any cross-replica sum, as the reference does per step.
#include "common.cuh"
#include "kernels.h"

namespace sgs {

__global__ void densify_stats_kernel(int P, const float* __restrict__ grad2d, const int* __restrict__ radii,
                                     float* __restrict__ accum, float* __restrict__ denom,
                                     float* __restrict__ max_radii) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const int r = radii[i];
    if (r <= 0) return;
    const float gx = grad2d[3 * (size_t)i], gy = grad2d[3 * (size_t)i + 1];
    accum[i] += sqrtf(gx * gx + gy * gy);
    denom[i] += 1.0f;
    max_radii[i] = fmaxf(max_radii[i], (float)r);
}

int launch_densify_stats(int P, const float* grad2d, const int* radii, float* accum, float* denom,
                         float* max_radii, cudaStream_t stream) {
    if (P <= 0) return 0;
    densify_stats_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, grad2d, radii, accum, denom, max_radii);
    SGS_LAUNCH_OK();
    return 0;
}

// Up-front clearing of a frame's state (sgs_raster_clear): up to three regions zeroed by ONE
// kernel -- one graph node with an overlapped launch instead of three serialised memset nodes.
// Regions are 16-byte aligned; sizes are rounded up to 16 bytes inside their allocations'
// 256-byte granularity (the caller guarantees it) except the last words, handled bytewise.
__global__ void __launch_bounds__(256) clear3_kernel(char* a, size_t na, char* b, size_t nb, char* c, size_t nc) {
    pdl_sync();
    const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    char* const ptr[3] = {a, b, c};
    const size_t len[3] = {na, nb, nc};
#pragma unroll
    for (int r = 0; r < 3; r++) {
        const unsigned n16 = (unsigned)(len[r] / 16);           // regions are far below 64 GB
        uint4* __restrict__ p = reinterpret_cast<uint4*>(ptr[r]);
#pragma unroll 4
        for (unsigned i = tid; i < n16; i += stride) p[i] = z;
        for (size_t i = (size_t)n16 * 16 + tid; i < len[r]; i += stride) ptr[r][i] = 0;
    }
}

int launch_clear3(void* a, size_t na, void* b, size_t nb, void* c, size_t nc, cudaStream_t stream) {
    const size_t total = na + nb + nc;
    if (total == 0) return 0;
    long long blocks = (long long)((total / 16 + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    launch_pdl(clear3_kernel, (unsigned)blocks, 256, 0, stream, (char*)a, na, (char*)b, nb, (char*)c, nc);
    SGS_LAUNCH_OK();
    return 0;
}

// Data-parallel step epilogue: after the all-reduce, fold this step's statistics
// (SUM-reduced accum / denom increments, MAX-reduced radii) into the persistent accumulators and
// clear the step buffers for the next view -- one launch instead of three adds and three fills.
__global__ void fold_stats_kernel(int P, float* __restrict__ step_accum, float* __restrict__ step_denom,
                                  float* __restrict__ step_max_radii, float* __restrict__ accum,
                                  float* __restrict__ denom, float* __restrict__ max_radii) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    accum[i] += step_accum[i];
    denom[i] += step_denom[i];
    max_radii[i] = fmaxf(max_radii[i], step_max_radii[i]);
    step_accum[i] = 0.0f;
    step_denom[i] = 0.0f;
    step_max_radii[i] = 0.0f;
}

int launch_fold_stats(int P, float* step_accum, float* step_denom, float* step_max_radii, float* accum,
                      float* denom, float* max_radii, cudaStream_t stream) {
    if (P <= 0) return 0;
    fold_stats_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, step_accum, step_denom, step_max_radii, accum, denom, max_radii);
    SGS_LAUNCH_OK();
    return 0;
}

}  // namespace sgs
