// hexplane.cu -- multi-scale tri-plane feature interpolation, forward and backward, fused
// (the first half of the canonical attribute decode, the step
// immediately upstream of the deformer).
//
// Replaces  /root/reference/sings/rec/models/modules/hexplane.py
//   :165-166  normalize_aabb            pts_n = (pts - aabb[0]) * (2 / (aabb[1] - aabb[0])) - 1
//   :44-68    grid_sample_wrapper       F.grid_sample(bilinear, align_corners=True, padding_mode='border')
//   :70-105   interpolate_ms_features   per scale: product over the three coordinate planes
//                                       (0,1), (0,2), (1,2); scales concatenated
// as called from sings_hybrid.py:252 over all N Gaussians every step: nine grid_sample launches on
// channel-major planes (a bilinear tap of 32 channels touches 32 different cache lines), their
// products, a concatenation and the matching autograd graph (nine scatter-add backward kernels).
//
// Here the planes are read CHANNEL-LAST ((H, W, C): a tap is one 128-byte line) -- the Python
// wrapper keeps that copy -- and one warp handles one point: lane = channel (C = 32 per round),
// 36 coalesced taps per point, products in registers, one 384-byte row out.  The backward
// recomputes the taps (L2-resident: all nine planes are 33 MB), scatters dL/dplane with one
// 128-byte reduction per tap and warp, and reduces dL/dpts over the lanes.
// HBM/L2-bound: 36 x 128 B gathered per point and direction.
#include "common.cuh"
#include "kernels.h"

namespace sgs {

constexpr int HEX_MAX_PLANES = 12;          // up to 4 scales x 3 planes
constexpr int HEX_WARPS = 8;

struct HexArgs {
    const float* plane[HEX_MAX_PLANES];     // (H, W, C) channel-last
    float* d_plane[HEX_MAX_PLANES];         // backward: same layout, zeroed by the caller
    int W[HEX_MAX_PLANES], H[HEX_MAX_PLANES];
    int S, C;                               // scales, channels per plane (multiple of 32)
    float a0[3], a1[3];                     // aabb[0], aabb[1]
};

// [upstream] grid_sampler_compute_source_index with align_corners=True + clip_coordinates (border):
// returns the clipped pixel coordinate and d(coordinate)/d(normalised input) (0 where clipped)
__device__ __forceinline__ float hex_pix(float v, int size, float& dmul) {
    float x = (v + 1.0f) * 0.5f * (float)(size - 1);
    dmul = 0.5f * (float)(size - 1);
    if (!(x > 0.0f)) { x = 0.0f; dmul = 0.0f; }                        // (NaN goes to 0, like fmin/fmax clamping)
    if (x >= (float)(size - 1)) { x = (float)(size - 1); dmul = 0.0f; }      // ([upstream] clip_coordinates_set_grad: no gradient on the border either)
    return x;
}

struct HexTap {
    int i00, i01, i10, i11;      // element offsets of the four taps' channel 0 (-1: outside, weight unused)
    float w00, w01, w10, w11;    // nw, ne, sw, se
    float fx, fy, dmx, dmy;
};

__device__ __forceinline__ HexTap hex_tap(float u, float v, int W, int H, int C) {
    HexTap t;
    const float ix = hex_pix(u, W, t.dmx), iy = hex_pix(v, H, t.dmy);
    const float x0f = floorf(ix), y0f = floorf(iy);
    const int x0 = (int)x0f, y0 = (int)y0f, x1 = x0 + 1, y1 = y0 + 1;
    t.fx = ix - x0f; t.fy = iy - y0f;
    t.w00 = (1.0f - t.fx) * (1.0f - t.fy); t.w01 = t.fx * (1.0f - t.fy);
    t.w10 = (1.0f - t.fx) * t.fy;          t.w11 = t.fx * t.fy;
    const bool xi = x1 < W, yi = y1 < H;                                  // (x0, y0 are always inside after the clip)
    t.i00 = (y0 * W + x0) * C;
    t.i01 = xi ? (y0 * W + x1) * C : -1;
    t.i10 = yi ? (y1 * W + x0) * C : -1;
    t.i11 = (xi && yi) ? (y1 * W + x1) * C : -1;
    return t;
}

__device__ __forceinline__ void hex_coords(const HexArgs& a, const float* __restrict__ pts, int n, float pn[3]) {
#pragma unroll
    for (int k = 0; k < 3; k++) pn[k] = (pts[3 * (size_t)n + k] - a.a0[k]) * (2.0f / (a.a1[k] - a.a0[k])) - 1.0f;
}

__global__ void __launch_bounds__(HEX_WARPS * 32)
hexplane_fwd_kernel(HexArgs a, int N, const float* __restrict__ pts, float* __restrict__ out) {
    const int lane = threadIdx.x & 31, n = blockIdx.x * HEX_WARPS + (threadIdx.x >> 5);
    if (n >= N) return;
    float pn[3];
    hex_coords(a, pts, n, pn);
    const int F = a.S * a.C;
    for (int s = 0; s < a.S; s++) {
        HexTap t[3];
#pragma unroll
        for (int p = 0; p < 3; p++) {
            const int q = 3 * s + p, cu = p == 2 ? 1 : 0, cv = p == 0 ? 1 : 2;      // planes (0,1), (0,2), (1,2)
            t[p] = hex_tap(pn[cu], pn[cv], a.W[q], a.H[q], a.C);
        }
        for (int c = lane; c < a.C; c += 32) {
            float prod = 1.0f;
#pragma unroll
            for (int p = 0; p < 3; p++) {
                const float* pl = a.plane[3 * s + p] + c;
                float v = t[p].w00 * __ldg(pl + t[p].i00);
                if (t[p].i01 >= 0) v += t[p].w01 * __ldg(pl + t[p].i01);
                if (t[p].i10 >= 0) v += t[p].w10 * __ldg(pl + t[p].i10);
                if (t[p].i11 >= 0) v += t[p].w11 * __ldg(pl + t[p].i11);
                prod *= v;
            }
            out[(size_t)n * F + s * a.C + c] = prod;
        }
    }
}

__global__ void __launch_bounds__(HEX_WARPS * 32)
hexplane_bwd_kernel(HexArgs a, int N, const float* __restrict__ pts, const float* __restrict__ d_out,
                    float* __restrict__ d_pts) {
    const int lane = threadIdx.x & 31, n = blockIdx.x * HEX_WARPS + (threadIdx.x >> 5);
    if (n >= N) return;
    float pn[3];
    hex_coords(a, pts, n, pn);
    const int F = a.S * a.C;
    float g[3] = {0.0f, 0.0f, 0.0f};                 // dL/d(normalised coordinate), this lane's channels
    for (int s = 0; s < a.S; s++) {
        HexTap t[3];
#pragma unroll
        for (int p = 0; p < 3; p++) {
            const int q = 3 * s + p, cu = p == 2 ? 1 : 0, cv = p == 0 ? 1 : 2;
            t[p] = hex_tap(pn[cu], pn[cv], a.W[q], a.H[q], a.C);
        }
        for (int c = lane; c < a.C; c += 32) {
            float v00[3], v01[3], v10[3], v11[3], val[3];
#pragma unroll
            for (int p = 0; p < 3; p++) {
                const float* pl = a.plane[3 * s + p] + c;
                v00[p] = __ldg(pl + t[p].i00);
                v01[p] = t[p].i01 >= 0 ? __ldg(pl + t[p].i01) : 0.0f;
                v10[p] = t[p].i10 >= 0 ? __ldg(pl + t[p].i10) : 0.0f;
                v11[p] = t[p].i11 >= 0 ? __ldg(pl + t[p].i11) : 0.0f;
                val[p] = t[p].w00 * v00[p] + t[p].w01 * v01[p] + t[p].w10 * v10[p] + t[p].w11 * v11[p];
            }
            const float go = d_out[(size_t)n * F + s * a.C + c];
#pragma unroll
            for (int p = 0; p < 3; p++) {
                const float gv = go * val[(p + 1) % 3] * val[(p + 2) % 3];        // dL/d(interpolated value of plane p)
                if (a.d_plane[3 * s + p]) {
                    float* dp = a.d_plane[3 * s + p] + c;
                    atomicAdd(dp + t[p].i00, gv * t[p].w00);
                    if (t[p].i01 >= 0) atomicAdd(dp + t[p].i01, gv * t[p].w01);
                    if (t[p].i10 >= 0) atomicAdd(dp + t[p].i10, gv * t[p].w10);
                    if (t[p].i11 >= 0) atomicAdd(dp + t[p].i11, gv * t[p].w11);
                }
                // [upstream] grid_sampler_2d_backward: gix = sum (+-) value * (y weights), giy likewise
                const float gix = (v01[p] - v00[p]) * (1.0f - t[p].fy) + (v11[p] - v10[p]) * t[p].fy;
                const float giy = (v10[p] - v00[p]) * (1.0f - t[p].fx) + (v11[p] - v01[p]) * t[p].fx;
                const int cu = p == 2 ? 1 : 0, cv = p == 0 ? 1 : 2;
                g[cu] += gv * gix * t[p].dmx;
                g[cv] += gv * giy * t[p].dmy;
            }
        }
    }
    if (d_pts) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const float s = warp_sum(g[k]);
            if (lane == 0) d_pts[3 * (size_t)n + k] = s * (2.0f / (a.a1[k] - a.a0[k]));
        }
    }
}

static int fill_args(HexArgs& a, int S, int C, const int* res, const float* const* planes, float* const* d_planes,
                     const float* aabb) {
    if (S < 1 || 3 * S > HEX_MAX_PLANES || C < 32 || (C & 31) || !res || !planes || !aabb) return SGS_ERR_BAD_ARG;
    a.S = S; a.C = C;
    for (int s = 0; s < S; s++) {
        const int rx = res[3 * s], ry = res[3 * s + 1], rz = res[3 * s + 2];
        if (rx < 1 || ry < 1 || rz < 1) return SGS_ERR_BAD_ARG;
        // plane (i, j) has shape (C, reso[j], reso[i]) in the reference (hexplane.py:33-35): W = reso[i], H = reso[j]
        const int Wd[3] = {rx, rx, ry}, Hd[3] = {ry, rz, rz};
        for (int p = 0; p < 3; p++) {
            const int q = 3 * s + p;
            if (!planes[q]) return SGS_ERR_BAD_ARG;
            a.plane[q] = planes[q];
            a.d_plane[q] = d_planes ? d_planes[q] : nullptr;
            a.W[q] = Wd[p]; a.H[q] = Hd[p];
        }
    }
    for (int k = 0; k < 3; k++) { a.a0[k] = aabb[k]; a.a1[k] = aabb[3 + k]; }
    return 0;
}

int launch_hexplane_fwd(int N, const float* pts, const float* aabb_host, int S, int C, const int* res_host,
                        const float* const* planes_host, float* out, cudaStream_t stream) {
    if (N <= 0) return 0;
    HexArgs a{};
    const int rc = fill_args(a, S, C, res_host, planes_host, nullptr, aabb_host);
    if (rc) return rc;
    hexplane_fwd_kernel<<<(N + HEX_WARPS - 1) / HEX_WARPS, HEX_WARPS * 32, 0, stream>>>(a, N, pts, out);
    SGS_LAUNCH_OK();
    return 0;
}

int launch_hexplane_bwd(int N, const float* pts, const float* aabb_host, int S, int C, const int* res_host,
                        const float* const* planes_host, const float* d_out, float* const* d_planes_host,
                        float* d_pts, cudaStream_t stream) {
    if (N <= 0) return 0;
    HexArgs a{};
    const int rc = fill_args(a, S, C, res_host, planes_host, d_planes_host, aabb_host);
    if (rc) return rc;
    hexplane_bwd_kernel<<<(N + HEX_WARPS - 1) / HEX_WARPS, HEX_WARPS * 32, 0, stream>>>(a, N, pts, d_out, d_pts);
    SGS_LAUNCH_OK();
    return 0;
}

}  // namespace sgs
