// raster_blend.cu -- tile ranges, front-to-back alpha blending (forward) and its backward.
//
// Replaces, with identical results, the reference rasterizer's
//   [upstream] rasterizer_impl.cu identifyTileRanges, forward.cu renderCUDA,
//   backward.cu renderCUDA (K7, K8, K9 of section 2.4),
// reached from /root/reference/sings/rec/renderer/gs_renderer_single.py:87-95.
//
// One CTA per 16x16 tile, one thread per pixel.  Each warp owns an 8x4 pixel block (not the
// reference's 16x2 strip) so whole warps fall outside small Gaussians more often.  The
// per-tile Gaussian list is consumed in batches of 256: every thread fetches one 48-byte
// geometry record (three 128-bit loads) into registers while the previous batch is being
// blended, then parks it in a double-buffered shared-memory stage, so the inner loop only
// does broadcast LDS.128 reads and there is one barrier per batch.  Forward exits a tile as
// soon as every pixel has saturated (__syncthreads_count).  The backward reduces the nine
// per-pixel partial gradients of a Gaussian across the warp with shuffles and issues three
// vector reductions (red.global.add.v4.f32) per warp instead of 9 x 32 scalar atomics.
#include "common.cuh"
#include "kernels.h"

namespace sgs {

// [upstream] identifyTileRanges: ranges[tile] = [start, end) in the sorted list (pre-zeroed)
__global__ void tile_ranges_kernel(const unsigned long long* __restrict__ keys, const int* n_ptr,
                                   long long n_cap, uint2* __restrict__ ranges) {
    long long n = min((long long)*n_ptr, n_cap);
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned cur = (unsigned)(keys[i] >> 32);
    if (i == 0) ranges[cur].x = 0;
    else {
        unsigned prev = (unsigned)(keys[i - 1] >> 32);
        if (prev != cur) {
            ranges[prev].y = (unsigned)i;
            ranges[cur].x = (unsigned)i;
        }
    }
    if (i == n - 1) ranges[cur].y = (unsigned)n;
}

static inline const unsigned long long* sorted_keys(const RasterLayout& lay, const char* bin) {
    return reinterpret_cast<const unsigned long long*>(bin + ((lay.passes & 1) ? lay.keys1_off : lay.keys0_off));
}
static inline const unsigned* sorted_vals(const RasterLayout& lay, const char* bin) {
    return reinterpret_cast<const unsigned*>(bin + ((lay.passes & 1) ? lay.vals1_off : lay.vals0_off));
}

int launch_tile_ranges(const RasterLayout& lay, long long L_cap, char* bin, cudaStream_t stream) {
    if (L_cap <= 0) return 0;
    const int* counters = reinterpret_cast<const int*>(bin + lay.cnt_off);
    long long blocks = (L_cap + 255) / 256;
    tile_ranges_kernel<<<(unsigned)blocks, 256, 0, stream>>>(
        sorted_keys(lay, bin), counters + CNT_NUM_RENDERED, L_cap,
        reinterpret_cast<uint2*>(bin + lay.ranges_off));
    SGS_LAUNCH_OK();
    return 0;
}

__device__ __forceinline__ void pixel_of_thread(int tid, int& lx, int& ly) {
    const int warp = tid >> 5, lane = tid & 31;
    lx = ((warp & 1) << 3) + (lane & 7);
    ly = ((warp >> 1) << 2) + (lane >> 3);
}

// ------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TILE_PIX)
blend_fwd_kernel(const uint2* __restrict__ ranges, const unsigned* __restrict__ point_list,
                 const float4* __restrict__ rec, const float* __restrict__ bg, int W, int H,
                 float* __restrict__ out_color, float* __restrict__ final_T,
                 unsigned* __restrict__ n_contrib, float* __restrict__ out_alpha,
                 float* __restrict__ out_depth) {
    __shared__ float4 s_q0[2][TILE_PIX];
    __shared__ float4 s_q1[2][TILE_PIX];
    __shared__ float4 s_q2[2][TILE_PIX];

    const int tid = threadIdx.x;
    int lx, ly;
    pixel_of_thread(tid, lx, ly);
    const int px = blockIdx.x * TILE + lx, py = blockIdx.y * TILE + ly;
    const bool inside = px < W && py < H;
    const float pxf = (float)px, pyf = (float)py;
    const uint2 range = ranges[blockIdx.y * gridDim.x + blockIdx.x];
    const int len = (int)(range.y - range.x);
    const int rounds = (len + TILE_PIX - 1) / TILE_PIX;

    bool done = !inside;
    float T = 1.0f, C0 = 0.0f, C1 = 0.0f, C2 = 0.0f, Dacc = 0.0f;
    unsigned contributor = 0, last = 0;

    float4 p0, p1, p2;
    p0 = p1 = p2 = make_float4(0, 0, 0, 0);
    if (tid < len) {
        const unsigned id = point_list[range.x + tid];
        p0 = rec[3 * (size_t)id]; p1 = rec[3 * (size_t)id + 1]; p2 = rec[3 * (size_t)id + 2];
    }
    for (int r = 0; r < rounds; r++) {
        const int buf = r & 1;
        s_q0[buf][tid] = p0; s_q1[buf][tid] = p1; s_q2[buf][tid] = p2;
        if (__syncthreads_count(done) == TILE_PIX) break;
        const int nxt = (r + 1) * TILE_PIX + tid;
        if (nxt < len) {
            const unsigned id = point_list[range.x + nxt];
            p0 = rec[3 * (size_t)id]; p1 = rec[3 * (size_t)id + 1]; p2 = rec[3 * (size_t)id + 2];
        }
        if (done) continue;
        const int batch = min(TILE_PIX, len - r * TILE_PIX);
        for (int j = 0; j < batch; j++) {
            contributor++;
            const float4 q0 = s_q0[buf][j];     // x, y, -a/2, -b
            const float4 q1 = s_q1[buf][j];     // -c/2, opacity, pmin, r
            const float dx = __fsub_rn(q0.x, pxf), dy = __fsub_rn(q0.y, pyf);
            const float u = __fmul_rn(q0.z, dx), v = __fmul_rn(q1.x, dy), w = __fmul_rn(q0.w, dx);
            const float power = __fmaf_rn(w, dy, __fmaf_rn(v, dy, __fmul_rn(u, dx)));
            if (power > 0.0f || power < q1.z) continue;
            const float alpha = fminf(0.99f, __fmul_rn(q1.y, expneg(power)));
            if (alpha < 1.0f / 255.0f) continue;
            const float test_T = __fmul_rn(T, __fsub_rn(1.0f, alpha));
            if (test_T < 0.0001f) { done = true; break; }
            const float4 q2 = s_q2[buf][j];     // g, b, depth, flags
            const float wgt = __fmul_rn(alpha, T);
            C0 = __fmaf_rn(q1.w, wgt, C0);
            C1 = __fmaf_rn(q2.x, wgt, C1);
            C2 = __fmaf_rn(q2.y, wgt, C2);
            Dacc = __fmaf_rn(q2.z, wgt, Dacc);
            T = test_T;
            last = contributor;
        }
    }
    if (inside) {
        const size_t pix = (size_t)py * W + px, plane = (size_t)W * H;
        final_T[pix] = T;
        n_contrib[pix] = last;
        out_color[pix] = __fmaf_rn(T, bg[0], C0);
        out_color[plane + pix] = __fmaf_rn(T, bg[1], C1);
        out_color[2 * plane + pix] = __fmaf_rn(T, bg[2], C2);
        if (out_alpha) out_alpha[pix] = __fsub_rn(1.0f, T);
        if (out_depth) out_depth[pix] = Dacc;
    }
}

int launch_blend_fwd(const RasterLayout& lay, int W, int H, const char* geom, const char* bin,
                     char* img, const float* bg, float* out_color, float* out_alpha,
                     float* out_depth, cudaStream_t stream) {
    dim3 grid(lay.gx, lay.gy);
    blend_fwd_kernel<<<grid, TILE_PIX, 0, stream>>>(
        reinterpret_cast<const uint2*>(bin + lay.ranges_off), sorted_vals(lay, bin),
        reinterpret_cast<const float4*>(geom + lay.rec_off), bg, W, H, out_color,
        reinterpret_cast<float*>(img + lay.finalT_off),
        reinterpret_cast<unsigned*>(img + lay.ncontrib_off), out_alpha, out_depth);
    SGS_LAUNCH_OK();
    return 0;
}

// ------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TILE_PIX)
blend_bwd_kernel(const uint2* __restrict__ ranges, const unsigned* __restrict__ point_list,
                 const float4* __restrict__ rec, const float* __restrict__ bg, int W, int H,
                 const float* __restrict__ final_T, const unsigned* __restrict__ n_contrib,
                 const float* __restrict__ dL_dpix, float* __restrict__ acc) {
    __shared__ float4 s_q0[2][TILE_PIX];
    __shared__ float4 s_q1[2][TILE_PIX];
    __shared__ float4 s_q2[2][TILE_PIX];
    __shared__ unsigned s_id[2][TILE_PIX];
    __shared__ unsigned s_max[TILE_PIX / 32];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int lx, ly;
    pixel_of_thread(tid, lx, ly);
    const int px = blockIdx.x * TILE + lx, py = blockIdx.y * TILE + ly;
    const bool inside = px < W && py < H;
    const float pxf = (float)px, pyf = (float)py;
    const uint2 range = ranges[blockIdx.y * gridDim.x + blockIdx.x];
    const size_t pix = (size_t)py * W + px, plane = (size_t)W * H;

    const float T_final = inside ? final_T[pix] : 0.0f;
    const unsigned last = inside ? n_contrib[pix] : 0u;
    float T = T_final;
    float dp0 = 0.0f, dp1 = 0.0f, dp2 = 0.0f;
    if (inside) { dp0 = dL_dpix[pix]; dp1 = dL_dpix[plane + pix]; dp2 = dL_dpix[2 * plane + pix]; }
    const float bg_dot = bg[0] * dp0 + bg[1] * dp1 + bg[2] * dp2;
    const float ddelx_dx = 0.5f * (float)W, ddely_dy = 0.5f * (float)H;

    // entries at list positions >= max(n_contrib) of the tile contribute to no pixel
    unsigned m = __reduce_max_sync(0xffffffffu, last);
    if (lane == 0) s_max[warp] = m;
    __syncthreads();
    unsigned len = 0;
#pragma unroll
    for (int w = 0; w < TILE_PIX / 32; w++) len = max(len, s_max[w]);
    if (len == 0) return;
    const int rounds = ((int)len + TILE_PIX - 1) / TILE_PIX;

    float acc_r0 = 0, acc_r1 = 0, acc_r2 = 0;       // accum_rec
    float last_alpha = 0, lc0 = 0, lc1 = 0, lc2 = 0;

    float4 p0, p1, p2;
    unsigned pid = 0;
    p0 = p1 = p2 = make_float4(0, 0, 0, 0);
    if (tid < (int)len) {
        pid = point_list[range.x + (len - 1 - tid)];
        p0 = rec[3 * (size_t)pid]; p1 = rec[3 * (size_t)pid + 1]; p2 = rec[3 * (size_t)pid + 2];
    }
    for (int r = 0; r < rounds; r++) {
        const int buf = r & 1;
        s_q0[buf][tid] = p0; s_q1[buf][tid] = p1; s_q2[buf][tid] = p2; s_id[buf][tid] = pid;
        __syncthreads();
        const int nxt = (r + 1) * TILE_PIX + tid;
        if (nxt < (int)len) {
            pid = point_list[range.x + (len - 1 - nxt)];
            p0 = rec[3 * (size_t)pid]; p1 = rec[3 * (size_t)pid + 1]; p2 = rec[3 * (size_t)pid + 2];
        }
        const int batch = min(TILE_PIX, (int)len - r * TILE_PIX);
        for (int j = 0; j < batch; j++) {
            const unsigned pos = len - 1 - (unsigned)(r * TILE_PIX + j);   // 0-based list position
            const float4 q0 = s_q0[buf][j];
            const float4 q1 = s_q1[buf][j];
            const float dx = __fsub_rn(q0.x, pxf), dy = __fsub_rn(q0.y, pyf);
            const float u = __fmul_rn(q0.z, dx), v = __fmul_rn(q1.x, dy), w = __fmul_rn(q0.w, dx);
            const float power = __fmaf_rn(w, dy, __fmaf_rn(v, dy, __fmul_rn(u, dx)));
            bool valid = pos < last && !(power > 0.0f) && !(power < q1.z);
            float G = 0.0f, alpha = 0.0f;
            if (valid) {
                G = expneg(power);
                alpha = fminf(0.99f, __fmul_rn(q1.y, G));
                valid = !(alpha < 1.0f / 255.0f);
            }
            if (!__any_sync(0xffffffffu, valid)) continue;
            float g_mx = 0, g_my = 0, g_ca = 0, g_cb = 0, g_cc = 0, g_op = 0, g_r = 0, g_g = 0, g_b = 0;
            if (valid) {
                const float4 q2 = s_q2[buf][j];
                T = T / (1.0f - alpha);
                const float dch = alpha * T;
                const float c0 = q1.w, c1 = q2.x, c2 = q2.y;
                acc_r0 = last_alpha * lc0 + (1.0f - last_alpha) * acc_r0;
                acc_r1 = last_alpha * lc1 + (1.0f - last_alpha) * acc_r1;
                acc_r2 = last_alpha * lc2 + (1.0f - last_alpha) * acc_r2;
                lc0 = c0; lc1 = c1; lc2 = c2;
                float dL_dalpha = (c0 - acc_r0) * dp0 + (c1 - acc_r1) * dp1 + (c2 - acc_r2) * dp2;
                g_r = dch * dp0; g_g = dch * dp1; g_b = dch * dp2;
                dL_dalpha *= T;
                last_alpha = alpha;
                dL_dalpha += (-T_final / (1.0f - alpha)) * bg_dot;
                const float dL_dG = q1.y * dL_dalpha;
                const float gdx = G * dx, gdy = G * dy;
                // conic entries: a = -2*q0.z, b = -q0.w, c = -2*q1.x
                const float ca = -2.0f * q0.z, cb = -q0.w, cc = -2.0f * q1.x;
                const float dG_ddelx = -gdx * ca - gdy * cb;
                const float dG_ddely = -gdy * cc - gdx * cb;
                g_mx = dL_dG * dG_ddelx * ddelx_dx;
                g_my = dL_dG * dG_ddely * ddely_dy;
                g_ca = -0.5f * gdx * dx * dL_dG;
                g_cb = -0.5f * gdx * dy * dL_dG;
                g_cc = -0.5f * gdy * dy * dL_dG;
                g_op = G * dL_dalpha;
            }
            g_mx = warp_sum(g_mx); g_my = warp_sum(g_my);
            g_ca = warp_sum(g_ca); g_cb = warp_sum(g_cb); g_cc = warp_sum(g_cc);
            g_op = warp_sum(g_op);
            g_r = warp_sum(g_r); g_g = warp_sum(g_g); g_b = warp_sum(g_b);
            if (lane == 0) {
                float* dst = acc + (size_t)s_id[buf][j] * ACC_FLOATS;
                red_add_f4(dst, g_mx, g_my, g_ca, g_cb);
                red_add_f4(dst + 4, g_cc, g_op, g_r, g_g);
                atomicAdd(dst + 8, g_b);
            }
        }
    }
}

int launch_blend_bwd(const RasterLayout& lay, int W, int H, const char* geom, const char* bin,
                     const char* img, const float* bg, const float* dL_dpix, float* acc,
                     cudaStream_t stream) {
    dim3 grid(lay.gx, lay.gy);
    blend_bwd_kernel<<<grid, TILE_PIX, 0, stream>>>(
        reinterpret_cast<const uint2*>(bin + lay.ranges_off), sorted_vals(lay, bin),
        reinterpret_cast<const float4*>(geom + lay.rec_off), bg, W, H,
        reinterpret_cast<const float*>(img + lay.finalT_off),
        reinterpret_cast<const unsigned*>(img + lay.ncontrib_off), dL_dpix, acc);
    SGS_LAUNCH_OK();
    return 0;
}

}  // namespace sgs
