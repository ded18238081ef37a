// image_loss.cu -- the image loss of a training step, fused: L1 + SSIM of the rendered image
// against the masked ground truth, forward and backward (the step
// immediately downstream of the rasterizer).
//
// Replaces (same formulas, same constants)
//   /root/reference/sings/rec/losses/loss.py:57-70       HumanLoss.forward: gt' = gt m + bg (1 - m);
//        l1 = sum |pred - gt'| / sum m;  ssim term = (1 - mean ssim_map) * (sum m / (H W))
//   /root/reference/sings/rec/losses/utils.py:16-20      l1_loss
//   /root/reference/sings/rec/losses/utils.py:27-70      gaussian / create_window / ssim / _ssim:
//        11x11 window, sigma 1.5, zero padding, C1 = 0.01^2, C2 = 0.03^2
// which run as five depthwise conv2d launches plus ~25 elementwise kernels and their autograd
// graph (and, before that, a host->device copy of a float32 ground truth: here the ground truth
// may also arrive as the dataset's uint8 HWC image, a quarter of the bytes).
//
// Forward: one CTA per 32x32 pixels and channel.  The 42x42 halo of pred and gt' goes to shared
// memory once; the window is separable, so the five moments (E x, E y, E xx, E yy, E xy) take a
// horizontal and a vertical 11-tap pass.  Per pixel it keeps the three derivatives of the SSIM
// map the backward needs (d/d mu1 with the sigma terms folded in, d/d E xx, d/d E xy) and the
// composited target; the three sums go to double accumulators.
// Backward: dL/dpred = w_l1 sign(pred - gt') / sum m
//                      - w_ssim (sum m / HW) / (3 HW) * [ conv(D mu1) + 2 pred conv(D Exx) + gt' conv(D Exy) ]
// -- the same separable pass over the three stored maps (the window is symmetric).
// HBM-bound: ~64 B per pixel and channel, forward + backward.
#include "common.cuh"
#include "kernels.h"


namespace sgs {

constexpr int LW = 11, LR = 5;                // window, radius
constexpr int LT = 32, LTH = LT + 2 * LR;     // output tile (32 x 32), tile + halo (42)
constexpr int LSTR = LTH + 2;                 // row stride of the input tile (even: a float4 holds two pixels' pairs)
constexpr int LTHREADS = 256;
// utils.py:27-29: exp(-(x - 5)^2 / (2 * 1.5^2)) as float32, divided by their float32 sum
__device__ constexpr float c_gauss[LW] = {1.028380124e-03f, 7.598758209e-03f, 3.600077331e-02f, 1.093606874e-01f,
                                          2.130055279e-01f, 2.660117149e-01f, 2.130055279e-01f, 1.093606874e-01f,
                                          3.600077331e-02f, 7.598758209e-03f, 1.028380124e-03f};

// Both kernels: the window is separable.  Horizontal pass: a thread takes four consecutive outputs of a
// row -- 14 inputs, seven LDS.128 -- vertical pass: four consecutive outputs of a column; the moments
// travel in packed pairs ((x, y), (xx, yy) + xy alone: three FMA-class instructions per tap instead
// of five).  The 32x32 tile keeps the halo at 1.7x.

template <bool GT_U8>
__global__ void __launch_bounds__(LTHREADS)
image_loss_fwd_kernel(int H, int W, const float* __restrict__ pred, const void* __restrict__ gt,
                      const float* __restrict__ mask, const float* __restrict__ bg,
                      float* __restrict__ part, float* __restrict__ gtc, double* __restrict__ sums) {
    __shared__ __align__(16) float2 s_in[LTH][LSTR];            // (pred, target)
    __shared__ float2 s_h01[LTH][LT + 1], s_h23[LTH][LT + 1];   // horizontal sums of (x, y), (xx, yy)
    __shared__ float s_h4[LTH][LT + 1];                         // ... of xy
    __shared__ float s_red[3][LTHREADS / 32];
    const int tid = threadIdx.x, c = blockIdx.z;
    const int x0 = (int)blockIdx.x * LT - LR, y0 = (int)blockIdx.y * LT - LR;
    const size_t plane = (size_t)H * W;
    const float bgc = bg[c];
    for (int i = tid; i < LTH * LSTR; i += LTHREADS) {
        const int ly = i / LSTR, lx = i - ly * LSTR, gx = x0 + lx, gy = y0 + ly;
        float x = 0.0f, y = 0.0f;
        if (lx < LTH && gx >= 0 && gx < W && gy >= 0 && gy < H) {        // (conv2d pads with zeros)
            const size_t p = (size_t)gy * W + gx;
            x = pred[c * plane + p];
            const float m = mask ? mask[p] : 1.0f;
            const float g = GT_U8 ? (float)reinterpret_cast<const unsigned char*>(gt)[p * 3 + c] / 255.0f
                                  : reinterpret_cast<const float*>(gt)[c * plane + p];
            y = g * m + bgc * (1.0f - m);
        }
        s_in[ly][lx] = make_float2(x, y);
    }
    __syncthreads();
    for (int o = tid; o < LTH * (LT / 4); o += LTHREADS) {      // horizontal pass: row r, outputs 4 q .. 4 q + 3
        const int r = o / (LT / 4), q = o - r * (LT / 4);
        float2 v[14], vv[14];
        float vxy[14];
        const float4* src = reinterpret_cast<const float4*>(&s_in[r][4 * q]);
#pragma unroll
        for (int j = 0; j < 7; j++) {
            const float4 t = src[j];
            v[2 * j] = make_float2(t.x, t.y); v[2 * j + 1] = make_float2(t.z, t.w);
        }
#pragma unroll
        for (int j = 0; j < 14; j++) { vv[j] = fmul2(v[j], v[j]); vxy[j] = v[j].x * v[j].y; }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            float2 a01 = splat2(0.0f), a23 = splat2(0.0f);
            float a4 = 0.0f;
#pragma unroll
            for (int k = 0; k < LW; k++) {
                const float2 w = splat2(c_gauss[k]);
                a01 = ffma2(w, v[u + k], a01); a23 = ffma2(w, vv[u + k], a23); a4 = fmaf(c_gauss[k], vxy[u + k], a4);
            }
            s_h01[r][4 * q + u] = a01; s_h23[r][4 * q + u] = a23; s_h4[r][4 * q + u] = a4;
        }
    }
    __syncthreads();
    // vertical pass: column tx, outputs rows 4 ty .. 4 ty + 3
    const int tx = tid & (LT - 1), ty = tid / LT;
    float2 m01[4], m23[4];
    float m4[4];
#pragma unroll
    for (int u = 0; u < 4; u++) { m01[u] = splat2(0.0f); m23[u] = splat2(0.0f); m4[u] = 0.0f; }
#pragma unroll
    for (int j = 0; j < 14; j++) {
        const float2 h01 = s_h01[4 * ty + j][tx], h23 = s_h23[4 * ty + j][tx];
        const float h4 = s_h4[4 * ty + j][tx];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int k = j - u;
            if (k >= 0 && k < LW) {
                const float2 w = splat2(c_gauss[k]);
                m01[u] = ffma2(w, h01, m01[u]); m23[u] = ffma2(w, h23, m23[u]); m4[u] = fmaf(c_gauss[k], h4, m4[u]);
            }
        }
    }
    float l1 = 0.0f, sm = 0.0f, ms = 0.0f;
    const int px = (int)blockIdx.x * LT + tx;
#pragma unroll
    for (int u = 0; u < 4; u++) {
        const int py = (int)blockIdx.y * LT + 4 * ty + u;
        if (px < W && py < H) {
            const float mu1 = m01[u].x, mu2 = m01[u].y, exx = m23[u].x, eyy = m23[u].y, exy = m4[u];
            const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
            const float s1 = exx - mu1 * mu1, s2 = eyy - mu2 * mu2, s12 = exy - mu1 * mu2;
            const float A1 = 2.0f * mu1 * mu2 + C1, A2 = 2.0f * s12 + C2;
            const float B1 = mu1 * mu1 + mu2 * mu2 + C1, B2 = s1 + s2 + C2;
            const float iB = 1.0f / (B1 * B2);
            const float map = A1 * A2 * iB;
            sm += map;
            // derivatives of the map: explicit in mu1, through sigma1^2 = Exx - mu1^2, through sigma12 = Exy - mu1 mu2
            const float d_s1 = -map / B2;                         // d map / d sigma1^2
            const float d_s12 = 2.0f * A1 * iB;                   // d map / d sigma12
            const float d_mu1 = 2.0f * mu2 * A2 * iB - map * 2.0f * mu1 / B1 - 2.0f * mu1 * d_s1 - mu2 * d_s12;
            const size_t p = (size_t)py * W + px, q = c * plane + p;
            part[q] = d_mu1; part[3 * plane + q] = d_s1; part[6 * plane + q] = d_s12;
            const float2 xy = s_in[4 * ty + u + LR][tx + LR];
            gtc[q] = xy.y;
            l1 += fabsf(xy.x - xy.y);
            if (c == 0) ms += mask ? mask[p] : 1.0f;
        }
    }
    l1 = warp_sum(l1); sm = warp_sum(sm); ms = warp_sum(ms);
    if ((tid & 31) == 0) { s_red[0][tid >> 5] = l1; s_red[1][tid >> 5] = sm; s_red[2][tid >> 5] = ms; }
    __syncthreads();
    if (tid < 3) {
        float v = 0.0f;
#pragma unroll
        for (int w = 0; w < LTHREADS / 32; w++) v += s_red[tid][w];
        if (v != 0.0f) atomicAdd(&sums[tid], (double)v);
    }
}

__global__ void __launch_bounds__(LTHREADS)
image_loss_bwd_kernel(int H, int W, const float* __restrict__ pred, const float* __restrict__ gtc,
                      const float* __restrict__ part, const double* __restrict__ sums,
                      float w_l1, float w_ssim, const float* __restrict__ dloss, float* __restrict__ dL_dpred,
                      float* __restrict__ loss_out) {
    __shared__ __align__(16) float2 s_ab[LTH][LSTR];            // (D mu1, D Exx)
    __shared__ __align__(16) float s_c[LTH][LSTR + 2];          // D Exy (row stride 46: rows stay 8-byte aligned)
    __shared__ float2 s_hab[LTH][LT + 1];
    __shared__ float s_hc[LTH][LT + 1];
    const int tid = threadIdx.x, c = blockIdx.z;
    const int x0 = (int)blockIdx.x * LT - LR, y0 = (int)blockIdx.y * LT - LR;
    const size_t plane = (size_t)H * W;
    if (loss_out && tid == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) {
        const double hw = (double)H * (double)W;
        *loss_out = (float)((double)w_l1 * sums[0] / sums[2] + (double)w_ssim * (1.0 - sums[1] / (3.0 * hw)) * (sums[2] / hw));
    }
    for (int i = tid; i < LTH * LSTR; i += LTHREADS) {
        const int ly = i / LSTR, lx = i - ly * LSTR, gx = x0 + lx, gy = y0 + ly;
        float a = 0.0f, b = 0.0f, d = 0.0f;
        if (lx < LTH && gx >= 0 && gx < W && gy >= 0 && gy < H) {
            const size_t q = c * plane + (size_t)gy * W + gx;
            a = part[q]; b = part[3 * plane + q]; d = part[6 * plane + q];
        }
        s_ab[ly][lx] = make_float2(a, b); s_c[ly][lx] = d;
    }
    __syncthreads();
    for (int o = tid; o < LTH * (LT / 4); o += LTHREADS) {
        const int r = o / (LT / 4), q = o - r * (LT / 4);
        float2 v[14];
        float vc[14];
        const float4* src = reinterpret_cast<const float4*>(&s_ab[r][4 * q]);
        const float2* srcc = reinterpret_cast<const float2*>(&s_c[r][4 * q]);
#pragma unroll
        for (int j = 0; j < 7; j++) {
            const float4 t = src[j];
            v[2 * j] = make_float2(t.x, t.y); v[2 * j + 1] = make_float2(t.z, t.w);
            const float2 tc = srcc[j];
            vc[2 * j] = tc.x; vc[2 * j + 1] = tc.y;
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            float2 a01 = splat2(0.0f);
            float a2 = 0.0f;
#pragma unroll
            for (int k = 0; k < LW; k++) {
                a01 = ffma2(splat2(c_gauss[k]), v[u + k], a01); a2 = fmaf(c_gauss[k], vc[u + k], a2);
            }
            s_hab[r][4 * q + u] = a01; s_hc[r][4 * q + u] = a2;
        }
    }
    __syncthreads();
    const int tx = tid & (LT - 1), ty = tid / LT;
    float2 g01[4];
    float g2[4];
#pragma unroll
    for (int u = 0; u < 4; u++) { g01[u] = splat2(0.0f); g2[u] = 0.0f; }
#pragma unroll
    for (int j = 0; j < 14; j++) {
        const float2 hab = s_hab[4 * ty + j][tx];
        const float hc = s_hc[4 * ty + j][tx];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int k = j - u;
            if (k >= 0 && k < LW) { g01[u] = ffma2(splat2(c_gauss[k]), hab, g01[u]); g2[u] = fmaf(c_gauss[k], hc, g2[u]); }
        }
    }
    const float msum = (float)sums[2], hw = (float)H * (float)W;
    const float k_l1 = msum > 0.0f ? w_l1 / msum : 0.0f;
    const float k_ss = -w_ssim * (msum / hw) / (3.0f * hw);
    const float up = dloss ? dloss[0] : 1.0f;
    const int px = (int)blockIdx.x * LT + tx;
#pragma unroll
    for (int u = 0; u < 4; u++) {
        const int py = (int)blockIdx.y * LT + 4 * ty + u;
        if (px < W && py < H) {
            const size_t q = c * plane + (size_t)py * W + px;
            const float x = pred[q], y = gtc[q];
            const float sgn = x > y ? 1.0f : (x < y ? -1.0f : 0.0f);
            dL_dpred[q] = up * (k_l1 * sgn + k_ss * (g01[u].x + 2.0f * x * g01[u].y + y * g2[u]));
        }
    }
}

// loss3 = { w_l1 * l1 + w_ssim * ssim_term, l1, ssim_term } from the three sums (one thread)
__global__ void image_loss_finalize_kernel(int H, int W, const double* __restrict__ sums, float w_l1, float w_ssim,
                                           float* __restrict__ loss3) {
    const double hw = (double)H * (double)W;
    const double l1 = sums[2] > 0.0 ? sums[0] / sums[2] : 0.0;
    const double ss = (1.0 - sums[1] / (3.0 * hw)) * (sums[2] / hw);
    loss3[0] = (float)((double)w_l1 * l1 + (double)w_ssim * ss);
    loss3[1] = (float)l1;
    loss3[2] = (float)ss;
}

int launch_image_loss_fwd(int H, int W, const float* pred, const void* gt, int gt_u8, const float* mask,
                          const float* bg, float* part, float* gtc, double* sums, float w_l1, float w_ssim,
                          float* loss3, cudaStream_t stream) {
    SGS_CUDA_OK(cudaMemsetAsync(sums, 0, 4 * sizeof(double), stream));
    const dim3 grid((W + LT - 1) / LT, (H + LT - 1) / LT, 3);
    if (gt_u8) image_loss_fwd_kernel<true><<<grid, LTHREADS, 0, stream>>>(H, W, pred, gt, mask, bg, part, gtc, sums);
    else image_loss_fwd_kernel<false><<<grid, LTHREADS, 0, stream>>>(H, W, pred, gt, mask, bg, part, gtc, sums);
    if (loss3) image_loss_finalize_kernel<<<1, 1, 0, stream>>>(H, W, sums, w_l1, w_ssim, loss3);
    SGS_LAUNCH_OK();
    return 0;
}

int launch_image_loss_bwd(int H, int W, const float* pred, const float* gtc, const float* part,
                          const double* sums, float w_l1, float w_ssim, const float* dloss,
                          float* dL_dpred, float* loss_out, cudaStream_t stream) {
    const dim3 grid((W + LT - 1) / LT, (H + LT - 1) / LT, 3);
    image_loss_bwd_kernel<<<grid, LTHREADS, 0, stream>>>(H, W, pred, gtc, part, sums, w_l1, w_ssim, dloss, dL_dpred, loss_out);
    SGS_LAUNCH_OK();
    return 0;
}

}  // namespace sgs
