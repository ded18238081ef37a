// regularizers.cu -- the non-image loss terms of a training step on the canonical Gaussians
// ("... scale-edge loss, Laplacians"), fused forward + backward.
//
// Replaces (same formulas)
//   /root/reference/sings/rec/losses/loss_items.py:173-190   RegionLaplacianLoss_v2.forward / forward_hands:
//        sum over regions of  w_region * mean((L_region x_region)^2)   -- per region one sparse matmul, a pow,
//        a mean and their autograd nodes (15 regions -> ~75 launches forward, as many backward); here the
//        regions' operators are ONE CSR matrix over all vertices with a per-row weight
//        w_region / (n_region * C), one kernel forward and one backward
//   /root/reference/sings/rec/losses/loss_items.py:205-214   pcd_laplacian_smoothing: mean_r |(L x)_r|_2
//   /root/reference/sings/rec/losses/loss_items.py:15-54     L2Norm.forward: four Frobenius norms over
//        xyz_offsets, scales[:, 0] (centred, and the part above a threshold) and the opacities below a threshold
//        (~20 elementwise / reduction / boolean-gather launches with a host sync per masked index); here one
//        pass forward (five sums in double), one pass backward
// L itself (pytorch3d.ops.laplacian: L = D^-1 A - I) is built once per densification by the host
// (sings_b200/regularizers.py); it is constant between densifications (gs_trainer.py:515-521).
// All of it is HBM-bound streaming over N <= a few 100k rows: 12 + 4 deg bytes per row and channel set.
#include "common.cuh"
#include "kernels.h"

namespace sgs {

constexpr int RG_THREADS = 256;

__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// one atomicAdd per CTA and accumulator
template <int K>
__device__ __forceinline__ void block_accumulate(double (&v)[K], double* __restrict__ sums) {
    __shared__ double s_part[K][RG_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < K; k++) {
        const double w = warp_sum_f64(v[k]);
        if (lane == 0) s_part[k][warp] = w;
    }
    __syncthreads();
    if (threadIdx.x < K) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < RG_THREADS / 32; w++) t += s_part[threadIdx.x][w];
        if (t != 0.0) atomicAdd(&sums[threadIdx.x], t);
    }
}

// ------------------------------------------------------------------------------------------------
// Laplacian terms
// ------------------------------------------------------------------------------------------------
// y_r = sum_e vals[e] x[col[e]]  (row r of the CSR operator);  sum += w_r f(y_r),
// f = |y|^2 (mode 0) or |y| (mode 1).  One thread per row, C channels (1..4) per thread.
template <int C>
__global__ void __launch_bounds__(RG_THREADS)
laplacian_loss_fwd_kernel(int n, const int* __restrict__ row_ptr, const int* __restrict__ col_idx,
                          const float* __restrict__ vals, const float* __restrict__ row_w, int mode,
                          const float* __restrict__ x, int ldx, float* __restrict__ y, double* __restrict__ sum) {
    const int r = blockIdx.x * RG_THREADS + threadIdx.x;
    double acc[1] = {0.0};
    if (r < n) {
        float yr[C];
#pragma unroll
        for (int c = 0; c < C; c++) yr[c] = 0.0f;
        const int e1 = row_ptr[r + 1];
        for (int e = row_ptr[r]; e < e1; e++) {
            const float v = vals[e];
            const float* xp = x + (size_t)col_idx[e] * ldx;
#pragma unroll
            for (int c = 0; c < C; c++) yr[c] = fmaf(v, xp[c], yr[c]);
        }
        float q = 0.0f;
#pragma unroll
        for (int c = 0; c < C; c++) { y[(size_t)r * C + c] = yr[c]; q = fmaf(yr[c], yr[c], q); }
        acc[0] = (double)row_w[r] * (double)(mode == 0 ? q : sqrtf(q));
    }
    block_accumulate<1>(acc, sum);
}

// dx_j = dloss * sum_{r : L[r][j] != 0} L[r][j] w_r f'(y_r),  f' = 2 y (mode 0) or y / |y| (0 at y = 0, as
// torch's norm backward) -- a gather over row j of the transposed operator, no atomics.
template <int C>
__global__ void __launch_bounds__(RG_THREADS)
laplacian_loss_bwd_kernel(int n, const int* __restrict__ t_ptr, const int* __restrict__ t_row,
                          const float* __restrict__ t_val, const float* __restrict__ row_w, int mode,
                          const float* __restrict__ y, const float* __restrict__ dloss, float* __restrict__ dx) {
    const int j = blockIdx.x * RG_THREADS + threadIdx.x;
    if (j >= n) return;
    const float up = dloss ? dloss[0] : 1.0f;
    float g[C];
#pragma unroll
    for (int c = 0; c < C; c++) g[c] = 0.0f;
    const int e1 = t_ptr[j + 1];
    for (int e = t_ptr[j]; e < e1; e++) {
        const int r = t_row[e];
        const float w = row_w[r];
        if (w == 0.0f) continue;
        const float* yp = y + (size_t)r * C;
        float yr[C], q = 0.0f;
#pragma unroll
        for (int c = 0; c < C; c++) { yr[c] = yp[c]; q = fmaf(yr[c], yr[c], q); }
        float k;
        if (mode == 0) k = 2.0f * w;
        else k = q > 0.0f ? w / sqrtf(q) : 0.0f;
        k *= t_val[e];
#pragma unroll
        for (int c = 0; c < C; c++) g[c] = fmaf(k, yr[c], g[c]);
    }
#pragma unroll
    for (int c = 0; c < C; c++) dx[(size_t)j * C + c] = up * g[c];
}

__global__ void sum_to_float_kernel(const double* __restrict__ sum, float* __restrict__ out) { out[0] = (float)sum[0]; }

int launch_laplacian_loss_fwd(int n, int C, const int* row_ptr, const int* col_idx, const float* vals,
                              const float* row_w, int mode, const float* x, int ldx, float* y, double* sum,
                              float* loss_out, cudaStream_t stream) {
    SGS_CUDA_OK(cudaMemsetAsync(sum, 0, sizeof(double), stream));
    if (n > 0) {
        const int grid = (n + RG_THREADS - 1) / RG_THREADS;
        switch (C) {
            case 1: laplacian_loss_fwd_kernel<1><<<grid, RG_THREADS, 0, stream>>>(n, row_ptr, col_idx, vals, row_w, mode, x, ldx, y, sum); break;
            case 2: laplacian_loss_fwd_kernel<2><<<grid, RG_THREADS, 0, stream>>>(n, row_ptr, col_idx, vals, row_w, mode, x, ldx, y, sum); break;
            case 3: laplacian_loss_fwd_kernel<3><<<grid, RG_THREADS, 0, stream>>>(n, row_ptr, col_idx, vals, row_w, mode, x, ldx, y, sum); break;
            default: laplacian_loss_fwd_kernel<4><<<grid, RG_THREADS, 0, stream>>>(n, row_ptr, col_idx, vals, row_w, mode, x, ldx, y, sum); break;
        }
    }
    if (loss_out) sum_to_float_kernel<<<1, 1, 0, stream>>>(sum, loss_out);
    SGS_LAUNCH_OK();
    return 0;
}

int launch_laplacian_loss_bwd(int n, int C, const int* t_ptr, const int* t_row, const float* t_val,
                              const float* row_w, int mode, const float* y, const float* dloss, float* dx,
                              cudaStream_t stream) {
    if (n <= 0) return 0;
    const int grid = (n + RG_THREADS - 1) / RG_THREADS;
    switch (C) {
        case 1: laplacian_loss_bwd_kernel<1><<<grid, RG_THREADS, 0, stream>>>(n, t_ptr, t_row, t_val, row_w, mode, y, dloss, dx); break;
        case 2: laplacian_loss_bwd_kernel<2><<<grid, RG_THREADS, 0, stream>>>(n, t_ptr, t_row, t_val, row_w, mode, y, dloss, dx); break;
        case 3: laplacian_loss_bwd_kernel<3><<<grid, RG_THREADS, 0, stream>>>(n, t_ptr, t_row, t_val, row_w, mode, y, dloss, dx); break;
        default: laplacian_loss_bwd_kernel<4><<<grid, RG_THREADS, 0, stream>>>(n, t_ptr, t_row, t_val, row_w, mode, y, dloss, dx); break;
    }
    SGS_LAUNCH_OK();
    return 0;
}

// ------------------------------------------------------------------------------------------------
// L2Norm
// ------------------------------------------------------------------------------------------------
// sums: [0] sum |xyz_offsets|^2   [1] sum s   [2] sum s^2   [3] sum_{s > thr_s} s^2   [4] sum_{o < thr_o} (0.5 - o)^2
__global__ void __launch_bounds__(RG_THREADS)
l2norm_fwd_kernel(int N, const float* __restrict__ off, const float* __restrict__ scales, int lds,
                  const float* __restrict__ opacity, float thr_s, float thr_o, double* __restrict__ sums) {
    double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    for (int i = blockIdx.x * RG_THREADS + threadIdx.x; i < N; i += gridDim.x * RG_THREADS) {
        if (off) {
            const float a = off[3 * (size_t)i], b = off[3 * (size_t)i + 1], c = off[3 * (size_t)i + 2];
            acc[0] += (double)a * a + (double)b * b + (double)c * c;
        }
        if (scales) {
            const double s = (double)scales[(size_t)i * lds];
            acc[1] += s;
            acc[2] += s * s;
            if ((float)s > thr_s) acc[3] += s * s;
        }
        if (opacity) {
            const float o = opacity[i];
            if (o < thr_o) { const double d = 0.5 - (double)o; acc[4] += d * d; }
        }
    }
    block_accumulate<5>(acc, sums);
}

// the four norms (double) from the five sums: sums[5..8] = |offsets|, |s - mean s|, |s[s > thr]|, |0.5 - o[o < thr]|;
// loss = sum_k lambda_k norm_k
__global__ void l2norm_finalize_kernel(int N, double* __restrict__ sums, float l_off, float l_diff, float l_max,
                                       float l_op, float* __restrict__ loss_out) {
    const double n0 = sqrt(sums[0]);
    const double var = N > 0 ? sums[2] - sums[1] * sums[1] / (double)N : 0.0;
    const double n1 = sqrt(var > 0.0 ? var : 0.0);
    const double n2 = sqrt(sums[3]), n3 = sqrt(sums[4]);
    sums[5] = n0; sums[6] = n1; sums[7] = n2; sums[8] = n3;
    if (loss_out) loss_out[0] = (float)((double)l_off * n0 + (double)l_diff * n1 + (double)l_max * n2 + (double)l_op * n3);
}

// d|v| / dv = v / |v| (0 where the norm is 0, as torch's norm backward); the mean inside |s - mean s| carries no
// gradient of its own: sum_j (s_j - mean) = 0.
__global__ void __launch_bounds__(RG_THREADS)
l2norm_bwd_kernel(int N, const float* __restrict__ off, const float* __restrict__ scales, int lds, int S,
                  const float* __restrict__ opacity, float thr_s, float thr_o, const double* __restrict__ sums,
                  float l_off, float l_diff, float l_max, float l_op, const float* __restrict__ dloss,
                  float* __restrict__ d_off, float* __restrict__ d_scales, float* __restrict__ d_opacity) {
    const float up = dloss ? dloss[0] : 1.0f;
    const double mean = N > 0 ? sums[1] / (double)N : 0.0;
    const float k_off = sums[5] > 0.0 ? (float)((double)(up * l_off) / sums[5]) : 0.0f;
    const double k_diff = sums[6] > 0.0 ? (double)(up * l_diff) / sums[6] : 0.0;
    const float k_max = sums[7] > 0.0 ? (float)((double)(up * l_max) / sums[7]) : 0.0f;
    const float k_op = sums[8] > 0.0 ? (float)((double)(up * l_op) / sums[8]) : 0.0f;
    for (int i = blockIdx.x * RG_THREADS + threadIdx.x; i < N; i += gridDim.x * RG_THREADS) {
        if (d_off) {
#pragma unroll
            for (int c = 0; c < 3; c++) d_off[3 * (size_t)i + c] = k_off * off[3 * (size_t)i + c];
        }
        if (d_scales) {
            const float s = scales[(size_t)i * lds];
            float g = (float)(k_diff * ((double)s - mean));
            if (s > thr_s) g = fmaf(k_max, s, g);
            d_scales[(size_t)i * S] = g;
            for (int c = 1; c < S; c++) d_scales[(size_t)i * S + c] = 0.0f;
        }
        if (d_opacity) {
            const float o = opacity[i];
            d_opacity[i] = o < thr_o ? -k_op * (0.5f - o) : 0.0f;
        }
    }
}

static int stream_grid(int N) {
    const int g = (N + RG_THREADS - 1) / RG_THREADS;
    return g < 1 ? 1 : (g > 148 * 8 ? 148 * 8 : g);
}

int launch_l2norm_fwd(int N, const float* off, const float* scales, int lds, const float* opacity, float thr_s,
                      float thr_o, float l_off, float l_diff, float l_max, float l_op, double* sums, float* loss_out,
                      cudaStream_t stream) {
    SGS_CUDA_OK(cudaMemsetAsync(sums, 0, 9 * sizeof(double), stream));
    if (N > 0) l2norm_fwd_kernel<<<stream_grid(N), RG_THREADS, 0, stream>>>(N, off, scales, lds, opacity, thr_s, thr_o, sums);
    l2norm_finalize_kernel<<<1, 1, 0, stream>>>(N, sums, l_off, l_diff, l_max, l_op, loss_out);
    SGS_LAUNCH_OK();
    return 0;
}

int launch_l2norm_bwd(int N, const float* off, const float* scales, int lds, int S, const float* opacity, float thr_s,
                      float thr_o, const double* sums, float l_off, float l_diff, float l_max, float l_op,
                      const float* dloss, float* d_off, float* d_scales, float* d_opacity, cudaStream_t stream) {
    if (N <= 0) return 0;
    l2norm_bwd_kernel<<<stream_grid(N), RG_THREADS, 0, stream>>>(N, off, scales, lds, S, opacity, thr_s, thr_o, sums, l_off,
                                                                 l_diff, l_max, l_op, dloss, d_off, d_scales, d_opacity);
    SGS_LAUNCH_OK();
    return 0;
}

}  // namespace sgs
