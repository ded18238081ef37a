// lbs.cu -- SMPL linear-blend-skinning deformation of every Gaussian, fused, forward + backward.
//
// Replaces, in one pass over HBM, the reference's per-frame deform segment
//   /root/reference/sings/rec/models/sings_hybrid.py:398-428 (forward) and :525-552 (chunked):
//     A_cano2pose = A_t2pose @ inv_A_t2cano                       (:399)
//     lbs_extra: T = W @ A.view(J,16); x' = T [x;1]               (utils/body_model/lbs.py:61-73)
//     * smpl_scale, + transl                                      (:411-416)
//     R' = T[:3,:3] @ R_canon; q = matrix_to_quaternion(R')       (:418-419,
//                                         utils/geometry/rotations.py:98-149, NOT normalised)
//     optional external similarity ext_tfs                        (:421-428)
//   and pose -> A: batch_rodrigues + batch_rigid_transform
//                                         (utils/body_model/smpl.py:415-446, 462-513).
// The backward is what torch autograd computes for that graph.
//
// Layout: one CTA = 256 consecutive Gaussians, one thread each, all B frames looped inside so
// the skinning-weight rows (the largest operand) are read from HBM once per call.  The CTA's
// contiguous block of W rows is fetched with ONE TMA bulk copy (cp.async.bulk) into shared
// memory; the B x J x 12 joint transforms sit in shared memory as float4 and are read as
// warp-broadcast LDS.128.  Tensor cores are deliberately unused: K = J <= 52, ~6 flop/byte.
#include "common.cuh"
#include "kernels_lbs.h"

namespace sgs {

constexpr int LBS_THREADS = 256;

// ------------------------------------------------------------------------------------------
// pose -> A
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void rodrigues(float rx, float ry, float rz, float* R) {
    // smpl.py:430-446: theta = ||r + 1e-8||, K = skew(r/theta), R = I + sin K + (1-cos) K^2
    const float ax = rx + 1e-8f, ay = ry + 1e-8f, az = rz + 1e-8f;
    const float angle = sqrtf(ax * ax + ay * ay + az * az);
    const float x = rx / angle, y = ry / angle, z = rz / angle;
    const float s = sinf(angle), c = 1.0f - cosf(angle);
    // K = [[0,-z,y],[z,0,-x],[-y,x,0]];  K^2 = [[-(y^2+z^2), xy, xz],[xy, -(x^2+z^2), yz],[xz, yz, -(x^2+y^2)]]
    R[0] = 1.0f + c * -(y * y + z * z); R[1] = s * -z + c * (x * y);        R[2] = s * y + c * (x * z);
    R[3] = s * z + c * (x * y);         R[4] = 1.0f + c * -(x * x + z * z); R[5] = s * -x + c * (y * z);
    R[6] = s * -y + c * (x * z);        R[7] = s * x + c * (y * z);         R[8] = 1.0f + c * -(x * x + y * y);
}

// affine 3x4 product: out = a @ b (both [R|t], implicit last row 0 0 0 1)
__device__ __forceinline__ void affine_mul(const float* a, const float* b, float* out) {
#pragma unroll
    for (int r = 0; r < 3; r++) {
#pragma unroll
        for (int c = 0; c < 4; c++) {
            float v = a[4 * r] * b[c] + a[4 * r + 1] * b[4 + c] + a[4 * r + 2] * b[8 + c];
            if (c == 3) v += a[4 * r + 3];
            out[4 * r + c] = v;
        }
    }
}

// one CTA per frame, one thread per joint; each thread multiplies its own ancestor chain
// from the root down (same association as the reference's sequential loop, smpl.py:495-501)
__global__ void pose_to_A_kernel(const float* __restrict__ pose, const float* __restrict__ rest,
                                 const int* __restrict__ parents, const float* __restrict__ inv_A,
                                 int J, float* __restrict__ A_out, float* __restrict__ G_out) {
    extern __shared__ float s_local[];     // J x 12 local transforms
    const int b = blockIdx.x, j = threadIdx.x;
    if (j < J) {
        float R[9];
        const float* p = pose + ((size_t)b * J + j) * 3;
        rodrigues(p[0], p[1], p[2], R);
        const int par = parents[j];
        float t[3];
#pragma unroll
        for (int k = 0; k < 3; k++) t[k] = rest[3 * j + k] - (par >= 0 ? rest[3 * par + k] : 0.0f);
        float* L = s_local + 12 * j;
#pragma unroll
        for (int r = 0; r < 3; r++) {
            L[4 * r] = R[3 * r]; L[4 * r + 1] = R[3 * r + 1]; L[4 * r + 2] = R[3 * r + 2]; L[4 * r + 3] = t[r];
        }
    }
    __syncthreads();
    if (j >= J) return;
    int chain[64];
    int depth = 0;
    for (int k = j; k >= 0 && depth < 64; k = parents[k]) chain[depth++] = k;
    float G[12], tmp[12];
#pragma unroll
    for (int k = 0; k < 12; k++) G[k] = s_local[12 * chain[depth - 1] + k];
    for (int d = depth - 2; d >= 0; d--) {
        affine_mul(G, s_local + 12 * chain[d], tmp);
#pragma unroll
        for (int k = 0; k < 12; k++) G[k] = tmp[k];
    }
    if (G_out) {
#pragma unroll
        for (int k = 0; k < 12; k++) G_out[((size_t)b * J + j) * 12 + k] = G[k];
    }
    // A = G - pad(G [rest_j; 0])  (smpl.py:510-511)
    float Arel[12];
#pragma unroll
    for (int r = 0; r < 3; r++) {
        Arel[4 * r] = G[4 * r]; Arel[4 * r + 1] = G[4 * r + 1]; Arel[4 * r + 2] = G[4 * r + 2];
        Arel[4 * r + 3] = G[4 * r + 3] - (G[4 * r] * rest[3 * j] + G[4 * r + 1] * rest[3 * j + 1] + G[4 * r + 2] * rest[3 * j + 2]);
    }
    float out[12];
    if (inv_A) {
        float Bm[12];
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
            for (int c = 0; c < 4; c++) Bm[4 * r + c] = inv_A[(size_t)j * 16 + 4 * r + c];
        affine_mul(Arel, Bm, out);
    } else {
#pragma unroll
        for (int k = 0; k < 12; k++) out[k] = Arel[k];
    }
    float* dst = A_out + ((size_t)b * J + j) * 16;
#pragma unroll
    for (int k = 0; k < 12; k++) dst[k] = out[k];
    dst[12] = 0.0f; dst[13] = 0.0f; dst[14] = 0.0f; dst[15] = 1.0f;
}

int launch_pose_to_A(const float* pose, const float* rest, const int* parents, const float* inv_A,
                     int B, int J, float* A_out, float* G_out, cudaStream_t stream) {
    if (B <= 0) return 0;
    if (J < 1 || J > 64) return SGS_ERR_BAD_JOINTS;
    pose_to_A_kernel<<<B, 64, (size_t)J * 12 * 4, stream>>>(pose, rest, parents, inv_A, J, A_out, G_out);
    SGS_LAUNCH_OK();
    return 0;
}

// backward of pose_to_A_kernel: dL/dA (B,J,16) -> dL/dpose (B,J,3).  One CTA per frame.
// Thread j seeds dL/dG_j, thread 0 walks the tree leaves-to-root (parents[j] < j), thread j
// then differentiates its Rodrigues formula.
__global__ void pose_to_A_bwd_kernel(const float* __restrict__ pose, const float* __restrict__ rest,
                                     const int* __restrict__ parents, const float* __restrict__ inv_A,
                                     const float* __restrict__ G_all, const float* __restrict__ dA,
                                     int J, float* __restrict__ d_pose) {
    extern __shared__ float s_mem[];
    float* s_L = s_mem;              // J x 12 local transforms
    float* s_dG = s_mem + 12 * J;    // J x 12 dL/dG, then dL/dL
    const int b = blockIdx.x, j = threadIdx.x;
    const float* G = G_all + (size_t)b * J * 12;
    float R[9];
    if (j < J) {
        const float* p = pose + ((size_t)b * J + j) * 3;
        rodrigues(p[0], p[1], p[2], R);
        const int par = parents[j];
        float* L = s_L + 12 * j;
#pragma unroll
        for (int r = 0; r < 3; r++) {
            L[4 * r] = R[3 * r]; L[4 * r + 1] = R[3 * r + 1]; L[4 * r + 2] = R[3 * r + 2];
            L[4 * r + 3] = rest[3 * j + r] - (par >= 0 ? rest[3 * par + r] : 0.0f);
        }
        // dOut -> dArel (through @ inv_A) -> dG
        float dO[12];
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
            for (int c = 0; c < 4; c++) dO[4 * r + c] = dA[((size_t)b * J + j) * 16 + 4 * r + c];
        float dAr[12];
        if (inv_A) {
            const float* Bm = inv_A + (size_t)j * 16;
#pragma unroll
            for (int r = 0; r < 3; r++) {
#pragma unroll
                for (int c = 0; c < 3; c++)
                    dAr[4 * r + c] = dO[4 * r] * Bm[4 * c] + dO[4 * r + 1] * Bm[4 * c + 1] +
                                     dO[4 * r + 2] * Bm[4 * c + 2] + dO[4 * r + 3] * Bm[4 * c + 3];
                dAr[4 * r + 3] = dO[4 * r + 3];
            }
        } else {
#pragma unroll
            for (int k = 0; k < 12; k++) dAr[k] = dO[k];
        }
        float* dG = s_dG + 12 * j;
#pragma unroll
        for (int r = 0; r < 3; r++) {
#pragma unroll
            for (int c = 0; c < 3; c++) dG[4 * r + c] = dAr[4 * r + c] - dAr[4 * r + 3] * rest[3 * j + c];
            dG[4 * r + 3] = dAr[4 * r + 3];
        }
    }
    __syncthreads();
    if (j == 0) {
        for (int k = J - 1; k >= 1; k--) {
            const int par = parents[k];
            const float* Gp = G + 12 * par;
            const float* L = s_L + 12 * k;
            float* dG = s_dG + 12 * k;
            float* dGp = s_dG + 12 * par;
            float dL[12];
            for (int r = 0; r < 3; r++)
                for (int c = 0; c < 4; c++)
                    dL[4 * r + c] = Gp[r] * dG[c] + Gp[4 + r] * dG[4 + c] + Gp[8 + r] * dG[8 + c];
            for (int r = 0; r < 3; r++) {
                for (int c = 0; c < 3; c++)
                    dGp[4 * r + c] += dG[4 * r] * L[4 * c] + dG[4 * r + 1] * L[4 * c + 1] +
                                      dG[4 * r + 2] * L[4 * c + 2] + dG[4 * r + 3] * L[4 * c + 3];
                dGp[4 * r + 3] += dG[4 * r + 3];
            }
            for (int q = 0; q < 12; q++) dG[q] = dL[q];
        }
    }
    __syncthreads();
    if (j >= J) return;
    // Rodrigues backward; E = dL/dR_j (rotation part of dL/dL_j; the root's L is G itself)
    float E[9];
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 3; c++) E[3 * r + c] = s_dG[12 * j + 4 * r + c];
    const float* p = pose + ((size_t)b * J + j) * 3;
    const float rx = p[0], ry = p[1], rz = p[2];
    const float ax = rx + 1e-8f, ay = ry + 1e-8f, az = rz + 1e-8f;
    const float th = sqrtf(ax * ax + ay * ay + az * az);
    const float x = rx / th, y = ry / th, z = rz / th;
    const float sn = sinf(th), cs = cosf(th), c2 = 1.0f - cs;
    const float K[9] = {0, -z, y, z, 0, -x, -y, x, 0};
    float K2[9], gK[9];
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 3; c++)
            K2[3 * r + c] = K[3 * r] * K[c] + K[3 * r + 1] * K[3 + c] + K[3 * r + 2] * K[6 + c];
    float g_s = 0, g_c2 = 0;
#pragma unroll
    for (int k = 0; k < 9; k++) { g_s += E[k] * K[k]; g_c2 += E[k] * K2[k]; }
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 3; c++) {
            // d(K K) = dK K + K dK  ->  gK = E K^T + K^T E
            float ekt = E[3 * r] * K[3 * c] + E[3 * r + 1] * K[3 * c + 1] + E[3 * r + 2] * K[3 * c + 2];
            float kte = K[r] * E[c] + K[3 + r] * E[3 + c] + K[6 + r] * E[6 + c];
            gK[3 * r + c] = sn * E[3 * r + c] + c2 * (ekt + kte);
        }
    const float gdx = gK[7] - gK[5], gdy = gK[2] - gK[6], gdz = gK[3] - gK[1];
    float g_th = g_s * cs + g_c2 * sn - (gdx * rx + gdy * ry + gdz * rz) / (th * th);
    float* out = d_pose + ((size_t)b * J + j) * 3;
    out[0] = gdx / th + g_th * ax / th;
    out[1] = gdy / th + g_th * ay / th;
    out[2] = gdz / th + g_th * az / th;
}

int launch_pose_to_A_bwd(const float* pose, const float* rest, const int* parents,
                         const float* inv_A, const float* G, const float* dA, int B, int J,
                         float* d_pose, cudaStream_t stream) {
    if (B <= 0) return 0;
    if (J < 1 || J > 64) return SGS_ERR_BAD_JOINTS;
    pose_to_A_bwd_kernel<<<B, 64, (size_t)J * 24 * 4, stream>>>(pose, rest, parents, inv_A, G, dA, J, d_pose);
    SGS_LAUNCH_OK();
    return 0;
}

// ------------------------------------------------------------------------------------------
// matrix -> quaternion (rotations.py:98-149) and its backward through the selected candidate
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int mat_to_quat(const float* m, float* q) {
    const float arg[4] = {1.0f + m[0] + m[4] + m[8], 1.0f + m[0] - m[4] - m[8],
                          1.0f - m[0] + m[4] - m[8], 1.0f - m[0] - m[4] + m[8]};
    float qa[4];
    int best = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) qa[i] = arg[i] > 0.0f ? sqrtf(arg[i]) : 0.0f;
#pragma unroll
    for (int i = 1; i < 4; i++)
        if (qa[i] > qa[best]) best = i;          // first maximum, like torch.argmax
    const float m01 = m[1], m02 = m[2], m10 = m[3], m12 = m[5], m20 = m[6], m21 = m[7];
    float c[4];
    const float sq = qa[best] * qa[best];
    if (best == 0)      { c[0] = sq;        c[1] = m21 - m12; c[2] = m02 - m20; c[3] = m10 - m01; }
    else if (best == 1) { c[0] = m21 - m12; c[1] = sq;        c[2] = m10 + m01; c[3] = m02 + m20; }
    else if (best == 2) { c[0] = m02 - m20; c[1] = m10 + m01; c[2] = sq;        c[3] = m12 + m21; }
    else                { c[0] = m10 - m01; c[1] = m20 + m02; c[2] = m21 + m12; c[3] = sq; }
    const float den = 2.0f * fmaxf(qa[best], 0.1f);
#pragma unroll
    for (int k = 0; k < 4; k++) q[k] = c[k] / den;
    return best;
}

// gradient of q w.r.t. the 3x3 matrix (row-major gm[9]) given dL/dq (g[4])
__device__ __forceinline__ void mat_to_quat_bwd(const float* m, const float* g, float* gm) {
    const float arg[4] = {1.0f + m[0] + m[4] + m[8], 1.0f + m[0] - m[4] - m[8],
                          1.0f - m[0] + m[4] - m[8], 1.0f - m[0] - m[4] + m[8]};
    float qa[4];
    int best = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) qa[i] = arg[i] > 0.0f ? sqrtf(arg[i]) : 0.0f;
#pragma unroll
    for (int i = 1; i < 4; i++)
        if (qa[i] > qa[best]) best = i;
    const float m01 = m[1], m02 = m[2], m10 = m[3], m12 = m[5], m20 = m[6], m21 = m[7];
    const float qb = qa[best];
    const float sq = qb * qb;
    float c[4];
    if (best == 0)      { c[0] = sq;        c[1] = m21 - m12; c[2] = m02 - m20; c[3] = m10 - m01; }
    else if (best == 1) { c[0] = m21 - m12; c[1] = sq;        c[2] = m10 + m01; c[3] = m02 + m20; }
    else if (best == 2) { c[0] = m02 - m20; c[1] = m10 + m01; c[2] = sq;        c[3] = m12 + m21; }
    else                { c[0] = m10 - m01; c[1] = m20 + m02; c[2] = m21 + m12; c[3] = sq; }
    const float den = 2.0f * fmaxf(qb, 0.1f);
    float gc[4];
    float gden = 0.0f;
#pragma unroll
    for (int k = 0; k < 4; k++) { gc[k] = g[k] / den; gden -= g[k] * c[k] / (den * den); }
    // den = 2 max(qb, 0.1): gradient reaches qb only above the floor; c[best] = qb^2
    float gqb = (qb > 0.1f ? 2.0f * gden : 0.0f) + 2.0f * qb * gc[best];
    const float garg = arg[best] > 0.0f ? gqb / (2.0f * qb) : 0.0f;   // zero sub-gradient at 0
#pragma unroll
    for (int k = 0; k < 9; k++) gm[k] = 0.0f;
    const float s0 = (best == 0 || best == 1) ? 1.0f : -1.0f;
    const float s1 = (best == 0 || best == 2) ? 1.0f : -1.0f;
    const float s2 = (best == 0 || best == 3) ? 1.0f : -1.0f;
    gm[0] = s0 * garg; gm[4] = s1 * garg; gm[8] = s2 * garg;
    // off-diagonal candidates: (index into m, sign) pairs per output slot
    if (best == 0) {
        gm[7] += gc[1]; gm[5] -= gc[1]; gm[2] += gc[2]; gm[6] -= gc[2]; gm[3] += gc[3]; gm[1] -= gc[3];
    } else if (best == 1) {
        gm[7] += gc[0]; gm[5] -= gc[0]; gm[3] += gc[2]; gm[1] += gc[2]; gm[2] += gc[3]; gm[6] += gc[3];
    } else if (best == 2) {
        gm[2] += gc[0]; gm[6] -= gc[0]; gm[3] += gc[1]; gm[1] += gc[1]; gm[5] += gc[3]; gm[7] += gc[3];
    } else {
        gm[3] += gc[0]; gm[1] -= gc[0]; gm[6] += gc[1]; gm[2] += gc[1]; gm[7] += gc[2]; gm[5] += gc[2];
    }
}

__device__ __forceinline__ void quat_mul(const float* a, const float* b, float* o) {
    o[0] = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
    o[1] = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
    o[2] = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
    o[3] = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
}

// ------------------------------------------------------------------------------------------
// shared staging: joint transforms of all frames + this CTA's skinning-weight rows (TMA)
// ------------------------------------------------------------------------------------------
struct LbsSmem {
    float4* A;        // [B][J][3]  rows 0..2 of each joint transform
    float* W;         // [256][J]
    float* frame;     // [B][16]: smpl_scale, transl(3), ext_scale, ext_trans(3), ext_quat(4) ...
    float* dT;        // backward only: [256][12]
    unsigned long long* bar;
};

__device__ __forceinline__ void stage_inputs(const LbsArgs& a, LbsSmem& s, int base, int rows) {
    const int tid = threadIdx.x;
    const size_t w_bytes = (size_t)rows * a.J * 4;
    const float* w_src = a.W + (size_t)base * a.J;
    const bool tma_ok = ((uintptr_t)w_src & 15) == 0 && (w_bytes & 15) == 0;
    if (tma_ok) {
        if (tid == 0) {
            mbar_init(s.bar, 1);
            mbar_fence_init();
        }
        __syncthreads();
        if (tid == 0) {
            mbar_arrive_expect_tx(s.bar, (unsigned)w_bytes);
            tma_bulk_g2s(s.W, w_src, (unsigned)w_bytes, s.bar);
        }
    } else {
        for (int f = tid; f < rows * a.J; f += LBS_THREADS) s.W[f] = w_src[f];
    }
    for (int f = tid; f < a.B * a.J * 3; f += LBS_THREADS) {
        int bj = f / 3, r = f - bj * 3;
        s.A[f] = reinterpret_cast<const float4*>(a.A)[(size_t)bj * 4 + r];
    }
    for (int b = tid; b < a.B; b += LBS_THREADS) {
        float* fr = s.frame + 16 * b;
        fr[0] = a.smpl_scale ? a.smpl_scale[b] : 1.0f;
        for (int k = 0; k < 3; k++) fr[1 + k] = a.transl ? a.transl[3 * b + k] : 0.0f;
        if (a.ext_rot) {
            fr[4] = a.ext_scale[b];
            for (int k = 0; k < 3; k++) fr[5 + k] = a.ext_trans[3 * b + k];
            float q[4];
            mat_to_quat(a.ext_rot + 9 * b, q);
            for (int k = 0; k < 4; k++) fr[8 + k] = q[k];
        }
    }
    __syncthreads();
    if (tma_ok) mbar_wait(s.bar, 0);
}

__device__ __forceinline__ LbsSmem carve(char* raw, int B, int J, bool bwd) {
    LbsSmem s;
    size_t o = 0;
    s.A = reinterpret_cast<float4*>(raw + o);   o += align_up((size_t)B * J * 3 * 16, 16);
    s.W = reinterpret_cast<float*>(raw + o);    o += align_up((size_t)LBS_THREADS * J * 4, 16);
    s.frame = reinterpret_cast<float*>(raw + o); o += align_up((size_t)B * 16 * 4, 16);
    s.dT = reinterpret_cast<float*>(raw + o);   if (bwd) o += (size_t)LBS_THREADS * 12 * 4;
    s.bar = reinterpret_cast<unsigned long long*>(raw + o);
    return s;
}

static size_t lbs_smem_bytes(int B, int J, bool bwd) {
    return align_up((size_t)B * J * 3 * 16, 16) + align_up((size_t)LBS_THREADS * J * 4, 16) +
           align_up((size_t)B * 16 * 4, 16) + (bwd ? (size_t)LBS_THREADS * 12 * 4 : 0) + 16;
}

// T (3x4, row-major 12 floats) = sum_j w_j A[b][j]
__device__ __forceinline__ void blend_T(const LbsSmem& s, int b, int J, int t, float* T) {
#pragma unroll
    for (int k = 0; k < 12; k++) T[k] = 0.0f;
    const float4* Ab = s.A + (size_t)b * J * 3;
    if ((J & 3) == 0) {      // rows are 16-byte aligned: vector reads of the weight row
        const float4* wrow = reinterpret_cast<const float4*>(s.W + (size_t)t * J);
        for (int j4 = 0; j4 < J / 4; j4++) {
            const float4 w4 = wrow[j4];
            const float w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const float4 r0 = Ab[(4 * j4 + u) * 3], r1 = Ab[(4 * j4 + u) * 3 + 1], r2 = Ab[(4 * j4 + u) * 3 + 2];
                T[0] = fmaf(w[u], r0.x, T[0]); T[1] = fmaf(w[u], r0.y, T[1]); T[2] = fmaf(w[u], r0.z, T[2]); T[3] = fmaf(w[u], r0.w, T[3]);
                T[4] = fmaf(w[u], r1.x, T[4]); T[5] = fmaf(w[u], r1.y, T[5]); T[6] = fmaf(w[u], r1.z, T[6]); T[7] = fmaf(w[u], r1.w, T[7]);
                T[8] = fmaf(w[u], r2.x, T[8]); T[9] = fmaf(w[u], r2.y, T[9]); T[10] = fmaf(w[u], r2.z, T[10]); T[11] = fmaf(w[u], r2.w, T[11]);
            }
        }
    } else {
        for (int j = 0; j < J; j++) {
            const float w = s.W[(size_t)t * J + j];
            const float4 r0 = Ab[j * 3], r1 = Ab[j * 3 + 1], r2 = Ab[j * 3 + 2];
            T[0] = fmaf(w, r0.x, T[0]); T[1] = fmaf(w, r0.y, T[1]); T[2] = fmaf(w, r0.z, T[2]); T[3] = fmaf(w, r0.w, T[3]);
            T[4] = fmaf(w, r1.x, T[4]); T[5] = fmaf(w, r1.y, T[5]); T[6] = fmaf(w, r1.z, T[6]); T[7] = fmaf(w, r1.w, T[7]);
            T[8] = fmaf(w, r2.x, T[8]); T[9] = fmaf(w, r2.y, T[9]); T[10] = fmaf(w, r2.z, T[10]); T[11] = fmaf(w, r2.w, T[11]);
        }
    }
}

__device__ __forceinline__ void compose_rot(const float* T, const float* Rc, bool iso, float* Rp) {
    if (iso) {
#pragma unroll
        for (int r = 0; r < 3; r++) { Rp[3 * r] = T[4 * r]; Rp[3 * r + 1] = T[4 * r + 1]; Rp[3 * r + 2] = T[4 * r + 2]; }
    } else {
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
            for (int c = 0; c < 3; c++)
                Rp[3 * r + c] = T[4 * r] * Rc[c] + T[4 * r + 1] * Rc[3 + c] + T[4 * r + 2] * Rc[6 + c];
    }
}

// ------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(LBS_THREADS) lbs_fwd_kernel(LbsArgs a, LbsOut o) {
    extern __shared__ __align__(16) char s_raw[];
    LbsSmem s = carve(s_raw, a.B, a.J, false);
    const int tid = threadIdx.x;
    const int base = blockIdx.x * LBS_THREADS;
    const int rows = min(LBS_THREADS, a.N - base);
    stage_inputs(a, s, base, rows);
    const int n = base + tid;
    if (n >= a.N) return;
    const float x = a.xyz[3 * (size_t)n], y = a.xyz[3 * (size_t)n + 1], z = a.xyz[3 * (size_t)n + 2];
    const float s0 = a.scales[3 * (size_t)n], s1 = a.scales[3 * (size_t)n + 1], s2 = a.scales[3 * (size_t)n + 2];
    const bool iso = a.rot == nullptr;
    float Rc[9];
    if (!iso) {
#pragma unroll
        for (int k = 0; k < 9; k++) Rc[k] = a.rot[9 * (size_t)n + k];
    }
    for (int b = 0; b < a.B; b++) {
        float T[12];
        blend_T(s, b, a.J, tid, T);
        const float* fr = s.frame + 16 * b;
        float vx = T[0] * x + T[1] * y + T[2] * z + T[3];
        float vy = T[4] * x + T[5] * y + T[6] * z + T[7];
        float vz = T[8] * x + T[9] * y + T[10] * z + T[11];
        float sc0 = s0, sc1 = s1, sc2 = s2;
        if (a.smpl_scale) { vx *= fr[0]; vy *= fr[0]; vz *= fr[0]; sc0 *= fr[0]; sc1 *= fr[0]; sc2 *= fr[0]; }
        if (a.transl) { vx += fr[1]; vy += fr[2]; vz += fr[3]; }
        float Rp[9], q[4];
        compose_rot(T, Rc, iso, Rp);
        mat_to_quat(Rp, q);
        if (a.ext_rot) {
            const float* eR = a.ext_rot + 9 * b;
            const float es = fr[4];
            const float rx = eR[0] * vx + eR[1] * vy + eR[2] * vz;
            const float ry = eR[3] * vx + eR[4] * vy + eR[5] * vz;
            const float rz = eR[6] * vx + eR[7] * vy + eR[8] * vz;
            vx = fr[5] + es * rx; vy = fr[6] + es * ry; vz = fr[7] + es * rz;
            sc0 *= es; sc1 *= es; sc2 *= es;
            float qo[4];
            quat_mul(fr + 8, q, qo);
            const float sg = qo[0] < 0.0f ? -1.0f : 1.0f;
#pragma unroll
            for (int k = 0; k < 4; k++) q[k] = sg * qo[k];
        }
        const size_t on = (size_t)b * a.N + n;
        o.xyz[3 * on] = vx; o.xyz[3 * on + 1] = vy; o.xyz[3 * on + 2] = vz;
        reinterpret_cast<float4*>(o.rotq)[on] = make_float4(q[0], q[1], q[2], q[3]);
        o.scales[3 * on] = sc0; o.scales[3 * on + 1] = sc1; o.scales[3 * on + 2] = sc2;
        if (o.T) {
            float4* Tn = reinterpret_cast<float4*>(o.T) + on * 4;
            Tn[0] = make_float4(T[0], T[1], T[2], T[3]);
            Tn[1] = make_float4(T[4], T[5], T[6], T[7]);
            Tn[2] = make_float4(T[8], T[9], T[10], T[11]);
            // row 3 = sum_j w_j (0,0,0,1): the reference's T[3,3] is the weight-row sum
            float wsum = 0.0f;
            for (int j = 0; j < a.J; j++) wsum += s.W[(size_t)tid * a.J + j];
            Tn[3] = make_float4(0.0f, 0.0f, 0.0f, wsum);
        }
    }
}

int launch_lbs_fwd(const LbsArgs& a, const LbsOut& o, cudaStream_t stream) {
    if (a.N <= 0 || a.B <= 0) return 0;
    if (a.J < 1 || a.J > 64) return SGS_ERR_BAD_JOINTS;
    if (((uintptr_t)a.A & 15) || ((uintptr_t)o.rotq & 15) || (o.T && ((uintptr_t)o.T & 15))) return SGS_ERR_MISALIGNED;
    const size_t smem = lbs_smem_bytes(a.B, a.J, false);
    if (smem > 220 * 1024) return SGS_ERR_CAPACITY;
    SGS_CUDA_OK(cudaFuncSetAttribute(lbs_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    lbs_fwd_kernel<<<(a.N + LBS_THREADS - 1) / LBS_THREADS, LBS_THREADS, smem, stream>>>(a, o);
    SGS_LAUNCH_OK();
    return 0;
}

// ------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(LBS_THREADS) lbs_bwd_kernel(LbsArgs a, LbsGrads g) {
    extern __shared__ __align__(16) char s_raw[];
    LbsSmem s = carve(s_raw, a.B, a.J, true);
    const int tid = threadIdx.x, lane = tid & 31;
    const int base = blockIdx.x * LBS_THREADS;
    const int rows = min(LBS_THREADS, a.N - base);
    stage_inputs(a, s, base, rows);
    const int n = base + tid;
    const bool live = n < a.N;
    const bool iso = a.rot == nullptr;
    float x = 0, y = 0, z = 0, s0 = 0, s1 = 0, s2 = 0;
    float Rc[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    if (live) {
        x = a.xyz[3 * (size_t)n]; y = a.xyz[3 * (size_t)n + 1]; z = a.xyz[3 * (size_t)n + 2];
        s0 = a.scales[3 * (size_t)n]; s1 = a.scales[3 * (size_t)n + 1]; s2 = a.scales[3 * (size_t)n + 2];
        if (!iso) {
#pragma unroll
            for (int k = 0; k < 9; k++) Rc[k] = a.rot[9 * (size_t)n + k];
        }
    }
    float dx = 0, dy = 0, dz = 0, ds0 = 0, ds1 = 0, ds2 = 0;
    float dRc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int b = 0; b < a.B; b++) {
        float dT[12];
#pragma unroll
        for (int k = 0; k < 12; k++) dT[k] = 0.0f;
        float gtr[3] = {0, 0, 0}, gss = 0.0f;
        if (live) {
            float T[12];
            blend_T(s, b, a.J, tid, T);
            const float* fr = s.frame + 16 * b;
            const size_t on = (size_t)b * a.N + n;
            float gx[3] = {g.g_xyz[3 * on], g.g_xyz[3 * on + 1], g.g_xyz[3 * on + 2]};
            const float4 gq4 = reinterpret_cast<const float4*>(g.g_rotq)[on];
            float gq[4] = {gq4.x, gq4.y, gq4.z, gq4.w};
            float gs[3] = {g.g_scales[3 * on], g.g_scales[3 * on + 1], g.g_scales[3 * on + 2]};
            float Rp[9];
            compose_rot(T, Rc, iso, Rp);
            if (a.ext_rot) {
                const float* eR = a.ext_rot + 9 * b;
                const float es = fr[4];
                const float t0 = gx[0], t1 = gx[1], t2 = gx[2];
                gx[0] = es * (eR[0] * t0 + eR[3] * t1 + eR[6] * t2);
                gx[1] = es * (eR[1] * t0 + eR[4] * t1 + eR[7] * t2);
                gx[2] = es * (eR[2] * t0 + eR[5] * t1 + eR[8] * t2);
                gs[0] *= es; gs[1] *= es; gs[2] *= es;
                float q[4], qo[4];
                mat_to_quat(Rp, q);
                quat_mul(fr + 8, q, qo);
                const float sg = qo[0] < 0.0f ? -1.0f : 1.0f;
                const float gr[4] = {sg * gq[0], sg * gq[1], sg * gq[2], sg * gq[3]};
                const float qc[4] = {fr[8], -fr[9], -fr[10], -fr[11]};
                quat_mul(qc, gr, gq);      // dL/dq = conj(q_ext) (x) dL/d(q_ext (x) q)
            }
            gtr[0] = gx[0]; gtr[1] = gx[1]; gtr[2] = gx[2];
            const float sm = a.smpl_scale ? fr[0] : 1.0f;
            const float v0 = T[0] * x + T[1] * y + T[2] * z + T[3];
            const float v1 = T[4] * x + T[5] * y + T[6] * z + T[7];
            const float v2 = T[8] * x + T[9] * y + T[10] * z + T[11];
            gss = gx[0] * v0 + gx[1] * v1 + gx[2] * v2 + gs[0] * s0 + gs[1] * s1 + gs[2] * s2;
            ds0 += gs[0] * sm; ds1 += gs[1] * sm; ds2 += gs[2] * sm;
            const float h0 = gx[0] * sm, h1 = gx[1] * sm, h2 = gx[2] * sm;   // dL/d(verts)
            float gR[9];
            mat_to_quat_bwd(Rp, gq, gR);
            // dT[:, :3] = h (x) x + gR Rc^T ; dT[:, 3] = h
            const float h[3] = {h0, h1, h2};
            const float xv[3] = {x, y, z};
#pragma unroll
            for (int r = 0; r < 3; r++) {
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    float rot = iso ? gR[3 * r + c]
                                    : gR[3 * r] * Rc[3 * c] + gR[3 * r + 1] * Rc[3 * c + 1] + gR[3 * r + 2] * Rc[3 * c + 2];
                    dT[4 * r + c] = h[r] * xv[c] + rot;
                }
                dT[4 * r + 3] = h[r];
            }
            if (g.g_T) {
                const float4* gT = reinterpret_cast<const float4*>(g.g_T) + on * 4;
#pragma unroll
                for (int r = 0; r < 3; r++) {
                    const float4 v = gT[r];
                    dT[4 * r] += v.x; dT[4 * r + 1] += v.y; dT[4 * r + 2] += v.z; dT[4 * r + 3] += v.w;
                }
            }
            // d xyz_canon += T3^T h ; d R_canon += T3^T gR
            dx += T[0] * h0 + T[4] * h1 + T[8] * h2;
            dy += T[1] * h0 + T[5] * h1 + T[9] * h2;
            dz += T[2] * h0 + T[6] * h1 + T[10] * h2;
            if (!iso) {
#pragma unroll
                for (int r = 0; r < 3; r++)
#pragma unroll
                    for (int c = 0; c < 3; c++)
                        dRc[3 * r + c] += T[r] * gR[c] + T[4 + r] * gR[3 + c] + T[8 + r] * gR[6 + c];
            }
        }
        // ---- frame-level sums: d transl, d smpl_scale ----
        if (g.d_transl) {
            const float t0 = warp_sum(gtr[0]), t1 = warp_sum(gtr[1]), t2 = warp_sum(gtr[2]);
            if (lane == 0) { atomicAdd(g.d_transl + 3 * b, t0); atomicAdd(g.d_transl + 3 * b + 1, t1); atomicAdd(g.d_transl + 3 * b + 2, t2); }
        }
        if (g.d_smpl_scale) {
            const float t = warp_sum(gss);
            if (lane == 0) atomicAdd(g.d_smpl_scale + b, t);
        }
        // ---- dA[b][j] = sum_n W[n][j] dT_n : CTA-level (J x 256) @ (256 x 12) ----
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 12; k++) s.dT[tid * 12 + k] = dT[k];
        __syncthreads();
        for (int oidx = tid; oidx < a.J * 12; oidx += LBS_THREADS) {
            const int j = oidx / 12, c = oidx - j * 12;
            float accv = 0.0f;
            for (int t = 0; t < rows; t++) accv = fmaf(s.W[(size_t)t * a.J + j], s.dT[t * 12 + c], accv);
            const int r = c >> 2, cc = c & 3;
            atomicAdd(g.d_A + ((size_t)b * a.J + j) * 16 + 4 * r + cc, accv);
        }
    }
    if (live) {
        g.d_xyz[3 * (size_t)n] = dx; g.d_xyz[3 * (size_t)n + 1] = dy; g.d_xyz[3 * (size_t)n + 2] = dz;
        g.d_scales[3 * (size_t)n] = ds0; g.d_scales[3 * (size_t)n + 1] = ds1; g.d_scales[3 * (size_t)n + 2] = ds2;
        if (g.d_rot) {
#pragma unroll
            for (int k = 0; k < 9; k++) g.d_rot[9 * (size_t)n + k] = dRc[k];
        }
    }
}

int launch_lbs_bwd(const LbsArgs& a, const LbsGrads& g, cudaStream_t stream) {
    if (a.N <= 0 || a.B <= 0) return 0;
    if (a.J < 1 || a.J > 64) return SGS_ERR_BAD_JOINTS;
    if (((uintptr_t)a.A & 15) || ((uintptr_t)g.g_rotq & 15) || (g.g_T && ((uintptr_t)g.g_T & 15))) return SGS_ERR_MISALIGNED;
    const size_t smem = lbs_smem_bytes(a.B, a.J, true);
    if (smem > 220 * 1024) return SGS_ERR_CAPACITY;
    SGS_CUDA_OK(cudaFuncSetAttribute(lbs_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    lbs_bwd_kernel<<<(a.N + LBS_THREADS - 1) / LBS_THREADS, LBS_THREADS, smem, stream>>>(a, g);
    SGS_LAUNCH_OK();
    return 0;
}

}  // namespace sgs
