// lbs_math.cuh -- per-Gaussian rotation math of the deformer, shared by lbs.cu and the fused
// LBS + rasterizer-geometry kernels.  References (paths under /root/reference):
//   sings/rec/utils/geometry/rotations.py:98-149 (matrix_to_quaternion), :393-407
//   (quaternion_multiply), :545-566 (rotation_6d_to_matrix), :514-542 (quaternion_to_axis_angle);
//   sings/rec/models/sings_hybrid.py:418-419 (R' = T[:3,:3] R_canon, q = matrix_to_quaternion(R')).
#pragma once
#include "common.cuh"
#include "kernels_lbs.h"

namespace sgs {

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
    // den >= 0.2: one reciprocal (numerator 1, operands in the normal range) instead of four
    // divisions whose zero numerators would take the slow IEEE path for the whole warp
    const float inv_den = 1.0f / (2.0f * fmaxf(qa[best], 0.1f));
#pragma unroll
    for (int k = 0; k < 4; k++) q[k] = c[k] * inv_den;
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
    const float inv_den = 1.0f / (2.0f * fmaxf(qb, 0.1f));
    float gc[4];
    float gden = 0.0f;
#pragma unroll
    for (int k = 0; k < 4; k++) { gc[k] = g[k] * inv_den; gden -= g[k] * c[k] * inv_den * inv_den; }
    // den = 2 max(qb, 0.1): gradient reaches qb only above the floor; c[best] = qb^2
    float gqb = (qb > 0.1f ? 2.0f * gden : 0.0f) + 2.0f * qb * gc[best];
    const float garg = arg[best] > 0.0f ? gqb * (0.5f / fmaxf(qb, 1e-30f)) : 0.0f;   // zero sub-gradient at 0
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
// 6D rotation (Zhou et al.) -> matrix by Gram-Schmidt, rows b1, b2, b3 (rotations.py:545-566:
// F.normalize(a1); a2 - (b1.a2) b1; F.normalize; cross; stack on dim -2) and its backward.
// F.normalize divides by max(||v||, 1e-12).
// ------------------------------------------------------------------------------------------
constexpr float NORMALIZE_EPS = 1e-12f;

__device__ __forceinline__ void rot6d_to_mat(const float* d6, float* R) {
    const float n1 = sqrtf(d6[0] * d6[0] + d6[1] * d6[1] + d6[2] * d6[2]);
    const float i1 = 1.0f / fmaxf(n1, NORMALIZE_EPS);
    const float b1[3] = {d6[0] * i1, d6[1] * i1, d6[2] * i1};
    const float d = b1[0] * d6[3] + b1[1] * d6[4] + b1[2] * d6[5];
    const float u[3] = {d6[3] - d * b1[0], d6[4] - d * b1[1], d6[5] - d * b1[2]};
    const float n2 = sqrtf(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
    const float i2 = 1.0f / fmaxf(n2, NORMALIZE_EPS);
    const float b2[3] = {u[0] * i2, u[1] * i2, u[2] * i2};
    R[0] = b1[0]; R[1] = b1[1]; R[2] = b1[2];
    R[3] = b2[0]; R[4] = b2[1]; R[5] = b2[2];
    R[6] = b1[1] * b2[2] - b1[2] * b2[1];
    R[7] = b1[2] * b2[0] - b1[0] * b2[2];
    R[8] = b1[0] * b2[1] - b1[1] * b2[0];
}

// v / max(||v||, eps) backward: g_v = (g - b (b.g)) / n above the floor, g / eps below it
__device__ __forceinline__ void normalize_bwd(const float* b, float n, const float* g, float* gv) {
    if (n > NORMALIZE_EPS) {
        const float bg = b[0] * g[0] + b[1] * g[1] + b[2] * g[2];
        const float inv = 1.0f / n;
#pragma unroll
        for (int k = 0; k < 3; k++) gv[k] = (g[k] - b[k] * bg) * inv;
    } else {
#pragma unroll
        for (int k = 0; k < 3; k++) gv[k] = g[k] * (1.0f / NORMALIZE_EPS);
    }
}

// dL/dR (row-major 9) -> dL/dd6 (6)
__device__ __forceinline__ void rot6d_to_mat_bwd(const float* d6, const float* gR, float* g6) {
    const float n1 = sqrtf(d6[0] * d6[0] + d6[1] * d6[1] + d6[2] * d6[2]);
    const float i1 = 1.0f / fmaxf(n1, NORMALIZE_EPS);
    const float b1[3] = {d6[0] * i1, d6[1] * i1, d6[2] * i1};
    const float a2[3] = {d6[3], d6[4], d6[5]};
    const float d = b1[0] * a2[0] + b1[1] * a2[1] + b1[2] * a2[2];
    const float u[3] = {a2[0] - d * b1[0], a2[1] - d * b1[1], a2[2] - d * b1[2]};
    const float n2 = sqrtf(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
    const float i2 = 1.0f / fmaxf(n2, NORMALIZE_EPS);
    const float b2[3] = {u[0] * i2, u[1] * i2, u[2] * i2};
    const float* g3 = gR + 6;
    // b3 = b1 x b2:  dL/db1 += b2 x g3,  dL/db2 += g3 x b1
    float gb1[3] = {gR[0] + (b2[1] * g3[2] - b2[2] * g3[1]), gR[1] + (b2[2] * g3[0] - b2[0] * g3[2]),
                    gR[2] + (b2[0] * g3[1] - b2[1] * g3[0])};
    const float gb2[3] = {gR[3] + (g3[1] * b1[2] - g3[2] * b1[1]), gR[4] + (g3[2] * b1[0] - g3[0] * b1[2]),
                          gR[5] + (g3[0] * b1[1] - g3[1] * b1[0])};
    float gu[3];
    normalize_bwd(b2, n2, gb2, gu);
    // u = a2 - (b1.a2) b1
    const float gub1 = gu[0] * b1[0] + gu[1] * b1[1] + gu[2] * b1[2];
#pragma unroll
    for (int k = 0; k < 3; k++) {
        g6[3 + k] = gu[k] - gub1 * b1[k];
        gb1[k] += -gub1 * a2[k] - d * gu[k];
    }
    normalize_bwd(b1, n1, gb1, g6);
}

// quaternion (real first) -> axis-angle (rotations.py:514-542): half = atan2(||v||, w),
// angle = 2 half, v / (sin(half)/angle), series 0.5 - angle^2/48 below 1e-6
__device__ __forceinline__ void quat_to_axis_angle(const float* q, float* aa) {
    const float n = sqrtf(q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    const float half = atan2f(n, q[0]);
    const float angle = 2.0f * half;
    const float s = fabsf(angle) < 1e-6f ? 0.5f - (angle * angle) / 48.0f : sinf(half) / angle;
    aa[0] = q[1] / s; aa[1] = q[2] / s; aa[2] = q[3] / s;
}

__device__ __forceinline__ void quat_to_axis_angle_bwd(const float* q, const float* g, float* gq) {
    const float n = sqrtf(q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    const float half = atan2f(n, q[0]);
    const float angle = 2.0f * half;
    const bool small = fabsf(angle) < 1e-6f;
    const float s = small ? 0.5f - (angle * angle) / 48.0f : sinf(half) / angle;
    // d s / d half: series -angle/24 per unit angle (x2); else (cos(half) angle - 2 sin(half)) / angle^2
    const float ds_dhalf = small ? -angle / 12.0f : (cosf(half) * angle - 2.0f * sinf(half)) / (angle * angle);
    const float gs = -(g[0] * q[1] + g[1] * q[2] + g[2] * q[3]) / (s * s);
    const float gh = gs * ds_dhalf;
    const float r2 = n * n + q[0] * q[0];
    const float gn = r2 > 0.0f ? gh * q[0] / r2 : 0.0f;
    gq[0] = r2 > 0.0f ? -gh * n / r2 : 0.0f;
    const float gn_over_n = n > 0.0f ? gn / n : 0.0f;      // torch.norm has a zero sub-gradient at 0
#pragma unroll
    for (int k = 0; k < 3; k++) gq[1 + k] = g[k] / s + gn_over_n * q[1 + k];
}


// R' = T[:3,:3] R_canon (row-major 3x3; T is 3x4 row-major); identity R_canon when iso
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

// ---- packed skinning weights of one Gaussian in registers ----
struct PackedW {
    float w[LBS_PACK_MAX_K];
    unsigned idx[LBS_PACK_MAX_K / 4];
};
// K is warp-uniform; loads are coalesced (slot-major tiles, kernels_lbs.h)
__device__ __forceinline__ void load_packed_w(const float* __restrict__ wq, const unsigned* __restrict__ iq, int K,
                                              int n, PackedW& p) {
    const size_t tile = (size_t)n / LBS_PACK_TILE, t = (size_t)n % LBS_PACK_TILE;
#pragma unroll
    for (int k = 0; k < LBS_PACK_MAX_K; k++)
        p.w[k] = k < K ? __ldg(wq + (tile * K + k) * LBS_PACK_TILE + t) : 0.0f;
#pragma unroll
    for (int k4 = 0; k4 < LBS_PACK_MAX_K / 4; k4++)
        p.idx[k4] = 4 * k4 < K ? __ldg(iq + (tile * (K / 4) + k4) * LBS_PACK_TILE + t) : 0u;
}
// T (3x4 row-major) = sum_k w_k A[j_k]; sA = rows 0..2 of every joint transform as float4 (shared memory)
__device__ __forceinline__ void blend_T_packed(const float4* sA, const PackedW& p, int K, float* T) {
#pragma unroll
    for (int k = 0; k < 12; k++) T[k] = 0.0f;
#pragma unroll
    for (int k = 0; k < LBS_PACK_MAX_K; k++) {
        if (k < K) {
            const float w = p.w[k];
            const unsigned j = (p.idx[k >> 2] >> (8 * (k & 3))) & 0xffu;
            const float4 r0 = sA[3 * j], r1 = sA[3 * j + 1], r2 = sA[3 * j + 2];
            T[0] = fmaf(w, r0.x, T[0]); T[1] = fmaf(w, r0.y, T[1]); T[2] = fmaf(w, r0.z, T[2]); T[3] = fmaf(w, r0.w, T[3]);
            T[4] = fmaf(w, r1.x, T[4]); T[5] = fmaf(w, r1.y, T[5]); T[6] = fmaf(w, r1.z, T[6]); T[7] = fmaf(w, r1.w, T[7]);
            T[8] = fmaf(w, r2.x, T[8]); T[9] = fmaf(w, r2.y, T[9]); T[10] = fmaf(w, r2.z, T[10]); T[11] = fmaf(w, r2.w, T[11]);
        }
    }
}

// canonical attributes of one Gaussian in registers (fused kernels: loaded ahead of the dependency wait)
struct CanonG {
    float x[3], s[3], Rc[9];
    PackedW pw;
};
__device__ __forceinline__ void load_canon(const LbsFuse& f, int n, CanonG& c) {
#pragma unroll
    for (int k = 0; k < 3; k++) { c.x[k] = __ldg(f.xyz + 3 * (size_t)n + k); c.s[k] = __ldg(f.scales + 3 * (size_t)n + k); }
    if (f.rot) {
        if (f.rot6d) {
            float d6[6];
#pragma unroll
            for (int k = 0; k < 6; k++) d6[k] = __ldg(f.rot + 6 * (size_t)n + k);
            rot6d_to_mat(d6, c.Rc);
        } else {
#pragma unroll
            for (int k = 0; k < 9; k++) c.Rc[k] = __ldg(f.rot + 9 * (size_t)n + k);
        }
    }
    load_packed_w(f.wq, f.iq, f.K, n, c.pw);
}
// the deform segment for one Gaussian (sings_hybrid.py:400-419; no ext_tfs in the fused path):
// v = (T [x;1]) * smpl_scale + transl, q = matrix_to_quaternion(T3 R_canon), s' = s * smpl_scale
__device__ __forceinline__ void deform_one(const LbsFuse& f, const float4* sA, const CanonG& c, float sm,
                                           const float* tr, float* v, float* q, float* sc, float* T) {
    blend_T_packed(sA, c.pw, f.K, T);
#pragma unroll
    for (int r = 0; r < 3; r++) v[r] = T[4 * r] * c.x[0] + T[4 * r + 1] * c.x[1] + T[4 * r + 2] * c.x[2] + T[4 * r + 3];
#pragma unroll
    for (int r = 0; r < 3; r++) sc[r] = c.s[r];
    if (f.smpl_scale) {
#pragma unroll
        for (int r = 0; r < 3; r++) { v[r] *= sm; sc[r] *= sm; }
    }
    if (f.transl) {
#pragma unroll
        for (int r = 0; r < 3; r++) v[r] += tr[r];
    }
    float Rp[9];
    compose_rot(T, c.Rc, f.rot == nullptr, Rp);
    mat_to_quat(Rp, q);
}

// ---- backward of the deform segment for one Gaussian: upstream gradients
// g_x (deformed mean), g_q (quaternion), g_s (deformed scale) -> canonical-parameter gradients
// (written to lf.d_xyz / d_scales / d_rot) and this Gaussian's dT (3x4) for dL/dA = sum_n W dT_n ----
__device__ __forceinline__ void lbs_bwd_one(const LbsFuse& lf, const float4* sA, const CanonG& cg, int idx,
                                            const float* g_x, const float* g_q, const float* g_s, float* dT) {
    const bool iso = lf.rot == nullptr;
    const float sm = lf.smpl_scale ? __ldg(lf.smpl_scale) : 1.0f;
    float T[12];
    blend_T_packed(sA, cg.pw, lf.K, T);
    float Rp[9], gR[9];
    compose_rot(T, cg.Rc, iso, Rp);
    mat_to_quat_bwd(Rp, g_q, gR);
    const float h[3] = {g_x[0] * sm, g_x[1] * sm, g_x[2] * sm};      // dL/d(verts)
#pragma unroll
    for (int r = 0; r < 3; r++) {
#pragma unroll
        for (int c = 0; c < 3; c++) {
            const float rot = iso ? gR[3 * r + c]
                                  : gR[3 * r] * cg.Rc[3 * c] + gR[3 * r + 1] * cg.Rc[3 * c + 1] + gR[3 * r + 2] * cg.Rc[3 * c + 2];
            dT[4 * r + c] = h[r] * cg.x[c] + rot;
        }
        dT[4 * r + 3] = h[r];
    }
    // d xyz_canon = T3^T h ; d scales = g_s * smpl_scale ; d R_canon = T3^T gR
    lf.d_xyz[3 * (size_t)idx] = T[0] * h[0] + T[4] * h[1] + T[8] * h[2];
    lf.d_xyz[3 * (size_t)idx + 1] = T[1] * h[0] + T[5] * h[1] + T[9] * h[2];
    lf.d_xyz[3 * (size_t)idx + 2] = T[2] * h[0] + T[6] * h[1] + T[10] * h[2];
    lf.d_scales[3 * (size_t)idx] = g_s[0] * sm; lf.d_scales[3 * (size_t)idx + 1] = g_s[1] * sm;
    lf.d_scales[3 * (size_t)idx + 2] = g_s[2] * sm;
    if (lf.d_rot && !iso) {
        float dRc[9];
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
            for (int c = 0; c < 3; c++)
                dRc[3 * r + c] = T[r] * gR[c] + T[4 + r] * gR[3 + c] + T[8 + r] * gR[6 + c];
        if (lf.rot6d) {
            float in6[6], g6[6];
#pragma unroll
            for (int k = 0; k < 6; k++) in6[k] = __ldg(lf.rot + 6 * (size_t)idx + k);
            rot6d_to_mat_bwd(in6, dRc, g6);
#pragma unroll
            for (int k = 0; k < 6; k++) lf.d_rot[6 * (size_t)idx + k] = g6[k];
        } else {
#pragma unroll
            for (int k = 0; k < 9; k++) lf.d_rot[9 * (size_t)idx + k] = dRc[k];
        }
    }
}

// Shared memory of the CTA-level reduction below, for a CTA of THREADS threads and J joints.
template <int THREADS>
__host__ __device__ constexpr size_t lbs_reduce_smem_bytes(int J, int K) {
    return (size_t)THREADS * 48 + (size_t)THREADS * K * 8 + (size_t)(THREADS / 32) * 2 * J * 48;      // two tiles per warp
}

// dL/dA[j] = sum_n W[n,j] dT_n and dL/dtransl = sum_n g_x over the CTA's Gaussians, with the
// packed weights: every warp folds its 32 rows into private J x 12 tiles (lane = (slot, float4
// group of the 3x4 block): the joints of one row are distinct, so a row is one conflict-free
// LDS.128 / 4 FMA / STS.128 per lane).  With up to five slots per row -- the usual case, K = 4 --
// the two halves of the warp take two rows at a time, each into its own tile; the CTA then adds
// the tiles and issues one atomic per non-zero entry.  `raw`: lbs_reduce_smem_bytes, 16-byte aligned.  All
// threads of the CTA call it (dT = 0 and weights 0 for threads without a Gaussian).
template <int THREADS>
__device__ __forceinline__ void lbs_bwd_reduce(const LbsFuse& lf, char* raw, const float* dT, const float* g_x,
                                               const PackedW& pw, bool live) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float4* const s_dT = reinterpret_cast<float4*>(raw);                                  // [THREADS][3]
    float2* const s_wj = reinterpret_cast<float2*>(raw + (size_t)THREADS * 48);           // [THREADS][K] (weight, joint bits)
    float* const s_part = reinterpret_cast<float*>(raw + (size_t)THREADS * 48 + (size_t)THREADS * lf.K * 8);   // [warps][J][12]
    for (int f = tid; f < (THREADS / 32) * 2 * lf.J * 12; f += THREADS) s_part[f] = 0.0f;
    if (lf.d_transl) {
        const float t0 = warp_sum(live ? g_x[0] : 0.0f), t1 = warp_sum(live ? g_x[1] : 0.0f), t2 = warp_sum(live ? g_x[2] : 0.0f);
        if (lane == 0) { atomicAdd(lf.d_transl, t0); atomicAdd(lf.d_transl + 1, t1); atomicAdd(lf.d_transl + 2, t2); }
    }
    s_dT[tid * 3] = make_float4(dT[0], dT[1], dT[2], dT[3]);
    s_dT[tid * 3 + 1] = make_float4(dT[4], dT[5], dT[6], dT[7]);
    s_dT[tid * 3 + 2] = make_float4(dT[8], dT[9], dT[10], dT[11]);
#pragma unroll
    for (int k = 0; k < LBS_PACK_MAX_K; k++)
        if (k < lf.K)
            s_wj[(size_t)tid * lf.K + k] = make_float2(live ? pw.w[k] : 0.0f,
                                                       __uint_as_float((pw.idx[k >> 2] >> (8 * (k & 3))) & 0xffu));
    __syncthreads();                           // the tiles are zero; (the rows below are this warp's own)
    const float4* const dT4 = s_dT + (size_t)warp * 32 * 3;
    const float2* const wj = s_wj + (size_t)warp * 32 * lf.K;
    if (lf.K <= 5) {
        // two rows per step: lanes 0..14 fold row 2 i, lanes 15..29 row 2 i + 1 (5 slots x 3 float4 groups each)
        const int h = lane / 15, q = lane - 15 * h, k = q / 3, g = q - 3 * k;
        const bool act = lane < 30 && k < lf.K;
        float4* const tile4 = reinterpret_cast<float4*>(s_part + ((size_t)warp * 2 + (h & 1)) * lf.J * 12);
#pragma unroll 4
        for (int i = 0; i < 16; i++) {
            const int l = 2 * i + h;
            if (act) {
                const float2 e = wj[l * lf.K + k];
                if (e.x != 0.0f) {
                    const unsigned j = __float_as_uint(e.y);
                    const float4 d = dT4[l * 3 + g];
                    float4 v = tile4[j * 3 + g];
                    v.x = fmaf(e.x, d.x, v.x); v.y = fmaf(e.x, d.y, v.y); v.z = fmaf(e.x, d.z, v.z); v.w = fmaf(e.x, d.w, v.w);
                    tile4[j * 3 + g] = v;
                }
            }
            __syncwarp();      // rows 2 i, 2 i + 1 are folded in before any lane reads a tile entry for the next two
        }
    } else {
        float4* const tile4 = reinterpret_cast<float4*>(s_part + (size_t)warp * 2 * lf.J * 12);
        const int kk = lane / 3, g = lane - 3 * kk;          // 10 slots x 3 float4 groups per round (lanes 30, 31 idle)
        for (int r0 = 0; r0 < lf.K; r0 += 10) {
            const int k = r0 + kk;
            const bool act = kk < 10 && k < lf.K;
#pragma unroll 4
            for (int l = 0; l < 32; l++) {
                if (act) {
                    const float2 e = wj[l * lf.K + k];
                    if (e.x != 0.0f) {
                        const unsigned j = __float_as_uint(e.y);
                        const float4 d = dT4[l * 3 + g];
                        float4 v = tile4[j * 3 + g];
                        v.x = fmaf(e.x, d.x, v.x); v.y = fmaf(e.x, d.y, v.y); v.z = fmaf(e.x, d.z, v.z); v.w = fmaf(e.x, d.w, v.w);
                        tile4[j * 3 + g] = v;
                    }
                }
                __syncwarp();      // row l is folded in before any lane reads a tile entry for row l + 1
            }
        }
    }
    __syncthreads();
    for (int oidx = tid; oidx < lf.J * 12; oidx += THREADS) {
        float accv = 0.0f;
#pragma unroll
        for (int w = 0; w < 2 * (THREADS / 32); w++) accv += s_part[(size_t)w * lf.J * 12 + oidx];
        const int j = oidx / 12, c = oidx - j * 12;
        if (accv != 0.0f) atomicAdd(lf.d_A + (size_t)j * 16 + 4 * (c >> 2) + (c & 3), accv);
    }
}

}  // namespace sgs
