// geom_math.cuh -- per-Gaussian projection math shared by the geometry forward and backward
// kernels.  Semantics: SURVEY.md Appendix A.2 / A.6 ([upstream] forward.cu computeCov3D,
// computeCov2D, computeColorFromSH; auxiliary.h getRect, ndc2Pix).  Operation order is the
// numeric contract of DESIGN.md and is spelled with explicit _rn intrinsics.
#pragma once
#include "common.cuh"

namespace sgs {

#define SH_C0 0.28209479177387814f
#define SH_C1 0.4886025119029199f
#define SH_C2_0 1.0925484305920792f
#define SH_C2_1 -1.0925484305920792f
#define SH_C2_2 0.31539156525252005f
#define SH_C2_3 -1.0925484305920792f
#define SH_C2_4 0.5462742152960396f
#define SH_C3_0 -0.5900435899266435f
#define SH_C3_1 2.890611442640554f
#define SH_C3_2 -0.4570457994644658f
#define SH_C3_3 0.3731763325901154f
#define SH_C3_4 -0.4570457994644658f
#define SH_C3_5 1.445305721320277f
#define SH_C3_6 -0.5900435899266435f

#define MUL(a, b) __fmul_rn((a), (b))
#define ADD(a, b) __fadd_rn((a), (b))
#define SUB(a, b) __fsub_rn((a), (b))
#define FMA(a, b, c) __fmaf_rn((a), (b), (c))
#define DIV(a, b) __fdiv_rn((a), (b))
#define SQRT(a) __fsqrt_rn((a))

// SH basis b_k(dir) for k < (D+1)^2
template <int D>
__device__ __forceinline__ void sh_basis(float x, float y, float z, float* b) {
    b[0] = SH_C0;
    if (D > 0) {
        b[1] = -MUL(SH_C1, y);
        b[2] = MUL(SH_C1, z);
        b[3] = -MUL(SH_C1, x);
    }
    if (D > 1) {
        float xx = MUL(x, x), yy = MUL(y, y), zz = MUL(z, z);
        float xy = MUL(x, y), yz = MUL(y, z), xz = MUL(x, z);
        b[4] = MUL(SH_C2_0, xy);
        b[5] = MUL(SH_C2_1, yz);
        b[6] = MUL(SH_C2_2, SUB(SUB(MUL(2.0f, zz), xx), yy));
        b[7] = MUL(SH_C2_3, xz);
        b[8] = MUL(SH_C2_4, SUB(xx, yy));
        if (D > 2) {
            b[9] = MUL(MUL(SH_C3_0, y), SUB(MUL(3.0f, xx), yy));
            b[10] = MUL(MUL(SH_C3_1, xy), z);
            b[11] = MUL(MUL(SH_C3_2, y), SUB(SUB(MUL(4.0f, zz), xx), yy));
            b[12] = MUL(MUL(SH_C3_3, z), SUB(SUB(MUL(2.0f, zz), MUL(3.0f, xx)), MUL(3.0f, yy)));
            b[13] = MUL(MUL(SH_C3_4, x), SUB(SUB(MUL(4.0f, zz), xx), yy));
            b[14] = MUL(MUL(SH_C3_5, z), SUB(xx, yy));
            b[15] = MUL(MUL(SH_C3_6, x), SUB(xx, MUL(3.0f, yy)));
        }
    }
}

// quaternion (r,x,y,z), NOT normalised -> rotation-matrix polynomial, row-major R[9]
__device__ __forceinline__ void quat_to_R(float r, float x, float y, float z, float* R) {
    R[0] = FMA(-2.0f, FMA(z, z, MUL(y, y)), 1.0f);
    R[1] = MUL(2.0f, FMA(x, y, -MUL(r, z)));
    R[2] = MUL(2.0f, FMA(x, z, MUL(r, y)));
    R[3] = MUL(2.0f, FMA(x, y, MUL(r, z)));
    R[4] = FMA(-2.0f, FMA(z, z, MUL(x, x)), 1.0f);
    R[5] = MUL(2.0f, FMA(y, z, -MUL(r, x)));
    R[6] = MUL(2.0f, FMA(x, z, -MUL(r, y)));
    R[7] = MUL(2.0f, FMA(y, z, MUL(r, x)));
    R[8] = FMA(-2.0f, FMA(y, y, MUL(x, x)), 1.0f);
}

// Sigma = (R diag(mod*s)) (R diag(mod*s))^T, upper triangle (00,01,02,11,12,22)
__device__ __forceinline__ void cov3d_from_scale_rot(float s0, float s1, float s2, float mod,
                                                     float qr, float qx, float qy, float qz,
                                                     float* c) {
    float R[9];
    quat_to_R(qr, qx, qy, qz, R);
    float sx = MUL(mod, s0), sy = MUL(mod, s1), sz = MUL(mod, s2);
    float N[9];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        N[3 * i + 0] = MUL(R[3 * i + 0], sx);
        N[3 * i + 1] = MUL(R[3 * i + 1], sy);
        N[3 * i + 2] = MUL(R[3 * i + 2], sz);
    }
#define SGS_DOT3(i, j) \
    FMA(N[3 * i + 2], N[3 * j + 2], FMA(N[3 * i + 1], N[3 * j + 1], MUL(N[3 * i], N[3 * j])))
    c[0] = SGS_DOT3(0, 0); c[1] = SGS_DOT3(0, 1); c[2] = SGS_DOT3(0, 2);
    c[3] = SGS_DOT3(1, 1); c[4] = SGS_DOT3(1, 2); c[5] = SGS_DOT3(2, 2);
#undef SGS_DOT3
}

struct Cov2D {
    float tx, ty, tz, xmul, ymul;
    float A0[3], A1[3], B0[3], B1[3];
    float a, b, c;
};

// EWA projection: cov2D = (J Rv) Sigma (J Rv)^T + 0.3 I.   V = column-major view matrix.
__device__ __forceinline__ void cov2d(float pvx, float pvy, float pvz, float fx, float fy,
                                      float tanx, float tany, const float* c3, const float* V,
                                      Cov2D& o) {
    float limx = MUL(1.3f, tanx), limy = MUL(1.3f, tany);
    float txtz = DIV(pvx, pvz), tytz = DIV(pvy, pvz);
    o.xmul = (txtz < -limx || txtz > limx) ? 0.0f : 1.0f;
    o.ymul = (tytz < -limy || tytz > limy) ? 0.0f : 1.0f;
    float tx = MUL(fminf(limx, fmaxf(-limx, txtz)), pvz);
    float ty = MUL(fminf(limy, fmaxf(-limy, tytz)), pvz);
    float tz = pvz;
    o.tx = tx; o.ty = ty; o.tz = tz;
    float tz2 = MUL(tz, tz);
    float J00 = DIV(fx, tz), J02 = DIV(-MUL(fx, tx), tz2);
    float J11 = DIV(fy, tz), J12 = DIV(-MUL(fy, ty), tz2);
#pragma unroll
    for (int cc = 0; cc < 3; cc++) {
        float r0 = V[4 * cc + 0], r1 = V[4 * cc + 1], r2 = V[4 * cc + 2];
        o.A0[cc] = FMA(J02, r2, MUL(J00, r0));
        o.A1[cc] = FMA(J12, r2, MUL(J11, r1));
    }
    const float s00 = c3[0], s01 = c3[1], s02 = c3[2], s11 = c3[3], s12 = c3[4], s22 = c3[5];
    o.B0[0] = FMA(s02, o.A0[2], FMA(s01, o.A0[1], MUL(s00, o.A0[0])));
    o.B0[1] = FMA(s12, o.A0[2], FMA(s11, o.A0[1], MUL(s01, o.A0[0])));
    o.B0[2] = FMA(s22, o.A0[2], FMA(s12, o.A0[1], MUL(s02, o.A0[0])));
    o.B1[0] = FMA(s02, o.A1[2], FMA(s01, o.A1[1], MUL(s00, o.A1[0])));
    o.B1[1] = FMA(s12, o.A1[2], FMA(s11, o.A1[1], MUL(s01, o.A1[0])));
    o.B1[2] = FMA(s22, o.A1[2], FMA(s12, o.A1[1], MUL(s02, o.A1[0])));
    o.a = ADD(FMA(o.A0[2], o.B0[2], FMA(o.A0[1], o.B0[1], MUL(o.A0[0], o.B0[0]))), 0.3f);
    o.b = FMA(o.A0[2], o.B1[2], FMA(o.A0[1], o.B1[1], MUL(o.A0[0], o.B1[0])));
    o.c = ADD(FMA(o.A1[2], o.B1[2], FMA(o.A1[1], o.B1[1], MUL(o.A1[0], o.B1[0]))), 0.3f);
}

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(hi, max(lo, v)); }

// [upstream] auxiliary.h getRect (float -> int conversion truncates, saturating)
__device__ __forceinline__ void get_rect(float px, float py, int radius, int gx, int gy, int& x0,
                                         int& y0, int& x1, int& y1) {
    float r = (float)radius;
    x0 = clampi(__float2int_rz(DIV(SUB(px, r), 16.0f)), 0, gx);
    y0 = clampi(__float2int_rz(DIV(SUB(py, r), 16.0f)), 0, gy);
    x1 = clampi(__float2int_rz(DIV(SUB(ADD(ADD(px, r), 16.0f), 1.0f), 16.0f)), 0, gx);
    y1 = clampi(__float2int_rz(DIV(SUB(ADD(ADD(py, r), 16.0f), 1.0f), 16.0f)), 0, gy);
}

// SH row stride in shared memory (in float4 units) that keeps per-thread LDS.128 reads
// bank-conflict free: stride/4 words must be odd in units of float4.
__host__ __device__ constexpr int sh_stride4(int nvec) { return (nvec % 2 == 0) ? nvec + 1 : nvec; }
// number of float4 that hold the (D+1)^2 * 3 active floats of a row
__host__ __device__ constexpr int sh_nvec(int D) { return ((D + 1) * (D + 1) * 3 + 3) / 4; }

}  // namespace sgs
