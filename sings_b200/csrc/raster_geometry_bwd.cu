// raster_geometry_bwd.cu -- per-Gaussian backward of the rasterizer's geometry stage.
//
// Replaces, fused into one pass over HBM, the reference rasterizer's
//   [upstream] backward.cu computeCov2DCUDA + preprocessCUDA (K10+K11),
// reached from autograd of /root/reference/sings/rec/renderer/gs_renderer_single.py:87-95.
//
// Input: the (P,12) accumulator the blend backward reduced into (dL/dmean2D, dL/dconic,
// dL/dopacity, dL/dcolor).  Output: dL/dmeans3D, dL/dmeans2D, dL/dopacity, dL/dcolors,
// dL/dcov3D, dL/dsh, dL/dscales, dL/drotations -- every element written (zeros for culled
// Gaussians and for inactive SH coefficients), so the caller allocates with empty().
// SH rows are staged with cp.async like the forward; dL/dsh is produced in the same shared
// rows and written back with coalesced 16-byte stores.
#include "geom_math.cuh"
#include "kernels.h"
#include "lbs_math.cuh"

namespace sgs {

#ifndef SGS_GB_THREADS
#define SGS_GB_THREADS 256
#endif
constexpr int GB_THREADS = SGS_GB_THREADS;

// FUSE: the backward of the deform segment (what autograd computes for sings_hybrid.py:398-419,
// SURVEY.md Appendix B) runs in this kernel's epilogue on the gradients it has just produced --
// dL/d(mean, scale, quaternion) never travel through global memory; out come the canonical-
// parameter gradients, dL/dA and dL/dtransl.  dL/dA[j] = sum_n W[n,j] dT_n uses the packed
// weights: every warp folds its 32 rows, one after the other, into a private J x 12 tile in
// shared memory (lanes = (slot, entry) of the row: distinct addresses, plain read-modify-write);
// the CTA then adds its warps' tiles and issues one atomic per non-zero entry.
template <int D, bool HAS_SH, bool VEC16, bool FUSE>
#ifndef SGS_GB_CTAS
#define SGS_GB_CTAS (512 / GB_THREADS)
#endif
__global__ void __launch_bounds__(GB_THREADS, FUSE ? SGS_GB_CTAS : 1)
geometry_bwd_kernel(GeomBwdArgs b, const float4* __restrict__ rec, LbsFuse lf) {
    constexpr int NB = (D + 1) * (D + 1);
    constexpr int NVEC = HAS_SH ? sh_nvec(D) : 0;
    constexpr int S4 = HAS_SH ? sh_stride4(NVEC) : 0;
    extern __shared__ float4 s_sh[];                 // SH rows; FUSE: then A (J x 3 float4), dT, per-warp dA tiles
    __shared__ float s_cam[36];
    const GeomArgs& a = b.fwd;
    const int tid = threadIdx.x;
    const int base = blockIdx.x * GB_THREADS;
    const int idx = base + tid;
    const bool in_range = idx < a.P;
    const int rows = min(GB_THREADS, a.P - base);
    const size_t row_floats = (size_t)a.M * 3;
    // SH rows: staged ahead of the dependency wait when the caller vouches for early_params (the
    // preceding kernel, the blend backward, only produces `acc`)
    auto stage_sh = [&]() {
        if constexpr (HAS_SH) {
            if (VEC16) {
                const int total = rows * NVEC;
                for (int f = tid; f < total; f += GB_THREADS) {
                    int row = f / NVEC, col = f - row * NVEC;
                    cp_async16(&s_sh[row * S4 + col], a.shs + (size_t)(base + row) * row_floats + col * 4);
                }
            } else {
                float* s_f = reinterpret_cast<float*>(s_sh);
                const int total = rows * NB * 3;
                for (int f = tid; f < total; f += GB_THREADS) {
                    int row = f / (NB * 3), col = f - row * (NB * 3);
                    cp_async4(&s_f[row * S4 * 4 + col], a.shs + (size_t)(base + row) * row_floats + col);
                }
            }
            cp_async_commit();
        }
    };
    if (a.early_params) stage_sh();
    // FUSE: canonical attributes (model parameters) and this frame's joint transforms (written by
    // the forward, many kernels ago) are fetched ahead of the wait too
    float4* const s_A = s_sh + (size_t)GB_THREADS * S4;                               // [J][3], then the reduction's arrays
    CanonG cg;
    if constexpr (FUSE) {
#pragma unroll
        for (int k = 0; k < 9; k++) cg.Rc[k] = (k % 4 == 0) ? 1.0f : 0.0f;
        if (in_range) load_canon(lf, idx, cg);
        for (int f = tid; f < lf.J * 3; f += GB_THREADS)
            s_A[f] = reinterpret_cast<const float4*>(lf.A)[(f / 3) * 4 + f % 3];
    }
    pdl_sync();
    if (tid < 16) s_cam[tid] = a.view[tid];
    else if (tid < 32) s_cam[tid] = a.proj[tid - 16];
    else if (tid < 35) s_cam[tid] = a.campos[tid - 32];
    if (!a.early_params) stage_sh();
    __syncthreads();
    const float* V = s_cam;
    const float* Mx = s_cam + 16;
    const float fx = (float)a.W / (2.0f * a.tanfovx), fy = (float)a.H / (2.0f * a.tanfovy);

    const bool live = in_range && b.radii[idx] > 0;
    float dmean[3] = {0, 0, 0}, dscale[3] = {0, 0, 0}, drot[4] = {0, 0, 0, 0};
    float dcov[6] = {0, 0, 0, 0, 0, 0};
    float g2x = 0, g2y = 0, gop = 0, gcol[3] = {0, 0, 0};
    float px = 0, py = 0, pz = 0;
    unsigned flags = 0;
    if (live) {
        const float4* ac = reinterpret_cast<const float4*>(b.acc) + (size_t)idx * 3;
        const float4 a0 = ac[0], a1 = ac[1], a2 = ac[2];
        g2x = a0.x; g2y = a0.y;
        const float gxx = a0.z, gxy = a0.w, gyy = a1.x;
        gop = a1.y; gcol[0] = a1.z; gcol[1] = a1.w; gcol[2] = a2.x;
        flags = __float_as_uint(rec[4 * (size_t)idx + 3].w);
        px = a.means3D[3 * idx]; py = a.means3D[3 * idx + 1]; pz = a.means3D[3 * idx + 2];
        // ---- cov2D backward ----
        float c3[6];
        float4 q = make_float4(1, 0, 0, 0);
        float sc[3] = {0, 0, 0};
        if (a.cov3D_precomp) {
#pragma unroll
            for (int k = 0; k < 6; k++) c3[k] = a.cov3D_precomp[6 * (size_t)idx + k];
        } else {
            q = reinterpret_cast<const float4*>(a.rotations)[idx];
            sc[0] = a.scales[3 * idx]; sc[1] = a.scales[3 * idx + 1]; sc[2] = a.scales[3 * idx + 2];
            cov3d_from_scale_rot(sc[0], sc[1], sc[2], a.scale_modifier, q.x, q.y, q.z, q.w, c3);
        }
        const float pvx = xform_row(V, 0, px, py, pz), pvy = xform_row(V, 1, px, py, pz);
        const float pvz = xform_row(V, 2, px, py, pz);
        Cov2D cv;
        cov2d(pvx, pvy, pvz, fx, fy, a.tanfovx, a.tanfovy, c3, V, cv);
        const float ca = cv.a, cb = cv.b, cc = cv.c;
        const float denom = ca * cc - cb * cb;
        const float denom2inv = 1.0f / (denom * denom + 0.0000001f);
        float dL_da = 0, dL_db = 0, dL_dc = 0;
        if (denom2inv != 0.0f) {
            dL_da = denom2inv * (-cc * cc * gxx + 2.0f * cb * cc * gxy + (denom - ca * cc) * gyy);
            dL_dc = denom2inv * (-ca * ca * gyy + 2.0f * ca * cb * gxy + (denom - ca * cc) * gxx);
            dL_db = denom2inv * 2.0f * (cb * cc * gxx - (denom + 2.0f * cb * cb) * gxy + ca * cb * gyy);
            const float* A0 = cv.A0; const float* A1 = cv.A1;
            dcov[0] = A0[0] * A0[0] * dL_da + A0[0] * A1[0] * dL_db + A1[0] * A1[0] * dL_dc;
            dcov[3] = A0[1] * A0[1] * dL_da + A0[1] * A1[1] * dL_db + A1[1] * A1[1] * dL_dc;
            dcov[5] = A0[2] * A0[2] * dL_da + A0[2] * A1[2] * dL_db + A1[2] * A1[2] * dL_dc;
            dcov[1] = 2.0f * A0[0] * A0[1] * dL_da + (A0[0] * A1[1] + A0[1] * A1[0]) * dL_db + 2.0f * A1[0] * A1[1] * dL_dc;
            dcov[2] = 2.0f * A0[0] * A0[2] * dL_da + (A0[0] * A1[2] + A0[2] * A1[0]) * dL_db + 2.0f * A1[0] * A1[2] * dL_dc;
            dcov[4] = 2.0f * A0[2] * A0[1] * dL_da + (A0[1] * A1[2] + A0[2] * A1[1]) * dL_db + 2.0f * A1[1] * A1[2] * dL_dc;
        }
        float dJ00 = 0, dJ02 = 0, dJ11 = 0, dJ12 = 0;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const float dA0 = 2.0f * cv.B0[k] * dL_da + cv.B1[k] * dL_db;
            const float dA1 = 2.0f * cv.B1[k] * dL_dc + cv.B0[k] * dL_db;
            dJ00 += V[4 * k + 0] * dA0;
            dJ02 += V[4 * k + 2] * dA0;
            dJ11 += V[4 * k + 1] * dA1;
            dJ12 += V[4 * k + 2] * dA1;
        }
        const float tz = 1.0f / cv.tz, tz2 = tz * tz, tz3 = tz2 * tz;
        const float dtx = cv.xmul * -fx * tz2 * dJ02;
        const float dty = cv.ymul * -fy * tz2 * dJ12;
        const float dtz = -fx * tz2 * dJ00 - fy * tz2 * dJ11 + (2.0f * fx * cv.tx) * tz3 * dJ02 +
                          (2.0f * fy * cv.ty) * tz3 * dJ12;
#pragma unroll
        for (int k = 0; k < 3; k++)
            dmean[k] = V[4 * k] * dtx + V[4 * k + 1] * dty + V[4 * k + 2] * dtz;
        // ---- mean2D -> mean3D through the projection ----
        {
            const float mhx = xform_row(Mx, 0, px, py, pz), mhy = xform_row(Mx, 1, px, py, pz);
            const float mhw = xform_row(Mx, 3, px, py, pz);
            const float m_w = 1.0f / (mhw + 0.0000001f);
            const float mul1 = mhx * m_w * m_w, mul2 = mhy * m_w * m_w;
#pragma unroll
            for (int k = 0; k < 3; k++)
                dmean[k] += (Mx[4 * k] * m_w - Mx[4 * k + 3] * mul1) * g2x +
                            (Mx[4 * k + 1] * m_w - Mx[4 * k + 3] * mul2) * g2y;
        }
        // ---- cov3D backward: scale and (un-normalised) quaternion ----
        if (!a.cov3D_precomp) {
            float R[9];
            quat_to_R(q.x, q.y, q.z, q.w, R);
            const float s[3] = {a.scale_modifier * sc[0], a.scale_modifier * sc[1], a.scale_modifier * sc[2]};
            const float Dm[9] = {dcov[0], 0.5f * dcov[1], 0.5f * dcov[2], 0.5f * dcov[1], dcov[3],
                                 0.5f * dcov[4], 0.5f * dcov[2], 0.5f * dcov[4], dcov[5]};
            float E[9];
#pragma unroll
            for (int i = 0; i < 3; i++)
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    float t = 0;
#pragma unroll
                    for (int j = 0; j < 3; j++) t += Dm[3 * i + j] * (R[3 * j + k] * s[k]);
                    E[3 * i + k] = 2.0f * t;   // dL/dN
                }
#pragma unroll
            for (int k = 0; k < 3; k++)   // upstream quirk kept: no scale_modifier factor here
                dscale[k] = R[k] * E[k] + R[3 + k] * E[3 + k] + R[6 + k] * E[6 + k];
#pragma unroll
            for (int i = 0; i < 3; i++)
#pragma unroll
                for (int k = 0; k < 3; k++) E[3 * i + k] *= s[k];   // dL/dR
            const float r = q.x, x = q.y, y = q.z, z = q.w;
            drot[0] = 2.0f * (-z * E[1] + y * E[2] + z * E[3] - x * E[5] - y * E[6] + x * E[7]);
            drot[1] = 2.0f * (y * E[1] + z * E[2] + y * E[3] - 2.0f * x * E[4] - r * E[5] + z * E[6] + r * E[7] - 2.0f * x * E[8]);
            drot[2] = 2.0f * (-2.0f * y * E[0] + x * E[1] + r * E[2] + x * E[3] + z * E[5] - r * E[6] + z * E[7] - 2.0f * y * E[8]);
            drot[3] = 2.0f * (-2.0f * z * E[0] - r * E[1] + x * E[2] + r * E[3] - 2.0f * z * E[4] + y * E[5] + x * E[6] + y * E[7]);
        }
    }

    // ---- SH backward, in place in the staged shared rows ----
    if constexpr (HAS_SH) {
        cp_async_wait_all();
        __syncthreads();
        if (in_range) {
            float4* row = s_sh + tid * S4;
            if (live) {
                const float dox = SUB(px, V[32]), doy = SUB(py, V[33]), doz = SUB(pz, V[34]);
                const float len = SQRT(FMA(doz, doz, FMA(doy, doy, MUL(dox, dox))));
                const float inv = DIV(1.0f, len);
                const float x = MUL(dox, inv), y = MUL(doy, inv), z = MUL(doz, inv);
                float bas[NB];
                sh_basis<D>(x, y, z, bas);
                float sh[NVEC * 4];
#pragma unroll
                for (int j = 0; j < NVEC; j++) {
                    float4 v = row[j];
                    sh[4 * j] = v.x; sh[4 * j + 1] = v.y; sh[4 * j + 2] = v.z; sh[4 * j + 3] = v.w;
                }
                float dRGB[3];
#pragma unroll
                for (int ch = 0; ch < 3; ch++) dRGB[ch] = ((flags >> ch) & 1u) ? 0.0f : gcol[ch];
                float ddx = 0, ddy = 0, ddz = 0;
                if (D > 0) {
                    const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
#pragma unroll
                    for (int ch = 0; ch < 3; ch++) {
#define S(k) sh[3 * (k) + ch]
                        float gx_ = -SH_C1 * S(3), gy_ = -SH_C1 * S(1), gz_ = SH_C1 * S(2);
                        if (D > 1) {
                            gx_ += SH_C2_0 * y * S(4) + SH_C2_2 * 2.0f * -x * S(6) + SH_C2_3 * z * S(7) + SH_C2_4 * 2.0f * x * S(8);
                            gy_ += SH_C2_0 * x * S(4) + SH_C2_1 * z * S(5) + SH_C2_2 * 2.0f * -y * S(6) + SH_C2_4 * 2.0f * -y * S(8);
                            gz_ += SH_C2_1 * y * S(5) + SH_C2_2 * 2.0f * 2.0f * z * S(6) + SH_C2_3 * x * S(7);
                        }
                        if (D > 2) {
                            gx_ += SH_C3_0 * S(9) * 3.0f * 2.0f * xy + SH_C3_1 * S(10) * yz +
                                   SH_C3_2 * S(11) * -2.0f * xy + SH_C3_3 * S(12) * -3.0f * 2.0f * xz +
                                   SH_C3_4 * S(13) * (-3.0f * xx + 4.0f * zz - yy) +
                                   SH_C3_5 * S(14) * 2.0f * xz + SH_C3_6 * S(15) * 3.0f * (xx - yy);
                            gy_ += SH_C3_0 * S(9) * 3.0f * (xx - yy) + SH_C3_1 * S(10) * xz +
                                   SH_C3_2 * S(11) * (-3.0f * yy + 4.0f * zz - xx) +
                                   SH_C3_3 * S(12) * -3.0f * 2.0f * yz + SH_C3_4 * S(13) * -2.0f * xy +
                                   SH_C3_5 * S(14) * -2.0f * yz + SH_C3_6 * S(15) * -3.0f * 2.0f * xy;
                            gz_ += SH_C3_1 * S(10) * xy + SH_C3_2 * S(11) * 4.0f * 2.0f * yz +
                                   SH_C3_3 * S(12) * 3.0f * (2.0f * zz - xx - yy) +
                                   SH_C3_4 * S(13) * 4.0f * 2.0f * xz + SH_C3_5 * S(14) * (xx - yy);
                        }
#undef S
                        ddx += gx_ * dRGB[ch]; ddy += gy_ * dRGB[ch]; ddz += gz_ * dRGB[ch];
                    }
                    // dnormvdv
                    const float sum2 = dox * dox + doy * doy + doz * doz;
                    const float invsum32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
                    dmean[0] += ((sum2 - dox * dox) * ddx - doy * dox * ddy - doz * dox * ddz) * invsum32;
                    dmean[1] += (-dox * doy * ddx + (sum2 - doy * doy) * ddy - doz * doy * ddz) * invsum32;
                    dmean[2] += (-dox * doz * ddx - doy * doz * ddy + (sum2 - doz * doz) * ddz) * invsum32;
                }
                // dL/dsh[k][ch] = basis_k * dRGB[ch], written over the staged row
                float out[NVEC * 4];
#pragma unroll
                for (int f = 0; f < NVEC * 4; f++) out[f] = (f < NB * 3) ? bas[f / 3] * dRGB[f % 3] : 0.0f;
#pragma unroll
                for (int j = 0; j < NVEC; j++)
                    row[j] = make_float4(out[4 * j], out[4 * j + 1], out[4 * j + 2], out[4 * j + 3]);
            } else {
#pragma unroll
                for (int j = 0; j < NVEC; j++) row[j] = make_float4(0, 0, 0, 0);
            }
        }
        __syncthreads();
        // coalesced write-out of the whole (rows, M*3) block: staged part + zero tail
        if (b.dL_dsh) {
            float* dst = b.dL_dsh + (size_t)base * row_floats;
            const float* s_f = reinterpret_cast<const float*>(s_sh);
            if (VEC16) {
                const int vec_per_row = (int)(row_floats / 4);
                const int total = rows * vec_per_row;
                for (int f = tid; f < total; f += GB_THREADS) {
                    int rr = f / vec_per_row, col = f - rr * vec_per_row;
                    float4 v = col < NVEC ? s_sh[rr * S4 + col] : make_float4(0, 0, 0, 0);
                    // the float4 that straddles the end of the active coefficients is already
                    // zero-padded in shared memory
                    reinterpret_cast<float4*>(dst)[f] = v;
                }
            } else {
                const int total = rows * (int)row_floats;
                for (int f = tid; f < total; f += GB_THREADS) {
                    int rr = f / (int)row_floats, col = f - rr * (int)row_floats;
                    dst[f] = col < NB * 3 ? s_f[rr * S4 * 4 + col] : 0.0f;
                }
            }
        }
    }

    if (live && b.stat_accum) {
        // densification statistics of this view, fused (sings_hybrid.py:1013-1015,
        // gs_trainer.py:487-490): norm of the screen-space gradient, count, max radius
        b.stat_accum[idx] += sqrtf(g2x * g2x + g2y * g2y);
        b.stat_denom[idx] += 1.0f;
        b.stat_max_radii[idx] = fmaxf(b.stat_max_radii[idx], (float)b.radii[idx]);
    }
    if (in_range) {
        b.dL_dmeans2D[3 * idx] = g2x; b.dL_dmeans2D[3 * idx + 1] = g2y; b.dL_dmeans2D[3 * idx + 2] = 0.0f;
        b.dL_dopacity[idx] = gop;
        if (b.dL_dcolors) {
            b.dL_dcolors[3 * idx] = gcol[0]; b.dL_dcolors[3 * idx + 1] = gcol[1]; b.dL_dcolors[3 * idx + 2] = gcol[2];
        }
        if (b.dL_dcov3D) {
#pragma unroll
            for (int k = 0; k < 6; k++) b.dL_dcov3D[6 * (size_t)idx + k] = dcov[k];
        }
    }
    if constexpr (!FUSE) {
        if (in_range) {
            b.dL_dmeans3D[3 * idx] = dmean[0]; b.dL_dmeans3D[3 * idx + 1] = dmean[1]; b.dL_dmeans3D[3 * idx + 2] = dmean[2];
            if (b.dL_dscales) {
                b.dL_dscales[3 * idx] = dscale[0]; b.dL_dscales[3 * idx + 1] = dscale[1]; b.dL_dscales[3 * idx + 2] = dscale[2];
            }
            if (b.dL_drots)
                reinterpret_cast<float4*>(b.dL_drots)[idx] = make_float4(drot[0], drot[1], drot[2], drot[3]);
        }
    } else {
        // ---- deform-segment backward on g_x = dmean, g_q = drot, g_s = dscale ----
        float dT[12];
#pragma unroll
        for (int k = 0; k < 12; k++) dT[k] = 0.0f;
        if (in_range) lbs_bwd_one(lf, s_A, cg, idx, dmean, drot, dscale, dT);
        lbs_bwd_reduce<GB_THREADS>(lf, reinterpret_cast<char*>(s_A + 64 * 3), dT, dmean, cg.pw, in_range);
    }
}

template <int D, bool HAS_SH>
static int launch_gb_t(const GeomBwdArgs& b, const float4* rec, int blocks, bool vec16, cudaStream_t st, const LbsFuse* lf) {
    size_t smem = HAS_SH ? (size_t)GB_THREADS * sh_stride4(sh_nvec(D)) * 16 : 0;
    const LbsFuse none{};
    if (lf) {
        if constexpr (HAS_SH) {
            if (!vec16) return SGS_ERR_MISALIGNED;
            smem += (size_t)64 * 3 * 16 + lbs_reduce_smem_bytes<GB_THREADS>(lf->J, lf->K);
            auto k = geometry_bwd_kernel<D, true, true, true>;
            SGS_CUDA_OK(set_max_smem(k, smem));
            SGS_CUDA_OK(launch_pdl(k, blocks, GB_THREADS, smem, st, b, rec, *lf));
            return 0;
        } else {
            return SGS_ERR_BAD_ARG;
        }
    }
    if (vec16) {
        auto k = geometry_bwd_kernel<D, HAS_SH, true, false>;
        SGS_CUDA_OK(set_max_smem(k, smem));
        SGS_CUDA_OK(launch_pdl(k, blocks, GB_THREADS, smem, st, b, rec, none));
    } else {
        auto k = geometry_bwd_kernel<D, HAS_SH, false, false>;
        SGS_CUDA_OK(set_max_smem(k, smem));
        SGS_CUDA_OK(launch_pdl(k, blocks, GB_THREADS, smem, st, b, rec, none));
    }
    return 0;
}

int launch_geometry_bwd(const GeomBwdArgs& b, const char* geom, cudaStream_t stream, const LbsFuse* lf) {
    const GeomArgs& a = b.fwd;
    if (a.P <= 0) return 0;
    const float4* rec = reinterpret_cast<const float4*>(geom);
    const int blocks = (a.P + GB_THREADS - 1) / GB_THREADS;
    const bool has_sh = a.colors_precomp == nullptr;
    if (!has_sh) return launch_gb_t<0, false>(b, rec, blocks, false, stream, lf);
    const bool vec16 = ((a.M * 3) % 4 == 0) && (((uintptr_t)a.shs & 15) == 0) &&
                       (((uintptr_t)b.dL_dsh & 15) == 0) && (a.M * 3 >= sh_nvec(a.D) * 4);
    switch (a.D) {
        case 0: return launch_gb_t<0, true>(b, rec, blocks, vec16, stream, lf);
        case 1: return launch_gb_t<1, true>(b, rec, blocks, vec16, stream, lf);
        case 2: return launch_gb_t<2, true>(b, rec, blocks, vec16, stream, lf);
        case 3: return launch_gb_t<3, true>(b, rec, blocks, vec16, stream, lf);
        default: return SGS_ERR_BAD_SH_DEGREE;
    }
}

}  // namespace sgs
