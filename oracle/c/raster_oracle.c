/*
 * raster_oracle.c -- CPU restatement of the 3DGS differentiable tile rasterizer.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under sings_b200/ or diff_gaussian_rasterization/
 * may import, link or execute this file; it is the checker for tests/, smoke() and the
 * cpu_baseline / --impl reference legs of bench.py.
 *
 * PARITY STATUS: "parity unpinned".  The algorithm restated here lives in a third-party
 * dependency of the reference that is NOT in /root/reference and has no pinned version:
 *   diff-gaussian-rasterization (graphdeco-inria), installed unpinned by
 *   /root/reference/install_all.sh:22, imported at
 *   /root/reference/sings/rec/renderer/gs_renderer_single.py:6-9, called at :69-95.
 * The reference holds no golden vectors, tests or fixtures for it.
 * This file restates its published algorithm ("[upstream]"
 * cuda_rasterizer/{forward,backward,rasterizer_impl}.cu, auxiliary.h) in plain C.
 * In-tree anchors that ARE checked (tests/test_oracle_*.py): the SH polynomial and
 * constants against sings/rec/utils/visualize/spherical_harmonics.py:30-47,61-125, the
 * camera conventions of sings/rec/utils/graphics.py:65-85 and datasets/utils.py:37-39, and
 * the gradients against a float64 torch-autograd restatement (oracle/raster_ref64.py).
 *
 * NUMERIC CONTRACT (DESIGN.md "Numeric contract"): every value that decides a sort key,
 * a tile rectangle, a tile range or a contributor count is produced by a fixed sequence
 * of IEEE-754 binary32 operations (mul, add, fma, div, sqrt, rint, ceil, trunc), written
 * out explicitly below.  Build with -ffp-contract=off so the only fused operations are
 * the fmaf() calls; the CUDA kernels are built with -fmad=false and spell the same
 * sequence, so the forward pass is reproducible bit for bit on CPU and GPU.  exp() is a
 * fixed polynomial (expneg below: magic-number rounding, Cody-Waite reduction, degree-5
 * Estrin polynomial, max relative error 2.0e-7), not libm, for the same reason.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define TILE 16

/* [upstream] auxiliary.h: SH constants (same values as
 * sings/rec/utils/visualize/spherical_harmonics.py:30-47). */
static const float SH_C0 = 0.28209479177387814f;
static const float SH_C1 = 0.4886025119029199f;
static const float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                               -1.0925484305920792f, 0.5462742152960396f};
static const float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                               0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                               -0.5900435899266435f};

int ro_version(void) { return 1; }

void ro_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int ro_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* exp(x) for x <= 0: Cody-Waite reduction + degree-5 polynomial (Cephes expf
 * coefficients), fixed op order.  Returns 0 below -80 (alpha would be < 1e-34). */
static inline float expneg(float x) {
    x = fmaxf(x, -80.0f);
    /* one fused multiply-add: x log2(e) rounded to the nearest integer by the 1.5 * 2^23 magic
     * addend (the packed f32x2 form of the CUDA kernels contracts a mul + add here anyway) */
    const float r = fmaf(x, 1.44269504088896341f, 12582912.0f);
    const float n = r + -12582912.0f;
    float g = fmaf(n, -0.693359375f, x);
    g = fmaf(n, 2.12194440e-4f, g);
    const float g2 = g * g;
    const float a = fmaf(9.9999970198e-01f, g, 1.0f);
    const float b = fmaf(1.6667643189e-01f, g, 4.9999141693e-01f);
    const float c = fmaf(8.2901455462e-03f, g, 4.1898854077e-02f);
    const float p = fmaf(fmaf(c, g2, b), g2, a);
    union { float f; int32_t i; uint32_t u; } up, ur;
    up.f = p;
    ur.f = r;
    up.u += ur.u << 23;
    return up.f;
}

float ro_expneg(float x) { return expneg(x); }

/* [upstream] auxiliary.h transformPoint4x3 / 4x4: column-major 4x4, row r of M*p. */
static inline float xform_row(const float* m, int r, float x, float y, float z) {
    float t = m[r] * x;
    t = fmaf(m[4 + r], y, t);
    t = fmaf(m[8 + r], z, t);
    return t + m[12 + r];
}

/* [upstream] rasterizer_impl.cu getHigherMsb: number of tile-id bits that take part
 * in the sort (4096 -> 13, 8160 -> 13, 16384 -> 15, 1024 -> 11). */
int ro_higher_msb(uint32_t n) {
    uint32_t msb = 16, step = 16;
    while (step > 1) {
        step /= 2;
        if (n >> msb) msb += step; else msb -= step;
    }
    if (n >> msb) msb++;
    return (int)msb;
}

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

static inline int f2i_trunc(float v) {
    /* C float->int truncation; saturate like the GPU cvt.rzi does, NaN -> 0 */
    if (!(v == v)) return 0;
    if (v >= 2147483520.0f) return 2147483647;
    if (v <= -2147483648.0f) return (-2147483647 - 1);
    return (int)v;
}

/* [upstream] auxiliary.h getRect */
static inline void get_rect(float px, float py, int radius, int gx, int gy, int* x0, int* y0,
                            int* x1, int* y1) {
    float r = (float)radius;
    *x0 = clampi(f2i_trunc((px - r) / 16.0f), 0, gx);
    *y0 = clampi(f2i_trunc((py - r) / 16.0f), 0, gy);
    *x1 = clampi(f2i_trunc((((px + r) + 16.0f) - 1.0f) / 16.0f), 0, gx);
    *y1 = clampi(f2i_trunc((((py + r) + 16.0f) - 1.0f) / 16.0f), 0, gy);
}

/* SH basis values b_k(dir), k < (D+1)^2 -- [upstream] forward.cu computeColorFromSH,
 * polynomial as in spherical_harmonics.py:87-113. */
static inline void sh_basis(int D, float x, float y, float z, float* b) {
    b[0] = SH_C0;
    if (D > 0) {
        b[1] = -(SH_C1 * y);
        b[2] = SH_C1 * z;
        b[3] = -(SH_C1 * x);
        if (D > 1) {
            float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
            b[4] = SH_C2[0] * xy;
            b[5] = SH_C2[1] * yz;
            b[6] = SH_C2[2] * ((2.0f * zz - xx) - yy);
            b[7] = SH_C2[3] * xz;
            b[8] = SH_C2[4] * (xx - yy);
            if (D > 2) {
                b[9] = (SH_C3[0] * y) * (3.0f * xx - yy);
                b[10] = (SH_C3[1] * xy) * z;
                b[11] = (SH_C3[2] * y) * ((4.0f * zz - xx) - yy);
                b[12] = (SH_C3[3] * z) * ((2.0f * zz - 3.0f * xx) - 3.0f * yy);
                b[13] = (SH_C3[4] * x) * ((4.0f * zz - xx) - yy);
                b[14] = (SH_C3[5] * z) * (xx - yy);
                b[15] = (SH_C3[6] * x) * (xx - 3.0f * yy);
            }
        }
    }
}

/* quaternion (r,x,y,z), NOT normalised -> R_std entries ([upstream] forward.cu computeCov3D;
 * glm's column-major constructor makes upstream's R the transpose of this one). */
static inline void quat_to_R(float r, float x, float y, float z, float* R) {
    R[0] = fmaf(-2.0f, fmaf(z, z, y * y), 1.0f);
    R[1] = 2.0f * fmaf(x, y, -(r * z));
    R[2] = 2.0f * fmaf(x, z, r * y);
    R[3] = 2.0f * fmaf(x, y, r * z);
    R[4] = fmaf(-2.0f, fmaf(z, z, x * x), 1.0f);
    R[5] = 2.0f * fmaf(y, z, -(r * x));
    R[6] = 2.0f * fmaf(x, z, -(r * y));
    R[7] = 2.0f * fmaf(y, z, r * x);
    R[8] = fmaf(-2.0f, fmaf(y, y, x * x), 1.0f);
}

/* Sigma = (R diag(s)) (R diag(s))^T, upper triangle (00,01,02,11,12,22). */
static inline void cov3d_from_scale_rot(const float* s, float mod, const float* q, float* c) {
    float R[9];
    quat_to_R(q[0], q[1], q[2], q[3], R);
    float sx = mod * s[0], sy = mod * s[1], sz = mod * s[2];
    float N[9];
    for (int i = 0; i < 3; i++) {
        N[3 * i + 0] = R[3 * i + 0] * sx;
        N[3 * i + 1] = R[3 * i + 1] * sy;
        N[3 * i + 2] = R[3 * i + 2] * sz;
    }
#define DOT3(i, j) fmaf(N[3 * i + 2], N[3 * j + 2], fmaf(N[3 * i + 1], N[3 * j + 1], N[3 * i] * N[3 * j]))
    c[0] = DOT3(0, 0); c[1] = DOT3(0, 1); c[2] = DOT3(0, 2);
    c[3] = DOT3(1, 1); c[4] = DOT3(1, 2); c[5] = DOT3(2, 2);
#undef DOT3
}

typedef struct {
    float tx, ty, tz;       /* clamped view-space point */
    float xmul, ymul;       /* 0 where the clamp was active (backward only) */
    float A0[3], A1[3];     /* rows of A = J * Rv */
    float B0[3], B1[3];     /* Sigma * A0^T, Sigma * A1^T */
    float a, b, c;          /* cov2D incl. +0.3 */
} cov2d_t;

/* [upstream] forward.cu computeCov2D: EWA projection, cov2D = (J Rv) Sigma (J Rv)^T + 0.3 I */
static inline void cov2d(const float* pv, float fx, float fy, float tanx, float tany,
                         const float* c3, const float* V, cov2d_t* o) {
    float limx = 1.3f * tanx, limy = 1.3f * tany;
    float txtz = pv[0] / pv[2], tytz = pv[1] / pv[2];
    o->xmul = (txtz < -limx || txtz > limx) ? 0.0f : 1.0f;
    o->ymul = (tytz < -limy || tytz > limy) ? 0.0f : 1.0f;
    float tx = fminf(limx, fmaxf(-limx, txtz)) * pv[2];
    float ty = fminf(limy, fmaxf(-limy, tytz)) * pv[2];
    float tz = pv[2];
    o->tx = tx; o->ty = ty; o->tz = tz;
    float tz2 = tz * tz;
    float J00 = fx / tz, J02 = -(fx * tx) / tz2;
    float J11 = fy / tz, J12 = -(fy * ty) / tz2;
    for (int cc = 0; cc < 3; cc++) {
        float r0 = V[4 * cc + 0], r1 = V[4 * cc + 1], r2 = V[4 * cc + 2];
        o->A0[cc] = fmaf(J02, r2, J00 * r0);
        o->A1[cc] = fmaf(J12, r2, J11 * r1);
    }
    const float s00 = c3[0], s01 = c3[1], s02 = c3[2], s11 = c3[3], s12 = c3[4], s22 = c3[5];
    const float* A0 = o->A0; const float* A1 = o->A1;
    o->B0[0] = fmaf(s02, A0[2], fmaf(s01, A0[1], s00 * A0[0]));
    o->B0[1] = fmaf(s12, A0[2], fmaf(s11, A0[1], s01 * A0[0]));
    o->B0[2] = fmaf(s22, A0[2], fmaf(s12, A0[1], s02 * A0[0]));
    o->B1[0] = fmaf(s02, A1[2], fmaf(s01, A1[1], s00 * A1[0]));
    o->B1[1] = fmaf(s12, A1[2], fmaf(s11, A1[1], s01 * A1[0]));
    o->B1[2] = fmaf(s22, A1[2], fmaf(s12, A1[1], s02 * A1[0]));
    o->a = fmaf(A0[2], o->B0[2], fmaf(A0[1], o->B0[1], A0[0] * o->B0[0])) + 0.3f;
    o->b = fmaf(A0[2], o->B1[2], fmaf(A0[1], o->B1[1], A0[0] * o->B1[0]));
    o->c = fmaf(A1[2], o->B1[2], fmaf(A1[1], o->B1[1], A1[0] * o->B1[0])) + 0.3f;
}

/* ------------------------------------------------------------------------------------
 * [upstream] forward.cu preprocessCUDA.  Outputs are zero/defined for
 * every Gaussian: radii=0 and tiles_touched=0 mean "not rendered".
 * shs: (P, M, 3) or NULL; colors_precomp: (P,3) or NULL; cov3D_precomp: (P,6) or NULL.
 * ---------------------------------------------------------------------------------- */
void ro_preprocess(int P, int D, int M, const float* means3D, const float* scales, float mod,
                   const float* rotations, const float* opacities, const float* shs,
                   const float* colors_precomp, const float* cov3D_precomp, const float* view,
                   const float* proj, const float* campos, int W, int H, float tanfovx,
                   float tanfovy, int* radii, float* xy, float* depths, float* cov3D,
                   float* conic_opacity, float* rgb, uint8_t* clamped, uint32_t* tiles_touched) {
    const float fx = (float)W / (2.0f * tanfovx), fy = (float)H / (2.0f * tanfovy);
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < P; i++) {
        radii[i] = 0; tiles_touched[i] = 0;
        xy[2 * i] = xy[2 * i + 1] = 0.0f; depths[i] = 0.0f;
        for (int k = 0; k < 6; k++) cov3D[6 * i + k] = 0.0f;
        for (int k = 0; k < 4; k++) conic_opacity[4 * i + k] = 0.0f;
        for (int k = 0; k < 3; k++) { rgb[3 * i + k] = 0.0f; clamped[3 * i + k] = 0; }
        const float px = means3D[3 * i], py = means3D[3 * i + 1], pz = means3D[3 * i + 2];
        float pv[3] = {xform_row(view, 0, px, py, pz), xform_row(view, 1, px, py, pz),
                       xform_row(view, 2, px, py, pz)};
        if (!(pv[2] > 0.2f)) continue;   /* in_frustum: p_view.z <= 0.2 culls (NaN culls too) */
        float phx = xform_row(proj, 0, px, py, pz), phy = xform_row(proj, 1, px, py, pz);
        float phw = xform_row(proj, 3, px, py, pz);
        float pw = 1.0f / (phw + 0.0000001f);
        float ppx = phx * pw, ppy = phy * pw;
        float c3[6];
        if (cov3D_precomp) memcpy(c3, cov3D_precomp + 6 * i, sizeof c3);
        else cov3d_from_scale_rot(scales + 3 * i, mod, rotations + 4 * i, c3);
        cov2d_t cv;
        cov2d(pv, fx, fy, tanfovx, tanfovy, c3, view, &cv);
        float det = fmaf(cv.a, cv.c, -(cv.b * cv.b));
        if (det == 0.0f) continue;
        float det_inv = 1.0f / det;
        float conA = cv.c * det_inv, conB = -cv.b * det_inv, conC = cv.a * det_inv;
        float mid = 0.5f * (cv.a + cv.c);
        float sq = sqrtf(fmaxf(0.1f, fmaf(mid, mid, -det)));
        float l1 = mid + sq, l2 = mid - sq;
        int rad = f2i_trunc(ceilf(3.0f * sqrtf(fmaxf(l1, l2))));
        float ix = fmaf(ppx + 1.0f, (float)W, -1.0f) * 0.5f;   /* ndc2Pix */
        float iy = fmaf(ppy + 1.0f, (float)H, -1.0f) * 0.5f;
        int x0, y0, x1, y1;
        get_rect(ix, iy, rad, gx, gy, &x0, &y0, &x1, &y1);
        if ((x1 - x0) * (y1 - y0) == 0) continue;
        if (colors_precomp) {
            for (int k = 0; k < 3; k++) rgb[3 * i + k] = colors_precomp[3 * i + k];
        } else {
            float dx = px - campos[0], dy = py - campos[1], dz = pz - campos[2];
            float len = sqrtf(fmaf(dz, dz, fmaf(dy, dy, dx * dx)));
            float inv = 1.0f / len;
            float b[16];
            sh_basis(D, dx * inv, dy * inv, dz * inv, b);
            const float* sh = shs + (size_t)i * M * 3;
            int nb = (D + 1) * (D + 1);
            for (int ch = 0; ch < 3; ch++) {
                float acc = b[0] * sh[ch];
                for (int k = 1; k < nb; k++) acc = fmaf(b[k], sh[3 * k + ch], acc);
                acc += 0.5f;
                clamped[3 * i + ch] = acc < 0.0f;
                rgb[3 * i + ch] = fmaxf(acc, 0.0f);
            }
        }
        depths[i] = pv[2];
        radii[i] = rad;
        xy[2 * i] = ix; xy[2 * i + 1] = iy;
        for (int k = 0; k < 6; k++) cov3D[6 * i + k] = c3[k];
        conic_opacity[4 * i] = conA; conic_opacity[4 * i + 1] = conB;
        conic_opacity[4 * i + 2] = conC; conic_opacity[4 * i + 3] = opacities[i];
        tiles_touched[i] = (uint32_t)((x1 - x0) * (y1 - y0));
    }
}

/* [upstream] cub::DeviceScan::InclusiveSum over tiles_touched; returns L = num_rendered. */
int64_t ro_scan(int P, const uint32_t* tiles_touched, uint32_t* offsets) {
    uint64_t acc = 0;
    for (int i = 0; i < P; i++) { acc += tiles_touched[i]; offsets[i] = (uint32_t)acc; }
    return (int64_t)acc;
}

/* [upstream] rasterizer_impl.cu duplicateWithKeys */
void ro_duplicate_with_keys(int P, int W, int H, const float* xy, const float* depths,
                            const uint32_t* offsets, const int* radii, uint64_t* keys,
                            uint32_t* vals) {
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < P; i++) {
        if (radii[i] <= 0) continue;
        uint32_t off = i == 0 ? 0 : offsets[i - 1];
        int x0, y0, x1, y1;
        get_rect(xy[2 * i], xy[2 * i + 1], radii[i], gx, gy, &x0, &y0, &x1, &y1);
        uint32_t dbits;
        memcpy(&dbits, depths + i, 4);
        for (int y = y0; y < y1; y++)
            for (int x = x0; x < x1; x++) {
                uint64_t key = (uint64_t)(uint32_t)(y * gx + x);
                keys[off] = (key << 32) | dbits;
                vals[off] = (uint32_t)i;
                off++;
            }
    }
}

/* [upstream] cub::DeviceRadixSort::SortPairs(keys, vals, L, 0, end_bit): stable LSD radix
 * sort on key bits [0, end_bit).  tmp arrays of L entries each are supplied by the caller. */
void ro_sort_pairs(int64_t L, uint64_t* keys, uint32_t* vals, uint64_t* tkeys, uint32_t* tvals,
                   int end_bit) {
    uint64_t *ka = keys, *kb = tkeys;
    uint32_t *va = vals, *vb = tvals;
    for (int shift = 0; shift < end_bit; shift += 8) {
        int bits = end_bit - shift < 8 ? end_bit - shift : 8;
        uint32_t mask = (1u << bits) - 1u;
        int64_t count[257];
        memset(count, 0, sizeof count);
        for (int64_t i = 0; i < L; i++) count[((ka[i] >> shift) & mask) + 1]++;
        for (int d = 0; d < 256; d++) count[d + 1] += count[d];
        for (int64_t i = 0; i < L; i++) {
            int64_t dst = count[(ka[i] >> shift) & mask]++;
            kb[dst] = ka[i]; vb[dst] = va[i];
        }
        uint64_t* tk = ka; ka = kb; kb = tk;
        uint32_t* tv = va; va = vb; vb = tv;
    }
    if (ka != keys) { memcpy(keys, ka, (size_t)L * 8); memcpy(vals, va, (size_t)L * 4); }
}

/* [upstream] rasterizer_impl.cu identifyTileRanges; ranges (tiles,2) pre-zeroed here. */
void ro_tile_ranges(int64_t L, const uint64_t* keys, int tiles, uint32_t* ranges) {
    memset(ranges, 0, (size_t)tiles * 8);
    for (int64_t i = 0; i < L; i++) {
        uint32_t cur = (uint32_t)(keys[i] >> 32);
        if (i == 0) ranges[2 * cur] = 0;
        else {
            uint32_t prev = (uint32_t)(keys[i - 1] >> 32);
            if (prev != cur) { ranges[2 * prev + 1] = (uint32_t)i; ranges[2 * cur] = (uint32_t)i; }
        }
        if (i == L - 1) ranges[2 * cur + 1] = (uint32_t)L;
    }
}

/* Gaussian falloff exponent for pixel (pxf,pyf); conic passed pre-scaled:
 * Ah = -0.5*A, Bn = -B, Ch = -0.5*C (exact scalings). */
static inline float power_of(float gx, float gy, float Ah, float Bn, float Ch, float pxf,
                             float pyf, float* dx, float* dy) {
    *dx = gx - pxf; *dy = gy - pyf;
    float u = Ah * *dx, v = Ch * *dy, w = Bn * *dx;
    return fmaf(w, *dy, fmaf(v, *dy, u * *dx));
}

/* ------------------------------------------------------------------------------------
 * [upstream] forward.cu renderCUDA.  out_color planar (3,H,W).
 * out_alpha (= 1 - final_T) and out_depth (= sum alpha*T*depth) may be NULL.
 * ---------------------------------------------------------------------------------- */
void ro_render_fwd(int W, int H, const uint32_t* ranges, const uint32_t* point_list,
                   const float* xy, const float* conic_opacity, const float* rgb,
                   const float* depths, const float* bg, float* out_color, float* final_T,
                   uint32_t* n_contrib, float* out_alpha, float* out_depth) {
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
#pragma omp parallel for schedule(dynamic, 1)
    for (int t = 0; t < gx * gy; t++) {
        const uint32_t r0 = ranges[2 * t], r1 = ranges[2 * t + 1];
        const int tx0 = (t % gx) * TILE, ty0 = (t / gx) * TILE;
        for (int ly = 0; ly < TILE; ly++)
            for (int lx = 0; lx < TILE; lx++) {
                const int px = tx0 + lx, py = ty0 + ly;
                if (px >= W || py >= H) continue;
                const float pxf = (float)px, pyf = (float)py;
                float T = 1.0f, C[3] = {0, 0, 0}, Dacc = 0.0f;
                uint32_t contributor = 0, last = 0;
                for (uint32_t k = r0; k < r1; k++) {
                    contributor++;
                    const uint32_t g = point_list[k];
                    const float* co = conic_opacity + 4 * g;
                    float dx, dy;
                    float power = power_of(xy[2 * g], xy[2 * g + 1], -0.5f * co[0], -co[1],
                                           -0.5f * co[2], pxf, pyf, &dx, &dy);
                    if (power > 0.0f) continue;
                    float alpha = fminf(0.99f, co[3] * expneg(power));
                    if (alpha < 1.0f / 255.0f) continue;
                    float test_T = T * (1.0f - alpha);
                    if (test_T < 0.0001f) break;
                    float w = alpha * T;
                    for (int ch = 0; ch < 3; ch++) C[ch] = fmaf(rgb[3 * g + ch], w, C[ch]);
                    Dacc = fmaf(depths[g], w, Dacc);
                    T = test_T;
                    last = contributor;
                }
                const size_t pix = (size_t)py * W + px;
                final_T[pix] = T;
                n_contrib[pix] = last;
                for (int ch = 0; ch < 3; ch++)
                    out_color[(size_t)ch * H * W + pix] = fmaf(T, bg[ch], C[ch]);
                if (out_alpha) out_alpha[pix] = 1.0f - T;
                if (out_depth) out_depth[pix] = Dacc;
            }
    }
}

static inline void atomic_addf(float* p, float v) {
#pragma omp atomic
    *p += v;
}

/* ------------------------------------------------------------------------------------
 * [upstream] backward.cu renderCUDA.  Outputs must be zeroed by caller:
 * dL_dmean2D (P,2) [already scaled by 0.5*W, 0.5*H], dL_dconic (P,3) [x, y(un-doubled), w],
 * dL_dopacity (P), dL_dcolor (P,3).
 * ---------------------------------------------------------------------------------- */
void ro_render_bwd(int W, int H, const uint32_t* ranges, const uint32_t* point_list,
                   const float* xy, const float* conic_opacity, const float* rgb,
                   const float* bg, const float* final_T, const uint32_t* n_contrib,
                   const float* dL_dpix, float* dL_dmean2D, float* dL_dconic,
                   float* dL_dopacity, float* dL_dcolor) {
    const int gx = (W + TILE - 1) / TILE, gy = (H + TILE - 1) / TILE;
    const float ddelx_dx = 0.5f * (float)W, ddely_dy = 0.5f * (float)H;
#pragma omp parallel for schedule(dynamic, 1)
    for (int t = 0; t < gx * gy; t++) {
        const uint32_t r0 = ranges[2 * t];
        const int tx0 = (t % gx) * TILE, ty0 = (t / gx) * TILE;
        for (int ly = 0; ly < TILE; ly++)
            for (int lx = 0; lx < TILE; lx++) {
                const int px = tx0 + lx, py = ty0 + ly;
                if (px >= W || py >= H) continue;
                const size_t pix = (size_t)py * W + px;
                const float pxf = (float)px, pyf = (float)py;
                const float T_final = final_T[pix];
                float T = T_final;
                float dpix[3], bg_dot = 0.0f;
                for (int ch = 0; ch < 3; ch++) {
                    dpix[ch] = dL_dpix[(size_t)ch * H * W + pix];
                    bg_dot = fmaf(bg[ch], dpix[ch], bg_dot);
                }
                float last_alpha = 0.0f, last_color[3] = {0, 0, 0}, accum_rec[3] = {0, 0, 0};
                for (int64_t k = (int64_t)n_contrib[pix] - 1; k >= 0; k--) {
                    const uint32_t g = point_list[r0 + k];
                    const float* co = conic_opacity + 4 * g;
                    float dx, dy;
                    float power = power_of(xy[2 * g], xy[2 * g + 1], -0.5f * co[0], -co[1],
                                           -0.5f * co[2], pxf, pyf, &dx, &dy);
                    if (power > 0.0f) continue;
                    const float G = expneg(power);
                    const float alpha = fminf(0.99f, co[3] * G);
                    if (alpha < 1.0f / 255.0f) continue;
                    T = T / (1.0f - alpha);
                    const float dchannel_dcolor = alpha * T;
                    float dL_dalpha = 0.0f;
                    for (int ch = 0; ch < 3; ch++) {
                        const float c = rgb[3 * g + ch];
                        accum_rec[ch] = fmaf(last_alpha, last_color[ch], (1.0f - last_alpha) * accum_rec[ch]);
                        last_color[ch] = c;
                        dL_dalpha = fmaf(c - accum_rec[ch], dpix[ch], dL_dalpha);
                        atomic_addf(dL_dcolor + 3 * g + ch, dchannel_dcolor * dpix[ch]);
                    }
                    dL_dalpha *= T;
                    last_alpha = alpha;
                    dL_dalpha = fmaf(-T_final / (1.0f - alpha), bg_dot, dL_dalpha);
                    const float dL_dG = co[3] * dL_dalpha;
                    const float gdx = G * dx, gdy = G * dy;
                    const float dG_ddelx = -gdx * co[0] - gdy * co[1];
                    const float dG_ddely = -gdy * co[2] - gdx * co[1];
                    atomic_addf(dL_dmean2D + 2 * g, dL_dG * dG_ddelx * ddelx_dx);
                    atomic_addf(dL_dmean2D + 2 * g + 1, dL_dG * dG_ddely * ddely_dy);
                    atomic_addf(dL_dconic + 3 * g, -0.5f * gdx * dx * dL_dG);
                    atomic_addf(dL_dconic + 3 * g + 1, -0.5f * gdx * dy * dL_dG);
                    atomic_addf(dL_dconic + 3 * g + 2, -0.5f * gdy * dy * dL_dG);
                    atomic_addf(dL_dopacity + g, G * dL_dalpha);
                }
            }
    }
}

/* ------------------------------------------------------------------------------------
 * [upstream] backward.cu computeCov2DCUDA + preprocessCUDA, one loop.
 * Inputs: the per-Gaussian sums of ro_render_bwd.  Outputs (all fully written):
 * dL_dmeans3D (P,3), dL_dscales (P,3), dL_drots (P,4), dL_dsh (P,M,3), dL_dcov3D (P,6).
 * Quirk kept from upstream: dL_dscale carries no scale_modifier factor.
 * ---------------------------------------------------------------------------------- */
void ro_preprocess_bwd(int P, int D, int M, const float* means3D, const float* scales, float mod,
                       const float* rotations, const float* shs, const float* cov3D_precomp,
                       const float* view, const float* proj, const float* campos, int W, int H,
                       float tanfovx, float tanfovy, const int* radii, const uint8_t* clamped,
                       const float* dL_dmean2D, const float* dL_dconic, const float* dL_dcolor,
                       float* dL_dmeans3D, float* dL_dscales, float* dL_drots, float* dL_dsh,
                       float* dL_dcov3D) {
    const float fx = (float)W / (2.0f * tanfovx), fy = (float)H / (2.0f * tanfovy);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < P; i++) {
        for (int k = 0; k < 3; k++) { dL_dmeans3D[3 * i + k] = 0.0f; dL_dscales[3 * i + k] = 0.0f; }
        for (int k = 0; k < 4; k++) dL_drots[4 * i + k] = 0.0f;
        for (int k = 0; k < 6; k++) dL_dcov3D[6 * i + k] = 0.0f;
        if (dL_dsh) for (int k = 0; k < 3 * M; k++) dL_dsh[(size_t)i * 3 * M + k] = 0.0f;
        if (radii[i] <= 0) continue;
        const float px = means3D[3 * i], py = means3D[3 * i + 1], pz = means3D[3 * i + 2];
        float dmean[3] = {0, 0, 0};
        /* ---- cov2D backward ---- */
        float c3[6];
        if (cov3D_precomp) memcpy(c3, cov3D_precomp + 6 * i, sizeof c3);
        else cov3d_from_scale_rot(scales + 3 * i, mod, rotations + 4 * i, c3);
        float pv[3] = {xform_row(view, 0, px, py, pz), xform_row(view, 1, px, py, pz),
                       xform_row(view, 2, px, py, pz)};
        cov2d_t cv;
        cov2d(pv, fx, fy, tanfovx, tanfovy, c3, view, &cv);
        const float a = cv.a, b = cv.b, c = cv.c;
        const float gxx = dL_dconic[3 * i], gxy = dL_dconic[3 * i + 1], gyy = dL_dconic[3 * i + 2];
        const float denom = a * c - b * b;
        const float denom2inv = 1.0f / (denom * denom + 0.0000001f);
        float dL_da = 0, dL_db = 0, dL_dc = 0;
        float dcov[6] = {0, 0, 0, 0, 0, 0};
        if (denom2inv != 0.0f) {
            dL_da = denom2inv * (-c * c * gxx + 2.0f * b * c * gxy + (denom - a * c) * gyy);
            dL_dc = denom2inv * (-a * a * gyy + 2.0f * a * b * gxy + (denom - a * c) * gxx);
            dL_db = denom2inv * 2.0f * (b * c * gxx - (denom + 2.0f * b * b) * gxy + a * b * gyy);
            const float* A0 = cv.A0; const float* A1 = cv.A1;
            dcov[0] = A0[0] * A0[0] * dL_da + A0[0] * A1[0] * dL_db + A1[0] * A1[0] * dL_dc;
            dcov[3] = A0[1] * A0[1] * dL_da + A0[1] * A1[1] * dL_db + A1[1] * A1[1] * dL_dc;
            dcov[5] = A0[2] * A0[2] * dL_da + A0[2] * A1[2] * dL_db + A1[2] * A1[2] * dL_dc;
            dcov[1] = 2.0f * A0[0] * A0[1] * dL_da + (A0[0] * A1[1] + A0[1] * A1[0]) * dL_db + 2.0f * A1[0] * A1[1] * dL_dc;
            dcov[2] = 2.0f * A0[0] * A0[2] * dL_da + (A0[0] * A1[2] + A0[2] * A1[0]) * dL_db + 2.0f * A1[0] * A1[2] * dL_dc;
            dcov[4] = 2.0f * A0[2] * A0[1] * dL_da + (A0[1] * A1[2] + A0[2] * A1[1]) * dL_db + 2.0f * A1[1] * A1[2] * dL_dc;
        }
        float dA0[3], dA1[3];
        for (int k = 0; k < 3; k++) {
            dA0[k] = 2.0f * cv.B0[k] * dL_da + cv.B1[k] * dL_db;
            dA1[k] = 2.0f * cv.B1[k] * dL_dc + cv.B0[k] * dL_db;
        }
        float dJ00 = 0, dJ02 = 0, dJ11 = 0, dJ12 = 0;
        for (int k = 0; k < 3; k++) {
            dJ00 += view[4 * k + 0] * dA0[k];
            dJ02 += view[4 * k + 2] * dA0[k];
            dJ11 += view[4 * k + 1] * dA1[k];
            dJ12 += view[4 * k + 2] * dA1[k];
        }
        const float tz = 1.0f / cv.tz, tz2 = tz * tz, tz3 = tz2 * tz;
        const float dtx = cv.xmul * -fx * tz2 * dJ02;
        const float dty = cv.ymul * -fy * tz2 * dJ12;
        const float dtz = -fx * tz2 * dJ00 - fy * tz2 * dJ11 + (2.0f * fx * cv.tx) * tz3 * dJ02 +
                          (2.0f * fy * cv.ty) * tz3 * dJ12;
        for (int k = 0; k < 3; k++)   /* transformVec4x3Transpose */
            dmean[k] = view[4 * k] * dtx + view[4 * k + 1] * dty + view[4 * k + 2] * dtz;
        /* ---- mean2D -> mean3D through the projection ---- */
        {
            float mhx = xform_row(proj, 0, px, py, pz), mhy = xform_row(proj, 1, px, py, pz);
            float mhw = xform_row(proj, 3, px, py, pz);
            float m_w = 1.0f / (mhw + 0.0000001f);
            float mul1 = mhx * m_w * m_w, mul2 = mhy * m_w * m_w;
            float g2x = dL_dmean2D[2 * i], g2y = dL_dmean2D[2 * i + 1];
            for (int k = 0; k < 3; k++)
                dmean[k] += (proj[4 * k] * m_w - proj[4 * k + 3] * mul1) * g2x +
                            (proj[4 * k + 1] * m_w - proj[4 * k + 3] * mul2) * g2y;
        }
        /* ---- SH backward ---- */
        if (shs && dL_dsh) {
            float dox = px - campos[0], doy = py - campos[1], doz = pz - campos[2];
            float len = sqrtf(fmaf(doz, doz, fmaf(doy, doy, dox * dox)));
            float inv = 1.0f / len;
            float x = dox * inv, y = doy * inv, z = doz * inv;
            float bas[16];
            sh_basis(D, x, y, z, bas);
            const float* sh = shs + (size_t)i * M * 3;
            float* dsh = dL_dsh + (size_t)i * M * 3;
            float dRGB[3];
            for (int ch = 0; ch < 3; ch++) dRGB[ch] = clamped[3 * i + ch] ? 0.0f : dL_dcolor[3 * i + ch];
            int nb = (D + 1) * (D + 1);
            for (int k = 0; k < nb; k++)
                for (int ch = 0; ch < 3; ch++) dsh[3 * k + ch] = bas[k] * dRGB[ch];
            float ddir[3] = {0, 0, 0};
            for (int ch = 0; ch < 3; ch++) {
#define S(k) sh[3 * (k) + ch]
                float gx_ = 0, gy_ = 0, gz_ = 0;
                if (D > 0) {
                    gx_ = -SH_C1 * S(3); gy_ = -SH_C1 * S(1); gz_ = SH_C1 * S(2);
                    if (D > 1) {
                        float xx = x * x, yy = y * y, zz = z * z, xy_ = x * y, yz = y * z, xz = x * z;
                        gx_ += SH_C2[0] * y * S(4) + SH_C2[2] * 2.0f * -x * S(6) + SH_C2[3] * z * S(7) + SH_C2[4] * 2.0f * x * S(8);
                        gy_ += SH_C2[0] * x * S(4) + SH_C2[1] * z * S(5) + SH_C2[2] * 2.0f * -y * S(6) + SH_C2[4] * 2.0f * -y * S(8);
                        gz_ += SH_C2[1] * y * S(5) + SH_C2[2] * 2.0f * 2.0f * z * S(6) + SH_C2[3] * x * S(7);
                        if (D > 2) {
                            gx_ += SH_C3[0] * S(9) * 3.0f * 2.0f * xy_ + SH_C3[1] * S(10) * yz +
                                   SH_C3[2] * S(11) * -2.0f * xy_ + SH_C3[3] * S(12) * -3.0f * 2.0f * xz +
                                   SH_C3[4] * S(13) * (-3.0f * xx + 4.0f * zz - yy) +
                                   SH_C3[5] * S(14) * 2.0f * xz + SH_C3[6] * S(15) * 3.0f * (xx - yy);
                            gy_ += SH_C3[0] * S(9) * 3.0f * (xx - yy) + SH_C3[1] * S(10) * xz +
                                   SH_C3[2] * S(11) * (-3.0f * yy + 4.0f * zz - xx) +
                                   SH_C3[3] * S(12) * -3.0f * 2.0f * yz + SH_C3[4] * S(13) * -2.0f * xy_ +
                                   SH_C3[5] * S(14) * -2.0f * yz + SH_C3[6] * S(15) * -3.0f * 2.0f * xy_;
                            gz_ += SH_C3[1] * S(10) * xy_ + SH_C3[2] * S(11) * 4.0f * 2.0f * yz +
                                   SH_C3[3] * S(12) * 3.0f * (2.0f * zz - xx - yy) +
                                   SH_C3[4] * S(13) * 4.0f * 2.0f * xz + SH_C3[5] * S(14) * (xx - yy);
                        }
                    }
                }
#undef S
                ddir[0] += gx_ * dRGB[ch]; ddir[1] += gy_ * dRGB[ch]; ddir[2] += gz_ * dRGB[ch];
            }
            /* dnormvdv */
            float sum2 = dox * dox + doy * doy + doz * doz;
            float invsum32 = 1.0f / sqrtf(sum2 * sum2 * sum2);
            dmean[0] += ((sum2 - dox * dox) * ddir[0] - doy * dox * ddir[1] - doz * dox * ddir[2]) * invsum32;
            dmean[1] += (-dox * doy * ddir[0] + (sum2 - doy * doy) * ddir[1] - doz * doy * ddir[2]) * invsum32;
            dmean[2] += (-dox * doz * ddir[0] - doy * doz * ddir[1] + (sum2 - doz * doz) * ddir[2]) * invsum32;
        }
        /* ---- cov3D backward (scale, rotation) ---- */
        if (!cov3D_precomp) {
            const float* q = rotations + 4 * i;
            const float r = q[0], x = q[1], y = q[2], z = q[3];
            float R[9];
            quat_to_R(r, x, y, z, R);
            float s[3] = {mod * scales[3 * i], mod * scales[3 * i + 1], mod * scales[3 * i + 2]};
            /* D = dL/dSigma as a symmetric matrix (off-diagonals halved) */
            float Dm[9] = {dcov[0], 0.5f * dcov[1], 0.5f * dcov[2], 0.5f * dcov[1], dcov[3],
                           0.5f * dcov[4], 0.5f * dcov[2], 0.5f * dcov[4], dcov[5]};
            /* N = R diag(s);  dN = 2 D N */
            float dN[9];
            for (int ii = 0; ii < 3; ii++)
                for (int k = 0; k < 3; k++) {
                    float acc = 0;
                    for (int j = 0; j < 3; j++) acc += Dm[3 * ii + j] * (R[3 * j + k] * s[k]);
                    dN[3 * ii + k] = 2.0f * acc;
                }
            for (int k = 0; k < 3; k++)
                dL_dscales[3 * i + k] = R[k] * dN[k] + R[3 + k] * dN[3 + k] + R[6 + k] * dN[6 + k];
            float E[9];
            for (int ii = 0; ii < 3; ii++)
                for (int k = 0; k < 3; k++) E[3 * ii + k] = dN[3 * ii + k] * s[k];
            dL_drots[4 * i + 0] = 2.0f * (-z * E[1] + y * E[2] + z * E[3] - x * E[5] - y * E[6] + x * E[7]);
            dL_drots[4 * i + 1] = 2.0f * (y * E[1] + z * E[2] + y * E[3] - 2.0f * x * E[4] - r * E[5] + z * E[6] + r * E[7] - 2.0f * x * E[8]);
            dL_drots[4 * i + 2] = 2.0f * (-2.0f * y * E[0] + x * E[1] + r * E[2] + x * E[3] + z * E[5] - r * E[6] + z * E[7] - 2.0f * y * E[8]);
            dL_drots[4 * i + 3] = 2.0f * (-2.0f * z * E[0] - r * E[1] + x * E[2] + r * E[3] - 2.0f * z * E[4] + y * E[5] + x * E[6] + y * E[7]);
        }
        for (int k = 0; k < 3; k++) dL_dmeans3D[3 * i + k] = dmean[k];
        for (int k = 0; k < 6; k++) dL_dcov3D[6 * i + k] = dcov[k];
    }
}

/* [upstream] rasterizer_impl.cu checkFrustum / markVisible */
void ro_mark_visible(int P, const float* means3D, const float* view, uint8_t* present) {
    for (int i = 0; i < P; i++) {
        float z = xform_row(view, 2, means3D[3 * i], means3D[3 * i + 1], means3D[3 * i + 2]);
        present[i] = z > 0.2f;
    }
}

/* Decision mask of ro_render_fwd for one tile: mask[(ly*16+lx)*K + k] = 1 when list entry
 * k of the tile contributed to that pixel (passed power<=0, alpha>=1/255, and lies before
 * the pixel's termination point).  Used by the float64 autograd restatement so that the
 * discrete decisions are exactly the fp32 ones. */
void ro_render_mask(int W, int H, const uint32_t* ranges, const uint32_t* point_list,
                    const float* xy, const float* conic_opacity, const uint32_t* n_contrib,
                    int tile, uint8_t* mask) {
    const int gx = (W + TILE - 1) / TILE;
    const uint32_t r0 = ranges[2 * tile], r1 = ranges[2 * tile + 1];
    const uint32_t K = r1 - r0;
    const int tx0 = (tile % gx) * TILE, ty0 = (tile / gx) * TILE;
    memset(mask, 0, (size_t)256 * K);
    for (int ly = 0; ly < TILE; ly++)
        for (int lx = 0; lx < TILE; lx++) {
            const int px = tx0 + lx, py = ty0 + ly;
            if (px >= W || py >= H) continue;
            const uint32_t last = n_contrib[(size_t)py * W + px];
            for (uint32_t k = 0; k < K && k < last; k++) {
                const uint32_t g = point_list[r0 + k];
                const float* co = conic_opacity + 4 * g;
                float dx, dy;
                float power = power_of(xy[2 * g], xy[2 * g + 1], -0.5f * co[0], -co[1],
                                       -0.5f * co[2], (float)px, (float)py, &dx, &dy);
                if (power > 0.0f) continue;
                float alpha = fminf(0.99f, co[3] * expneg(power));
                if (alpha < 1.0f / 255.0f) continue;
                mask[(size_t)(ly * TILE + lx) * K + k] = 1;
            }
        }
}
