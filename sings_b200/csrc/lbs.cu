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
#include "lbs_math.cuh"

namespace sgs {

// Gaussians per CTA (A/B knob, tools/sweep.sh)
#ifndef SGS_LBS_THREADS
#define SGS_LBS_THREADS 256
#endif
constexpr int LBS_THREADS = SGS_LBS_THREADS;
#ifndef SGS_LBS_BWD_MINB          // resident CTAs per SM the backward is compiled for (register cap)
#define SGS_LBS_BWD_MINB 2
#endif

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

// pose -> A in two steps with a block barrier between them:
//   pose_local: joint j's local transform [R(pose_j) | rest_j - rest_parent] into s_local
//   pose_chain: each thread multiplies its own ancestor chain from the root down (same
//     association as the reference's sequential loop, smpl.py:495-501), then
//     A = G - pad(G [rest_j; 0]) (smpl.py:510-511) and A @ inv_A_t2cano (sings_hybrid.py:399)
__device__ __forceinline__ void pose_local(const float* __restrict__ pose_bj, const float* __restrict__ rest,
                                           const int* __restrict__ parents, int j, float* s_local, int* s_par) {
    float R[9];
    rodrigues(pose_bj[0], pose_bj[1], pose_bj[2], R);
    const int par = parents[j];
    s_par[j] = par;
    float t[3];
#pragma unroll
    for (int k = 0; k < 3; k++) t[k] = rest[3 * j + k] - (par >= 0 ? rest[3 * par + k] : 0.0f);
    float* L = s_local + 12 * j;
#pragma unroll
    for (int r = 0; r < 3; r++) {
        L[4 * r] = R[3 * r]; L[4 * r + 1] = R[3 * r + 1]; L[4 * r + 2] = R[3 * r + 2]; L[4 * r + 3] = t[r];
    }
}

__device__ __forceinline__ void pose_chain(const float* __restrict__ rest, const float* __restrict__ inv_A, int j,
                                           const float* s_local, const int* s_par, float* G, float* out) {
    int chain[64];
    int depth = 0;
    for (int k = j; k >= 0 && depth < 64; k = s_par[k]) chain[depth++] = k;
    float tmp[12];
#pragma unroll
    for (int k = 0; k < 12; k++) G[k] = s_local[12 * chain[depth - 1] + k];
    for (int d = depth - 2; d >= 0; d--) {
        affine_mul(G, s_local + 12 * chain[d], tmp);
#pragma unroll
        for (int k = 0; k < 12; k++) G[k] = tmp[k];
    }
    float Arel[12];
#pragma unroll
    for (int r = 0; r < 3; r++) {
        Arel[4 * r] = G[4 * r]; Arel[4 * r + 1] = G[4 * r + 1]; Arel[4 * r + 2] = G[4 * r + 2];
        Arel[4 * r + 3] = G[4 * r + 3] - (G[4 * r] * rest[3 * j] + G[4 * r + 1] * rest[3 * j + 1] + G[4 * r + 2] * rest[3 * j + 2]);
    }
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
}

__device__ __forceinline__ void pose_store(float* __restrict__ A_out, float* __restrict__ G_out, size_t bj,
                                           const float* G, const float* out) {
    if (G_out) {
#pragma unroll
        for (int k = 0; k < 12; k++) G_out[bj * 12 + k] = G[k];
    }
    float* dst = A_out + bj * 16;
#pragma unroll
    for (int k = 0; k < 12; k++) dst[k] = out[k];
    dst[12] = 0.0f; dst[13] = 0.0f; dst[14] = 0.0f; dst[15] = 1.0f;
}

// one CTA per frame, one thread per joint
__global__ void pose_to_A_kernel(const float* __restrict__ pose, const float* __restrict__ rest,
                                 const int* __restrict__ parents, const float* __restrict__ inv_A,
                                 int J, float* __restrict__ A_out, float* __restrict__ G_out) {
    extern __shared__ float s_local[];     // J x 12 local transforms
    __shared__ int s_par[64];
    const int b = blockIdx.x, j = threadIdx.x;
    pdl_sync();
    if (j < J) pose_local(pose + ((size_t)b * J + j) * 3, rest, parents, j, s_local, s_par);
    __syncthreads();
    if (j >= J) return;
    float G[12], out[12];
    pose_chain(rest, inv_A, j, s_local, s_par, G, out);
    pose_store(A_out, G_out, (size_t)b * J + j, G, out);
}

int launch_pose_to_A(const float* pose, const float* rest, const int* parents, const float* inv_A,
                     int B, int J, float* A_out, float* G_out, cudaStream_t stream) {
    if (B <= 0) return 0;
    if (J < 1 || J > 64) return SGS_ERR_BAD_JOINTS;
    launch_pdl(pose_to_A_kernel, B, 64, (size_t)J * 12 * 4, stream, pose, rest, parents, inv_A, J, A_out, G_out);
    SGS_LAUNCH_OK();
    return 0;
}

// backward of pose_to_A_kernel: dL/dA (B,J,16) -> dL/dpose (B,J,3).  One CTA per frame.
// The kernel is one short dependent chain (it sits alone at the end of the frame), so the chain
// is cut two ways: (1) everything that does not depend on dL/dA -- Rodrigues, the local
// transforms, the forward's G, the children lists, tree depths -- is done by one thread per
// joint BEFORE the grid-dependency wait, i.e. while the LBS backward is still running (all of
// it is at least two kernels old, common.cuh); (2) after the wait the 3x4 matrices are handled
// one ELEMENT per thread (12 threads per joint): seed dL/dG_j, then walk the kinematic tree
// leaves-to-root one depth level at a time (a parent gathers dG_k L_k^T from its children, whose
// dG are final after the previous level), then dL/dR_j = G_parent^T dL/dG_j.  The last step,
// the derivative of the Rodrigues formula, is per joint again.  All sums keep the order of the
// per-joint formulation (children highest index first), so results do not depend on the mapping.
__global__ void pose_to_A_bwd_kernel(const float* __restrict__ pose, const float* __restrict__ rest,
                                     const int* __restrict__ parents, const float* __restrict__ inv_A,
                                     const float* __restrict__ G_all, const float* __restrict__ dA,
                                     int J, float* __restrict__ d_pose) {
    extern __shared__ float s_mem[];
    float* s_L = s_mem;              // J x 12 local transforms
    float* s_dG = s_mem + 12 * J;    // J x 12 dL/dG
    float* s_G = s_mem + 24 * J;     // J x 12 global transforms (forward's G_out)
    float* s_dO = s_mem + 36 * J;    // J x 12 dL/dA rows 0..2, later dL/dArel
    float* s_E = s_mem + 48 * J;     // J x 9  dL/dR
    float* s_B = s_mem + 57 * J;     // J x 12 inv_A rows 0..2
    __shared__ int s_par[64], s_depth[64], s_maxd;
    __shared__ unsigned char s_nchild[64], s_child[64][64];
    const int b = blockIdx.x, t = threadIdx.x;
    const int j = t;                 // joint role: threads 0..J-1
    const int ej = t / 12, ee = t - ej * 12, er = ee >> 2, ec = ee & 3;   // element role
    const bool elem = ej < J;
    // ---------------- before the wait: nothing here depends on dL/dA ----------------
    if (t == 0) s_maxd = 0;
    float rx = 0, ry = 0, rz = 0;
    if (j < J) {
        const float* p = pose + ((size_t)b * J + j) * 3;
        rx = p[0]; ry = p[1]; rz = p[2];
        float R[9];
        rodrigues(rx, ry, rz, R);
        const int par = parents[j];
        s_par[j] = par;
        float* L = s_L + 12 * j;
#pragma unroll
        for (int r = 0; r < 3; r++) {
            L[4 * r] = R[3 * r]; L[4 * r + 1] = R[3 * r + 1]; L[4 * r + 2] = R[3 * r + 2];
            L[4 * r + 3] = rest[3 * j + r] - (par >= 0 ? rest[3 * par + r] : 0.0f);
        }
    }
    if (elem) {
        s_G[12 * ej + ee] = G_all[((size_t)b * J + ej) * 12 + ee];
        if (inv_A) s_B[12 * ej + ee] = inv_A[(size_t)ej * 16 + ee];
    }
    __syncthreads();
    int depth = 0;
    if (j < J) {
        for (int k = s_par[j]; k >= 0 && depth < 64; k = s_par[k]) depth++;
        s_depth[j] = depth;
        atomicMax(&s_maxd, depth);
        int nc = 0;
        for (int k = J - 1; k > j; k--)             // children, highest index first
            if (s_par[k] == j) s_child[j][nc++] = (unsigned char)k;
        s_nchild[j] = (unsigned char)nc;
    }
    pdl_sync();
    // ---------------- after the wait ----------------
    if (elem) s_dO[12 * ej + ee] = dA[((size_t)b * J + ej) * 16 + ee];
    __syncthreads();
    const int maxd = s_maxd;
    const float rest_c = (elem && ec < 3) ? rest[3 * ej + ec] : 0.0f;
    float dAr = 0.0f;
    if (elem) {
        // dOut -> dArel (through @ inv_A)
        const float* dO = s_dO + 12 * ej + 4 * er;
        if (inv_A && ec < 3) {
            const float* Bm = s_B + 12 * ej + 4 * ec;      // row ec of inv_A
            dAr = dO[0] * Bm[0] + dO[1] * Bm[1] + dO[2] * Bm[2] + dO[3] * Bm[3];
        } else {
            dAr = dO[ec];
        }
    }
    __syncthreads();
    if (elem) s_dO[12 * ej + ee] = dAr;
    __syncthreads();
    if (elem) {
        // dArel -> dG:  Arel[:, 3] = G[:, 3] - G[:, :3] rest_j
        const float d3 = s_dO[12 * ej + 4 * er + 3];
        s_dG[12 * ej + ee] = ec < 3 ? dAr - d3 * rest_c : d3;
    }
    __syncthreads();
    const int edepth = elem ? s_depth[ej] : -1;
    for (int d = maxd - 1; d >= 0; d--) {
        if (edepth == d) {
            float acc = s_dG[12 * ej + ee];
            const int nc = s_nchild[ej];
            for (int q = 0; q < nc; q++) {
                const int k = s_child[ej][q];
                const float* dG = s_dG + 12 * k + 4 * er;
                if (ec < 3) {
                    const float* L = s_L + 12 * k + 4 * ec;
                    acc += dG[0] * L[0] + dG[1] * L[1] + dG[2] * L[2] + dG[3] * L[3];
                } else {
                    acc += dG[3];
                }
            }
            s_dG[12 * ej + ee] = acc;      // no other thread reads this joint's dG at this level
        }
        __syncthreads();
    }
    // E = dL/dR_j: rotation part of dL/dL_j = G_parent^T dL/dG_j (the root's L is G itself)
    if (elem && ee < 9) {
        const int r = ee / 3, c = ee - 3 * r;
        const float* dG = s_dG + 12 * ej;
        const int par = s_par[ej];
        float v;
        if (par >= 0) {
            const float* Gp = s_G + 12 * par;
            v = Gp[r] * dG[c] + Gp[4 + r] * dG[4 + c] + Gp[8 + r] * dG[8 + c];
        } else {
            v = dG[4 * r + c];
        }
        s_E[9 * ej + ee] = v;
    }
    __syncthreads();
    if (j >= J) return;
    float E[9];
#pragma unroll
    for (int k = 0; k < 9; k++) E[k] = s_E[9 * j + k];
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
    launch_pdl(pose_to_A_bwd_kernel, B, (J * 12 + 31) / 32 * 32, (size_t)J * 69 * 4, stream, pose, rest, parents, inv_A, G, dA, J, d_pose);
    SGS_LAUNCH_OK();
    return 0;
}

// ------------------------------------------------------------------------------------------
// shared-memory tile of one CTA (256 consecutive Gaussians).  Every per-Gaussian array is a
// contiguous block in global memory, so it is moved as a block: one TMA bulk copy
// (cp.async.bulk + mbarrier) when the block is 16-byte aligned and a multiple of 16 bytes,
// a cooperative coalesced copy otherwise (partial last CTA, unaligned views).  Threads then
// read their own row from shared memory (strides 3, 4, 9, 12 words: bank-conflict free).
// Results leave the same way (cp.async.bulk shared -> global, or coalesced stores).
// ------------------------------------------------------------------------------------------
struct LbsTile {
    float4* A;        // [B][J][3]  rows 0..2 of each joint transform
    float* frame;     // [B][16]: smpl_scale, transl(3), ext_scale, ext_trans(3), ext_quat(4)
    float* W;         // [256][J]
    float* xyz;       // [256][3]   (backward: reused for d_xyz at the end)
    float* scl;       // [256][3]   (backward: reused for d_scales)
    float* rot;       // [256][9]   (backward: reused for d_rot)
    float* f_xyz;     // [256][3]   forward: output staging; backward: this frame's g_xyz
    float* f_q;       // [256][4]
    float* f_scl;     // [256][3]
    float* dT;        // backward: [256][12]
    float* part;      // backward: [8][J][12] per-warp partial dA
    unsigned long long* bar;
};

__host__ __device__ inline size_t lbs_tile_bytes(int B, int J, bool iso, bool bwd, LbsTile* t = nullptr,
                                                 char* raw = nullptr) {
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t at = o; o += align_up(bytes, 16); return at; };
    const size_t oA = take((size_t)B * J * 3 * 16), oF = take((size_t)B * 16 * 4);
    const size_t oW = take((size_t)LBS_THREADS * J * 4);
    const size_t oX = take(LBS_THREADS * 12), oS = take(LBS_THREADS * 12);
    const size_t oR = take(iso ? 0 : LBS_THREADS * 36);
    const size_t oFX = take(LBS_THREADS * 12), oFQ = take(LBS_THREADS * 16), oFS = take(LBS_THREADS * 12);
    const size_t oT = take(bwd ? LBS_THREADS * 48 : 0);
    const size_t oP = take(bwd ? (size_t)(LBS_THREADS / 32) * J * 48 : 0);
    const size_t oB = take(16);
    if (t) {
        t->A = reinterpret_cast<float4*>(raw + oA);  t->frame = reinterpret_cast<float*>(raw + oF);
        t->W = reinterpret_cast<float*>(raw + oW);   t->xyz = reinterpret_cast<float*>(raw + oX);
        t->scl = reinterpret_cast<float*>(raw + oS); t->rot = reinterpret_cast<float*>(raw + oR);
        t->f_xyz = reinterpret_cast<float*>(raw + oFX); t->f_q = reinterpret_cast<float*>(raw + oFQ);
        t->f_scl = reinterpret_cast<float*>(raw + oFS); t->dT = reinterpret_cast<float*>(raw + oT);
        t->part = reinterpret_cast<float*>(raw + oP);
        t->bar = reinterpret_cast<unsigned long long*>(raw + oB);
    }
    return o;
}

__device__ __forceinline__ bool bulk_ok(const void* g, unsigned nfloats) {
    return (((uintptr_t)g) & 15) == 0 && (nfloats & 3) == 0 && nfloats > 0;
}

// A set of block loads into shared memory completing on one mbarrier phase.  Usage (all
// threads): begin(); add(...)...; end(); -- thread 0 arms the barrier with the byte total of
// the TMA-eligible blocks and issues them, everyone copies the others, end() waits for both.
struct BlockLoads {
    unsigned long long* bar;
    unsigned parity;
    __device__ __forceinline__ void load(const float* const* src, float* const* dst, const unsigned* n, int count) {
        const int tid = threadIdx.x;
        if (tid == 0) {
            unsigned bytes = 0;
            for (int k = 0; k < count; k++)
                if (bulk_ok(src[k], n[k])) bytes += n[k] * 4;
            mbar_arrive_expect_tx(bar, bytes);
            for (int k = 0; k < count; k++)
                if (bulk_ok(src[k], n[k])) tma_bulk_g2s(dst[k], src[k], n[k] * 4, bar);
        }
        for (int k = 0; k < count; k++)
            if (!bulk_ok(src[k], n[k]))
                for (unsigned f = tid; f < n[k]; f += LBS_THREADS) dst[k][f] = src[k][f];
    }
    __device__ __forceinline__ void wait() {
        __syncthreads();
        mbar_wait(bar, parity);
        parity ^= 1;
    }
};

__device__ __forceinline__ void bulk_s2g(float* gmem_dst, const float* smem_src, unsigned bytes) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem_src);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(sa), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// Store blocks from shared memory.  Call with all threads after the block's writers have
// passed fence_proxy_async() + __syncthreads().  Thread 0 must call bulk_wait_read() (then a
// __syncthreads) before the source is overwritten.
__device__ __forceinline__ void block_store(float* dst, const float* src, unsigned nfloats) {
    if (bulk_ok(dst, nfloats)) {
        if (threadIdx.x == 0) bulk_s2g(dst, src, nfloats * 4);
    } else {
        for (unsigned f = threadIdx.x; f < nfloats; f += LBS_THREADS) dst[f] = src[f];
    }
}

__device__ __forceinline__ void stage_frames(const LbsArgs& a, LbsTile& s) {
    const int tid = threadIdx.x;
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
}

// T (3x4, row-major 12 floats) = sum_j w_j A[b][j]
__device__ __forceinline__ void blend_T(const LbsTile& s, int b, int J, int t, float* T) {
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

// ------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(LBS_THREADS) lbs_fwd_kernel(LbsArgs a, LbsOut o) {
    extern __shared__ __align__(16) char s_raw[];
    const bool iso = a.rot == nullptr;
    const int rw = a.rot6d ? 6 : 9;      // floats per canonical rotation: 6D (Gram-Schmidt here) or matrix
    LbsTile s;
    lbs_tile_bytes(a.B, a.J, iso, false, &s, s_raw);
    const int tid = threadIdx.x;
    const int base = blockIdx.x * LBS_THREADS;
    const int rows = min(LBS_THREADS, a.N - base);
    if (tid == 0) {
        mbar_init(s.bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (!a.early_params) pdl_sync();
    BlockLoads ld{s.bar, 0};
    {
        const float* src[4] = {a.W + (size_t)base * a.J, a.xyz + (size_t)base * 3, a.scales + (size_t)base * 3,
                               iso ? nullptr : a.rot + (size_t)base * rw};
        float* dst[4] = {s.W, s.xyz, s.scl, s.rot};
        const unsigned n[4] = {(unsigned)(rows * a.J), (unsigned)rows * 3, (unsigned)rows * 3, iso ? 0u : (unsigned)(rows * rw)};
        ld.load(src, dst, n, 4);
    }
    if (a.early_params) pdl_sync();
    stage_frames(a, s);
    ld.wait();
    const int n = base + tid;
    const bool live = n < a.N;
    float x = 0, y = 0, z = 0, s0 = 0, s1 = 0, s2 = 0;
    float Rc[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    float d6[6] = {1, 0, 0, 0, 1, 0};
    if (live) {
        x = s.xyz[3 * tid]; y = s.xyz[3 * tid + 1]; z = s.xyz[3 * tid + 2];
        s0 = s.scl[3 * tid]; s1 = s.scl[3 * tid + 1]; s2 = s.scl[3 * tid + 2];
        if (!iso) {
            if (a.rot6d) {      // rotation_6d_to_matrix (sings_hybrid.py:354-356)
#pragma unroll
                for (int k = 0; k < 6; k++) d6[k] = s.rot[6 * tid + k];
                rot6d_to_mat(d6, Rc);
            } else {
#pragma unroll
                for (int k = 0; k < 9; k++) Rc[k] = s.rot[9 * tid + k];
            }
        }
    }
    for (int b = 0; b < a.B; b++) {
        if (live) {
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
            s.f_xyz[3 * tid] = vx; s.f_xyz[3 * tid + 1] = vy; s.f_xyz[3 * tid + 2] = vz;
            reinterpret_cast<float4*>(s.f_q)[tid] = make_float4(q[0], q[1], q[2], q[3]);
            s.f_scl[3 * tid] = sc0; s.f_scl[3 * tid + 1] = sc1; s.f_scl[3 * tid + 2] = sc2;
            if (o.T) {
                const size_t on = (size_t)b * a.N + n;
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
        fence_proxy_async();
        __syncthreads();
        const size_t ob = (size_t)b * a.N + base;
        block_store(o.xyz + ob * 3, s.f_xyz, (unsigned)rows * 3);
        block_store(o.rotq + ob * 4, s.f_q, (unsigned)rows * 4);
        block_store(o.scales + ob * 3, s.f_scl, (unsigned)rows * 3);
        if (tid == 0) {
            bulk_commit();
            if (b + 1 < a.B) bulk_wait_read();
        }
        if (b + 1 < a.B) __syncthreads();
    }
    if (tid == 0) bulk_wait_read();      // shared memory must outlive the outstanding bulk stores
}

int launch_lbs_fwd(const LbsArgs& a, const LbsOut& o, cudaStream_t stream) {
    if (a.N <= 0 || a.B <= 0) return 0;
    if (a.J < 1 || a.J > 64) return SGS_ERR_BAD_JOINTS;
    if (((uintptr_t)a.A & 15) || (o.T && ((uintptr_t)o.T & 15))) return SGS_ERR_MISALIGNED;
    const size_t smem = lbs_tile_bytes(a.B, a.J, a.rot == nullptr, false);
    if (smem > 220 * 1024) return SGS_ERR_CAPACITY;
    SGS_CUDA_OK(set_max_smem(lbs_fwd_kernel, smem));
    launch_pdl(lbs_fwd_kernel, (a.N + LBS_THREADS - 1) / LBS_THREADS, LBS_THREADS, smem, stream, a, o);
    SGS_LAUNCH_OK();
    return 0;
}

// ------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(LBS_THREADS, SGS_LBS_BWD_MINB) lbs_bwd_kernel(LbsArgs a, LbsGrads g) {
    extern __shared__ __align__(16) char s_raw[];
    const bool iso = a.rot == nullptr;
    const int rw = a.rot6d ? 6 : 9;      // floats per canonical rotation: 6D (Gram-Schmidt here) or matrix
    LbsTile s;
    lbs_tile_bytes(a.B, a.J, iso, true, &s, s_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int base = blockIdx.x * LBS_THREADS;
    const int rows = min(LBS_THREADS, a.N - base);
    if (tid == 0) {
        mbar_init(s.bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    pdl_sync();
    BlockLoads ld{s.bar, 0};
    {
        const float* src[7] = {a.W + (size_t)base * a.J, a.xyz + (size_t)base * 3, a.scales + (size_t)base * 3,
                               iso ? nullptr : a.rot + (size_t)base * rw, g.g_xyz + (size_t)base * 3,
                               g.g_rotq + (size_t)base * 4, g.g_scales + (size_t)base * 3};
        float* dst[7] = {s.W, s.xyz, s.scl, s.rot, s.f_xyz, s.f_q, s.f_scl};
        const unsigned n[7] = {(unsigned)(rows * a.J), (unsigned)rows * 3, (unsigned)rows * 3, iso ? 0u : (unsigned)(rows * rw),
                               (unsigned)rows * 3, (unsigned)rows * 4, (unsigned)rows * 3};
        ld.load(src, dst, n, 7);
    }
    stage_frames(a, s);
    ld.wait();
    const int n = base + tid;
    const bool live = n < a.N;
    float x = 0, y = 0, z = 0, s0 = 0, s1 = 0, s2 = 0;
    float Rc[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    float d6[6] = {1, 0, 0, 0, 1, 0};
    if (live) {
        x = s.xyz[3 * tid]; y = s.xyz[3 * tid + 1]; z = s.xyz[3 * tid + 2];
        s0 = s.scl[3 * tid]; s1 = s.scl[3 * tid + 1]; s2 = s.scl[3 * tid + 2];
        if (!iso) {
            if (a.rot6d) {      // rotation_6d_to_matrix (sings_hybrid.py:354-356)
#pragma unroll
                for (int k = 0; k < 6; k++) d6[k] = s.rot[6 * tid + k];
                rot6d_to_mat(d6, Rc);
            } else {
#pragma unroll
                for (int k = 0; k < 9; k++) Rc[k] = s.rot[9 * tid + k];
            }
        }
    }
    float dx = 0, dy = 0, dz = 0, ds0 = 0, ds1 = 0, ds2 = 0;
    float dRc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    const int warp_rows = max(0, min(32, rows - warp * 32));
    for (int b = 0; b < a.B; b++) {
        if (b > 0) {           // this frame's incoming gradients (frame 0 came with the tile)
            __syncthreads();   // everyone is done with the previous frame's f_* blocks
            const size_t ob = (size_t)b * a.N + base;
            const float* src[3] = {g.g_xyz + ob * 3, g.g_rotq + ob * 4, g.g_scales + ob * 3};
            float* dst[3] = {s.f_xyz, s.f_q, s.f_scl};
            const unsigned nn[3] = {(unsigned)rows * 3, (unsigned)rows * 4, (unsigned)rows * 3};
            ld.load(src, dst, nn, 3);
            ld.wait();
        }
        float dT[12];
#pragma unroll
        for (int k = 0; k < 12; k++) dT[k] = 0.0f;
        float gtr[3] = {0, 0, 0}, gss = 0.0f;
        if (live) {
            float T[12];
            blend_T(s, b, a.J, tid, T);
            const float* fr = s.frame + 16 * b;
            const size_t on = (size_t)b * a.N + n;
            float gx[3] = {s.f_xyz[3 * tid], s.f_xyz[3 * tid + 1], s.f_xyz[3 * tid + 2]};
            const float4 gq4 = reinterpret_cast<const float4*>(s.f_q)[tid];
            float gq[4] = {gq4.x, gq4.y, gq4.z, gq4.w};
            float gs[3] = {s.f_scl[3 * tid], s.f_scl[3 * tid + 1], s.f_scl[3 * tid + 2]};
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
        // ---- dA[b][j] = sum_n W[n][j] dT_n.  Each warp multiplies its own 32 rows: lane =
        // joint (two joints per lane when J > 32), 12 accumulators per joint, the dT row is a
        // broadcast LDS.128 x3 and the weight column a conflict-free LDS per row.  The eight
        // per-warp partials are summed through shared memory; one atomic per CTA and entry.
        float4* dTrow = reinterpret_cast<float4*>(s.dT) + tid * 3;
        dTrow[0] = make_float4(dT[0], dT[1], dT[2], dT[3]);
        dTrow[1] = make_float4(dT[4], dT[5], dT[6], dT[7]);
        dTrow[2] = make_float4(dT[8], dT[9], dT[10], dT[11]);
        __syncwarp();
#pragma unroll 1
        for (int jb = 0; jb < a.J; jb += 32) {
            const int j = jb + lane;
            float acc[12];
#pragma unroll
            for (int k = 0; k < 12; k++) acc[k] = 0.0f;
            if (j < a.J) {
                const float* wcol = s.W + (size_t)(warp * 32) * a.J + j;
                const float4* drow = reinterpret_cast<const float4*>(s.dT) + (warp * 32) * 3;
#pragma unroll 4
                for (int l = 0; l < warp_rows; l++) {
                    const float w = wcol[(size_t)l * a.J];
                    const float4 d0 = drow[3 * l], d1 = drow[3 * l + 1], d2 = drow[3 * l + 2];
                    acc[0] = fmaf(w, d0.x, acc[0]); acc[1] = fmaf(w, d0.y, acc[1]); acc[2] = fmaf(w, d0.z, acc[2]); acc[3] = fmaf(w, d0.w, acc[3]);
                    acc[4] = fmaf(w, d1.x, acc[4]); acc[5] = fmaf(w, d1.y, acc[5]); acc[6] = fmaf(w, d1.z, acc[6]); acc[7] = fmaf(w, d1.w, acc[7]);
                    acc[8] = fmaf(w, d2.x, acc[8]); acc[9] = fmaf(w, d2.y, acc[9]); acc[10] = fmaf(w, d2.z, acc[10]); acc[11] = fmaf(w, d2.w, acc[11]);
                }
                float4* pr = reinterpret_cast<float4*>(s.part) + ((size_t)warp * a.J + j) * 3;
                pr[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
                pr[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
                pr[2] = make_float4(acc[8], acc[9], acc[10], acc[11]);
            }
        }
        __syncthreads();
        for (int oidx = tid; oidx < a.J * 12; oidx += LBS_THREADS) {
            float accv = 0.0f;
#pragma unroll
            for (int w = 0; w < LBS_THREADS / 32; w++) accv += s.part[(size_t)w * a.J * 12 + oidx];
            const int j = oidx / 12, c = oidx - j * 12;
            atomicAdd(g.d_A + ((size_t)b * a.J + j) * 16 + 4 * (c >> 2) + (c & 3), accv);
        }
    }
    // ---- canonical-parameter gradients leave through the input blocks' shared memory ----
    __syncthreads();
    if (live) {
        s.xyz[3 * tid] = dx; s.xyz[3 * tid + 1] = dy; s.xyz[3 * tid + 2] = dz;
        s.scl[3 * tid] = ds0; s.scl[3 * tid + 1] = ds1; s.scl[3 * tid + 2] = ds2;
        if (g.d_rot && !iso) {
            if (a.rot6d) {
                float in6[6], g6[6];      // re-read: keeps the 6 inputs out of the frame loop's registers
#pragma unroll
                for (int k = 0; k < 6; k++) in6[k] = s.rot[6 * tid + k];
                rot6d_to_mat_bwd(in6, dRc, g6);
#pragma unroll
                for (int k = 0; k < 6; k++) s.rot[6 * tid + k] = g6[k];
            } else {
#pragma unroll
                for (int k = 0; k < 9; k++) s.rot[9 * tid + k] = dRc[k];
            }
        }
    }
    fence_proxy_async();
    __syncthreads();
    block_store(g.d_xyz + (size_t)base * 3, s.xyz, (unsigned)rows * 3);
    block_store(g.d_scales + (size_t)base * 3, s.scl, (unsigned)rows * 3);
    if (g.d_rot && !iso) block_store(g.d_rot + (size_t)base * rw, s.rot, (unsigned)(rows * rw));
    if (tid == 0) {
        bulk_commit();
        bulk_wait_read();
    }
}

int launch_lbs_bwd(const LbsArgs& a, const LbsGrads& g, cudaStream_t stream) {
    if (a.N <= 0 || a.B <= 0) return 0;
    if (a.J < 1 || a.J > 64) return SGS_ERR_BAD_JOINTS;
    if (((uintptr_t)a.A & 15) || (g.g_T && ((uintptr_t)g.g_T & 15))) return SGS_ERR_MISALIGNED;
    const size_t smem = lbs_tile_bytes(a.B, a.J, a.rot == nullptr, true);
    if (smem > 220 * 1024) return SGS_ERR_CAPACITY;
    SGS_CUDA_OK(set_max_smem(lbs_bwd_kernel, smem));
    launch_pdl(lbs_bwd_kernel, (a.N + LBS_THREADS - 1) / LBS_THREADS, LBS_THREADS, smem, stream, a, g);
    SGS_LAUNCH_OK();
    return 0;
}

// ------------------------------------------------------------------------------------------
// packed skinning weights (kernels_lbs.h): one thread per Gaussian collects the non-zero entries
// of its row in ascending joint order; *max_nnz receives the longest row (the caller checks it
// against K once -- the buffer only changes at densification).
// ------------------------------------------------------------------------------------------
__global__ void lbs_pack_weights_kernel(int N, int J, const float* __restrict__ W, int K, float* __restrict__ wq,
                                        unsigned* __restrict__ iq, int* __restrict__ max_nnz) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    int nnz = 0;
    if (n < N) {
        const size_t tile = (size_t)n / LBS_PACK_TILE, t = (size_t)n % LBS_PACK_TILE;
        unsigned word = 0;
        for (int j = 0; j < J; j++) {
            const float w = W[(size_t)n * J + j];
            if (w != 0.0f) {
                if (nnz < K) {
                    wq[(tile * K + nnz) * LBS_PACK_TILE + t] = w;
                    word |= (unsigned)j << (8 * (nnz & 3));
                    if ((nnz & 3) == 3) { iq[(tile * (K / 4) + nnz / 4) * LBS_PACK_TILE + t] = word; word = 0; }
                }
                nnz++;
            }
        }
        for (int k = nnz; k < K; k++) {           // unused slots: weight 0 (joint 0)
            wq[(tile * K + k) * LBS_PACK_TILE + t] = 0.0f;
            if ((k & 3) == 3) { iq[(tile * (K / 4) + k / 4) * LBS_PACK_TILE + t] = word; word = 0; }
        }
    }
    const int m = __reduce_max_sync(0xffffffffu, nnz);
    if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(max_nnz, m);
}

int launch_lbs_pack_weights(int N, int J, const float* W, int K, float* wq, unsigned* iq, int* max_nnz,
                            cudaStream_t stream) {
    if (N <= 0) return 0;
    if (J < 1 || J > 64) return SGS_ERR_BAD_JOINTS;
    if (K < 4 || K > LBS_PACK_MAX_K || (K & 3)) return SGS_ERR_BAD_ARG;
    lbs_pack_weights_kernel<<<(N + 255) / 256, 256, 0, stream>>>(N, J, W, K, wq, iq, max_nnz);
    SGS_LAUNCH_OK();
    return 0;
}

// ------------------------------------------------------------------------------------------
// stand-alone 6D-rotation conversions (one thread per rotation): the pose parameters are stored
// as 6D rotations and turned into axis-angle every frame (sings_hybrid.py:370-376:
// rotation_6d_to_axis_angle = 6D -> matrix -> quaternion -> axis-angle, rotations.py:601-603).
// mode 0: matrix out (n,9); mode 1: axis-angle out (n,3).
// ------------------------------------------------------------------------------------------
__global__ void rot6d_convert_kernel(const float* __restrict__ d6_all, int n, int mode, float* __restrict__ out) {
    pdl_sync();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float d6[6], R[9];
#pragma unroll
    for (int k = 0; k < 6; k++) d6[k] = d6_all[(size_t)i * 6 + k];
    rot6d_to_mat(d6, R);
    if (mode == 0) {
#pragma unroll
        for (int k = 0; k < 9; k++) out[(size_t)i * 9 + k] = R[k];
    } else {
        float q[4], aa[3];
        mat_to_quat(R, q);
        quat_to_axis_angle(q, aa);
#pragma unroll
        for (int k = 0; k < 3; k++) out[(size_t)i * 3 + k] = aa[k];
    }
}

__global__ void rot6d_convert_bwd_kernel(const float* __restrict__ d6_all, const float* __restrict__ g_out, int n,
                                         int mode, float* __restrict__ g_d6) {
    pdl_sync();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float d6[6], R[9], gR[9], g6[6];
#pragma unroll
    for (int k = 0; k < 6; k++) d6[k] = d6_all[(size_t)i * 6 + k];
    if (mode == 0) {
#pragma unroll
        for (int k = 0; k < 9; k++) gR[k] = g_out[(size_t)i * 9 + k];
    } else {
        rot6d_to_mat(d6, R);
        float q[4], gq[4];
        mat_to_quat(R, q);
        const float g[3] = {g_out[(size_t)i * 3], g_out[(size_t)i * 3 + 1], g_out[(size_t)i * 3 + 2]};
        quat_to_axis_angle_bwd(q, g, gq);
        mat_to_quat_bwd(R, gq, gR);
    }
    rot6d_to_mat_bwd(d6, gR, g6);
#pragma unroll
    for (int k = 0; k < 6; k++) g_d6[(size_t)i * 6 + k] = g6[k];
}

int launch_rot6d_convert(const float* d6, int n, int mode, float* out, cudaStream_t stream) {
    if (n <= 0) return 0;
    launch_pdl(rot6d_convert_kernel, (n + 127) / 128, 128, 0, stream, d6, n, mode, out);
    SGS_LAUNCH_OK();
    return 0;
}

int launch_rot6d_convert_bwd(const float* d6, const float* g_out, int n, int mode, float* g_d6,
                             cudaStream_t stream) {
    if (n <= 0) return 0;
    launch_pdl(rot6d_convert_bwd_kernel, (n + 127) / 128, 128, 0, stream, d6, g_out, n, mode, g_d6);
    SGS_LAUNCH_OK();
    return 0;
}

}  // namespace sgs
