// kernels_lbs.h -- host-side launch functions of lbs.cu (internal, C++).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sgs {

struct LbsArgs {
    int B, N, J;
    const float* A;           // (B,J,16) cano->pose joint transforms, row-major 4x4
    const float* xyz;         // (N,3) canonical means
    const float* W;           // (N,J) skinning weights
    const float* rot;         // (N,9) canonical rotation matrices -- (N,6) 6D rotations when rot6d -- or null (= identity)
    const float* scales;      // (N,3)
    const float* smpl_scale;  // (B) or null
    const float* transl;      // (B,3) or null
    const float* ext_trans;   // (B,3) |
    const float* ext_rot;     // (B,9) | all three set or all null
    const float* ext_scale;   // (B)   |
    // 1: the canonical arrays (xyz, W, rot, scales) were final before the PREVIOUS kernel of
    // the stream started (that kernel being one of ours, e.g. pose_to_A), so their tiles may
    // be fetched ahead of the programmatic-dependency wait (common.cuh, PDL)
    int early_params = 0;
    int rot6d = 0;            // rot holds 6D rotations; d_rot is (N,6)
};

struct LbsOut {
    float* xyz;      // (B,N,3)
    float* rotq;     // (B,N,4) real part first, not normalised
    float* scales;   // (B,N,3)
    float* T;        // (B,N,16) or null
};

struct LbsGrads {
    const float* g_xyz;     // (B,N,3)
    const float* g_rotq;    // (B,N,4)
    const float* g_scales;  // (B,N,3)
    const float* g_T;       // (B,N,16) gradient w.r.t. the optional T output, or null
    float* d_xyz;           // (N,3)   written
    float* d_rot;           // (N,9) -- (N,6) when rot6d -- written, or null
    float* d_scales;        // (N,3)   written
    float* d_A;             // (B,J,16) accumulated (caller zeroes)
    float* d_smpl_scale;    // (B) accumulated (caller zeroes) or null
    float* d_transl;        // (B,3) accumulated (caller zeroes) or null
};

// Packed skinning weights.  lbs_weights is a constant buffer between densifications
// (sings_hybrid.py:724) and each row has only a handful of non-zero entries (<= 4 on SMPL
// vertices, up to ~12 after repeated midpoint subdivision, geometry_ops.py:65-73), so the fused
// per-frame kernels read a compact copy: per tile of 256 Gaussians, K slots, slot-major so a
// warp's loads are contiguous --
//   wq[(tile * K + k) * 256 + t]        k-th non-zero weight of Gaussian tile * 256 + t (0 = unused slot)
//   iq[(tile * K/4 + k/4) * 256 + t]    four joint indices, 8 bits each (byte k % 4)
// in ascending joint order.  Skipping exact zeros leaves T = sum_j w_j A_j unchanged.
constexpr int LBS_PACK_TILE = 256;
constexpr int LBS_PACK_MAX_K = 16;
inline size_t lbs_packed_tiles(int N) { return (size_t)((N > 0 ? N : 1) + LBS_PACK_TILE - 1) / LBS_PACK_TILE; }
int launch_lbs_pack_weights(int N, int J, const float* W, int K, float* wq, unsigned* iq, int* max_nnz,
                            cudaStream_t stream);

// What the fused LBS + rasterizer-geometry kernels need of the deformer (one frame, B = 1).
struct LbsFuse {
    int J, K, rot6d;
    const float* A;           // (J,16) cano->pose joint transforms of this frame
    const float* xyz;         // (N,3) canonical means
    const float* scales;      // (N,3)
    const float* rot;         // (N,9) / (N,6) when rot6d / null (= identity)
    const float* wq;          // packed weights
    const unsigned* iq;       // packed joint indices
    const float* smpl_scale;  // (1) or null
    const float* transl;      // (3) or null
    // forward outputs = the rasterizer's inputs (kept for inspection and for the drop-in boundary)
    float* xyz_out;           // (N,3)
    float* rotq_out;          // (N,4)
    float* scales_out;        // (N,3)
    // backward outputs
    float* d_xyz;             // (N,3)
    float* d_rot;             // (N,9) / (N,6) or null
    float* d_scales;          // (N,3)
    float* d_A;               // (J,16) accumulated (caller zeroes)
    float* d_transl;          // (3) accumulated (caller zeroes) or null
};

int launch_pose_to_A(const float* pose, const float* rest, const int* parents, const float* inv_A,
                     int B, int J, float* A_out, float* G_out, cudaStream_t stream);
int launch_pose_to_A_bwd(const float* pose, const float* rest, const int* parents,
                         const float* inv_A, const float* G, const float* dA, int B, int J,
                         float* d_pose, cudaStream_t stream);
int launch_rot6d_convert(const float* d6, int n, int mode, float* out, cudaStream_t stream);
int launch_rot6d_convert_bwd(const float* d6, const float* g_out, int n, int mode, float* g_d6,
                             cudaStream_t stream);
int launch_lbs_fwd(const LbsArgs& a, const LbsOut& o, cudaStream_t stream);
int launch_lbs_bwd(const LbsArgs& a, const LbsGrads& g, cudaStream_t stream);

}  // namespace sgs
