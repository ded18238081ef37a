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
