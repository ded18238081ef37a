// frame_out.cu -- the device half of the animation output path.
//
// The reference turns every rendered frame into an 8-bit image on the HOST:
//   /root/reference/sings/rec/trainer/gs_trainer.py:716-719
//       img_np = (image.detach().cpu().clamp(0, 1).permute(1, 2, 0).numpy() * 255).astype('uint8')
//       img_np = cv2.cvtColor(img_np, cv2.COLOR_RGB2BGR);  cv2.imwrite(..., img_np)
// i.e. a blocking 4-bytes-per-value device->host copy of the float image, then clamp, transpose,
// scale and truncation on one CPU thread.  Here clamp + scale + truncation + CHW->HWC (+ the
// RGB->BGR swap cv2 wants) are one pass over the image on the device, so the copy that follows
// moves a quarter of the bytes and the host only encodes.  Same bits as the reference's
// expression: float32 multiply by 255, conversion toward zero.
#include "common.cuh"
#include "kernels.h"

namespace sgs {

// one thread = 4 consecutive pixels of a row: three coalesced 16-byte loads (one per plane), one
// 12-byte interleaved store
__global__ void __launch_bounds__(256) frame_to_u8_kernel(const float* __restrict__ img, int H, int W, int bgr,
                                                          unsigned char* __restrict__ out) {
    pdl_sync();
    const size_t plane = (size_t)H * W;
    const size_t n4 = (plane + 3) / 4;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += (size_t)gridDim.x * blockDim.x) {
        const size_t p0 = q * 4;
        float v[3][4];
        const bool full = p0 + 3 < plane && (plane & 3) == 0;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            if (full) {
                const float4 f = ldg_stream_f4(reinterpret_cast<const float4*>(img + c * plane + p0));
                v[c][0] = f.x; v[c][1] = f.y; v[c][2] = f.z; v[c][3] = f.w;
            } else {
#pragma unroll
                for (int k = 0; k < 4; k++) v[c][k] = p0 + k < plane ? img[c * plane + p0 + k] : 0.0f;
            }
        }
        unsigned char b[12];
#pragma unroll
        for (int k = 0; k < 4; k++)
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const float x = __fmul_rn(fminf(fmaxf(v[c][k], 0.0f), 1.0f), 255.0f);      // clamp(0,1) * 255 ...
                b[3 * k + (bgr ? 2 - c : c)] = (unsigned char)__float2int_rz(x);            // ... .astype('uint8')
            }
        unsigned char* dst = out + 3 * p0;
        if (full) {         // 12 bytes, 4-byte aligned (3 * 4 q)
            unsigned w[3];
#pragma unroll
            for (int k = 0; k < 3; k++) w[k] = b[4 * k] | (b[4 * k + 1] << 8) | (b[4 * k + 2] << 16) | ((unsigned)b[4 * k + 3] << 24);
            reinterpret_cast<unsigned*>(dst)[0] = w[0];
            reinterpret_cast<unsigned*>(dst)[1] = w[1];
            reinterpret_cast<unsigned*>(dst)[2] = w[2];
        } else {
            for (int k = 0; k < 4 && p0 + k < plane; k++)
                for (int c = 0; c < 3; c++) dst[3 * k + c] = b[3 * k + c];
        }
    }
}

int launch_frame_to_u8(const float* img, int H, int W, int bgr, unsigned char* out, cudaStream_t stream) {
    if (H <= 0 || W <= 0) return 0;
    const size_t n4 = ((size_t)H * W + 3) / 4;
    long long blocks = (long long)((n4 + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    SGS_CUDA_OK(launch_pdl(frame_to_u8_kernel, (unsigned)blocks, 256, 0, stream, img, H, W, bgr, out));
    return 0;
}

}  // namespace sgs
