// api.cu -- the C ABI of libsings_b200.so (include/sings_b200.h).
#include "../../include/sings_b200.h"
#include "common.cuh"
#include "kernels.h"
#include "kernels_lbs.h"

#include <stdlib.h>

using namespace sgs;

#include <mutex>
#include <vector>

bool sgs::pdl_enabled() {
    static const bool v = getenv("SGS_NO_PDL") == nullptr;      // C++11: initialised once, thread-safe
    return v;
}

// Per-(kernel, device) launch attributes are process state of the CUDA runtime, not of this
// library; what we cache about them is guarded by one mutex so that concurrent host threads
// (each on its own stream) may call every entry point.
namespace {
struct FuncState { const void* f; int dev; size_t smem; int threads; size_t occ_smem; long long resident; };
std::mutex g_mu;
std::vector<FuncState> g_funcs;
FuncState& func_state(const void* f, int dev) {
    for (auto& s : g_funcs)
        if (s.f == f && s.dev == dev) return s;
    g_funcs.push_back(FuncState{f, dev, 0, 0, 0, -1});
    return g_funcs.back();
}
}  // namespace

cudaError_t sgs::ensure_max_smem(const void* func, size_t bytes) {
    // (no shortcut below 48 KB: the limit without opt-in counts the kernel's STATIC shared memory too)
    if (bytes == 0) return cudaSuccess;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lk(g_mu);
    FuncState& s = func_state(func, dev);
    if (bytes <= s.smem) return cudaSuccess;
    e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) s.smem = bytes;
    return e;
}

long long sgs::resident_ctas(const void* func, int threads, size_t smem) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    std::lock_guard<std::mutex> lk(g_mu);
    FuncState& s = func_state(func, dev);
    if (s.resident >= 0 && s.threads == threads && s.occ_smem == smem) return s.resident;
    int sms = 0, per_sm = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, func, threads, smem) != cudaSuccess) return 0;
    s.threads = threads; s.occ_smem = smem; s.resident = (long long)sms * per_sm;
    return s.resident;
}

struct Timing {
    int n;
    cudaEvent_t* ev;
    unsigned mask;        // events that are recorded (bit i = event i); the others are skipped
};
// Inside a stream capture the record becomes an "external" event-record node, so the stage
// events keep working (cudaEventElapsedTime) when the frame is replayed as a CUDA graph.
static inline cudaError_t record_event(cudaEvent_t ev, cudaStream_t s) {
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(s, &st) == cudaSuccess && st == cudaStreamCaptureStatusActive)
        return cudaEventRecordWithFlags(ev, s, cudaEventRecordExternal);
    return cudaEventRecord(ev, s);
}
static inline void tick(void* timing, int i, cudaStream_t s) {
    if (!timing) return;
    Timing* t = (Timing*)timing;
    if (i < t->n && ((t->mask >> i) & 1u)) record_event(t->ev[i], s);
}

extern "C" {

int sgs_version(void) { return 210; }

int sgs_timing_create(int n_events, void** handle) {
    if (n_events < 1 || !handle) return SGS_ERR_BAD_ARG;
    Timing* t = new Timing;
    t->n = n_events;
    t->mask = 0xffffffffu;
    t->ev = new cudaEvent_t[n_events];
    for (int i = 0; i < n_events; i++) SGS_CUDA_OK(cudaEventCreate(&t->ev[i]));
    *handle = t;
    return 0;
}
int sgs_timing_destroy(void* handle) {
    if (!handle) return 0;
    Timing* t = (Timing*)handle;
    for (int i = 0; i < t->n; i++) cudaEventDestroy(t->ev[i]);
    delete[] t->ev;
    delete t;
    return 0;
}
int sgs_timing_record(void* handle, int i, sgs_stream_t stream) {
    if (!handle || i < 0 || i >= ((Timing*)handle)->n) return SGS_ERR_BAD_ARG;
    if (!((((Timing*)handle)->mask >> i) & 1u)) return 0;
    SGS_CUDA_OK(record_event(((Timing*)handle)->ev[i], (cudaStream_t)stream));
    return 0;
}
int sgs_timing_set_mask(void* handle, unsigned mask) {
    if (!handle) return SGS_ERR_BAD_ARG;
    ((Timing*)handle)->mask = mask;
    return 0;
}
int sgs_timing_elapsed_ms(void* handle, int i, int j, float* ms) {
    if (!handle || !ms) return SGS_ERR_BAD_ARG;
    Timing* t = (Timing*)handle;
    if (i < 0 || j < 0 || i >= t->n || j >= t->n) return SGS_ERR_BAD_ARG;
    SGS_CUDA_OK(cudaEventSynchronize(t->ev[j]));
    SGS_CUDA_OK(cudaEventElapsedTime(ms, t->ev[i], t->ev[j]));
    return 0;
}

// A frame recorded once and replayed as one launch.  Plain CUDA graph capture of whatever the
// caller enqueues on `stream` between begin and end (library calls only -- nothing here knows
// about torch's allocator); used by AvatarStep.capture for the pure C-ABI frame, whose replay
// then costs a single cudaGraphLaunch (torch.cuda.CUDAGraph.replay adds two RNG-state fill
// kernels in front of every launch).
int sgs_graph_begin(sgs_stream_t stream) {
    SGS_CUDA_OK(cudaStreamBeginCapture((cudaStream_t)stream, cudaStreamCaptureModeRelaxed));
    return 0;
}
int sgs_graph_end(sgs_stream_t stream, void** graph_exec) {
    if (!graph_exec) return SGS_ERR_BAD_ARG;
    cudaGraph_t g = nullptr;
    SGS_CUDA_OK(cudaStreamEndCapture((cudaStream_t)stream, &g));
    cudaGraphExec_t e = nullptr;
    cudaError_t err = cudaGraphInstantiate(&e, g, 0);
    cudaGraphDestroy(g);
    if (err != cudaSuccess) return (int)err;
    *graph_exec = e;
    return 0;
}
int sgs_graph_launch(void* graph_exec, sgs_stream_t stream) {
    if (!graph_exec) return SGS_ERR_BAD_ARG;
    SGS_CUDA_OK(cudaGraphLaunch((cudaGraphExec_t)graph_exec, (cudaStream_t)stream));
    return 0;
}
int sgs_graph_destroy(void* graph_exec) {
    if (graph_exec) cudaGraphExecDestroy((cudaGraphExec_t)graph_exec);
    return 0;
}

const char* sgs_error_string(int code) {
    switch (code) {
        case 0: return "ok";
        case SGS_ERR_BAD_ARG: return "sings_b200: bad argument";
        case SGS_ERR_BAD_SH_DEGREE: return "sings_b200: SH degree must be 0..3 and M >= (D+1)^2";
        case SGS_ERR_BAD_JOINTS: return "sings_b200: joint count must be 1..64";
        case SGS_ERR_MISALIGNED: return "sings_b200: pointer not 16-byte aligned";
        case SGS_ERR_CAPACITY: return "sings_b200: capacity exceeded";
        default: break;
    }
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "sings_b200: unknown error";
}

int sgs_raster_sizes(int P, int W, int H, long long L_cap, size_t* geom_bytes,
                     size_t* binning_bytes, size_t* img_bytes, size_t* acc_bytes) {
    if (P < 0 || W <= 0 || H <= 0 || L_cap < 0) return SGS_ERR_BAD_ARG;
    RasterLayout l = raster_layout(P, W, H, L_cap);
    if (geom_bytes) *geom_bytes = l.geom_bytes;
    if (binning_bytes) *binning_bytes = l.bin_bytes;
    if (img_bytes) *img_bytes = l.img_bytes;
    if (acc_bytes) *acc_bytes = acc_total_bytes(P);
    return 0;
}

int sgs_raster_clear(int P, int W, int H, long long L_cap, void* binning, void* acc, void* extra,
                     size_t extra_bytes, sgs_stream_t stream) {
    if (P < 0 || W <= 0 || H <= 0 || L_cap < 1 || !binning) return SGS_ERR_BAD_ARG;
    if (((uintptr_t)binning & 15) || ((uintptr_t)acc & 15) || ((uintptr_t)extra & 15)) return SGS_ERR_MISALIGNED;
    RasterLayout l = raster_layout(P, W, H, L_cap);
    return launch_clear3(binning, l.zero_bytes, acc, acc ? acc_total_bytes(P) : 0, extra, extra ? extra_bytes : 0,
                         (cudaStream_t)stream);
}

int sgs_raster_layout_info(int P, int W, int H, long long L_cap, long long* info) {
    if (!info || P < 0 || W <= 0 || H <= 0 || L_cap < 0) return SGS_ERR_BAD_ARG;
    RasterLayout l = raster_layout(P, W, H, L_cap);
    info[0] = (long long)l.cnt_off;
    info[1] = (long long)l.keys0_off;
    info[2] = (long long)l.vals0_off;
    info[3] = (long long)(l.sorted_in_1() ? l.keys1_off : l.keys0_off);
    info[4] = (long long)(l.sorted_in_1() ? l.vals1_off : l.vals0_off);
    info[5] = (long long)l.ranges_off;
    info[6] = (long long)l.finalT_off;
    info[7] = (long long)l.ncontrib_off;
    info[8] = l.tiles;
    info[9] = l.end_bit;
    info[10] = l.passes;
    info[11] = REC_FLOATS;
    return 0;
}

static int fill_geom_args(GeomArgs& a, int P, int D, int M, int W, int H, const float* means3D,
                          const float* colors_precomp, const float* opacities,
                          const float* scales, float scale_modifier, const float* rotations,
                          const float* cov3D_precomp, const float* view, const float* proj,
                          const float* campos, float tanfovx, float tanfovy, const float* shs,
                          int prefiltered) {
    if (P < 0 || W <= 0 || H <= 0) return SGS_ERR_BAD_ARG;
    if (P >= (1 << 24)) return SGS_ERR_CAPACITY;      // a pair-list entry = Gaussian id (24 bits) | reach mask (8 bits)
    if (P > 0) {
        if ((shs == nullptr) == (colors_precomp == nullptr)) return SGS_ERR_BAD_ARG;
        if ((cov3D_precomp == nullptr) == (scales == nullptr || rotations == nullptr)) return SGS_ERR_BAD_ARG;
        if (shs && (D < 0 || D > 3 || M < (D + 1) * (D + 1))) return SGS_ERR_BAD_SH_DEGREE;
        if (rotations && ((uintptr_t)rotations & 15)) return SGS_ERR_MISALIGNED;
        if (!means3D || !view || !proj || !campos) return SGS_ERR_BAD_ARG;
    }
    a.P = P; a.D = shs ? D : 0; a.M = M; a.W = W; a.H = H;
    a.tanfovx = tanfovx; a.tanfovy = tanfovy; a.scale_modifier = scale_modifier;
    a.means3D = means3D; a.scales = scales; a.rotations = rotations; a.opacities = opacities;
    a.shs = shs; a.colors_precomp = colors_precomp; a.cov3D_precomp = cov3D_precomp;
    a.view = view; a.proj = proj; a.campos = campos; a.prefiltered = prefiltered;
    return 0;
}

}  // extern "C"

// lf != null: the deform segment runs inside the geometry kernel (sgs_avatar_forward)
static int raster_forward_impl(int P, int D, int M, int W, int H, const float* bg, const float* means3D,
                       const float* colors_precomp, const float* opacities, const float* scales,
                       float scale_modifier, const float* rotations, const float* cov3D_precomp,
                       const float* viewmatrix, const float* projmatrix, const float* campos,
                       float tanfovx, float tanfovy, const float* shs, int prefiltered,
                       long long L_cap, void* geom, void* binning, void* img, float* out_color,
                       int* radii, float* out_alpha, float* out_depth, int* host_counters,
                       sgs_stream_t stream_, int debug, void* timing, const LbsFuse* lf) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GeomArgs a;
    int rc = fill_geom_args(a, P, D, M, W, H, means3D, colors_precomp, opacities, scales,
                            scale_modifier, rotations, cov3D_precomp, viewmatrix, projmatrix,
                            campos, tanfovx, tanfovy, shs, prefiltered);
    if (rc) return rc;
    if (!bg || !geom || !binning || !img || !out_color || (P > 0 && (!radii || !opacities)) || L_cap < 1)
        return SGS_ERR_BAD_ARG;
    if (L_cap >= (1ll << 30)) return SGS_ERR_CAPACITY;
    if (((uintptr_t)geom & 15) || ((uintptr_t)binning & 15) || ((uintptr_t)img & 15)) return SGS_ERR_MISALIGNED;
    RasterLayout lay = raster_layout(P, W, H, L_cap);
    char* g = (char*)geom; char* b = (char*)binning; char* im = (char*)img;
    const bool precleared = (debug & SGS_FLAG_PRECLEARED) != 0;
    const bool forward_only = (debug & SGS_FLAG_FORWARD_ONLY) != 0;
    a.early_params = (debug & SGS_FLAG_EARLY_PARAMS) != 0;
    debug &= SGS_FLAG_SYNC_CHECK;
    // {num_rendered, overflow} for the host: by the emission kernel when the memory is mapped
    int* host_dev = nullptr;
    if (host_counters && P > 0 && cudaHostGetDevicePointer((void**)&host_dev, host_counters, 0) != cudaSuccess) {
        host_dev = nullptr;
        (void)cudaGetLastError();
    }
    tick(timing, 0, stream);
    rc = launch_geometry(a, lay, L_cap, radii, g, b, stream, !precleared, lf);
    if (rc) return rc;
    if (debug) SGS_CUDA_OK(cudaStreamSynchronize(stream));
    tick(timing, 1, stream);
    if (P > 0) {
        rc = launch_depth_sort(P, lay, b, stream, debug);     // Gaussians by depth (N items)
        if (rc) return rc;
        rc = launch_emit_pairs(P, lay, L_cap, g, b, host_dev, stream);  // scan + (tile|depth, id|reach mask) pairs in depth order
        if (rc) return rc;
        rc = launch_tile_sort(lay, L_cap, g, b, stream, debug);  // stable passes over the tile-id digits (L items)
        if (rc) return rc;
    }
    tick(timing, 2, stream);
    rc = launch_tile_ranges(lay, L_cap, b, stream);      // ranges + tiles bucketed by list length
    if (rc) return rc;
    if (debug) SGS_CUDA_OK(cudaStreamSynchronize(stream));
    tick(timing, 3, stream);
    rc = launch_blend_fwd(lay, W, H, g, b, im, bg, out_color, out_alpha, out_depth, !forward_only, stream);
    if (rc) return rc;
    tick(timing, 4, stream);
    if (debug) SGS_CUDA_OK(cudaStreamSynchronize(stream));
    if (host_counters && !host_dev)
        SGS_CUDA_OK(cudaMemcpyAsync(host_counters, b + lay.cnt_off, 2 * sizeof(int),
                                    cudaMemcpyDeviceToHost, stream));
    return 0;
}

static int raster_backward_impl(int P, int D, int M, int W, int H, const float* bg, const float* means3D,
                        const float* colors_precomp, const float* scales, float scale_modifier,
                        const float* rotations, const float* cov3D_precomp,
                        const float* viewmatrix, const float* projmatrix, const float* campos,
                        float tanfovx, float tanfovy, const float* shs, const int* radii,
                        const float* dL_dout_color, long long L_cap, const void* geom,
                        const void* binning, const void* img, void* acc, float* dL_dmeans3D,
                        float* dL_dmeans2D, float* dL_dcolors, float* dL_dopacity,
                        float* dL_dcov3D, float* dL_dsh, float* dL_dscales, float* dL_drots,
                        float* xyz_gradient_accum, float* denom, float* max_radii2D,
                        sgs_stream_t stream_, int debug, void* timing, const LbsFuse* lf) {
    cudaStream_t stream = (cudaStream_t)stream_;
    GeomBwdArgs b;
    int rc = fill_geom_args(b.fwd, P, D, M, W, H, means3D, colors_precomp, nullptr, scales,
                            scale_modifier, rotations, cov3D_precomp, viewmatrix, projmatrix,
                            campos, tanfovx, tanfovy, shs, 0);
    if (rc) return rc;
    if (!bg || !geom || !binning || !img || !acc || !dL_dout_color || !dL_dmeans2D || !dL_dopacity ||
        (P > 0 && !radii) || (shs && !dL_dsh) || L_cap < 1)
        return SGS_ERR_BAD_ARG;
    if (!lf && (!dL_dmeans3D || (colors_precomp && !dL_dcolors))) return SGS_ERR_BAD_ARG;
    if (((uintptr_t)acc & 15) || (dL_drots && ((uintptr_t)dL_drots & 15))) return SGS_ERR_MISALIGNED;
    const int n_stat = (xyz_gradient_accum != nullptr) + (denom != nullptr) + (max_radii2D != nullptr);
    if (n_stat != 0 && n_stat != 3) return SGS_ERR_BAD_ARG;
    if (P == 0) return 0;
    RasterLayout lay = raster_layout(P, W, H, L_cap);
    const bool precleared = (debug & SGS_FLAG_PRECLEARED) != 0;
    b.fwd.early_params = (debug & SGS_FLAG_EARLY_PARAMS) != 0;
    debug &= SGS_FLAG_SYNC_CHECK;
    tick(timing, 5, stream);
    if (!precleared) SGS_CUDA_OK(cudaMemsetAsync(acc, 0, acc_total_bytes(P), stream));
    rc = launch_blend_bwd(lay, W, H, (const char*)geom, (const char*)binning, (const char*)img, bg,
                          dL_dout_color, (float*)acc, stream);
    if (rc) return rc;
    if (debug) SGS_CUDA_OK(cudaStreamSynchronize(stream));
    tick(timing, 6, stream);
    b.radii = radii; b.acc = (const float*)acc;
    b.dL_dmeans3D = dL_dmeans3D; b.dL_dmeans2D = dL_dmeans2D; b.dL_dcolors = dL_dcolors;
    b.dL_dopacity = dL_dopacity; b.dL_dcov3D = dL_dcov3D; b.dL_dsh = dL_dsh;
    b.dL_dscales = dL_dscales; b.dL_drots = dL_drots;
    b.stat_accum = xyz_gradient_accum; b.stat_denom = denom; b.stat_max_radii = max_radii2D;
    rc = launch_geometry_bwd(b, (const char*)geom, stream, lf);
    if (rc) return rc;
    tick(timing, 7, stream);
    if (debug) SGS_CUDA_OK(cudaStreamSynchronize(stream));
    return 0;
}

extern "C" {

int sgs_raster_forward(int P, int D, int M, int W, int H, const float* bg, const float* means3D,
                       const float* colors_precomp, const float* opacities, const float* scales,
                       float scale_modifier, const float* rotations, const float* cov3D_precomp,
                       const float* viewmatrix, const float* projmatrix, const float* campos,
                       float tanfovx, float tanfovy, const float* shs, int prefiltered,
                       long long L_cap, void* geom, void* binning, void* img, float* out_color,
                       int* radii, float* out_alpha, float* out_depth, int* host_counters,
                       sgs_stream_t stream, int debug, void* timing) {
    return raster_forward_impl(P, D, M, W, H, bg, means3D, colors_precomp, opacities, scales, scale_modifier,
                               rotations, cov3D_precomp, viewmatrix, projmatrix, campos, tanfovx, tanfovy, shs,
                               prefiltered, L_cap, geom, binning, img, out_color, radii, out_alpha, out_depth,
                               host_counters, stream, debug, timing, nullptr);
}

int sgs_raster_backward(int P, int D, int M, int W, int H, const float* bg, const float* means3D,
                        const float* colors_precomp, const float* scales, float scale_modifier,
                        const float* rotations, const float* cov3D_precomp,
                        const float* viewmatrix, const float* projmatrix, const float* campos,
                        float tanfovx, float tanfovy, const float* shs, const int* radii,
                        const float* dL_dout_color, long long L_cap, const void* geom,
                        const void* binning, const void* img, void* acc, float* dL_dmeans3D,
                        float* dL_dmeans2D, float* dL_dcolors, float* dL_dopacity,
                        float* dL_dcov3D, float* dL_dsh, float* dL_dscales, float* dL_drots,
                        float* xyz_gradient_accum, float* denom, float* max_radii2D,
                        sgs_stream_t stream, int debug, void* timing) {
    return raster_backward_impl(P, D, M, W, H, bg, means3D, colors_precomp, scales, scale_modifier, rotations,
                                cov3D_precomp, viewmatrix, projmatrix, campos, tanfovx, tanfovy, shs, radii,
                                dL_dout_color, L_cap, geom, binning, img, acc, dL_dmeans3D, dL_dmeans2D, dL_dcolors,
                                dL_dopacity, dL_dcov3D, dL_dsh, dL_dscales, dL_drots, xyz_gradient_accum, denom,
                                max_radii2D, stream, debug, timing, nullptr);
}

static int fill_fuse(LbsFuse& f, const sgs_deform_args* d) {
    if (!d || d->N < 0) return SGS_ERR_BAD_ARG;
    if (d->J < 1 || d->J > 64) return SGS_ERR_BAD_JOINTS;
    if (d->K < 4 || d->K > LBS_PACK_MAX_K || (d->K & 3)) return SGS_ERR_BAD_ARG;
    if (!d->pose || !d->rest || !d->parents || !d->A || !d->G || !d->xyz_canon || !d->scales || !d->wq || !d->iq ||
        !d->xyz || !d->rotq || !d->scales_out)
        return SGS_ERR_BAD_ARG;
    if (((uintptr_t)d->A & 15) || ((uintptr_t)d->rotq & 15)) return SGS_ERR_MISALIGNED;
    f = LbsFuse{};
    f.J = d->J; f.K = d->K; f.rot6d = d->rot6d;
    f.A = d->A; f.xyz = d->xyz_canon; f.scales = d->scales; f.rot = d->rot_canon;
    f.wq = d->wq; f.iq = d->iq; f.smpl_scale = d->smpl_scale; f.transl = d->transl;
    f.xyz_out = d->xyz; f.rotq_out = d->rotq; f.scales_out = d->scales_out;
    f.d_xyz = d->d_xyz_canon; f.d_rot = d->d_rot_canon; f.d_scales = d->d_scales; f.d_A = d->d_A;
    f.d_transl = d->d_transl;
    return 0;
}

size_t sgs_lbs_packed_bytes(int N, int K) {
    return lbs_packed_tiles(N) * (size_t)(K > 0 ? K : 0) * LBS_PACK_TILE * 4;
}

int sgs_lbs_pack_weights(int N, int J, const float* W, int K, float* wq, unsigned int* iq, int* max_nnz,
                         sgs_stream_t stream) {
    if (N < 0 || (N > 0 && (!W || !wq || !iq || !max_nnz))) return SGS_ERR_BAD_ARG;
    return launch_lbs_pack_weights(N, J, W, K, wq, iq, max_nnz, (cudaStream_t)stream);
}

int sgs_avatar_forward(const sgs_deform_args* d, int D, int M, int W, int H, const float* bg,
                       const float* opacities, float scale_modifier, const float* viewmatrix,
                       const float* projmatrix, const float* campos, float tanfovx, float tanfovy,
                       const float* shs, long long L_cap, void* geom, void* binning, void* img,
                       float* out_color, int* radii, float* out_alpha, float* out_depth, int* host_counters,
                       sgs_stream_t stream, int debug, void* timing) {
    LbsFuse f;
    int rc = fill_fuse(f, d);
    if (rc) return rc;
    if (!shs) return SGS_ERR_BAD_ARG;
    tick(timing, 8, (cudaStream_t)stream);
    rc = launch_pose_to_A(d->pose, d->rest, d->parents, d->inv_A_t2cano, 1, d->J, d->A, d->G, (cudaStream_t)stream);
    if (rc) return rc;
    // the kernel in front of the fused geometry kernel is pose -> A: parameters may be fetched early
    return raster_forward_impl(d->N, D, M, W, H, bg, d->xyz, nullptr, opacities, d->scales_out, scale_modifier,
                               d->rotq, nullptr, viewmatrix, projmatrix, campos, tanfovx, tanfovy, shs, 0, L_cap,
                               geom, binning, img, out_color, radii, out_alpha, out_depth, host_counters, stream,
                               debug | SGS_FLAG_EARLY_PARAMS, timing, &f);
}

int sgs_avatar_backward(const sgs_deform_args* d, int D, int M, int W, int H, const float* bg,
                        float scale_modifier, const float* viewmatrix, const float* projmatrix,
                        const float* campos, float tanfovx, float tanfovy, const float* shs, const int* radii,
                        const float* dL_dout_color, long long L_cap, const void* geom, const void* binning,
                        const void* img, void* acc, float* dL_dmeans2D, float* dL_dopacity, float* dL_dsh,
                        float* xyz_gradient_accum, float* denom, float* max_radii2D, sgs_stream_t stream,
                        int debug, void* timing) {
    LbsFuse f;
    int rc = fill_fuse(f, d);
    if (rc) return rc;
    if (!d->d_xyz_canon || !d->d_scales || !d->d_A || !d->d_pose || (d->rot_canon && !d->d_rot_canon))
        return SGS_ERR_BAD_ARG;
    // (A split arrangement -- the rasterizer's per-Gaussian backward and a packed-weights LBS backward as
    // two kernels -- was measured slower than the fused kernel: 101 us against 84 us for the stage at
    // 200k Gaussians; profiles/README.md, round 2.)
    rc = raster_backward_impl(d->N, D, M, W, H, bg, d->xyz, nullptr, d->scales_out, scale_modifier, d->rotq, nullptr,
                              viewmatrix, projmatrix, campos, tanfovx, tanfovy, shs, radii, dL_dout_color, L_cap,
                              geom, binning, img, acc, nullptr, dL_dmeans2D, nullptr, dL_dopacity, nullptr, dL_dsh,
                              nullptr, nullptr, xyz_gradient_accum, denom, max_radii2D, stream, debug, timing, &f);
    if (rc) return rc;
    tick(timing, 10, (cudaStream_t)stream);
    rc = launch_pose_to_A_bwd(d->pose, d->rest, d->parents, d->inv_A_t2cano, d->G, d->d_A, 1, d->J, d->d_pose,
                              (cudaStream_t)stream);
    if (rc) return rc;
    tick(timing, 11, (cudaStream_t)stream);
    return 0;
}

int sgs_mark_visible(int P, const float* means3D, const float* viewmatrix,
                     unsigned char* present, sgs_stream_t stream) {
    if (P < 0 || (P > 0 && (!means3D || !viewmatrix || !present))) return SGS_ERR_BAD_ARG;
    return launch_mark_visible(P, means3D, viewmatrix, present, (cudaStream_t)stream);
}

int sgs_densify_stats(int P, const float* grad_means2D, const int* radii, float* xyz_gradient_accum,
                      float* denom, float* max_radii2D, sgs_stream_t stream) {
    if (P < 0 || (P > 0 && (!grad_means2D || !radii || !xyz_gradient_accum || !denom || !max_radii2D)))
        return SGS_ERR_BAD_ARG;
    return launch_densify_stats(P, grad_means2D, radii, xyz_gradient_accum, denom, max_radii2D,
                                (cudaStream_t)stream);
}

int sgs_fold_stats(int P, float* step_accum, float* step_denom, float* step_max_radii,
                   float* xyz_gradient_accum, float* denom, float* max_radii2D, sgs_stream_t stream) {
    if (P < 0 || (P > 0 && (!step_accum || !step_denom || !step_max_radii || !xyz_gradient_accum || !denom || !max_radii2D)))
        return SGS_ERR_BAD_ARG;
    return launch_fold_stats(P, step_accum, step_denom, step_max_radii, xyz_gradient_accum, denom, max_radii2D,
                             (cudaStream_t)stream);
}

int sgs_frame_to_u8(const float* image, int H, int W, int bgr, unsigned char* out, sgs_stream_t stream) {
    if (H < 0 || W < 0 || (H > 0 && W > 0 && (!image || !out))) return SGS_ERR_BAD_ARG;
    if ((uintptr_t)out & 3) return SGS_ERR_MISALIGNED;
    return launch_frame_to_u8(image, H, W, bgr, out, (cudaStream_t)stream);
}

size_t sgs_image_loss_scratch_floats(int H, int W) {
    return H > 0 && W > 0 ? (size_t)12 * H * W : 0;     // 9 planes of map derivatives + 3 of the composited target
}

int sgs_image_loss_fwd(int H, int W, const float* pred, const void* gt, int gt_is_u8_hwc, const float* mask,
                       const float* bg, float* scratch, double* sums, float w_l1, float w_ssim, float* loss3,
                       sgs_stream_t stream) {
    if (H <= 0 || W <= 0 || !pred || !gt || !bg || !scratch || !sums) return SGS_ERR_BAD_ARG;
    if ((uintptr_t)sums & 7) return SGS_ERR_MISALIGNED;
    const size_t plane = (size_t)H * W;
    return launch_image_loss_fwd(H, W, pred, gt, gt_is_u8_hwc, mask, bg, scratch, scratch + 9 * plane, sums, w_l1, w_ssim,
                                 loss3, (cudaStream_t)stream);
}

int sgs_image_loss_bwd(int H, int W, const float* pred, const float* scratch, const double* sums,
                       float w_l1, float w_ssim, const float* dloss, float* dL_dpred, float* loss_out,
                       sgs_stream_t stream) {
    if (H <= 0 || W <= 0 || !pred || !scratch || !sums || !dL_dpred) return SGS_ERR_BAD_ARG;
    const size_t plane = (size_t)H * W;
    return launch_image_loss_bwd(H, W, pred, scratch + 9 * plane, scratch, sums, w_l1, w_ssim, dloss, dL_dpred,
                                 loss_out, (cudaStream_t)stream);
}

int sgs_hexplane_fwd(int N, const float* pts, const float* aabb_host, int n_scales, int C, const int* res_host,
                     const float* const* planes_host, float* out, sgs_stream_t stream) {
    if (N < 0 || (N > 0 && (!pts || !out))) return SGS_ERR_BAD_ARG;
    return launch_hexplane_fwd(N, pts, aabb_host, n_scales, C, res_host, planes_host, out, (cudaStream_t)stream);
}

int sgs_hexplane_bwd(int N, const float* pts, const float* aabb_host, int n_scales, int C, const int* res_host,
                     const float* const* planes_host, const float* d_out, float* const* d_planes_host, float* d_pts,
                     sgs_stream_t stream) {
    if (N < 0 || (N > 0 && (!pts || !d_out)) || (!d_planes_host && !d_pts)) return SGS_ERR_BAD_ARG;
    return launch_hexplane_bwd(N, pts, aabb_host, n_scales, C, res_host, planes_host, d_out, d_planes_host, d_pts,
                               (cudaStream_t)stream);
}

size_t sgs_knn_scratch_bytes(int N) {
    int max_cells = 0;
    knn_grid_resolution(N, &max_cells);
    return knn_scratch_bytes(N, max_cells);
}

int sgs_knn_mean_dist(int N, const float* xyz, int K, void* scratch, size_t scratch_bytes, float* mean_dist,
                      int* idx, float* dist2, sgs_stream_t stream) {
    if (N < 0 || (N > 0 && (!xyz || !scratch)) || (!mean_dist && !idx && !dist2)) return SGS_ERR_BAD_ARG;
    return launch_knn(N, xyz, K, (char*)scratch, scratch_bytes, mean_dist, idx, dist2, (cudaStream_t)stream);
}

int sgs_laplacian_loss_fwd(int n, int C, const int* row_ptr, const int* col_idx, const float* vals,
                           const float* row_w, int mode, const float* x, int ldx, float* y, double* sum,
                           float* loss_out, sgs_stream_t stream) {
    if (n < 0 || C < 1 || C > 4 || (mode != 0 && mode != 1) || ldx < C || !sum) return SGS_ERR_BAD_ARG;
    if (n > 0 && (!row_ptr || !col_idx || !vals || !row_w || !x || !y)) return SGS_ERR_BAD_ARG;
    if ((uintptr_t)sum & 7) return SGS_ERR_MISALIGNED;
    return launch_laplacian_loss_fwd(n, C, row_ptr, col_idx, vals, row_w, mode, x, ldx, y, sum, loss_out,
                                     (cudaStream_t)stream);
}

int sgs_laplacian_loss_bwd(int n, int C, const int* t_ptr, const int* t_row, const float* t_val,
                           const float* row_w, int mode, const float* y, const float* dloss, float* dx,
                           sgs_stream_t stream) {
    if (n < 0 || C < 1 || C > 4 || (mode != 0 && mode != 1)) return SGS_ERR_BAD_ARG;
    if (n > 0 && (!t_ptr || !t_row || !t_val || !row_w || !y || !dx)) return SGS_ERR_BAD_ARG;
    return launch_laplacian_loss_bwd(n, C, t_ptr, t_row, t_val, row_w, mode, y, dloss, dx, (cudaStream_t)stream);
}

int sgs_l2norm_fwd(int N, const float* xyz_offsets, const float* scales, int lds, const float* opacity,
                   float max_scale_threshold, float min_opacity_threshold, float lambda_xyz_offsets,
                   float lambda_scales_diff, float lambda_max_scale, float lambda_min_opacity, double* sums,
                   float* loss_out, sgs_stream_t stream) {
    if (N < 0 || !sums || (scales && lds < 1)) return SGS_ERR_BAD_ARG;
    if ((uintptr_t)sums & 7) return SGS_ERR_MISALIGNED;
    return launch_l2norm_fwd(N, xyz_offsets, scales, lds, opacity, max_scale_threshold, min_opacity_threshold,
                             lambda_xyz_offsets, lambda_scales_diff, lambda_max_scale, lambda_min_opacity, sums,
                             loss_out, (cudaStream_t)stream);
}

int sgs_l2norm_bwd(int N, const float* xyz_offsets, const float* scales, int lds, int scale_cols,
                   const float* opacity, float max_scale_threshold, float min_opacity_threshold,
                   const double* sums, float lambda_xyz_offsets, float lambda_scales_diff, float lambda_max_scale,
                   float lambda_min_opacity, const float* dloss, float* d_xyz_offsets, float* d_scales,
                   float* d_opacity, sgs_stream_t stream) {
    if (N < 0 || !sums) return SGS_ERR_BAD_ARG;
    if ((d_xyz_offsets && !xyz_offsets) || (d_scales && (!scales || lds < 1 || scale_cols < 1)) ||
        (d_opacity && !opacity))
        return SGS_ERR_BAD_ARG;
    return launch_l2norm_bwd(N, xyz_offsets, scales, lds, scale_cols, opacity, max_scale_threshold,
                             min_opacity_threshold, sums, lambda_xyz_offsets, lambda_scales_diff, lambda_max_scale,
                             lambda_min_opacity, dloss, d_xyz_offsets, d_scales, d_opacity, (cudaStream_t)stream);
}

size_t sgs_sort_scratch_bytes(long long n) { return sort_scratch_bytes(n < 0 ? 0 : n); }

int sgs_sort_pairs_u64(unsigned long long* keys, unsigned int* vals,
                       unsigned long long* keys_tmp, unsigned int* vals_tmp, void* scratch,
                       size_t scratch_bytes, long long n, int end_bit, int* result_in_tmp,
                       sgs_stream_t stream) {
    if (n > 0 && (!keys || !vals || !keys_tmp || !vals_tmp || !scratch)) return SGS_ERR_BAD_ARG;
    return launch_sort_pairs_u64(keys, vals, keys_tmp, vals_tmp, (char*)scratch, scratch_bytes, n,
                                 end_bit, result_in_tmp, (cudaStream_t)stream);
}

int sgs_pose_to_A(const float* pose, const float* rest, const int* parents,
                  const float* inv_A_t2cano, int B, int J, float* A_out, float* G_out,
                  sgs_stream_t stream) {
    if (B < 0 || (B > 0 && (!pose || !rest || !parents || !A_out))) return SGS_ERR_BAD_ARG;
    return launch_pose_to_A(pose, rest, parents, inv_A_t2cano, B, J, A_out, G_out, (cudaStream_t)stream);
}

int sgs_pose_to_A_bwd(const float* pose, const float* rest, const int* parents,
                      const float* inv_A_t2cano, const float* G, const float* dL_dA, int B, int J,
                      float* dL_dpose, sgs_stream_t stream) {
    if (B < 0 || (B > 0 && (!pose || !rest || !parents || !G || !dL_dA || !dL_dpose))) return SGS_ERR_BAD_ARG;
    return launch_pose_to_A_bwd(pose, rest, parents, inv_A_t2cano, G, dL_dA, B, J, dL_dpose, (cudaStream_t)stream);
}

static int fill_lbs(LbsArgs& a, int B, int N, int J, const float* A, const float* xyz, const float* W,
                    const float* rot, const float* scales, const float* smpl_scale, const float* transl,
                    const float* ext_trans, const float* ext_rot, const float* ext_scale) {
    if (B < 0 || N < 0) return SGS_ERR_BAD_ARG;
    if (B > 0 && N > 0 && (!A || !xyz || !W || !scales)) return SGS_ERR_BAD_ARG;
    const int n_ext = (ext_trans != nullptr) + (ext_rot != nullptr) + (ext_scale != nullptr);
    if (n_ext != 0 && n_ext != 3) return SGS_ERR_BAD_ARG;
    a.B = B; a.N = N; a.J = J; a.A = A; a.xyz = xyz; a.W = W; a.rot = rot; a.scales = scales;
    a.smpl_scale = smpl_scale; a.transl = transl;
    a.ext_trans = ext_trans; a.ext_rot = ext_rot; a.ext_scale = ext_scale;
    return 0;
}

int sgs_lbs_fwd(int B, int N, int J, const float* A, const float* xyz_canon, const float* W,
                const float* rot_canon, const float* scales, const float* smpl_scale,
                const float* transl, const float* ext_trans, const float* ext_rot,
                const float* ext_scale, float* xyz_out, float* rotq_out, float* scales_out,
                float* T_out, sgs_stream_t stream) {
    LbsArgs a;
    int rc = fill_lbs(a, B, N, J, A, xyz_canon, W, rot_canon, scales, smpl_scale, transl, ext_trans, ext_rot, ext_scale);
    if (rc) return rc;
    if (B > 0 && N > 0 && (!xyz_out || !rotq_out || !scales_out)) return SGS_ERR_BAD_ARG;
    LbsOut o{xyz_out, rotq_out, scales_out, T_out};
    return launch_lbs_fwd(a, o, (cudaStream_t)stream);
}

int sgs_pose_lbs_fwd(const float* pose, const float* rest, const int* parents,
                     const float* inv_A_t2cano, int B, int N, int J, float* A_out, float* G_out,
                     const float* xyz_canon, const float* W, const float* rot_canon,
                     const float* scales, const float* smpl_scale, const float* transl,
                     float* xyz_out, float* rotq_out, float* scales_out, sgs_stream_t stream) {
    if (B < 0 || (B > 0 && (!pose || !rest || !parents || !A_out))) return SGS_ERR_BAD_ARG;
    LbsArgs a;
    int rc = fill_lbs(a, B, N, J, A_out, xyz_canon, W, rot_canon, scales, smpl_scale, transl, nullptr, nullptr, nullptr);
    if (rc) return rc;
    if (B > 0 && N > 0 && (!xyz_out || !rotq_out || !scales_out)) return SGS_ERR_BAD_ARG;
    LbsOut o{xyz_out, rotq_out, scales_out, nullptr};
    // (pose -> A recomputed by every LBS CTA in its prologue, i.e. one kernel instead of two, was
    // measured at 25.0 us against 24.7 us for this pair: the small kernel is already hidden behind
    // the LBS kernel's early tile prefetch, and the extra code cost the LBS kernel instruction-cache
    // misses; removed)
    rc = launch_pose_to_A(pose, rest, parents, inv_A_t2cano, B, J, A_out, G_out, (cudaStream_t)stream);
    if (rc) return rc;
    a.early_params = 1;      // the preceding kernel is pose_to_A, which never writes them
    return launch_lbs_fwd(a, o, (cudaStream_t)stream);
}

int sgs_lbs_bwd(int B, int N, int J, const float* A, const float* xyz_canon, const float* W,
                const float* rot_canon, const float* scales, const float* smpl_scale,
                const float* transl, const float* ext_trans, const float* ext_rot,
                const float* ext_scale, const float* g_xyz, const float* g_rotq,
                const float* g_scales, const float* g_T, float* d_xyz_canon, float* d_rot_canon,
                float* d_scales, float* d_A, float* d_smpl_scale, float* d_transl,
                sgs_stream_t stream) {
    LbsArgs a;
    int rc = fill_lbs(a, B, N, J, A, xyz_canon, W, rot_canon, scales, smpl_scale, transl, ext_trans, ext_rot, ext_scale);
    if (rc) return rc;
    if (B > 0 && N > 0 && (!g_xyz || !g_rotq || !g_scales || !d_xyz_canon || !d_scales || !d_A)) return SGS_ERR_BAD_ARG;
    LbsGrads g{g_xyz, g_rotq, g_scales, g_T, d_xyz_canon, d_rot_canon, d_scales, d_A, d_smpl_scale, d_transl};
    return launch_lbs_bwd(a, g, (cudaStream_t)stream);
}

int sgs_lbs_fwd_rot6d(int B, int N, int J, const float* A, const float* xyz_canon, const float* W,
                      const float* rot6d_canon, const float* scales, const float* smpl_scale,
                      const float* transl, const float* ext_trans, const float* ext_rot,
                      const float* ext_scale, float* xyz_out, float* rotq_out, float* scales_out,
                      float* T_out, sgs_stream_t stream) {
    if (B > 0 && N > 0 && !rot6d_canon) return SGS_ERR_BAD_ARG;
    LbsArgs a;
    int rc = fill_lbs(a, B, N, J, A, xyz_canon, W, rot6d_canon, scales, smpl_scale, transl, ext_trans, ext_rot, ext_scale);
    if (rc) return rc;
    if (B > 0 && N > 0 && (!xyz_out || !rotq_out || !scales_out)) return SGS_ERR_BAD_ARG;
    a.rot6d = 1;
    LbsOut o{xyz_out, rotq_out, scales_out, T_out};
    return launch_lbs_fwd(a, o, (cudaStream_t)stream);
}

int sgs_lbs_bwd_rot6d(int B, int N, int J, const float* A, const float* xyz_canon, const float* W,
                      const float* rot6d_canon, const float* scales, const float* smpl_scale,
                      const float* transl, const float* ext_trans, const float* ext_rot,
                      const float* ext_scale, const float* g_xyz, const float* g_rotq,
                      const float* g_scales, const float* g_T, float* d_xyz_canon,
                      float* d_rot6d_canon, float* d_scales, float* d_A, float* d_smpl_scale,
                      float* d_transl, sgs_stream_t stream) {
    if (B > 0 && N > 0 && !rot6d_canon) return SGS_ERR_BAD_ARG;
    LbsArgs a;
    int rc = fill_lbs(a, B, N, J, A, xyz_canon, W, rot6d_canon, scales, smpl_scale, transl, ext_trans, ext_rot, ext_scale);
    if (rc) return rc;
    if (B > 0 && N > 0 && (!g_xyz || !g_rotq || !g_scales || !d_xyz_canon || !d_scales || !d_A)) return SGS_ERR_BAD_ARG;
    a.rot6d = 1;
    LbsGrads g{g_xyz, g_rotq, g_scales, g_T, d_xyz_canon, d_rot6d_canon, d_scales, d_A, d_smpl_scale, d_transl};
    return launch_lbs_bwd(a, g, (cudaStream_t)stream);
}

int sgs_rot6d_to_matrix(const float* d6, int n, float* R_out, sgs_stream_t stream) {
    if (n < 0 || (n > 0 && (!d6 || !R_out))) return SGS_ERR_BAD_ARG;
    return launch_rot6d_convert(d6, n, 0, R_out, (cudaStream_t)stream);
}
int sgs_rot6d_to_matrix_bwd(const float* d6, const float* dL_dR, int n, float* dL_dd6, sgs_stream_t stream) {
    if (n < 0 || (n > 0 && (!d6 || !dL_dR || !dL_dd6))) return SGS_ERR_BAD_ARG;
    return launch_rot6d_convert_bwd(d6, dL_dR, n, 0, dL_dd6, (cudaStream_t)stream);
}
int sgs_rot6d_to_axis_angle(const float* d6, int n, float* aa_out, sgs_stream_t stream) {
    if (n < 0 || (n > 0 && (!d6 || !aa_out))) return SGS_ERR_BAD_ARG;
    return launch_rot6d_convert(d6, n, 1, aa_out, (cudaStream_t)stream);
}
int sgs_rot6d_to_axis_angle_bwd(const float* d6, const float* dL_daa, int n, float* dL_dd6, sgs_stream_t stream) {
    if (n < 0 || (n > 0 && (!d6 || !dL_daa || !dL_dd6))) return SGS_ERR_BAD_ARG;
    return launch_rot6d_convert_bwd(d6, dL_daa, n, 1, dL_dd6, (cudaStream_t)stream);
}

}  // extern "C"
