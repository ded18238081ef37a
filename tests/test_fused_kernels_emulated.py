"""CPU: the SOURCE of image_loss.cu (fused L1 + SSIM) and hexplane.cu (multi-scale tri-plane interpolation)
executed under the SIMT emulation of tests/cuda_emu against the golden vectors produced by the reference's
own functions (tests/golden/make_loss_golden.py, make_hexplane_golden.py).  Complements
tests/test_gpu_image_loss.py and tests/test_gpu_hexplane.py; test infrastructure only."""
import ctypes as C
import glob
import os
import shutil

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "sings_b200", "csrc")
GOLD = os.path.join(ROOT, "tests", "golden")
vp, i32, f32 = C.c_void_p, C.c_int, C.c_float


def p(a):
    return None if a is None else a.ctypes.data


@pytest.fixture(scope="module")
def image_loss():
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    from cuda_emu import build
    L = build(os.path.join(CSRC, "image_loss.cu"), r'''
extern "C" {
int emu_image_loss_fwd(int H, int W, const float* pred, const void* gt, int gt_u8, const float* mask, const float* bg,
                       float* scratch, double* sums, float w_l1, float w_ssim, float* loss3) {
    const size_t plane = (size_t)H * W;          // the split of sgs_image_loss_fwd (api.cu)
    return sgs::launch_image_loss_fwd(H, W, pred, gt, gt_u8, mask, bg, scratch, scratch + 9 * plane, sums, w_l1, w_ssim, loss3, nullptr);
}
int emu_image_loss_bwd(int H, int W, const float* pred, const float* scratch, const double* sums, float w_l1, float w_ssim,
                       const float* dloss, float* dL_dpred, float* loss_out) {
    const size_t plane = (size_t)H * W;
    return sgs::launch_image_loss_bwd(H, W, pred, scratch + 9 * plane, scratch, sums, w_l1, w_ssim, dloss, dL_dpred, loss_out, nullptr);
}
}
''')
    L.emu_image_loss_fwd.argtypes = [i32, i32, vp, vp, i32, vp, vp, vp, vp, f32, f32, vp]
    L.emu_image_loss_bwd.argtypes = [i32, i32, vp, vp, vp, f32, f32, vp, vp, vp]
    return L


@pytest.fixture(scope="module")
def hexplane():
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    from cuda_emu import build
    L = build(os.path.join(CSRC, "hexplane.cu"), r'''
extern "C" {
int emu_hexplane_fwd(int N, const float* pts, const float* aabb, int S, int Cc, const int* res, const float* const* planes, float* out) {
    return sgs::launch_hexplane_fwd(N, pts, aabb, S, Cc, res, planes, out, nullptr);
}
int emu_hexplane_bwd(int N, const float* pts, const float* aabb, int S, int Cc, const int* res, const float* const* planes,
                     const float* d_out, float* const* d_planes, float* d_pts) {
    return sgs::launch_hexplane_bwd(N, pts, aabb, S, Cc, res, planes, d_out, d_planes, d_pts, nullptr);
}
}
''')
    L.emu_hexplane_fwd.argtypes = [i32, vp, vp, i32, i32, vp, vp, vp]
    L.emu_hexplane_bwd.argtypes = [i32, vp, vp, i32, i32, vp, vp, vp, vp, vp]
    return L


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "loss_golden_*_f32.npz"))))
@pytest.mark.parametrize("as_u8", [False, True])
def test_image_loss_kernels(image_loss, path, as_u8):
    z, z64 = np.load(path), np.load(path.replace("_f32", "_f64"))
    pred = np.ascontiguousarray(z["pred"])
    _, H, W = pred.shape
    gt = np.ascontiguousarray(z["gt_u8"]) if as_u8 else np.ascontiguousarray((z["gt_u8"].astype(np.float32) / np.float32(255)).transpose(2, 0, 1))
    mask = np.ascontiguousarray(z["mask"].astype(np.float32)) if z["mask"].size else None
    bg = np.ascontiguousarray(z["bg"].astype(np.float32))
    scratch = np.full(12 * H * W, np.nan, np.float32)
    sums, loss3 = np.full(4, np.nan), np.full(3, np.nan, np.float32)
    assert image_loss.emu_image_loss_fwd(H, W, p(pred), p(gt), int(as_u8), p(mask), p(bg), p(scratch), p(sums), 0.8, 0.2, p(loss3)) == 0
    assert abs(float(loss3[0]) - float(z64["loss"])) <= 1e-5 * abs(float(z64["loss"]))
    assert abs(0.8 * float(loss3[1]) - float(z64["l1"])) <= 1e-5 * abs(float(z64["l1"]))
    assert abs(0.2 * float(loss3[2]) - float(z64["ssim"])) <= 1e-5 * max(abs(float(z64["ssim"])), 0.05)
    grad = np.full((3, H, W), np.nan, np.float32)
    dl = np.asarray([1.0], np.float32)
    assert image_loss.emu_image_loss_bwd(H, W, p(pred), p(scratch), p(sums), 0.8, 0.2, p(dl), p(grad), None) == 0
    ref = z64["grad"]
    assert np.abs(grad - ref).max() <= 1e-4 * np.abs(ref).max()


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "hexplane_golden_*.npz"))))
def test_hexplane_kernels(hexplane, path):
    z = np.load(path)
    Cc, mult, reso = int(z["C"]), [int(v) for v in z["multires"]], [int(v) for v in z["reso"]]
    S = len(mult)
    res = np.asarray([int(r * m) for m in mult for r in reso[:3]], np.int32)
    # the reference's (1, C, H, W) parameters as the kernels read them: channel-last (H, W, C)
    planes = [np.ascontiguousarray(z[f"plane_{i}"][0].transpose(1, 2, 0)) for i in range(3 * S)]
    pts = np.ascontiguousarray(z["pts"].reshape(-1, 3).astype(np.float32))
    N = pts.shape[0]
    aabb = np.ascontiguousarray(z["aabb"].reshape(-1).astype(np.float32))
    ptrs = (vp * (3 * S))(*[a.ctypes.data for a in planes])
    out = np.full((N, S * Cc), np.nan, np.float32)
    assert hexplane.emu_hexplane_fwd(N, p(pts), p(aabb), S, Cc, p(res), ptrs, p(out)) == 0
    ref = z["feats"].reshape(N, -1)
    assert np.abs(out - ref).max() <= 1e-5 * np.abs(ref).max()
    d_out = np.ascontiguousarray(z["d_out"].reshape(N, -1).astype(np.float32))
    d_planes = [np.zeros_like(a) for a in planes]
    d_ptrs = (vp * (3 * S))(*[a.ctypes.data for a in d_planes])
    d_pts = np.full((N, 3), np.nan, np.float32)
    assert hexplane.emu_hexplane_bwd(N, p(pts), p(aabb), S, Cc, p(res), ptrs, p(d_out), d_ptrs, p(d_pts)) == 0
    ref = z["d_pts"].reshape(N, 3)
    assert np.abs(d_pts - ref).max() <= 1e-5 * np.abs(ref).max()
    for i in range(3 * S):
        ref = z[f"d_plane_{i}"][0].transpose(1, 2, 0)
        assert np.abs(d_planes[i] - ref).max() <= 1e-5 * np.abs(ref).max(), i


@pytest.mark.parametrize("H,W,masked", [(1, 1, False), (5, 7, True), (11, 3, False), (33, 65, True)])
def test_image_loss_edge_sizes_vs_oracle(image_loss, H, W, masked):
    """Images smaller than the 11 x 11 window and ragged against the 32 x 32 tile, against float64 autograd of
    the oracle (the reference's l1_loss + ssim composed as HumanLoss.forward)."""
    import torch
    from oracle import loss_oracle as llo
    rng = np.random.default_rng(H * 100 + W)
    pred, gt = rng.random((3, H, W)).astype(np.float32), rng.random((3, H, W)).astype(np.float32)
    mask = None
    if masked:
        mask = (rng.random((H, W)) > 0.3).astype(np.float32)
        mask[0, 0] = 1.0
    bg = np.array([1.0, 0.5, 0.2], np.float32)
    scratch, sums, loss3 = np.full(12 * H * W, np.nan, np.float32), np.full(4, np.nan), np.full(3, np.nan, np.float32)
    assert image_loss.emu_image_loss_fwd(H, W, p(pred), p(gt), 0, p(mask), p(bg), p(scratch), p(sums), 0.8, 0.2, p(loss3)) == 0
    grad = np.full((3, H, W), np.nan, np.float32)
    dl = np.asarray([1.0], np.float32)
    assert image_loss.emu_image_loss_bwd(H, W, p(pred), p(scratch), p(sums), 0.8, 0.2, p(dl), p(grad), None) == 0
    pt = torch.from_numpy(pred).double().requires_grad_(True)
    ref = llo.human_image_loss(pt, torch.from_numpy(gt).double(), None if mask is None else torch.from_numpy(mask).double(),
                               torch.from_numpy(bg).double())[0]
    ref.backward()
    assert abs(float(loss3[0]) - float(ref.detach())) <= 1e-5 * abs(float(ref.detach()))
    assert np.abs(grad - pt.grad.numpy()).max() <= 1e-4 * np.abs(pt.grad.numpy()).max()
