"""CPU: the SOURCE of sings_b200/csrc/regularizers.cu executed thread for thread under the SIMT emulation
of tests/cuda_emu (g++, every CUDA thread a host thread), driven through the same host code as on the GPU
(sings_b200/regularizers.py builds the operators), against the golden vectors of the reference's own
classes.  This does not replace the GPU parity tests (tests/test_gpu_regularizers.py): it checks indexing,
reductions and the backward formulas of the kernel code where no GPU is available.  Test infrastructure
only -- the product has no CPU path."""
import ctypes as C
import os
import shutil

import numpy as np
import pytest
import torch

from sings_b200 import regularizers as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")

EXPORTS = r'''
extern "C" {
int emu_laplacian_loss_fwd(int n, int C, const int* row_ptr, const int* col_idx, const float* vals, const float* row_w,
                           int mode, const float* x, int ldx, float* y, double* sum, float* loss_out) {
    return sgs::launch_laplacian_loss_fwd(n, C, row_ptr, col_idx, vals, row_w, mode, x, ldx, y, sum, loss_out, nullptr);
}
int emu_laplacian_loss_bwd(int n, int C, const int* t_ptr, const int* t_row, const float* t_val, const float* row_w,
                           int mode, const float* y, const float* dloss, float* dx) {
    return sgs::launch_laplacian_loss_bwd(n, C, t_ptr, t_row, t_val, row_w, mode, y, dloss, dx, nullptr);
}
int emu_l2norm_fwd(int N, const float* off, const float* scales, int lds, const float* opacity, float thr_s, float thr_o,
                   float l_off, float l_diff, float l_max, float l_op, double* sums, float* loss_out) {
    return sgs::launch_l2norm_fwd(N, off, scales, lds, opacity, thr_s, thr_o, l_off, l_diff, l_max, l_op, sums, loss_out, nullptr);
}
int emu_l2norm_bwd(int N, const float* off, const float* scales, int lds, int S, const float* opacity, float thr_s,
                   float thr_o, const double* sums, float l_off, float l_diff, float l_max, float l_op, const float* dloss,
                   float* d_off, float* d_scales, float* d_opacity) {
    return sgs::launch_l2norm_bwd(N, off, scales, lds, S, opacity, thr_s, thr_o, sums, l_off, l_diff, l_max, l_op, dloss,
                                  d_off, d_scales, d_opacity, nullptr);
}
}
'''

POSITION_W = {'head-neck': 0.5, 'spine': 0.75, 'leftUpArm': 1., 'rightUpArm': 1., 'leftDownArm': 1., 'rightDownArm': 1.,
              'leftHand': 1.5, 'rightHand': 1.5, 'hips': 1., 'leftUpLeg': 1., 'rightUpLeg': 1., 'leftDownLeg': 1.,
              'rightDownLeg': 1., 'leftFoot': 0.75, 'rightFoot': 0.75}
COLOR_W = {k: (1.0 if k in ('leftDownArm', 'rightDownArm', 'leftHand', 'rightHand') else 0.0) for k in POSITION_W}


@pytest.fixture(scope="module")
def emu():
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    from cuda_emu import build
    L = build(os.path.join(ROOT, "sings_b200", "csrc", "regularizers.cu"), EXPORTS)
    vp, i, f = C.c_void_p, C.c_int, C.c_float
    L.emu_laplacian_loss_fwd.argtypes = [i, i, vp, vp, vp, vp, i, vp, i, vp, vp, vp]
    L.emu_laplacian_loss_bwd.argtypes = [i, i, vp, vp, vp, vp, i, vp, vp, vp]
    L.emu_l2norm_fwd.argtypes = [i, vp, vp, i, vp, f, f, f, f, f, f, vp, vp]
    L.emu_l2norm_bwd.argtypes = [i, vp, vp, i, i, vp, f, f, vp, f, f, f, f, vp, vp, vp, vp]
    return L


def p(a):
    return None if a is None else a.ctypes.data


def lap_fwd_bwd(L, op, x_full, ldx, C_, row_w, mode, dloss=1.0):
    n = op.n
    rp, ci, va = op.row_ptr.numpy(), op.col_idx.numpy(), op.vals.numpy()
    tp, tr, tv = op.t_ptr.numpy(), op.t_row.numpy(), op.t_val.numpy()
    y = np.full((n, C_), np.nan, np.float32)
    acc = np.full(1, np.nan, np.float64)
    loss = np.full(1, np.nan, np.float32)
    assert L.emu_laplacian_loss_fwd(n, C_, p(rp), p(ci), p(va), p(row_w), mode, p(x_full), ldx, p(y), p(acc), p(loss)) == 0
    dx = np.full((n, C_), np.nan, np.float32)
    dl = np.asarray([dloss], np.float32)
    assert L.emu_laplacian_loss_bwd(n, C_, p(tp), p(tr), p(tv), p(row_w), mode, p(y), p(dl), p(dx)) == 0
    assert float(loss[0]) == np.float32(acc[0])
    return float(acc[0]), dx


def rel(a, b):
    return np.abs(np.asarray(a, np.float64) - b).max() / np.abs(b).max()


@pytest.mark.parametrize("name", ["region_a", "region_b"])
def test_region_laplacian_kernels(emu, name):
    z = np.load(os.path.join(GOLD, f"reg_golden_{name}.npz"))
    labels, edges = torch.from_numpy(z["labels"]), torch.from_numpy(z["edges"])
    pos = R.RegionLaplacianLoss_v2(torch.from_numpy(z["verts"]), edges, labels, region_weights=POSITION_W)
    col = R.RegionLaplacianLoss_v2(torch.from_numpy(z["verts"]), edges, labels, region_weights=COLOR_W)
    x = np.ascontiguousarray(z["xyz"])
    loss, dx = lap_fwd_bwd(emu, pos.operator, x, 3, 3, pos._weights_for("all", pos.weights[:15], 3).numpy(), 0)
    assert abs(loss - z["loss_pos_f64"]) <= 1e-5 * z["loss_pos_f64"] and rel(dx, z["grad_pos_f64"]) <= 1e-5
    w = np.zeros(15)
    w[[6, 7]] = 1000
    loss, dx = lap_fwd_bwd(emu, pos.operator, x, 3, 3, pos._weights_for(("hands", 1000.0), w, 3).numpy(), 0, dloss=0.5)
    assert abs(loss - z["loss_hand_f64"]) <= 1e-5 * z["loss_hand_f64"] and rel(dx, 0.5 * z["grad_hand_f64"]) <= 1e-5
    shs = np.ascontiguousarray(z["shs"])                     # (V, 16, 3): colours = shs[:, 0], row stride 48
    loss, dx = lap_fwd_bwd(emu, col.operator, shs, 48, 3, col._weights_for("all", col.weights[:15], 3).numpy(), 0)
    assert abs(loss - z["loss_col_f64"]) <= 1e-5 * z["loss_col_f64"] and rel(dx, z["grad_col_f64"][:, 0]) <= 1e-5


def test_pcd_kernels_and_channel_counts(emu):
    z = np.load(os.path.join(GOLD, "reg_golden_pcd.npz"))
    pts, edges = torch.from_numpy(z["pts"]), torch.from_numpy(z["edges"])
    op = R.laplacian(pts, edges)
    w = np.full(300, 1.0 / 300, np.float32)
    loss, dx = lap_fwd_bwd(emu, op, np.ascontiguousarray(z["pts"]), 3, 3, w, 1)
    assert abs(loss - z["loss_f64"]) <= 1e-5 * z["loss_f64"] and rel(dx, z["grad_f64"]) <= 1e-4
    # zero field: loss 0, gradient 0 (not NaN)
    loss, dx = lap_fwd_bwd(emu, op, np.zeros((300, 3), np.float32), 3, 3, w, 1)
    assert loss == 0.0 and np.abs(dx).max() == 0.0
    # C = 1, 2, 4 against a dense float64 evaluation
    D = op.to_dense().double().numpy()
    rng = np.random.default_rng(0)
    for C_ in (1, 2, 4):
        x = rng.standard_normal((300, C_)).astype(np.float32)
        for mode in (0, 1):
            loss, dx = lap_fwd_bwd(emu, op, x, C_, C_, w, mode)
            y = D @ x.astype(np.float64)
            if mode == 0:
                ref, gy = (w * (y ** 2).sum(1)).sum(), 2 * w[:, None] * y
            else:
                nrm = np.linalg.norm(y, axis=1)
                ref, gy = (w * nrm).sum(), w[:, None] * y / nrm[:, None]
            assert abs(loss - ref) <= 1e-5 * ref and rel(dx, D.T @ gy) <= 1e-5


@pytest.mark.parametrize("name", ["l2_a", "l2_b", "l2_c"])
def test_l2norm_kernels(emu, name):
    z = np.load(os.path.join(GOLD, f"reg_golden_{name}.npz"))
    c = z["cfg"]        # lambda_xyz_offsets, lambda_scales_diff, lambda_max_scale, max_scale_threshold, lambda_min_opacity, min_opacity_threshold
    off, sc = np.ascontiguousarray(z["xyz_offsets"]), np.ascontiguousarray(z["scales"])
    op = np.ascontiguousarray(z["opacity"]).reshape(-1) if bool(z["has_opacity"]) else None
    N = off.shape[0]
    sums = np.full(9, np.nan)
    loss = np.full(1, np.nan, np.float32)
    assert emu.emu_l2norm_fwd(N, p(off), p(sc), 3, p(op), c[3], c[5], c[0], c[1], c[2], c[4], p(sums), p(loss)) == 0
    assert abs(float(loss[0]) - z["loss_f64"]) <= 1e-5 * z["loss_f64"]
    d_off, d_sc = np.full((N, 3), np.nan, np.float32), np.full((N, 3), np.nan, np.float32)
    d_op = np.full(N, np.nan, np.float32) if op is not None else None
    dl = np.asarray([2.0], np.float32)
    assert emu.emu_l2norm_bwd(N, p(off), p(sc), 3, 3, p(op), c[3], c[5], p(sums), c[0], c[1], c[2], c[4], p(dl), p(d_off),
                              p(d_sc), p(d_op)) == 0
    assert rel(d_off, 2 * z["grad_off_f64"]) <= 1e-5 and rel(d_sc, 2 * z["grad_scales_f64"]) <= 1e-5
    if op is not None:
        ref = 2 * z["grad_opacity_f64"].reshape(-1)
        assert (np.abs(d_op - ref).max() <= 1e-5 * np.abs(ref).max()) if np.abs(ref).max() > 0 else np.abs(d_op).max() == 0.0
