"""Shared builders for the parity tests: synthetic scenes for the oracle and the CUDA path."""
from __future__ import annotations

import numpy as np
import torch

from oracle import lbs_oracle as lo
from oracle import raster_oracle as ro
from sings_b200 import synthetic as syn


def make_scene(N=2000, H=128, W=160, J=24, seed=0, scale_range=(0.004, 0.03), isotropic=False,
               yaw=0.0, fill=0.85):
    """A posed synthetic avatar in front of the SinGS camera: returns a dict of numpy float32
    arrays at the rasterizer boundary (deformed by the CPU oracle) plus the camera."""
    av = syn.make_avatar(N, J, seed=seed, scale_range=scale_range, isotropic=isotropic)
    pose = torch.from_numpy(syn.random_pose(J, seed=seed + 2))[None]
    A = lo.pose_to_A(pose, torch.from_numpy(av.rest), av.parents, torch.from_numpy(av.inv_A_t2cano))
    focal = 5000.0 * (H / 896.0)
    transl = torch.from_numpy(syn.default_transl(H, focal=focal, fill=fill))[None]
    xyz, q, sc, _ = lo.deform(A, torch.from_numpy(av.xyz_canon), torch.from_numpy(av.lbs_weights),
                              torch.from_numpy(av.scales),
                              None if isotropic else torch.from_numpy(av.rotmat_canon), None, transl)
    view = syn.make_view(H, W, yaw=yaw, centre=(0.0, 0.0, float(transl[0, 2])))
    return dict(avatar=av, view=view, A=A[0].numpy(), pose=pose[0].numpy(), transl=transl[0].numpy(),
                means3D=np.ascontiguousarray(xyz[0].numpy()), rotations=np.ascontiguousarray(q[0].numpy()),
                scales=np.ascontiguousarray(sc[0].numpy()), opacity=av.opacity, shs=av.shs)


def oracle_camera(view: syn.View) -> ro.Camera:
    return ro.Camera(W=view.image_width, H=view.image_height, tanfovx=view.tanfovx,
                     tanfovy=view.tanfovy, view=view.world_view_transform.reshape(-1),
                     proj=view.full_proj_transform.reshape(-1), campos=view.camera_center)


def raster_settings(view: syn.View, bg, sh_degree, device="cuda", scale_modifier=1.0, debug=False):
    from diff_gaussian_rasterization import GaussianRasterizationSettings
    t = lambda a: torch.as_tensor(np.asarray(a, np.float32), device=device)
    return GaussianRasterizationSettings(
        image_height=view.image_height, image_width=view.image_width, tanfovx=view.tanfovx,
        tanfovy=view.tanfovy, bg=t(bg), scale_modifier=scale_modifier,
        viewmatrix=t(view.world_view_transform), projmatrix=t(view.full_proj_transform),
        sh_degree=sh_degree, campos=t(view.camera_center), prefiltered=False, debug=debug)


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-20))


def grad_errors(a, b):
    """Three views of how far gradient tensor a is from the reference b:
      max_rel  max|a-b| / max|b|            (the north star's "relative error", dominated by the largest entries)
      l2_rel   ||a-b||_2 / ||b||_2          (per-tensor L2-relative error)
      bad_frac fraction of ELEMENTS outside |a-b| <= 1e-3 |b| + 1e-6 max|b|  -- a small entry that is
               100 % wrong passes the first two; this one catches it."""
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    d = np.abs(a - b)
    mb = np.abs(b).max() + 1e-300
    return dict(max_rel=float(d.max() / mb), l2_rel=float(np.linalg.norm(d) / (np.linalg.norm(b) + 1e-300)),
                bad_frac=float((d > 1e-3 * np.abs(b) + 1e-6 * mb).mean()))


def assert_grad_close(a, b, name="", tol=1e-3, pct=99.9):
    """Gradient parity bar of the tests: tensor-relative and L2-relative error <= tol, and at least
    `pct` % of the elements individually within 1e-3 relative (+ 1e-6 of the tensor's scale)."""
    e = grad_errors(a, b)
    assert e["max_rel"] <= tol, f"grad {name}: max-relative error {e}"
    assert e["l2_rel"] <= tol, f"grad {name}: L2-relative error {e}"
    assert e["bad_frac"] <= 1.0 - pct / 100.0, f"grad {name}: element-wise check {e}"
    return e


def inspect_state(ctx_tensors, P, W, H, L_cap):
    """Pull keys / point list / ranges / final_T / n_contrib out of the scratch buffers a
    forward saved (parity tests only)."""
    from sings_b200.rasterizer import layout_info
    geom, binning, img = ctx_tensors
    info = layout_info(P, W, H, L_cap)
    b = binning.cpu().numpy()
    im = img.cpu().numpy()
    cnt = b[info["counters"]:info["counters"] + 8].view(np.int32)
    L = int(cnt[0])
    out = dict(num_rendered=L, overflow=int(cnt[1]))
    out["keys_unsorted"] = b[info["keys_unsorted"]:info["keys_unsorted"] + 8 * L].view(np.uint64).copy()
    out["vals_unsorted"] = b[info["vals_unsorted"]:info["vals_unsorted"] + 4 * L].view(np.uint32).copy() & np.uint32(0xffffff)
    out["keys"] = b[info["keys_sorted"]:info["keys_sorted"] + 8 * L].view(np.uint64).copy()
    # a list entry = Gaussian id (low 24 bits) | reach mask (high 8 bits, include/sings_b200.h)
    entries = b[info["vals_sorted"]:info["vals_sorted"] + 4 * L].view(np.uint32).copy()
    out["point_list"] = entries & np.uint32(0xffffff)
    out["reach_masks"] = (entries >> np.uint32(24)).astype(np.uint8)
    out["ranges"] = b[info["ranges"]:info["ranges"] + 8 * info["tiles"]].view(np.uint32).reshape(-1, 2).copy()
    out["final_T"] = im[info["final_T"]:info["final_T"] + 4 * W * H].view(np.float32).reshape(H, W).copy()
    out["n_contrib"] = im[info["n_contrib"]:info["n_contrib"] + 4 * W * H].view(np.uint32).reshape(H, W).copy()
    nf = info["rec_floats"]
    rec = geom.cpu().numpy()[:P * nf * 4].view(np.float32).reshape(P, nf).copy()
    out["rec"] = rec
    return out
