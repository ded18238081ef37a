"""Synthetic avatars, poses and cameras for tests and benchmarks.

SMPL / SMPL-H assets and AMASS need registration and are not available offline, so the
benchmark avatar is synthesised: a fixed humanoid rest skeleton with SMPL's joint order and
kinematic tree (/root/reference/sings/rec/models/modules/smpl_layer.py:272), Gaussians
sampled on capsules around the bones, SMPL-shaped skinning weights (rows sum to 1, <= 4
non-zeros), random axis-angle poses, and the SinGS pinhole camera
(/root/reference/sings/rec/datasets/AnimDataset_opt.py:70-102: identity extrinsic,
fx = fy = 5000 px at 896 px height, znear 0.01, zfar 100; projection from
/root/reference/sings/rec/utils/graphics.py:65-85).

Everything here is numpy and deterministic in its seeds.  Nothing here is timed.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

# kinematic tree of SMPL (smpl_layer.py:272)
SMPL_PARENTS = np.array([-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18,
                         19, 20, 21], dtype=np.int32)

# rest joints of a 1.75 m humanoid, y up, pelvis at the origin, T-pose (metres)
_REST24 = np.array([
    [0.00, 0.00, 0.00], [0.07, -0.09, 0.00], [-0.07, -0.09, 0.00], [0.00, 0.11, 0.00],
    [0.10, -0.47, 0.00], [-0.10, -0.47, 0.00], [0.00, 0.25, 0.00], [0.09, -0.87, -0.03],
    [-0.09, -0.87, -0.03], [0.00, 0.30, 0.00], [0.11, -0.93, 0.09], [-0.11, -0.93, 0.09],
    [0.00, 0.51, -0.02], [0.07, 0.42, 0.00], [-0.07, 0.42, 0.00], [0.00, 0.60, 0.02],
    [0.19, 0.45, 0.00], [-0.19, 0.45, 0.00], [0.45, 0.45, 0.00], [-0.45, 0.45, 0.00],
    [0.70, 0.45, 0.00], [-0.70, 0.45, 0.00], [0.78, 0.45, 0.00], [-0.78, 0.45, 0.00],
], dtype=np.float64)
# capsule radius of the body part driven by each joint
_RADIUS24 = np.array([0.12, 0.08, 0.08, 0.12, 0.06, 0.06, 0.12, 0.045, 0.045, 0.12, 0.04, 0.04,
                      0.05, 0.06, 0.06, 0.10, 0.05, 0.05, 0.04, 0.04, 0.035, 0.035, 0.03, 0.03])


def skeleton(J: int = 24):
    """(rest_joints (J,3) float64, parents (J,) int32, radius (J,)) for J in {24, 52}.

    J=52 follows SMPL-H's order: 0-21 body, 22-36 left hand, 37-51 right hand
    (/root/reference/sings/rec/models/modules/smplh_layer.py:79-81): each hand is five
    3-joint finger chains hanging off the wrist.
    """
    if J == 24:
        return _REST24.copy(), SMPL_PARENTS.copy(), _RADIUS24.copy()
    if J != 52:
        raise ValueError("synthetic skeleton supports J=24 (SMPL) or J=52 (SMPL-H)")
    rest = [r for r in _REST24[:22]]
    parents = list(SMPL_PARENTS[:22])
    radius = list(_RADIUS24[:22])
    for side, wrist in ((1.0, 20), (-1.0, 21)):
        for f in range(5):
            base = np.array([side * 0.76, 0.45, -0.03 + 0.015 * f])
            prev = wrist
            for k in range(3):
                rest.append(base + np.array([side * 0.025 * (k + 1), -0.005 * f, 0.0]))
                parents.append(prev)
                radius.append(0.009)
                prev = len(rest) - 1
    return np.array(rest), np.array(parents, np.int32), np.array(radius)


def rodrigues_np(rvec: np.ndarray) -> np.ndarray:
    """(J,3) axis-angle -> (J,3,3); same formula as smplx batch_rodrigues
    (/root/reference/sings/rec/utils/body_model/smpl.py:415-446)."""
    rvec = np.asarray(rvec, np.float64).reshape(-1, 3)
    angle = np.linalg.norm(rvec + 1e-8, axis=1, keepdims=True)
    d = rvec / angle
    K = np.zeros((rvec.shape[0], 3, 3))
    K[:, 0, 1], K[:, 0, 2] = -d[:, 2], d[:, 1]
    K[:, 1, 0], K[:, 1, 2] = d[:, 2], -d[:, 0]
    K[:, 2, 0], K[:, 2, 1] = -d[:, 1], d[:, 0]
    s, c = np.sin(angle)[:, :, None], np.cos(angle)[:, :, None]
    return np.eye(3)[None] + s * K + (1 - c) * (K @ K)


def pose_to_A_np(pose: np.ndarray, rest: np.ndarray, parents: np.ndarray) -> np.ndarray:
    """Host-side (init-time) pose -> relative joint transforms A (J,4,4), float64; the
    arithmetic of smplx batch_rigid_transform (smpl.py:462-513).  The per-frame version is
    the CUDA kernel sgs_pose_to_A."""
    Rm = rodrigues_np(pose)
    J = rest.shape[0]
    G = np.zeros((J, 4, 4))
    for j in range(J):
        loc = np.eye(4)
        loc[:3, :3] = Rm[j]
        loc[:3, 3] = rest[j] - (rest[parents[j]] if parents[j] >= 0 else 0.0)
        G[j] = loc if parents[j] < 0 else G[parents[j]] @ loc
    A = G.copy()
    A[:, :3, 3] -= np.einsum("jab,jb->ja", G[:, :3, :3], rest)
    return A


def _segments(rest, parents):
    """Segment driven by joint j: from joint j towards the mean of its children (stub for
    leaves)."""
    J = rest.shape[0]
    a = rest.copy()
    b = rest.copy()
    for j in range(J):
        ch = np.where(parents == j)[0]
        if len(ch):
            b[j] = rest[ch].mean(0)
        elif parents[j] >= 0:
            d = rest[j] - rest[parents[j]]
            b[j] = rest[j] + 0.6 * d
        else:
            b[j] = rest[j] + np.array([0, 0.05, 0])
    b[15] = rest[15] + np.array([0.0, 0.14, 0.0]) if J >= 16 else b[15]   # head
    return a, b


@dataclass
class Avatar:
    """A synthetic canonical avatar: the per-frame hot path's resident inputs."""

    J: int
    rest: np.ndarray          # (J,3) float32 rest joints (T-pose)
    parents: np.ndarray       # (J,) int32
    inv_A_t2cano: np.ndarray  # (J,4,4) float32
    xyz_canon: np.ndarray     # (N,3) float32
    rotmat_canon: np.ndarray  # (N,3,3) float32 (identity when isotropic)
    scales: np.ndarray        # (N,3) float32
    opacity: np.ndarray       # (N,1) float32
    shs: np.ndarray           # (N,16,3) float32
    lbs_weights: np.ndarray   # (N,J) float32
    isotropic: bool

    @property
    def N(self):
        return self.xyz_canon.shape[0]


def _random_rotations(rng, n):
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    r, x, y, z = q.T
    R = np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
                  2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
                  2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], 1)
    return R.reshape(n, 3, 3)


def make_avatar(N: int, J: int = 24, seed: int = 0, isotropic: bool = False,
                smooth_weights: int = 0, scale_range=(0.002, 0.012)) -> Avatar:
    """Random canonical Gaussians + SMPL-shaped skinning weights."""
    rng = np.random.default_rng(seed)
    rest, parents, radius = skeleton(J)
    a, b = _segments(rest, parents)
    seglen = np.linalg.norm(b - a, axis=1)
    area = radius * (seglen + 2 * radius)
    part = rng.choice(J, size=N, p=area / area.sum())
    t = rng.uniform(-0.15, 1.15, size=N).clip(0, 1)
    centre = a[part] + (b[part] - a[part]) * t[:, None]
    nrm = rng.normal(size=(N, 3))
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    pts = centre + nrm * radius[part][:, None] + rng.normal(scale=0.005, size=(N, 3))

    # skinning weights: 4 nearest driven segments, w ~ exp(-d^2 / 2 sigma^2)
    wgt = np.zeros((N, J), np.float64)
    rng_w = np.random.default_rng(seed + 1)
    chunk = 65536
    for s in range(0, N, chunk):
        p = pts[s:s + chunk]
        ab = (b - a)[None]
        ap = p[:, None, :] - a[None]
        tt = ((ap * ab).sum(-1) / np.maximum((ab * ab).sum(-1), 1e-12)).clip(0, 1)
        d = np.linalg.norm(ap - tt[..., None] * ab, axis=-1)
        idx = np.argpartition(d, 3, axis=1)[:, :4]
        dd = np.take_along_axis(d, idx, 1)
        dd = dd - dd.min(1, keepdims=True)
        w = np.exp(-dd * dd / (2 * 0.06 ** 2)) + 1e-4 * rng_w.uniform(size=dd.shape)
        w /= w.sum(1, keepdims=True)
        np.put_along_axis(wgt[s:s + chunk], idx, w, 1)
    for _ in range(smooth_weights):   # mimic midpoint subdivision (geometry_ops.py:65-73)
        perm = rng_w.permutation(N)
        wgt = 0.5 * (wgt + wgt[perm] * (part == part[perm])[:, None]
                     + wgt * (part != part[perm])[:, None])
        wgt /= wgt.sum(1, keepdims=True)

    # canonical ("da") pose: legs apart; Gaussians live in canonical space
    cano_pose = np.zeros((J, 3))
    cano_pose[1] = [0, 0, 0.35]
    cano_pose[2] = [0, 0, -0.35]
    A_t2cano = pose_to_A_np(cano_pose, rest, parents)
    T = np.einsum("nj,jab->nab", wgt, A_t2cano)
    xyz_canon = np.einsum("nab,nb->na", T[:, :3, :3], pts) + T[:, :3, 3]

    lo, hi = math.log(scale_range[0]), math.log(scale_range[1])
    if isotropic:
        scales = np.repeat(np.exp(rng.uniform(lo, hi, size=(N, 1))), 3, 1)
        rot = np.repeat(np.eye(3)[None], N, 0)
    else:
        scales = np.exp(rng.uniform(lo, hi, size=(N, 3)))
        rot = _random_rotations(rng, N)
    opacity = 1.0 / (1.0 + np.exp(-rng.normal(1.5, 1.0, size=(N, 1))))
    shs = rng.normal(scale=0.05, size=(N, 16, 3))
    shs[:, 0, :] = (rng.uniform(size=(N, 3)) - 0.5) / 0.28209479177387814
    f32 = np.float32
    return Avatar(J=J, rest=rest.astype(f32), parents=parents,
                  inv_A_t2cano=np.linalg.inv(A_t2cano).astype(f32),
                  xyz_canon=xyz_canon.astype(f32), rotmat_canon=rot.astype(f32),
                  scales=scales.astype(f32), opacity=opacity.astype(f32), shs=shs.astype(f32),
                  lbs_weights=wgt.astype(f32), isotropic=isotropic)


def random_pose(J: int = 24, seed: int = 2, sigma: float = 0.35, neutral: bool = False):
    """(J,3) float32 axis-angle.  Joint 0 is the global orientation R_x(pi) R_y(yaw) -- the
    dataset convention (/root/reference/sings/rec/datasets/motion_utils.py:38-43)."""
    rng = np.random.default_rng(seed)
    pose = np.zeros((J, 3)) if neutral else rng.normal(scale=sigma, size=(J, 3))
    if J == 52 and not neutral:
        pose[22:] *= 0.5
    yaw = 0.0 if neutral else rng.uniform(0, 2 * math.pi)
    Rx = np.array([[1, 0, 0], [0, -1, 0], [0, 0, -1.0]])
    Ry = np.array([[math.cos(yaw), 0, math.sin(yaw)], [0, 1, 0], [-math.sin(yaw), 0, math.cos(yaw)]])
    pose[0] = rotmat_to_axis_angle(Rx @ Ry)
    return pose.astype(np.float32)


def rotmat_to_axis_angle(R: np.ndarray) -> np.ndarray:
    """Robust 3x3 -> axis-angle (handles angle = pi)."""
    R = np.asarray(R, np.float64)
    c = np.clip((np.trace(R) - 1) / 2, -1, 1)
    ang = math.acos(c)
    if ang < 1e-8:
        return np.zeros(3)
    if math.pi - ang < 1e-6:
        w, v = np.linalg.eigh((R + R.T) / 2)
        ax = v[:, np.argmax(w)]
        return ax * ang
    ax = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]]) / (2 * math.sin(ang))
    return ax * ang


def pose_sequence(frames: int, J: int = 24, seed: int = 4, sigma: float = 0.35):
    """AMASS-shaped sequence: smoothed random walk, (F,J,3) float32, plus transl (F,3)."""
    rng = np.random.default_rng(seed)
    base = random_pose(J, seed).astype(np.float64)
    steps = rng.normal(scale=0.04, size=(frames, J, 3))
    walk = np.cumsum(steps, 0)
    k = np.ones(9) / 9
    for j in range(J):
        for c in range(3):
            walk[:, j, c] = np.convolve(walk[:, j, c], k, mode="same")
    walk[:, 0, :] *= 0.1
    poses = base[None] + walk.clip(-3 * sigma, 3 * sigma)
    transl = np.zeros((frames, 3))
    transl[:, 0] = 0.15 * np.sin(np.linspace(0, 2 * math.pi, frames))
    return poses.astype(np.float32), transl.astype(np.float32)


def projection_matrix(znear, zfar, fovx, fovy) -> np.ndarray:
    """graphics.py:65-85 get_projection_matrix (z_sign = +1)."""
    ty, tx = math.tan(fovy / 2), math.tan(fovx / 2)
    top, right = ty * znear, tx * znear
    Pm = np.zeros((4, 4))
    Pm[0, 0] = 2 * znear / (2 * right)
    Pm[1, 1] = 2 * znear / (2 * top)
    Pm[3, 2] = 1.0
    Pm[2, 2] = zfar / (zfar - znear)
    Pm[2, 3] = -(zfar * znear) / (zfar - znear)
    return Pm


@dataclass
class View:
    """The camera half of the reference's per-frame data dict."""

    image_height: int
    image_width: int
    fovx: float
    fovy: float
    world_view_transform: np.ndarray   # (4,4) float32 = W2C^T
    full_proj_transform: np.ndarray    # (4,4) float32 = W2C^T P^T
    camera_center: np.ndarray          # (3,) float32

    @property
    def tanfovx(self):
        return math.tan(self.fovx * 0.5)

    @property
    def tanfovy(self):
        return math.tan(self.fovy * 0.5)


def make_view(H: int, W: int, yaw: float = 0.0, focal: float | None = None, znear=0.01,
              zfar=100.0, centre=(0.0, 0.0, 0.0)) -> View:
    """SinGS pinhole (AnimDataset_opt.py:70-102: identity extrinsic when yaw == 0) with an
    optional orbit by `yaw` about `centre` (turn-around views in the spirit of
    datasets/utils.py:60-120): p_cam = R_y(yaw) (p - centre) + centre."""
    if focal is None:
        focal = 5000.0 * (H / 896.0)
    fovx = 2 * math.atan(W / (2 * focal))
    fovy = 2 * math.atan(H / (2 * focal))
    c, s = math.cos(yaw), math.sin(yaw)
    Ry = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])
    centre = np.asarray(centre, np.float64)
    W2C = np.eye(4)
    W2C[:3, :3] = Ry
    W2C[:3, 3] = centre - Ry @ centre
    wvt = W2C.T
    Pm = projection_matrix(znear, zfar, fovx, fovy).T
    full = wvt @ Pm
    cam_center = np.linalg.inv(wvt)[3, :3]
    f32 = np.float32
    return View(H, W, fovx, fovy, wvt.astype(f32), full.astype(f32), cam_center.astype(f32))


def default_transl(H: int, focal: float | None = None, fill: float = 0.85, height: float = 1.75):
    """Translation that puts the avatar in front of the SinGS camera so it spans `fill` of the
    image height (the shipped kits sit at z ~ 10.2-11.9 m with fx = 5000 at 512x896)."""
    if focal is None:
        focal = 5000.0 * (H / 896.0)
    z = focal * height / (fill * H)
    # the global orientation flips y (R_x(pi)), so the body centre (~ -0.07 m) maps to +0.07
    return np.array([0.0, -0.07, z], np.float32)
