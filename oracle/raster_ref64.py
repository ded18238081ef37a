"""float64, autograd-differentiable restatement of the rasterizer forward (gradient truth).

TEST INFRASTRUCTURE ONLY (see oracle/raster_oracle.py).  Small problems only: it loops over
tiles in Python and holds a (256, K) matrix per tile.

The continuous arithmetic of SURVEY.md Appendix A.2/A.4 is re-expressed in torch so that
torch.autograd provides gradients that were NOT hand-derived; the discrete decisions
(culling, tile lists, sort order, per-pair accept/reject, per-pixel termination) are taken
from the fp32 C oracle's FwdState so both compute the same piecewise-smooth function.
Algorithm source: un-vendored diff-gaussian-rasterization, call site
/root/reference/sings/rec/renderer/gs_renderer_single.py:69-95; SH polynomial as in
/root/reference/sings/rec/utils/visualize/spherical_harmonics.py:61-125.
"""
from __future__ import annotations

import numpy as np
import torch

from . import raster_oracle as ro

C0 = 0.28209479177387814
C1 = 0.4886025119029199
C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792,
      0.5462742152960396]
C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154,
      -0.4570457994644658, 1.445305721320277, -0.5900435899266435]


def sh_to_rgb(D, sh, dirs):
    """sh (P,M,3), dirs (P,3) unit -> (P,3) before the +0.5 / clamp."""
    x, y, z = dirs[:, 0:1], dirs[:, 1:2], dirs[:, 2:3]
    res = C0 * sh[:, 0]
    if D > 0:
        res = res - C1 * y * sh[:, 1] + C1 * z * sh[:, 2] - C1 * x * sh[:, 3]
    if D > 1:
        xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
        res = (res + C2[0] * xy * sh[:, 4] + C2[1] * yz * sh[:, 5]
               + C2[2] * (2 * zz - xx - yy) * sh[:, 6] + C2[3] * xz * sh[:, 7]
               + C2[4] * (xx - yy) * sh[:, 8])
    if D > 2:
        res = (res + C3[0] * y * (3 * xx - yy) * sh[:, 9] + C3[1] * xy * z * sh[:, 10]
               + C3[2] * y * (4 * zz - xx - yy) * sh[:, 11]
               + C3[3] * z * (2 * zz - 3 * xx - 3 * yy) * sh[:, 12]
               + C3[4] * x * (4 * zz - xx - yy) * sh[:, 13] + C3[5] * z * (xx - yy) * sh[:, 14]
               + C3[6] * x * (xx - 3 * yy) * sh[:, 15])
    return res


def quat_to_R(q):
    r, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = torch.stack([
        1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
        2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
        2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], dim=-1)
    return R.reshape(-1, 3, 3)


def render(st: ro.FwdState, means3D, opacities, means2D=None, shs=None, colors_precomp=None,
           scales=None, rotations=None, cov3D_precomp=None, scale_grad_has_modifier=True):
    """Differentiable float64 image (3,H,W) for the decisions recorded in `st`.

    means2D (P,3) is the zero-valued gradient carrier of the reference call site
    (gs_renderer_single.py:50-56): it is added to the NDC position so its autograd gradient
    equals upstream's dL_dmean2D (pixel gradient times 0.5*W, 0.5*H).
    """
    inp = st.inputs
    cam: ro.Camera = inp["cam"]
    dt = means3D.dtype
    V = torch.as_tensor(np.asarray(cam.view, np.float64).reshape(4, 4), dtype=dt)   # = W2C^T
    Mx = torch.as_tensor(np.asarray(cam.proj, np.float64).reshape(4, 4), dtype=dt)
    campos = torch.as_tensor(np.asarray(cam.campos, np.float64), dtype=dt)
    bg = torch.as_tensor(np.asarray(inp["bg"], np.float64), dtype=dt)
    W, H = cam.W, cam.H
    mod = float(inp["scale_modifier"])
    P = means3D.shape[0]
    ones = torch.ones(P, 1, dtype=dt)
    ph = torch.cat([means3D, ones], 1)
    pv = (ph @ V)[:, :3]                       # row-vector convention == column-major M*p
    phom = ph @ Mx
    pw = 1.0 / (phom[:, 3:4] + 1e-7)
    ndc = phom[:, :2] * pw
    if means2D is not None:
        ndc = ndc + means2D[:, :2]
    pix = torch.stack([((ndc[:, 0] + 1) * W - 1) * 0.5, ((ndc[:, 1] + 1) * H - 1) * 0.5], 1)

    if cov3D_precomp is not None:
        c = cov3D_precomp
        Sigma = torch.stack([c[:, 0], c[:, 1], c[:, 2], c[:, 1], c[:, 3], c[:, 4], c[:, 2],
                             c[:, 4], c[:, 5]], 1).reshape(P, 3, 3)
    else:
        R = quat_to_R(rotations)
        s = scales * mod if scale_grad_has_modifier else scales * mod
        N = R * s[:, None, :]
        Sigma = N @ N.transpose(1, 2)
    fx, fy = W / (2.0 * cam.tanfovx), H / (2.0 * cam.tanfovy)
    limx, limy = 1.3 * cam.tanfovx, 1.3 * cam.tanfovy
    tz = pv[:, 2]
    txtz, tytz = pv[:, 0] / tz, pv[:, 1] / tz
    # upstream multiplies the clamped ratio by t.z but back-propagates as if t.x were simply
    # masked (x_grad_mul): reproduce that by detaching inside the clamp region
    inx = ((txtz >= -limx) & (txtz <= limx)).to(dt)
    iny = ((tytz >= -limy) & (tytz <= limy)).to(dt)
    tx = inx * pv[:, 0] + (1 - inx) * (txtz.clamp(-limx, limx) * tz).detach()
    ty = iny * pv[:, 1] + (1 - iny) * (tytz.clamp(-limy, limy) * tz).detach()
    zero = torch.zeros_like(tz)
    J = torch.stack([fx / tz, zero, -(fx * tx) / (tz * tz),
                     zero, fy / tz, -(fy * ty) / (tz * tz)], 1).reshape(P, 2, 3)
    Rv = V[:3, :3].t()                         # p_view = Rv p + t
    A = J @ Rv
    cov = A @ Sigma @ A.transpose(1, 2)
    a = cov[:, 0, 0] + 0.3
    b = cov[:, 0, 1]
    cc = cov[:, 1, 1] + 0.3
    det = a * cc - b * b
    conA, conB, conC = cc / det, -b / det, a / det

    if colors_precomp is not None:
        rgb = colors_precomp
    else:
        d = means3D - campos
        d = d / d.norm(dim=1, keepdim=True)
        rgb = sh_to_rgb(st.D, shs, d) + 0.5
        clamped = torch.as_tensor(st.clamped.astype(bool))
        rgb = torch.where(clamped, torch.zeros_like(rgb), rgb)
    op = opacities.reshape(-1)

    gx, gy = cam.grid
    out = torch.zeros(3, H, W, dtype=dt)
    plist = torch.as_tensor(st.point_list.astype(np.int64))
    ly, lx = torch.meshgrid(torch.arange(16), torch.arange(16), indexing="ij")
    for t in range(gx * gy):
        r0, r1 = int(st.ranges[t, 0]), int(st.ranges[t, 1])
        px = (t % gx) * 16 + lx.reshape(-1)
        py = (t // gx) * 16 + ly.reshape(-1)
        inside = (px < W) & (py < H)
        pxc, pyc = px.clamp(max=W - 1), py.clamp(max=H - 1)
        if r1 <= r0:
            col = bg[:, None].expand(3, 256)
        else:
            ids = plist[r0:r1]
            mask = torch.as_tensor(ro.render_mask(st, t).astype(np.float64), dtype=dt)
            dx = pix[ids, 0][None, :] - px[:, None].to(dt)
            dy = pix[ids, 1][None, :] - py[:, None].to(dt)
            power = -0.5 * (conA[ids] * dx * dx + conC[ids] * dy * dy) - conB[ids] * dx * dy
            alpha = torch.clamp(op[ids][None, :] * torch.exp(power), max=0.99) * mask
            # upstream propagates through the 0.99 clamp as if it were inactive
            raw = op[ids][None, :] * torch.exp(power) * mask
            alpha = raw + (alpha - raw).detach()
            one_m = 1.0 - alpha
            Tincl = torch.cumprod(one_m, dim=1)
            Texcl = torch.cat([torch.ones(256, 1, dtype=dt), Tincl[:, :-1]], 1)
            wgt = alpha * Texcl
            col = (wgt @ rgb[ids]).t() + Tincl[:, -1][None, :] * bg[:, None]
        sel = inside.nonzero().reshape(-1)
        out[:, pyc[sel], pxc[sel]] = col[:, sel]
    return out
