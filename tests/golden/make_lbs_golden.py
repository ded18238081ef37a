"""Generate golden vectors for the deformer by running the REFERENCE's own code on CPU.

Run in the build container only (needs /root/reference):
    python tests/golden/make_lbs_golden.py
Writes tests/golden/lbs_golden_<case>.npz.  The reference functions executed are
  lbs_extra                /root/reference/sings/rec/utils/body_model/lbs.py:16-74
  matrix_to_quaternion     /root/reference/sings/rec/utils/geometry/rotations.py:98-149
  quaternion_multiply      /root/reference/sings/rec/utils/geometry/rotations.py:393-407
  batch_rodrigues          /root/reference/sings/rec/utils/body_model/smpl.py:415-446
  batch_rigid_transform    /root/reference/sings/rec/utils/body_model/smpl.py:462-513
composed exactly as SinGS.forward_chunk composes them (sings_hybrid.py:525-552).  lbs.py
imports `smplx.lbs`, which is not installed; the identical functions are vendored in the
reference's smpl.py, so a stand-in module forwards to those.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_reference():
    smpl = _load("ref_smpl", f"{REF}/sings/rec/utils/body_model/smpl.py")
    shim = types.ModuleType("smplx.lbs")
    for n in ("batch_rodrigues", "batch_rigid_transform", "blend_shapes", "vertices2joints"):
        setattr(shim, n, getattr(smpl, n))
    pkg = types.ModuleType("smplx")
    pkg.lbs = shim
    sys.modules["smplx"] = pkg
    sys.modules["smplx.lbs"] = shim
    lbs = _load("ref_lbs", f"{REF}/sings/rec/utils/body_model/lbs.py")
    rot = _load("ref_rot", f"{REF}/sings/rec/utils/geometry/rotations.py")
    return smpl, lbs, rot


def reference_deform(smpl, lbs, rot, pose, rest, parents, inv_A, xyz, W, scales, Rc, smpl_scale,
                     transl, ext):
    """sings_hybrid.py:525-552 with the body model replaced by its pose->A chain
    (batch_rodrigues + batch_rigid_transform, smpl.py / lbs.py:126-171)."""
    B, J = pose.shape[:2]
    Rm = smpl.batch_rodrigues(pose.reshape(-1, 3), dtype=pose.dtype).view(B, J, 3, 3)
    _, A_t2pose = smpl.batch_rigid_transform(Rm, rest[None].expand(B, -1, -1), parents,
                                             dtype=pose.dtype)
    A = A_t2pose @ inv_A.unsqueeze(0)
    xyz_b = xyz.unsqueeze(0).expand(B, -1, -1)
    xyz_d, _, T, _, _ = lbs.lbs_extra(A, xyz_b, posedirs=None, lbs_weights=W,
                                      pose=pose.reshape(B, -1), disable_posedirs=True,
                                      pose2rot=True)
    sc = scales.unsqueeze(0).expand(B, -1, -1)
    if smpl_scale is not None:
        xyz_d = xyz_d * smpl_scale.unsqueeze(-1)
        sc = sc * smpl_scale.unsqueeze(-1)
    if transl is not None:
        xyz_d = xyz_d + transl.unsqueeze(1)
    Rdef = T[..., :3, :3] @ Rc.unsqueeze(0).expand(B, -1, -1, -1)
    q = rot.matrix_to_quaternion(Rdef)
    if ext is not None:
        trans, rotmat, scale = ext
        xyz_d = (trans[:, None, :] + (scale[:, None] * (rotmat[:, None, ...] @ xyz_d[..., None]).squeeze(-1)))
        sc = scale[..., None] * sc
        q = rot.quaternion_multiply(rot.matrix_to_quaternion(rotmat)[:, None, :], q)
    return A, xyz_d, q, sc, T


def main():
    from sings_b200 import synthetic as syn
    smpl, lbs, rot = load_reference()
    cases = [
        dict(name="j24_aniso_ext", N=301, J=24, B=3, iso=False, ext=True, seed=11),
        dict(name="j24_iso", N=257, J=24, B=1, iso=True, ext=False, seed=12),
        dict(name="j52_aniso", N=203, J=52, B=2, iso=False, ext=False, seed=13),
        dict(name="j52_iso_ext_smooth", N=129, J=52, B=2, iso=True, ext=True, seed=14, smooth=2),
    ]
    for c in cases:
        for dt, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
            av = syn.make_avatar(c["N"], c["J"], seed=c["seed"], isotropic=c["iso"],
                                 smooth_weights=c.get("smooth", 0))
            g = torch.Generator().manual_seed(c["seed"])
            B, J = c["B"], c["J"]
            pose = torch.stack([torch.from_numpy(syn.random_pose(J, seed=c["seed"] + b)) for b in range(B)]).to(dt)
            rest = torch.from_numpy(av.rest).to(dt)
            parents = torch.from_numpy(av.parents.astype(np.int64))
            inv_A = torch.from_numpy(av.inv_A_t2cano).to(dt)
            xyz = torch.from_numpy(av.xyz_canon).to(dt)
            W = torch.from_numpy(av.lbs_weights).to(dt)
            scales = torch.from_numpy(av.scales).to(dt)
            Rc = torch.from_numpy(av.rotmat_canon).to(dt)
            smpl_scale = (1.0 + 0.1 * torch.rand(B, 1, generator=g)).to(dt)
            transl = torch.randn(B, 3, generator=g).to(dt)
            ext = None
            if c["ext"]:
                er = smpl.batch_rodrigues(torch.randn(B, 3, generator=g).to(dt), dtype=dt)
                ext = (torch.randn(B, 3, generator=g).to(dt), er,
                       (0.5 + torch.rand(B, 1, generator=g)).to(dt))
            leaves = [pose, xyz, scales, Rc, smpl_scale, transl] + (list(ext) if ext else [])
            for t in leaves:
                t.requires_grad_(True)
            A, xyz_d, q, sc, T = reference_deform(smpl, lbs, rot, pose, rest, parents, inv_A, xyz,
                                                  W, scales, Rc, smpl_scale, transl, ext)
            gx = torch.randn(xyz_d.shape, generator=g).to(dt)
            gq = torch.randn(q.shape, generator=g).to(dt)
            gs = torch.randn(sc.shape, generator=g).to(dt)
            loss = (xyz_d * gx).sum() + (q * gq).sum() + (sc * gs).sum()
            grads = torch.autograd.grad(loss, leaves, allow_unused=True)
            # gradient w.r.t. the joint transforms themselves (what sgs_lbs_bwd emits)
            A_leaf = A.detach().clone().requires_grad_(True)
            xyz2, _, T2, _, _ = lbs.lbs_extra(A_leaf, xyz.detach().unsqueeze(0).expand(B, -1, -1),
                                              posedirs=None, lbs_weights=W, pose=pose.detach().reshape(B, -1),
                                              disable_posedirs=True, pose2rot=True)
            q2 = rot.matrix_to_quaternion(T2[..., :3, :3] @ Rc.detach().unsqueeze(0).expand(B, -1, -1, -1))
            gA = torch.autograd.grad((xyz2 * gx).sum() + (q2 * gq).sum(), A_leaf)[0]
            out = dict(
                pose=pose, rest=rest, parents=parents, inv_A_t2cano=inv_A, xyz_canon=xyz,
                lbs_weights=W, scales=scales, rotmat_canon=Rc, smpl_scale=smpl_scale, transl=transl,
                A_cano2pose=A, xyz=xyz_d, rotq=q, scales_out=sc, T=T, gx=gx, gq=gq, gs=gs,
                d_pose=grads[0], d_xyz_canon=grads[1], d_scales=grads[2], d_rotmat_canon=grads[3],
                d_smpl_scale=grads[4], d_transl=grads[5], dA_plain_lbs=gA,
                isotropic=torch.tensor(int(c["iso"])))
            if ext:
                out.update(ext_trans=ext[0], ext_rotmat=ext[1], ext_scale=ext[2],
                           d_ext_trans=grads[6], d_ext_rotmat=grads[7], d_ext_scale=grads[8])
            np.savez_compressed(os.path.join(HERE, f"lbs_golden_{c['name']}_{tag}.npz"),
                                **{k: v.detach().numpy() for k, v in out.items()})
            print("wrote", c["name"], tag, {k: tuple(v.shape) for k, v in out.items() if k in ("xyz", "rotq", "T")})


if __name__ == "__main__":
    main()
