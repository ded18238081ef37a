"""Golden vectors for the SMPL-H pose assembly + rest-joint regression + pose -> A glue, from the
REFERENCE's own code on CPU.

Run in the build container only (needs /root/reference):
    python tests/golden/make_smplh_golden.py
Writes tests/golden/smplh_golden.npz.  The reference's `lbs()` (utils/body_model/lbs.py:77-188, with
the in-tree smplx.lbs stand-in of SURVEY.md 8c) is executed on a small synthetic body model (the
real SMPL-H pickle needs registration), wrapped in the statements of SMPLH.forward that the hot path
depends on (modules/smplh_layer.py:293-349: default zero hand PCA coefficients, PCA expansion,
concatenation, `full_pose += pose_mean`, lbs(...), `A[..., :3, 3] += transl`), followed by
`A_t2pose @ inv_A_t2cano` (models/sings_hybrid.py:399).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from make_lbs_golden import load_reference  # noqa: E402
from sings_b200 import synthetic as syn      # noqa: E402


def main():
    smpl, lbs, rot = load_reference()
    g = torch.Generator().manual_seed(91)
    dt = torch.float64
    J, V, NB, NPCA, B = 52, 300, 16, 6, 3
    rest0, parents, _ = syn.skeleton(J)
    parents_t = torch.from_numpy(np.asarray(parents)).long()
    v_template = torch.randn(V, 3, generator=g, dtype=dt) * 0.4
    shapedirs = torch.randn(V, 3, NB, generator=g, dtype=dt) * 0.02
    Jr = torch.rand(J, V, generator=g, dtype=dt)
    Jr = Jr / Jr.sum(1, keepdim=True)
    posedirs = torch.zeros((J - 1) * 9, V * 3, dtype=dt)          # disabled pose blend shapes (smpl_layer.py:351)
    lbs_w = torch.softmax(torch.randn(V, J, generator=g, dtype=dt), 1)
    pose_mean = torch.cat([torch.zeros(66, dtype=dt), 0.2 * torch.randn(90, generator=g, dtype=dt)])
    lh_comp = torch.randn(NPCA, 45, generator=g, dtype=dt) * 0.3
    rh_comp = torch.randn(NPCA, 45, generator=g, dtype=dt) * 0.3
    betas = torch.randn(1, NB, generator=g, dtype=dt) * 0.5
    global_orient = torch.randn(B, 3, generator=g, dtype=dt)
    body_pose = torch.randn(B, 63, generator=g, dtype=dt) * 0.3
    transl = torch.randn(B, 3, generator=g, dtype=dt)
    inv_A = torch.linalg.inv(torch.eye(4, dtype=dt)[None].repeat(J, 1, 1) + 0.0)
    inv_A[:, :3, 3] = 0.05 * torch.randn(J, 3, generator=g, dtype=dt)
    out = {}
    for tag, lhp, rhp in (("default_hands", None, None),
                          ("given_hands", torch.randn(B, NPCA, generator=g, dtype=dt), torch.randn(B, NPCA, generator=g, dtype=dt))):
        # ---- smplh_layer.py:293-317 ----
        left = lhp if lhp is not None else torch.zeros(1, NPCA, dtype=dt).expand(B, -1)
        right = rhp if rhp is not None else torch.zeros(1, NPCA, dtype=dt).expand(B, -1)
        left = torch.einsum('bi,ij->bj', [left, lh_comp])
        right = torch.einsum('bi,ij->bj', [right, rh_comp])
        full_pose = torch.cat([global_orient, body_pose, left, right], dim=1)
        full_pose = full_pose + pose_mean
        # ---- the reference's lbs() ----
        res = lbs.lbs(betas.expand(B, -1), full_pose, v_template, shapedirs, posedirs, Jr, parents_t, lbs_w, pose2rot=True,
                      disable_posedirs=True)
        A = res[2].clone()
        A[..., :3, 3] += transl.unsqueeze(dim=1)                 # smplh_layer.py:342-349
        A_c2p = A @ inv_A.unsqueeze(0)                            # sings_hybrid.py:399
        out[f"{tag}_full_pose"] = full_pose.numpy()
        out[f"{tag}_A"] = A_c2p.numpy()
        if lhp is not None:
            out["left_hand_pose"], out["right_hand_pose"] = lhp.numpy(), rhp.numpy()
    v_shaped = v_template + smpl.blend_shapes(betas, shapedirs)[0]
    out["rest_joints"] = smpl.vertices2joints(Jr, v_shaped[None])[0].numpy()
    out.update(parents=np.asarray(parents, np.int32), v_template=v_template.numpy(), shapedirs=shapedirs.numpy(),
               J_regressor=Jr.numpy(), pose_mean=pose_mean.numpy(), lh_comp=lh_comp.numpy(), rh_comp=rh_comp.numpy(),
               betas=betas.numpy(), global_orient=global_orient.numpy(), body_pose=body_pose.numpy(),
               transl=transl.numpy(), inv_A=inv_A.numpy())
    np.savez_compressed(os.path.join(HERE, "smplh_golden.npz"), **out)
    print("wrote smplh_golden.npz", {k: v.shape for k, v in out.items() if k.endswith("_A") or k == "rest_joints"})


if __name__ == "__main__":
    main()
