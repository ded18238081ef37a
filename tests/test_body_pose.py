"""SMPL-H pose assembly, rest-joint cache and pose -> A glue (sings_b200/body.py) against golden
vectors produced by the reference's own lbs() and the statements of SMPLH.forward
(tests/golden/make_smplh_golden.py)."""
import os

import numpy as np
import pytest
import torch

from sings_b200.body import BodyPoseToA

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "smplh_golden.npz"))


def _body(dev, dtype):
    t = lambda k: torch.tensor(G[k], device=dev, dtype=dtype)
    return BodyPoseToA(torch.from_numpy(G["parents"]), t("J_regressor"), t("v_template"), t("shapedirs"), t("pose_mean"),
                       t("lh_comp"), t("rh_comp"), use_pca=True, inv_A_t2cano=t("inv_A")), t


def test_full_pose_and_rest_joints_match_the_reference_cpu():
    body, t = _body("cpu", torch.float64)
    fp = body.full_pose(t("global_orient"), t("body_pose"))
    assert np.abs(fp.numpy() - G["default_hands_full_pose"]).max() < 1e-14
    fp2 = body.full_pose(t("global_orient"), t("body_pose"), t("left_hand_pose"), t("right_hand_pose"))
    assert np.abs(fp2.numpy() - G["given_hands_full_pose"]).max() < 1e-14
    betas = t("betas")
    r1 = body.rest_joints(betas)
    assert np.abs(r1.numpy() - G["rest_joints"]).max() < 1e-13
    assert body.rest_joints(betas) is r1                       # cached: same betas, no recomputation
    betas.mul_(1.5)                                            # in-place update bumps the version: recomputed
    assert body.rest_joints(betas) is not r1
    with pytest.raises(ValueError):
        BodyPoseToA(torch.from_numpy(G["parents"][:24]), rest_joints=torch.zeros(24, 3)).full_pose(
            torch.zeros(1, 3), torch.zeros(1, 63))             # 63 values are not an SMPL body pose (69)


@pytest.mark.gpu
@pytest.mark.parametrize("hands", ["default_hands", "given_hands"])
def test_A_matches_the_reference_gpu(hands):
    body, t = _body("cuda", torch.float32)
    lh = t("left_hand_pose") if hands == "given_hands" else None
    rh = t("right_hand_pose") if hands == "given_hands" else None
    go, bp = t("global_orient").requires_grad_(True), t("body_pose").requires_grad_(True)
    A, fp = body(t("betas"), go, bp, lh, rh, transl=t("transl"))
    assert np.abs(A.detach().cpu().numpy() - G[f"{hands}_A"]).max() < 2e-5
    assert np.abs(fp.detach().cpu().numpy() - G[f"{hands}_full_pose"]).max() < 1e-6
    A.square().sum().backward()                                # differentiable through the CUDA chain into the poses
    assert go.grad is not None and bp.grad is not None and float(bp.grad.abs().max()) > 0
