"""pose -> A for SMPL / SMPL-H WITHOUT the template forward: the part of the body-model call the
hot path consumes.

The reference runs the whole body model on its 110k-vertex subdivided template every frame
(`self.smpl_template(...)`, /root/reference/sings/rec/models/sings_hybrid.py:390-396 ->
SMPLH.forward, modules/smplh_layer.py:268-367 -> lbs(), utils/body_model/lbs.py:77-188, including
a 607 MB zero pose-blend-shape matmul) and keeps only `.A` and `.full_pose`.  `.A` depends on
  * the rest joints  J = J_regressor (v_template + shapedirs beta)      (lbs.py:130-135) -- a function
    of `betas` alone, which are fixed during training (`optim_betas: false`, human_complex.yaml:47):
    computed once per distinct betas and cached;
  * the full pose    cat(global_orient, body_pose, left_hand, right_hand) + pose_mean, the hands
    expanded from their PCA coefficients when use_pca (smplh_layer.py:307-317);
  * the kinematic chain  batch_rodrigues + batch_rigid_transform (lbs.py:159-171), then
    `A_t2pose @ inv_A_t2cano` (sings_hybrid.py:399) and, if a translation is applied inside the body
    model, `A[..., :3, 3] += transl` (smplh_layer.py:342-349 / smpl_layer.py:579-585)
-- the last item is one CUDA kernel per call (sgs_pose_to_A, differentiable in the pose).
SMPL assets need registration and are not available offline: the class takes the model tensors
from whoever loaded them (names as in the smplx pickles).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import deform


class BodyPoseToA:
    def __init__(self, parents: torch.Tensor, J_regressor: Optional[torch.Tensor] = None,
                 v_template: Optional[torch.Tensor] = None, shapedirs: Optional[torch.Tensor] = None,
                 pose_mean: Optional[torch.Tensor] = None, left_hand_components: Optional[torch.Tensor] = None,
                 right_hand_components: Optional[torch.Tensor] = None, use_pca: bool = True,
                 inv_A_t2cano: Optional[torch.Tensor] = None, rest_joints: Optional[torch.Tensor] = None):
        """parents (J,) int; J_regressor (J,V); v_template (V,3); shapedirs (V,3,num_betas);
        pose_mean (J*3,) or None (SMPL); hand components (num_pca,45) each (SMPL-H with use_pca);
        inv_A_t2cano (J,4,4) or None.  `rest_joints` (J,3) may be given instead of the three
        regression tensors when the joints are already known."""
        self.parents = parents.to(torch.int32)
        self.J = int(parents.numel())
        self.J_regressor, self.v_template, self.shapedirs = J_regressor, v_template, shapedirs
        self.pose_mean = pose_mean
        self.lh, self.rh, self.use_pca = left_hand_components, right_hand_components, use_pca
        self.inv_A = inv_A_t2cano
        self._rest_fixed = rest_joints
        self._rest_key, self._rest = None, None
        # (SMPL-H called without hand poses uses the module's own left/right_hand_pose parameters,
        # zero-initialised PCA coefficients, smplh_layer.py:196-209,296-299: full_pose() does the same)

    # ---- rest joints: cached per distinct betas --------------------------------------------------
    def rest_joints(self, betas: Optional[torch.Tensor]) -> torch.Tensor:
        if self._rest_fixed is not None:
            return self._rest_fixed
        b = betas.reshape(1, -1) if betas.dim() == 1 else betas[:1]
        key = (b.data_ptr(), b._version, tuple(b.shape))
        if key != self._rest_key:
            nb = b.shape[1]
            # blend_shapes + vertices2joints (lbs.py:130-135; einsum forms of smpl.py:370-412)
            v_shaped = self.v_template + torch.einsum("bl,mkl->bmk", b.to(self.v_template.dtype), self.shapedirs[..., :nb])[0]
            self._rest = torch.einsum("ik,ji->jk", v_shaped, self.J_regressor).contiguous()
            self._rest_key = key
        return self._rest

    # ---- full pose -------------------------------------------------------------------------------
    def full_pose(self, global_orient: torch.Tensor, body_pose: torch.Tensor,
                  left_hand_pose: Optional[torch.Tensor] = None, right_hand_pose: Optional[torch.Tensor] = None) -> torch.Tensor:
        """(B,3), (B,63 | 69), optional hands -> (B, J*3), smplh_layer.py:293-317 (SMPL: smpl_layer.py:547-549)."""
        B = body_pose.shape[0]
        parts = [global_orient.reshape(B, -1), body_pose.reshape(B, -1)]
        n_hand = self.J * 3 - sum(p.shape[1] for p in parts)
        if n_hand != 0 and n_hand != 90:
            raise ValueError(f"global_orient + body_pose have {self.J * 3 - n_hand} values; a skeleton of {self.J} joints "
                             f"needs {self.J * 3} (SMPL: 3 + 69) or 90 fewer (SMPL-H: 3 + 63 + two hands)")
        if n_hand > 0:                                  # SMPL-H: two hands of 15 joints
            def hand(h, comp):
                if h is None:
                    h = body_pose.new_zeros(B, comp.shape[0] if (self.use_pca and comp is not None) else n_hand // 2)
                if self.use_pca and comp is not None:
                    h = torch.einsum("bi,ij->bj", h, comp.to(h.dtype))
                return h
            parts += [hand(left_hand_pose, self.lh), hand(right_hand_pose, self.rh)]
        fp = torch.cat(parts, dim=1)
        if fp.shape[1] != self.J * 3:
            raise ValueError(f"pose has {fp.shape[1]} values, the skeleton needs {self.J * 3}")
        if self.pose_mean is not None:
            fp = fp + self.pose_mean.to(fp.dtype)
        return fp

    # ---- the call --------------------------------------------------------------------------------
    def __call__(self, betas, global_orient, body_pose, left_hand_pose=None, right_hand_pose=None,
                 transl: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """-> (A_cano2pose (B,J,4,4), full_pose (B,J*3)); A is differentiable w.r.t. the poses."""
        fp = self.full_pose(global_orient, body_pose, left_hand_pose, right_hand_pose)
        B = fp.shape[0]
        rest = self.rest_joints(betas)
        A = deform.pose_to_A(fp.reshape(B, self.J, 3), rest, self.parents, self.inv_A)
        if transl is not None:
            # the body model adds it to A_t2pose[..., :3, 3] (smplh_layer.py:342-349) BEFORE `@ inv_A_t2cano`;
            # the last row of inv_A_t2cano is (0,0,0,1), so adding it to the product's column 3 is the same
            A = A.clone()
            A[..., :3, 3] = A[..., :3, 3] + transl.reshape(B, 1, 3)
        return A, fp
