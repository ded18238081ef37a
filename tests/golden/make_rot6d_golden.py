"""Golden vectors for the 6D-rotation conversions, from the REFERENCE's own code on CPU.

Run in the build container only (needs /root/reference):
    python tests/golden/make_rot6d_golden.py
Writes tests/golden/rot6d_golden_{f32,f64}.npz.  Reference functions executed:
  rotation_6d_to_matrix      /root/reference/sings/rec/utils/geometry/rotations.py:545-566
  rotation_6d_to_axis_angle  /root/reference/sings/rec/utils/geometry/rotations.py:601-603
     (= matrix_to_quaternion :98-149 + quaternion_to_axis_angle :514-542)
as SinGS.forward uses them (sings_hybrid.py:354-356 canonical rotations, :370-376 pose).
Values and autograd gradients for random upstream gradients.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_lbs_golden import load_reference  # noqa: E402


def inputs(dt):
    g = torch.Generator().manual_seed(61)
    d6 = torch.randn(96, 6, generator=g, dtype=torch.float64)
    d6[:8] *= 5.0                                   # far from unit length
    d6[8:16] *= 0.05
    eye = torch.tensor([1.0, 0, 0, 0, 1, 0], dtype=torch.float64)
    d6[16:24] = eye + 1e-3 * torch.randn(8, 6, generator=g, dtype=torch.float64)     # small angles
    d6[24] = eye                                                                     # exactly identity
    d6[25:33] = torch.tensor([1.0, 0, 0, 0, -1, 0], dtype=torch.float64) + 1e-2 * torch.randn(8, 6, generator=g, dtype=torch.float64)  # near pi about x
    d6[33:41] = torch.tensor([-1.0, 0, 0, 0, 1, 0], dtype=torch.float64) + 1e-2 * torch.randn(8, 6, generator=g, dtype=torch.float64)  # near pi about y
    d6[41:49] = torch.tensor([-1.0, 0, 0, 0, -1, 0], dtype=torch.float64) + 1e-2 * torch.randn(8, 6, generator=g, dtype=torch.float64) # near pi about z
    gR = torch.randn(96, 3, 3, generator=g, dtype=torch.float64)
    gaa = torch.randn(96, 3, generator=g, dtype=torch.float64)
    return d6.to(dt), gR.to(dt), gaa.to(dt)


def main():
    _, _, rot = load_reference()
    for dt, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
        d6, gR, gaa = inputs(dt)
        d6.requires_grad_(True)
        R = rot.rotation_6d_to_matrix(d6)
        aa = rot.rotation_6d_to_axis_angle(d6)
        d_R = torch.autograd.grad((R * gR).sum(), d6, retain_graph=True)[0]
        d_aa = torch.autograd.grad((aa * gaa).sum(), d6)[0]
        out = dict(d6=d6, gR=gR, gaa=gaa, R=R, aa=aa, d_d6_from_R=d_R, d_d6_from_aa=d_aa)
        np.savez_compressed(os.path.join(HERE, f"rot6d_golden_{tag}.npz"),
                            **{k: v.detach().numpy() for k, v in out.items()})
        print("wrote rot6d", tag, "max |aa| =", float(aa.abs().max()))


if __name__ == "__main__":
    main()
