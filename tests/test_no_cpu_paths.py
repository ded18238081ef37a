"""The modules around the per-frame path have no CPU fallback either: CPU tensors raise SgsError
before any library call (runs without a GPU)."""
import pytest
import torch

from sings_b200._lib import SgsError


def test_image_loss_refuses_cpu_tensors():
    from sings_b200.losses import image_loss
    with pytest.raises(SgsError):
        image_loss(torch.rand(3, 8, 8), torch.rand(3, 8, 8), None, torch.ones(3))


def test_knn_and_edge_loss_refuse_cpu_tensors():
    from sings_b200.losses import GaussiansEdgeLoss, knn_points
    with pytest.raises(SgsError):
        knn_points(torch.rand(10, 3), 3)
    with pytest.raises(SgsError):
        GaussiansEdgeLoss()({"xyz_canon": torch.rand(10, 3), "scales": torch.rand(10, 3)})


def test_triplane_field_mirrors_the_reference_layout_and_refuses_cpu_tensors():
    from sings_b200.triplane import HexPlaneField
    cfg = {"grid_dimensions": 2, "input_coordinate_dim": 3, "output_coordinate_dim": 32, "resolution": [8, 6, 4], "multires": [1, 2]}
    f = HexPlaneField(cfg, bounds=1.5, device="cpu")
    # (like the reference's `nn.Parameter(aabb).to(device)`, aabb is a registered parameter only when no copy is made: on CPU)
    shapes = {k: tuple(v.shape) for k, v in f.state_dict().items() if k != "aabb"}
    # init_grid_param (hexplane.py:30-41): plane (i, j) is (1, C, reso[j], reso[i]); scales multiply the resolution
    assert shapes == {"grids.0.0": (1, 32, 6, 8), "grids.0.1": (1, 32, 4, 8), "grids.0.2": (1, 32, 4, 6),
                      "grids.1.0": (1, 32, 12, 16), "grids.1.1": (1, 32, 8, 16), "grids.1.2": (1, 32, 8, 12)}
    assert f.feat_dim == 64 and all(p.is_contiguous(memory_format=torch.channels_last) for p in f.parameters() if p.dim() == 4)
    assert all(0.1 <= float(p.min()) and float(p.max()) <= 0.5 for gp in f.grids for p in gp)
    with pytest.raises(SgsError):
        f(torch.rand(5, 3))
    with pytest.raises(SgsError):
        HexPlaneField({**cfg, "output_coordinate_dim": 16}, device="cpu")


def test_avatar_renderer_refuses_cpu_and_non_contiguous_parameters():
    from sings_b200.fused import AvatarRenderer
    n, J = 16, 4
    args = dict(rotmat_canon=None, scales=torch.rand(n, 3), opacity=torch.rand(n, 1), shs=torch.rand(n, 16, 3),
                lbs_weights=torch.rand(n, J), rest_joints=torch.rand(J, 3), parents=torch.tensor([-1, 0, 1, 2]),
                inv_A_t2cano=torch.eye(4).repeat(J, 1, 1), H=32, W=32, sh_degree=3)
    with pytest.raises(SgsError):
        AvatarRenderer(torch.rand(n, 3), **args)                               # CPU tensors
    with pytest.raises(SgsError):
        AvatarRenderer(torch.rand(3, n).t(), **args)                           # read in place: must be contiguous
