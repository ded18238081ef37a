"""CPU restatement of the reference's multi-scale tri-plane interpolation (TEST INFRASTRUCTURE: only
tests/ and bench.py's comparison legs may import it).

Follows /root/reference/sings/rec/models/modules/hexplane.py
  :44-68    grid_sample_wrapper     :70-105  interpolate_ms_features     :165-166  normalize_aabb
PINNED: tests/golden/hexplane_golden_*.npz hold outputs and gradients of the reference's own
HexPlaneField (tests/golden/make_hexplane_golden.py); tests/test_hexplane_oracle.py checks this file
against them.
"""
import itertools

import torch
import torch.nn.functional as F


def grid_sample_wrapper(grid, coords, align_corners=True):
    grid_dim = coords.shape[-1]
    if grid.dim() == grid_dim + 1:
        grid = grid.unsqueeze(0)
    if coords.dim() == 2:
        coords = coords.unsqueeze(0)
    coords = coords.view([coords.shape[0]] + [1] * (grid_dim - 1) + list(coords.shape[1:]))
    B, feature_dim = grid.shape[:2]
    n = coords.shape[-2]
    interp = F.grid_sample(grid, coords, align_corners=align_corners, mode="bilinear", padding_mode="border")
    return interp.view(B, feature_dim, n).transpose(-1, -2).squeeze()


def interpolate_ms_features(pts, ms_grids):
    coo_combs = list(itertools.combinations(range(pts.shape[-1]), 2))
    out = []
    for grid in ms_grids:
        interp_space = 1.0
        for ci, comb in enumerate(coo_combs):
            feature_dim = grid[ci].shape[1]
            interp_space = interp_space * grid_sample_wrapper(grid[ci], pts[..., comb]).view(-1, feature_dim)
        out.append(interp_space)
    return torch.cat(out, dim=-1)


def hexplane_features(pts, aabb, ms_grids):
    """aabb (2, 3) as HexPlaneField keeps it; ms_grids: list over scales of the three (1, C, H, W) planes."""
    pn = (pts - aabb[0]) * (2.0 / (aabb[1] - aabb[0])) - 1.0
    return interpolate_ms_features(pn.reshape(-1, pn.shape[-1]), ms_grids)
