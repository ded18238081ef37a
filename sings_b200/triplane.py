"""Multi-scale tri-plane features on the device (the first half of the canonical
attribute decode, sings_hybrid.py:252).

Host-side mirror of /root/reference/sings/rec/models/modules/hexplane.py `HexPlaneField`: same
constructor (planeconfig dict with grid_dimensions=2, input_coordinate_dim=3, output_coordinate_dim,
resolution, multires; bounds), same parameters under the same state_dict keys (`grids.<scale>.<plane>`,
shape (1, C, H, W), initialised uniform(0.1, 0.5) like init_grid_param), same forward(pts) ->
(N, len(multires) * C) features.  The nine grid_sample launches, their products, the concatenation
and their autograd graph are one forward and one backward kernel (sgs_hexplane_fwd / _bwd) reading
channel-last copies of the planes.  There is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import itertools
from typing import List, Optional, Sequence

import torch
import torch.nn as nn

from . import _lib
from ._lib import SgsError, raw_stream


def _ptr_array(tensors: Sequence[Optional[torch.Tensor]]):
    arr = (C.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = None if t is None else t.data_ptr()
    return arr


class _Interp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pts, aabb6, res, Cc, *planes):
        if not pts.is_cuda:
            raise SgsError("sings_b200.triplane needs CUDA tensors (there is no CPU path)")
        S = len(res) // 3
        p = pts.detach().reshape(-1, 3).to(torch.float32).contiguous()
        N = p.shape[0]
        # channel-last (H, W, C): a bilinear tap is one contiguous row of C floats.  For a channels-last
        # parameter (HexPlaneField below creates them so) this is a view, not a copy.
        cl = [g.detach()[0].permute(1, 2, 0).contiguous() for g in planes]
        out = torch.empty(N, S * Cc, device=p.device, dtype=torch.float32)
        aabb = (C.c_float * 6)(*aabb6)
        resa = (C.c_int * len(res))(*res)
        with torch.cuda.device(p.device):
            _lib.check(_lib.lib().sgs_hexplane_fwd(N, p.data_ptr(), aabb, S, Cc, resa, _ptr_array(cl), out.data_ptr(),
                                                   raw_stream(p.device)), "sgs_hexplane_fwd")
        ctx.save_for_backward(p, *cl)
        ctx.meta = (aabb6, res, Cc, pts.shape)
        return out

    @staticmethod
    def backward(ctx, d_out):
        p, *cl = ctx.saved_tensors
        aabb6, res, Cc, pts_shape = ctx.meta
        S = len(res) // 3
        need_pts = ctx.needs_input_grad[0]
        need_pl = [ctx.needs_input_grad[4 + i] for i in range(len(cl))]
        d_cl: List[Optional[torch.Tensor]] = [torch.zeros_like(t) if n else None for t, n in zip(cl, need_pl)]
        d_pts = torch.empty_like(p) if need_pts else None
        if need_pts or any(need_pl):
            aabb = (C.c_float * 6)(*aabb6)
            resa = (C.c_int * len(res))(*res)
            g = d_out.to(torch.float32).contiguous()
            with torch.cuda.device(p.device):
                _lib.check(_lib.lib().sgs_hexplane_bwd(p.shape[0], p.data_ptr(), aabb, S, Cc, resa, _ptr_array(cl),
                                                       g.data_ptr(), _ptr_array(d_cl), None if d_pts is None else d_pts.data_ptr(),
                                                       raw_stream(p.device)), "sgs_hexplane_bwd")
        grads = [None if t is None else t.permute(2, 0, 1).unsqueeze(0) for t in d_cl]
        return (None if d_pts is None else d_pts.reshape(pts_shape), None, None, None, *grads)


class HexPlaneField(nn.Module):
    """hexplane.py:108-189 (tri-plane case: grid_dimensions = 2, input_coordinate_dim = 3)."""

    def __init__(self, planeconfig, bounds: float = 1.0, device="cuda"):
        super().__init__()
        if planeconfig["grid_dimensions"] != 2 or planeconfig["input_coordinate_dim"] != 3:
            raise SgsError("sings_b200.triplane implements the tri-plane case (grid_dimensions=2, input_coordinate_dim=3)")
        Cc = int(planeconfig["output_coordinate_dim"])
        if Cc % 32:
            raise SgsError("output_coordinate_dim must be a multiple of 32")
        aabb = torch.tensor([[bounds, bounds, bounds], [-bounds, -bounds, -bounds]], dtype=torch.float32)
        self.aabb = nn.Parameter(aabb, requires_grad=False).to(device)
        self.grid_config = [planeconfig]
        self.multiscale_res_multipliers = list(planeconfig["multires"])
        self.concat_features = True
        self.grids = nn.ModuleList()
        self.feat_dim = 0
        self._res: List[int] = []
        for mult in self.multiscale_res_multipliers:
            reso = [int(r * mult) for r in planeconfig["resolution"][:3]]
            self._res += reso
            gp = nn.ParameterList()
            for comb in itertools.combinations(range(3), 2):          # (0,1), (0,2), (1,2): init_grid_param, hexplane.py:30-41
                # same shape as the reference's parameter, stored channels-last: the kernels then read the
                # parameter itself (its (H, W, C) view is contiguous) and no layout copy is made per step
                t = torch.empty([1, Cc] + [reso[cc] for cc in comb[::-1]], device=device).contiguous(memory_format=torch.channels_last)
                nn.init.uniform_(t, a=0.1, b=0.5)
                gp.append(nn.Parameter(t))
            self.feat_dim += Cc
            self.grids.append(gp)
        self._C = Cc

    @property
    def get_aabb(self):
        return self.aabb[0], self.aabb[1]

    def set_aabb(self, xyz_max, xyz_min):
        self.aabb = nn.Parameter(torch.tensor([xyz_max, xyz_min], dtype=torch.float32), requires_grad=False)

    def forward(self, pts: torch.Tensor, timestamps: Optional[torch.Tensor] = None) -> torch.Tensor:
        if timestamps is not None:
            raise SgsError("the tri-plane field takes no timestamps (input_coordinate_dim = 3)")
        aabb6 = tuple(float(v) for v in self.aabb.detach().reshape(-1).tolist())
        planes = [g for gp in self.grids for g in gp]
        return _Interp.apply(pts, aabb6, tuple(self._res), self._C, *planes)
