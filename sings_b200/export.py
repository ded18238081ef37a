"""On-disk Gaussian formats of the reference: the 3DGS `.ply` the trainer
saves and the `.splat` file the bundled viewer loads.

Host-side mirror of
  save_ply                 /root/reference/sings/rec/utils/visualize/vis.py:22-61
  process_ply_to_splat     /root/reference/playground/display/convert.py:10-50
with the same attribute order, the same value transforms (inverse sigmoid of the opacity, log of
the scales, SH rows flattened channel-major) and the same byte layout.  The reference builds the
PLY through a Python tuple per vertex and the .splat through a Python loop per vertex (tens of
seconds at 200k Gaussians); here both are vectorised numpy and the .splat is written straight
from the model tensors, without the detour through a PLY file.  No plyfile dependency.
"""
from __future__ import annotations

import os
from typing import Dict

import numpy as np
import torch

SH_C0 = 0.28209479177387814


def ply_attributes() -> list:
    """vis.py:22-35: x y z nx ny nz f_dc_0..2 f_rest_0..44 opacity scale_0..2 rot_0..3 (62 floats)."""
    names = ["x", "y", "z", "nx", "ny", "nz"]
    names += [f"f_dc_{i}" for i in range(3)]
    names += [f"f_rest_{i}" for i in range(45)]
    names.append("opacity")
    names += [f"scale_{i}" for i in range(3)]
    names += [f"rot_{i}" for i in range(4)]
    return names


def _np(t) -> np.ndarray:
    return t.detach().to("cpu", torch.float32).numpy() if isinstance(t, torch.Tensor) else np.asarray(t, np.float32)


def ply_table(human_gs_out: Dict[str, torch.Tensor], pose: str = "canonical") -> np.ndarray:
    """The (N, 62) float32 table save_ply writes (vis.py:38-56).  Keys as in the model's output
    dict: xyz_canon | xyz, shs (N, 16, 3), opacity (N, 1) in (0, 1), scales_canon (N, 3) > 0,
    rotq_canon (N, 4)."""
    if pose not in ("canonical", "deformed"):
        raise ValueError("pose must be 'canonical' or 'deformed'")
    xyz = _np(human_gs_out["xyz_canon" if pose == "canonical" else "xyz"])
    shs = human_gs_out["shs"]
    shs = shs.detach().to("cpu", torch.float32) if isinstance(shs, torch.Tensor) else torch.as_tensor(np.asarray(shs, np.float32))
    if shs.shape[1:] != (16, 3):
        raise ValueError("shs must be (N, 16, 3): the reference's PLY layout has 3 + 45 SH attributes")
    f_dc = shs[:, :1].transpose(1, 2).flatten(start_dim=1).contiguous().numpy()
    f_rest = shs[:, 1:].transpose(1, 2).flatten(start_dim=1).contiguous().numpy()
    op = human_gs_out["opacity"]
    op = op.detach().to("cpu", torch.float32) if isinstance(op, torch.Tensor) else torch.as_tensor(np.asarray(op, np.float32))
    opacities = torch.log(op / (1 - op)).numpy().reshape(-1, 1)           # inverse_sigmoid (general.py)
    sc = human_gs_out["scales_canon"]
    sc = sc.detach().to("cpu", torch.float32) if isinstance(sc, torch.Tensor) else torch.as_tensor(np.asarray(sc, np.float32))
    scale = torch.log(sc).numpy()
    rotation = _np(human_gs_out["rotq_canon"])
    return np.concatenate((xyz, np.zeros_like(xyz), f_dc, f_rest, opacities, scale, rotation), axis=1).astype(np.float32)


def save_ply(human_gs_out: Dict[str, torch.Tensor], path: str, pose: str = "canonical", text: bool = True) -> None:
    """vis.py:38-61.  text=True writes the ASCII PLY the reference writes (`PlyData(..., text=True)`);
    text=False the binary little-endian PLY most 3DGS tools exchange (same header otherwise)."""
    d = os.path.dirname(path)
    if d:
        os.makedirs(d, exist_ok=True)
    table = ply_table(human_gs_out, pose)
    names = ply_attributes()
    header = ["ply", "format ascii 1.0" if text else "format binary_little_endian 1.0", f"element vertex {table.shape[0]}"]
    header += [f"property float {n}" for n in names]
    header.append("end_header")
    with open(path, "wb") as f:
        f.write(("\n".join(header) + "\n").encode("ascii"))
        if text:
            np.savetxt(f, table, fmt="%.9g")          # float32 round-trips through 9 significant digits
        else:
            f.write(np.ascontiguousarray(table, dtype="<f4").tobytes())


def load_ply(path: str) -> Dict[str, np.ndarray]:
    """Reader for the two variants save_ply writes (round-trip tests, .splat conversion)."""
    with open(path, "rb") as f:
        names, n, fmt = [], 0, None
        while True:
            line = f.readline().decode("ascii").strip()
            if line.startswith("format"):
                fmt = line.split()[1]
            elif line.startswith("element vertex"):
                n = int(line.split()[-1])
            elif line.startswith("property"):
                names.append(line.split()[-1])
            elif line == "end_header":
                break
        if fmt == "ascii":
            table = np.loadtxt(f, dtype=np.float32, ndmin=2)
        else:
            table = np.frombuffer(f.read(), dtype="<f4").reshape(n, len(names))
    assert table.shape == (n, len(names))
    return {name: table[:, i] for i, name in enumerate(names)}


def splat_bytes(vert: Dict[str, np.ndarray]) -> bytes:
    """convert.py:10-50 on PLY columns: 32 bytes per Gaussian -- position 3 x f32, exp(scale) 3 x f32,
    (0.5 + SH_C0 f_dc, sigmoid(opacity)) * 255 as 4 x u8, normalised quaternion * 128 + 128 as
    4 x u8 -- most significant (large, opaque) Gaussians first."""
    s0, s1, s2 = (np.asarray(vert[f"scale_{i}"], np.float32) for i in range(3))
    opacity = np.asarray(vert["opacity"], np.float32)
    order = np.argsort(-np.exp(s0 + s1 + s2) / (1 + np.exp(-opacity)))
    n = order.shape[0]
    out = np.zeros(n, dtype=[("pos", "<f4", 3), ("scale", "<f4", 3), ("rgba", "u1", 4), ("rot", "u1", 4)])
    out["pos"] = np.stack([vert["x"], vert["y"], vert["z"]], 1).astype(np.float32)[order]
    out["scale"] = np.exp(np.stack([s0, s1, s2], 1).astype(np.float32))[order]
    color = np.stack([0.5 + SH_C0 * np.asarray(vert["f_dc_0"], np.float32), 0.5 + SH_C0 * np.asarray(vert["f_dc_1"], np.float32),
                      0.5 + SH_C0 * np.asarray(vert["f_dc_2"], np.float32), 1 / (1 + np.exp(-opacity))], 1)     # (float64, like the reference)
    out["rgba"] = (color * 255).clip(0, 255).astype(np.uint8)[order]
    rot = np.stack([vert[f"rot_{i}"] for i in range(4)], 1).astype(np.float32)
    rot = rot / np.linalg.norm(rot, axis=1, keepdims=True)
    out["rot"] = (rot * 128 + 128).clip(0, 255).astype(np.uint8)[order]
    return out.tobytes()


def save_splat(human_gs_out: Dict[str, torch.Tensor], path: str, pose: str = "canonical") -> None:
    """The .splat file of the model's Gaussians, straight from the tensors (save_ply + convert.py in one step)."""
    table = ply_table(human_gs_out, pose)
    vert = {name: table[:, i] for i, name in enumerate(ply_attributes())}
    d = os.path.dirname(path)
    if d:
        os.makedirs(d, exist_ok=True)
    with open(path, "wb") as f:
        f.write(splat_bytes(vert))
