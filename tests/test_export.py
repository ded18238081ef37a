"""PLY / .splat export (sings_b200/export.py) against a literal restatement of the reference's
per-vertex code (vis.py:38-61, convert.py:10-50) on a small model: byte-identical .splat, identical
PLY columns, both PLY variants round-trip."""
import os

import numpy as np
import torch

from sings_b200 import export as ex


def model(n=137, seed=3):
    g = torch.Generator().manual_seed(seed)
    q = torch.randn(n, 4, generator=g)
    return {"xyz_canon": torch.randn(n, 3, generator=g), "xyz": torch.randn(n, 3, generator=g),
            "shs": 0.5 * torch.randn(n, 16, 3, generator=g), "opacity": torch.rand(n, 1, generator=g) * 0.98 + 0.01,
            "scales_canon": torch.exp(torch.randn(n, 3, generator=g) - 4.0), "rotq_canon": q}


def reference_table(out, pose):
    """vis.py:41-56, statement by statement."""
    xyz = out["xyz_canon" if pose == "canonical" else "xyz"].cpu().numpy()
    normals = np.zeros_like(xyz)
    f_dc = out["shs"][:, :1].transpose(1, 2).flatten(start_dim=1).contiguous().cpu().numpy()
    f_rest = out["shs"][:, 1:].transpose(1, 2).flatten(start_dim=1).contiguous().cpu().numpy()
    opacities = torch.log(out["opacity"] / (1 - out["opacity"])).cpu().numpy()
    scale = torch.log(out["scales_canon"]).cpu().numpy()
    rotation = out["rotq_canon"].cpu().numpy()
    return np.concatenate((xyz, normals, f_dc, f_rest, opacities, scale, rotation), axis=1)


def reference_splat(vert):
    """convert.py:10-50: the per-vertex loop."""
    from io import BytesIO
    n = len(vert["x"])
    sorted_indices = np.argsort(-np.exp(vert["scale_0"] + vert["scale_1"] + vert["scale_2"]) / (1 + np.exp(-vert["opacity"])))
    buffer = BytesIO()
    for idx in sorted_indices:
        v = {k: a[idx] for k, a in vert.items()}
        position = np.array([v["x"], v["y"], v["z"]], dtype=np.float32)
        scales = np.exp(np.array([v["scale_0"], v["scale_1"], v["scale_2"]], dtype=np.float32))
        rot = np.array([v["rot_0"], v["rot_1"], v["rot_2"], v["rot_3"]], dtype=np.float32)
        SH_C0 = 0.28209479177387814
        color = np.array([0.5 + SH_C0 * v["f_dc_0"], 0.5 + SH_C0 * v["f_dc_1"], 0.5 + SH_C0 * v["f_dc_2"],
                          1 / (1 + np.exp(-v["opacity"]))])
        buffer.write(position.tobytes())
        buffer.write(scales.tobytes())
        buffer.write((color * 255).clip(0, 255).astype(np.uint8).tobytes())
        buffer.write(((rot / np.linalg.norm(rot)) * 128 + 128).clip(0, 255).astype(np.uint8).tobytes())
    assert n * 32 == buffer.tell()
    return buffer.getvalue()


def test_ply_table_and_attribute_order():
    out = model()
    names = ex.ply_attributes()
    assert len(names) == 62 and names[:6] == ["x", "y", "z", "nx", "ny", "nz"] and names[6] == "f_dc_0"
    assert names[9] == "f_rest_0" and names[54] == "opacity" and names[55] == "scale_0" and names[58] == "rot_0"
    for pose in ("canonical", "deformed"):
        assert np.array_equal(ex.ply_table(out, pose), reference_table(out, pose).astype(np.float32))


def test_ply_round_trip_text_and_binary(tmp_path):
    out = model()
    table = ex.ply_table(out)
    for text in (True, False):
        p = os.path.join(tmp_path, "sub", f"m_{int(text)}.ply")
        ex.save_ply(out, p, text=text)
        cols = ex.load_ply(p)
        assert list(cols) == ex.ply_attributes()
        back = np.stack([cols[n] for n in ex.ply_attributes()], 1)
        assert np.array_equal(back, table)                      # %.9g round-trips binary32 exactly
    head = open(os.path.join(tmp_path, "sub", "m_1.ply"), "rb").read(64).decode("ascii", "ignore")
    assert head.startswith("ply\nformat ascii 1.0\nelement vertex 137\n")


def test_splat_bytes_identical_to_the_reference_loop(tmp_path):
    out = model(n=211, seed=8)
    table = ex.ply_table(out)
    vert = {n: table[:, i] for i, n in enumerate(ex.ply_attributes())}
    assert ex.splat_bytes(vert) == reference_splat(vert)
    p = os.path.join(tmp_path, "m.splat")
    ex.save_splat(out, p)
    assert open(p, "rb").read() == reference_splat(vert)
