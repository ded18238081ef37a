"""Generate golden vectors for the tri-plane interpolation by running the REFERENCE's own module on CPU.

Run in the build container only (needs /root/reference):  python tests/golden/make_hexplane_golden.py
Executed reference code: HexPlaneField (init_grid_param, normalize_aabb, interpolate_ms_features,
grid_sample_wrapper), /root/reference/sings/rec/models/modules/hexplane.py:18-189.
Small planes (the fixture holds them), points inside, on and outside the box (border padding)."""
import importlib.util
import os

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    spec = importlib.util.spec_from_file_location("ref_hexplane", f"{REF}/sings/rec/models/modules/hexplane.py")
    H = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(H)
    cases = [dict(name="a", reso=[8, 6, 5], multires=[1, 2], C=32, bounds=1.0, N=300, seed=1),
             dict(name="b", reso=[4, 3, 2], multires=[1, 2, 4], C=64, bounds=1.6, N=257, seed=2)]
    for c in cases:
        torch.manual_seed(c["seed"])
        cfg = {"grid_dimensions": 2, "input_coordinate_dim": 3, "output_coordinate_dim": c["C"],
               "resolution": c["reso"], "multires": c["multires"]}
        field = H.HexPlaneField(planeconfig=cfg, bounds=c["bounds"], device="cpu")
        g = torch.Generator().manual_seed(c["seed"] + 10)
        pts = (torch.rand(c["N"], 3, generator=g) * 2 - 1) * c["bounds"] * 1.15        # ~13 % outside the box
        pts[:5] = torch.tensor([[c["bounds"], 0, 0], [-c["bounds"], c["bounds"], 0], [0, 0, -c["bounds"]],
                                [0, 0, 0], [c["bounds"], c["bounds"], c["bounds"]]], dtype=torch.float32)     # on the faces / centre
        pts.requires_grad_(True)
        feats = field(pts)
        d_out = torch.randn(feats.shape, generator=g)
        params = [p for gp in field.grids for p in gp]
        grads = torch.autograd.grad((feats * d_out).sum(), [pts] + params)
        out = dict(pts=pts.detach().numpy(), aabb=field.aabb.detach().numpy(), feats=feats.detach().numpy(), d_out=d_out.numpy(),
                   d_pts=grads[0].numpy(), reso=np.array(c["reso"]), multires=np.array(c["multires"]), C=np.array(c["C"]),
                   bounds=np.array(c["bounds"], np.float32))
        for i, (p, gp) in enumerate(zip(params, grads[1:])):
            out[f"plane_{i}"] = p.detach().numpy()
            out[f"d_plane_{i}"] = gp.numpy()
        np.savez_compressed(os.path.join(HERE, f"hexplane_golden_{c['name']}.npz"), **out)
        print(c["name"], feats.shape, float(feats.abs().mean()))


if __name__ == "__main__":
    main()
