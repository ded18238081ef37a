"""Golden vectors for SH -> RGB from the reference's own evaluator
(/root/reference/sings/rec/utils/visualize/spherical_harmonics.py:30-47 constants, :61-125
eval_sh).  The module moves its constants to CUDA at import time, so its source is executed
with `.cuda()` stripped and without the TorchScript decorator (CPU container).
Run in the build container:  python tests/golden/make_sh_golden.py
"""
import os
import re

import numpy as np
import torch

SRC = "/root/reference/sings/rec/utils/visualize/spherical_harmonics.py"
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    src = open(SRC).read().replace(".cuda()", "").replace("@torch.jit.script", "")
    ns = {}
    exec(compile(src, SRC, "exec"), ns)
    g = torch.Generator().manual_seed(5)
    P = 64
    dirs = torch.randn(P, 3, generator=g, dtype=torch.float64)
    dirs = dirs / dirs.norm(dim=1, keepdim=True)
    sh = torch.randn(P, 16, 3, generator=g, dtype=torch.float64)
    out = {}
    for D in range(4):
        # reference layout: sh [..., C, coeffs]
        res = ns["eval_sh"](D, sh.transpose(1, 2), dirs, ns["C0"].double(), ns["C1"].double(),
                            ns["C2"].double(), ns["C3"].double(), ns["C4"].double())
        out[f"rgb_deg{D}"] = res.numpy()
    # camera convention: the reference's projection matrix (utils/graphics.py:65-85)
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_graphics", "/root/reference/sings/rec/utils/graphics.py")
    gm = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gm)
    out["proj_a"] = gm.get_projection_matrix(0.01, 100.0, 0.4, 0.4).numpy()
    out["proj_b"] = gm.get_projection_matrix(0.01, 100.0, 0.2276, 0.3962).numpy()
    consts = dict(C0=ns["C0"].numpy(), C1=ns["C1"].numpy(), C2=ns["C2"].numpy(), C3=ns["C3"].numpy())
    np.savez_compressed(os.path.join(HERE, "sh_golden.npz"), dirs=dirs.numpy(), sh=sh.numpy(), **out, **consts)
    print("wrote sh_golden.npz")


if __name__ == "__main__":
    main()
