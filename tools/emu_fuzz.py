"""Random small scenes through the emulated library (tests/cuda_emu: the kernel sources compiled for the host, real
C ABI) against the C oracle: keys / ranges / contributor counts / image bit-equal, gradients within the bars.
Odd image sizes, every SH degree, tiny and huge Gaussians, opacities at the ends of (0, 1], turned cameras.
    python tools/emu_fuzz.py [n_scenes] [first_seed]
CPU only (about 10 s per scene); a divergence prints the scene's parameters and exits 1."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np

import test_raster_emulated as T
from helpers import grad_errors, make_scene, oracle_camera
from oracle import raster_oracle as ro


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    from cuda_emu import build_library
    from sings_b200 import _lib
    L = build_library()
    for name, (res, args) in _lib._SIGNATURES.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    worst = {}
    for seed in range(first, first + n):
        rng = np.random.default_rng(seed)
        N = int(rng.integers(40, 700))
        H, W = int(rng.integers(17, 90)), int(rng.integers(17, 90))
        D = int(rng.integers(0, 4))
        lo = float(rng.choice([0.001, 0.004, 0.02]))
        hi = lo * float(rng.choice([2.0, 8.0, 40.0]))
        par = dict(seed=seed, N=N, H=H, W=W, D=D, scale_range=(lo, hi), yaw=float(rng.uniform(0, 6.28)),
                   fill=float(rng.choice([0.5, 0.85, 1.6])), iso=bool(rng.integers(0, 2)))
        sc = make_scene(N=N, H=H, W=W, seed=seed, scale_range=(lo, hi), isotropic=par["iso"], yaw=par["yaw"], fill=par["fill"])
        op = sc["opacity"].copy()
        k = rng.integers(0, N, max(N // 10, 1))
        op[k] = rng.choice(np.array([1.0, 1e-3, 0.0039, 0.00393, 0.999], np.float32), k.size).reshape(-1, 1)
        sc["opacity"] = op
        bg = rng.uniform(size=3).astype(np.float32)
        try:
            st = ro.forward(oracle_camera(sc["view"]), sc["means3D"], sc["opacity"], bg, sh_degree=D, shs=sc["shs"],
                            scales=sc["scales"], rotations=sc["rotations"])
            fr = T.Frame(L, sc, bg, D, L_cap=max(int(st.num_rendered * 1.2) + 1024, 1 << 12))
            T.check_forward(fr, st)
            if st.num_rendered:
                G = rng.normal(size=st.color.shape).astype(np.float32)
                got, ref = fr.backward(G), ro.backward(st, G)
                for key in ("means3D", "means2D", "opacities", "sh", "scales", "rotations"):
                    b = np.asarray(ref[key]).reshape(got[key].shape)
                    if np.abs(b).max() > 0:
                        # the bars of the parity tests (1e-3 tensor- and L2-relative); element-wise, sub-pixel Gaussians leave
                        # a few entries that are pure cancellation (1e-6 of the tensor's scale): at most 0.5 % may miss 1e-3
                        e = grad_errors(got[key], b)
                        worst[key] = max(worst.get(key, 0.0), e["max_rel"])
                        assert e["max_rel"] <= T.GRAD_TOL and e["l2_rel"] <= T.GRAD_TOL and e["bad_frac"] <= 5e-3, (key, e)
                    else:
                        assert not got[key].any(), key
        except AssertionError as e:
            print("DIVERGENCE", par, "pairs", getattr(st, "num_rendered", None), "\n", str(e)[:600], flush=True)
            sys.exit(1)
        print("ok", par, "pairs", st.num_rendered, flush=True)
    print(f"{n} scenes: emulated kernels == oracle (forward bit for bit); largest tensor-relative gradient errors:",
          {k: float(f"{v:.2e}") for k, v in worst.items()})


if __name__ == "__main__":
    main()
