"""Random small scenes through the emulated library (tests/cuda_emu: the kernel sources compiled for the host, real
C ABI) against the C oracle: keys / ranges / contributor counts / image bit-equal, gradients within the bars.
Odd image sizes, every SH degree, tiny and huge Gaussians, opacities at the ends of (0, 1], turned cameras.
    python tools/emu_fuzz.py [n_scenes] [first_seed]              the rasterizer (about 10-20 s per scene)
    python tools/emu_fuzz.py --deform [n_cases] [first_seed]      the deformer: pose -> A, LBS forward / backward (matrix and 6D
                                                                  canonical rotations, isotropic, smpl_scale, transl, ext_tfs, B frames,
                                                                  J = 24 / 52) against the float64 oracle of the reference's lbs_extra path
    python tools/emu_fuzz.py --stress                             four fixed shapes: long per-tile lists (every Gaussian covers the
                                                                  image), many tiles, a 1936-pixel-wide strip (121 tile columns), a tall strip
CPU only; a divergence prints the case's parameters and exits 1."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np

import test_raster_emulated as T
from helpers import grad_errors, make_scene, oracle_camera
from oracle import raster_oracle as ro


def load():
    from cuda_emu import build_library
    from sings_b200 import _lib
    L = build_library()
    for name, (res, args) in _lib._SIGNATURES.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    return L


def deform_fuzz(n, first):
    import torch
    from oracle import lbs_oracle as lo
    from oracle.raster_ref64 import quat_to_R
    from sings_b200 import synthetic as syn
    L, p, c32 = load(), T.p, T.c32
    worst = {}
    for seed in range(first, first + n):
        rng = np.random.default_rng(seed)
        gen = torch.Generator().manual_seed(seed)
        N, J, B = int(rng.choice([1, 31, 255, 256, 257, 600, 901])), int(rng.choice([24, 52])), int(rng.integers(1, 4))
        mode = str(rng.choice(["matrix", "rot6d", "iso"]))
        use_ss, use_tr, use_ext = bool(rng.integers(0, 2)), bool(rng.integers(0, 2)), bool(rng.integers(0, 2))
        par = dict(seed=seed, N=N, J=J, B=B, mode=mode, smpl_scale=use_ss, transl=use_tr, ext=use_ext)
        av = syn.make_avatar(N, J, seed=seed % 89, smooth_weights=int(rng.integers(0, 3)))
        t = torch.from_numpy
        pose = torch.stack([t(syn.random_pose(J, seed=seed * 7 + b)) for b in range(B)])
        d6 = torch.randn(N, 6, generator=gen)
        ss = 1.0 + 0.1 * torch.rand(B, 1, generator=gen) if use_ss else None
        tr = torch.randn(B, 3, generator=gen) if use_tr else None
        ext = (torch.randn(B, 3, generator=gen), lo.batch_rodrigues(torch.randn(B, 3, generator=gen)),
               0.5 + torch.rand(B, 1, generator=gen)) if use_ext else None
        gx, gq, gs = (torch.randn(B, N, k, generator=gen) for k in (3, 4, 3))
        # float64 truth: the oracle of the reference's path, autograd
        dd = lambda a: None if a is None else a.double()
        pose_o = dd(pose).requires_grad_(True)
        A_o = lo.pose_to_A(pose_o, dd(t(av.rest)), av.parents, dd(t(av.inv_A_t2cano)))
        x_o, s_o = dd(t(av.xyz_canon)).requires_grad_(True), dd(t(av.scales)).requires_grad_(True)
        r_o = None if mode == "iso" else (dd(d6) if mode == "rot6d" else dd(t(av.rotmat_canon))).requires_grad_(True)
        kw = dict(rot6d_canon=r_o) if mode == "rot6d" else {}
        xo, qo, sco, _ = lo.deform(A_o, x_o, dd(t(av.lbs_weights)), s_o, r_o if mode == "matrix" else None, dd(ss), dd(tr),
                                   tuple(dd(e) for e in ext) if ext else None, **kw)
        ((xo * gx).sum() + (qo * gq).sum() + (sco * gs).sum()).backward()
        # emulated kernels through the C ABI
        f = lambda a: None if a is None else c32(a.detach().numpy())
        z = lambda *sh: np.zeros(sh, np.float32)
        pose_n, rest, inv_A, par_n = f(pose), c32(av.rest), c32(av.inv_A_t2cano), np.ascontiguousarray(av.parents, np.int32)
        xyz_c, W_c, scl_c = c32(av.xyz_canon), c32(av.lbs_weights), c32(av.scales)
        rot_c = None if mode == "iso" else (f(d6) if mode == "rot6d" else c32(av.rotmat_canon))
        ss_n, tr_n = (None if ss is None else f(ss).reshape(B)), f(tr)
        et, er, es = (f(ext[0]), f(ext[1]), f(ext[2]).reshape(B)) if ext else (None, None, None)
        A, Gm = z(B, J, 4, 4), z(B, J, 12)
        try:
            assert L.sgs_pose_to_A(p(pose_n), p(rest), p(par_n), p(inv_A), B, J, p(A), p(Gm), None) == 0
            assert np.abs(A - A_o.detach().numpy()).max() < 1e-5, "A"
            xo_n, qo_n, so_n = z(B, N, 3), z(B, N, 4), z(B, N, 3)
            fwd = L.sgs_lbs_fwd_rot6d if mode == "rot6d" else L.sgs_lbs_fwd
            assert fwd(B, N, J, p(A), p(xyz_c), p(W_c), p(rot_c), p(scl_c), p(ss_n), p(tr_n), p(et), p(er), p(es), p(xo_n), p(qo_n),
                       p(so_n), None, None) == 0
            assert np.abs(xo_n - xo.detach().numpy()).max() < 3e-5, "xyz"
            assert np.abs(so_n - sco.detach().numpy()).max() < 1e-5, "scales"
            Rk, Ro = quat_to_R(t(qo_n).reshape(-1, 4).double()), quat_to_R(qo.detach().reshape(-1, 4))
            assert (Rk - Ro).abs().max() < 1e-4, "rotation"
            d_xyz, d_scl = z(N, 3), z(N, 3)
            d_rot = None if mode == "iso" else (z(N, 6) if mode == "rot6d" else z(N, 3, 3))
            d_A, d_ss, d_tr, d_pose = z(B, J, 4, 4), (z(B) if use_ss else None), (z(B, 3) if use_tr else None), z(B, J, 3)
            bwd = L.sgs_lbs_bwd_rot6d if mode == "rot6d" else L.sgs_lbs_bwd
            assert bwd(B, N, J, p(A), p(xyz_c), p(W_c), p(rot_c), p(scl_c), p(ss_n), p(tr_n), p(et), p(er), p(es), p(f(gx)), p(f(gq)),
                       p(f(gs)), None, p(d_xyz), p(d_rot), p(d_scl), p(d_A), p(d_ss), p(d_tr), None) == 0
            assert L.sgs_pose_to_A_bwd(p(pose_n), p(rest), p(par_n), p(inv_A), p(Gm), p(d_A), B, J, p(d_pose), None) == 0
            checks = [("d_pose", d_pose, pose_o.grad), ("d_xyz", d_xyz, x_o.grad), ("d_scales", d_scl, s_o.grad)]
            if d_rot is not None:
                checks.append(("d_rot", d_rot, r_o.grad))
            for name, got, ref in checks:
                e = grad_errors(got, ref.numpy().reshape(got.shape))
                worst[name] = max(worst.get(name, 0.0), e["max_rel"])
                assert e["max_rel"] <= 1e-3 and e["l2_rel"] <= 1e-3, (name, e)
        except AssertionError as e:
            print("DIVERGENCE", par, "\n", str(e)[:600], flush=True)
            sys.exit(1)
        print("ok", par, flush=True)
    print(f"{n} cases: emulated deformer kernels == oracle; largest tensor-relative gradient errors:",
          {k: float(f"{v:.2e}") for k, v in worst.items()})


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--deform":
        return deform_fuzz(int(sys.argv[2]) if len(sys.argv) > 2 else 10, int(sys.argv[3]) if len(sys.argv) > 3 else 0)
    stress = len(sys.argv) > 1 and sys.argv[1] == "--stress"
    fixed = [(400, 64, 96, 3, 0.2, 0.4), (2500, 144, 208, 1, 0.002, 0.01), (900, 32, 1936, 2, 0.004, 0.03), (700, 1100, 16, 0, 0.004, 0.03)]
    n = len(fixed) if stress else (int(sys.argv[1]) if len(sys.argv) > 1 else 10)
    first = 0 if stress else (int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    L = load()
    worst = {}
    for seed in range(first, first + n):
        rng = np.random.default_rng(seed)
        N = int(rng.integers(40, 700))
        H, W = int(rng.integers(17, 90)), int(rng.integers(17, 90))
        D = int(rng.integers(0, 4))
        lo = float(rng.choice([0.001, 0.004, 0.02]))
        hi = lo * float(rng.choice([2.0, 8.0, 40.0]))
        if stress:
            N, H, W, D, lo, hi = fixed[seed]
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
