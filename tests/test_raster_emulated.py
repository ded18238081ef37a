"""CPU: the WHOLE library -- every .cu of sings_b200/csrc behind the real C ABI of include/sings_b200.h --
compiled against the SIMT emulation of tests/cuda_emu and driven with host buffers: rasterizer forward
(geometry, depth + tile radix passes, pair emission with look-back scan, tile ranges, blend) bit-exact
against the C oracle, backward within the gradient bars, the fused avatar path, the stand-alone sort.
The same assertions as tests/test_gpu_raster.py at sizes the emulation finishes in seconds.  Test
infrastructure only: the product has no CPU path and cannot reach this build."""
import ctypes as C
import os
import shutil

import numpy as np
import pytest

from helpers import assert_grad_close, make_scene, oracle_camera
from oracle import raster_oracle as ro

IMG_TOL, GRAD_TOL = 1e-4, 1e-3
FLAG_PRECLEARED, FLAG_FORWARD_ONLY = 2, 8


@pytest.fixture(scope="module")
def L():
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    from cuda_emu import build_library
    from sings_b200 import _lib
    lib = build_library()
    for name, (res, args) in _lib._SIGNATURES.items():          # the binding table of the real library
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    return lib


def p(a):
    return None if a is None else a.ctypes.data


def buf(nbytes):
    """zero-initialised byte buffer, 256-byte aligned (what torch's allocator gives the real library)"""
    raw = np.zeros(nbytes + 256, np.uint8)
    off = (-raw.ctypes.data) % 256
    return raw[off:off + nbytes]


def c32(a):
    return None if a is None else np.ascontiguousarray(a, np.float32)


def layout(L, P, W, H, L_cap):
    info = (C.c_longlong * 16)()
    assert L.sgs_raster_layout_info(P, W, H, L_cap, info) == 0
    keys = ["counters", "keys_unsorted", "vals_unsorted", "keys_sorted", "vals_sorted", "ranges", "final_T", "n_contrib",
            "tiles", "end_bit", "passes", "rec_floats"]
    return {k: int(info[i]) for i, k in enumerate(keys)}


class Frame:
    """One forward through sgs_raster_clear + sgs_raster_forward, buffers kept for the backward."""

    def __init__(self, L, sc, bg, D, colors=None, L_cap=None, expect_rc=0, flags=0, preclear=True):
        view = sc["view"]
        self.L, self.sc, self.D = L, sc, D
        self.P = P = sc["means3D"].shape[0]
        self.W, self.H = W, H = view.image_width, view.image_height
        self.M = M = 0 if colors is not None else sc["shs"].shape[1]
        self.L_cap = L_cap = L_cap or max(4 * P, 1 << 16)
        s = [C.c_size_t() for _ in range(4)]
        assert L.sgs_raster_sizes(P, W, H, L_cap, *[C.byref(x) for x in s]) == 0
        self.geom, self.binning, self.img, self.acc = (buf(int(x.value)) for x in s)
        for b_ in (self.geom, self.binning, self.img, self.acc):
            b_[:] = 0xA5                         # scratch arrives dirty: whatever must be zero is cleared by the library
        self.bg, self.m3, self.opa = c32(bg), c32(sc["means3D"]), c32(sc["opacity"])
        self.sca, self.rot = c32(sc["scales"]), c32(sc["rotations"])
        self.shs, self.col = (None if colors is not None else c32(sc["shs"])), c32(colors)
        self.viewm, self.proj = c32(view.world_view_transform), c32(view.full_proj_transform)
        self.campos = c32(view.camera_center)
        self.tfx, self.tfy = float(view.tanfovx), float(view.tanfovy)
        self.color = np.full((3, H, W), np.nan, np.float32)
        self.radii = np.full(P, -7, np.int32)
        self.alpha, self.depth = np.full((H, W), np.nan, np.float32), np.full((H, W), np.nan, np.float32)
        self.counters = np.zeros(2, np.int32)
        if preclear:
            assert L.sgs_raster_clear(P, W, H, L_cap, p(self.binning), p(self.acc), None, 0, None) == 0
            flags |= FLAG_PRECLEARED
        rc = L.sgs_raster_forward(P, D, M, W, H, p(self.bg), p(self.m3), p(self.col), p(self.opa), p(self.sca), 1.0,
                                  p(self.rot), None, p(self.viewm), p(self.proj), p(self.campos), self.tfx, self.tfy,
                                  p(self.shs), 0, L_cap, p(self.geom), p(self.binning), p(self.img), p(self.color),
                                  p(self.radii), p(self.alpha), p(self.depth), p(self.counters), None, flags, None)
        assert rc == expect_rc, L.sgs_error_string(rc)

    def state(self):
        info = layout(self.L, self.P, self.W, self.H, self.L_cap)
        b, im, W, H = self.binning, self.img, self.W, self.H
        n = int(self.counters[0])
        out = dict(num_rendered=n, overflow=int(self.counters[1]))
        out["keys"] = b[info["keys_sorted"]:info["keys_sorted"] + 8 * n].view(np.uint64).copy()
        entries = b[info["vals_sorted"]:info["vals_sorted"] + 4 * n].view(np.uint32).copy()
        out["point_list"] = entries & np.uint32(0xffffff)
        out["ranges"] = b[info["ranges"]:info["ranges"] + 8 * info["tiles"]].view(np.uint32).reshape(-1, 2).copy()
        out["final_T"] = im[info["final_T"]:info["final_T"] + 4 * W * H].view(np.float32).reshape(H, W).copy()
        out["n_contrib"] = im[info["n_contrib"]:info["n_contrib"] + 4 * W * H].view(np.uint32).reshape(H, W).copy()
        return out

    def backward(self, G, precleared=True):
        P, M = self.P, self.M
        z = lambda *s: np.full(s, np.nan, np.float32)
        g = dict(means3D=z(P, 3), means2D=z(P, 3), colors=z(P, 3), opacities=z(P, 1), cov=z(P, 6),
                 sh=z(P, max(M, 1), 3), scales=z(P, 3), rotations=z(P, 4))
        rc = self.L.sgs_raster_backward(P, self.D, M, self.W, self.H, p(self.bg), p(self.m3), p(self.col), p(self.sca), 1.0,
                                        p(self.rot), None, p(self.viewm), p(self.proj), p(self.campos), self.tfx, self.tfy,
                                        p(self.shs), p(self.radii), p(c32(G)), self.L_cap, p(self.geom), p(self.binning),
                                        p(self.img), p(self.acc), p(g["means3D"]), p(g["means2D"]), p(g["colors"]),
                                        p(g["opacities"]), p(g["cov"]), p(g["sh"]) if M else None, p(g["scales"]),
                                        p(g["rotations"]), None, None, None, None, FLAG_PRECLEARED if precleared else 0, None)
        assert rc == 0, self.L.sgs_error_string(rc)
        return g


def check_forward(fr, st):
    ins = fr.state()
    assert ins["overflow"] == 0 and ins["num_rendered"] == st.num_rendered
    assert np.array_equal(fr.radii, st.radii)
    assert np.array_equal(ins["keys"], st.keys), "sorted (tile|depth) keys must be bit-exact"
    assert np.array_equal(ins["point_list"], st.point_list), "sorted Gaussian ids must be bit-exact"
    assert np.array_equal(ins["ranges"], st.ranges), "tile ranges must be bit-exact"
    assert np.array_equal(ins["n_contrib"], st.n_contrib), "contributor counts must be bit-exact"
    assert np.array_equal(fr.color, st.color) and np.array_equal(ins["final_T"], st.final_T)
    assert np.abs(fr.alpha - st.alpha).max() <= IMG_TOL
    assert np.abs(fr.depth - st.depth).max() <= IMG_TOL * max(1.0, float(st.depth.max()))


@pytest.mark.parametrize("N,H,W,D", [(600, 48, 64, 3), (900, 70, 45, 0)])
def test_rasterizer_forward_backward_through_the_c_abi(L, N, H, W, D):
    sc = make_scene(N=N, H=H, W=W, seed=N + D)
    sc["shs"], sc["opacity"] = sc["shs"], sc["opacity"]
    bg = np.array([0.2, 0.4, 0.6], np.float32)
    st = ro.forward(oracle_camera(sc["view"]), sc["means3D"], sc["opacity"], bg, sh_degree=D, shs=sc["shs"],
                    scales=sc["scales"], rotations=sc["rotations"])
    assert st.num_rendered > N            # a real workload: several tiles per Gaussian
    fr = Frame(L, sc, bg, D)
    check_forward(fr, st)
    # without the up-front clear kernel the entry points clear what they need themselves (dirty scratch either way),
    # and a second backward of the same forward finds the accumulator dirty
    own = Frame(L, sc, bg, D, preclear=False)
    check_forward(own, st)
    # animation frames (SGS_FLAG_FORWARD_ONLY: nothing is left for a backward) render the same bits
    fo = Frame(L, sc, bg, D, flags=FLAG_FORWARD_ONLY)
    assert np.array_equal(fo.color, fr.color) and np.array_equal(fo.radii, fr.radii) and np.array_equal(fo.alpha, fr.alpha)
    G = np.random.default_rng(1).normal(size=st.color.shape).astype(np.float32)
    got, ref = fr.backward(G), ro.backward(st, G)
    again = own.backward(G, precleared=False)
    second = own.backward(G, precleared=False)
    for k in ("means3D", "means2D", "opacities", "sh", "scales", "rotations"):
        assert np.abs(again[k] - got[k]).max() <= 2e-5 * np.abs(got[k]).max(), k
        assert np.abs(second[k] - got[k]).max() <= 2e-5 * np.abs(got[k]).max(), k
    for k, name in (("means3D", "means3D"), ("means2D", "means2D"), ("opacities", "opacities"), ("sh", "sh"),
                    ("scales", "scales"), ("rotations", "rotations")):
        assert_grad_close(got[k], np.asarray(ref[name]).reshape(got[k].shape), k, tol=GRAD_TOL)


def test_standalone_sort_matches_a_stable_sort(L):
    rng = np.random.default_rng(0)
    for n, end_bit in ((1, 13), (5000, 45), (9001, 64)):
        keys = rng.integers(0, 1 << 62, n, dtype=np.uint64) >> np.uint64(64 - end_bit) if end_bit < 64 else rng.integers(0, 1 << 63, n, dtype=np.uint64)
        keys[::7] = keys[0]                                      # duplicates: stability matters
        vals = np.arange(n, dtype=np.uint32)
        k0, v0 = keys.copy(), vals.copy()
        k1, v1 = np.zeros_like(keys), np.zeros_like(vals)
        scratch = buf(int(L.sgs_sort_scratch_bytes(n)))
        which = C.c_int(-1)
        assert L.sgs_sort_pairs_u64(p(k0), p(v0), p(k1), p(v1), p(scratch), scratch.nbytes, n, end_bit, C.byref(which), None) == 0
        ks, vs = (k1, v1) if which.value else (k0, v0)
        order = np.argsort(keys, kind="stable")
        assert np.array_equal(ks, keys[order]) and np.array_equal(vs, vals[order])


@pytest.mark.parametrize("iso", [False, True])
def test_fused_avatar_path_equals_the_separate_kernels(L, iso):
    """sgs_avatar_forward / _backward (LBS fused into the per-Gaussian rasterizer kernels, packed skinning
    weights) against pose -> A, stand-alone LBS and rasterizer kernels on the same inputs: same deformed
    Gaussians, same image bits, gradients to binary32 re-association; the deformation itself against the
    oracle of the reference's lbs_extra."""
    from oracle import lbs_oracle as lo
    from sings_b200 import _lib, synthetic as syn
    import torch
    H, W, N, J, D = 48, 64, 700, 24, 3
    av = syn.make_avatar(N, J, seed=31, isotropic=iso)
    view = syn.make_view(H, W)
    pose = c32(syn.random_pose(J, seed=33)).reshape(1, J, 3)
    transl = c32(syn.default_transl(H, focal=5000.0 * H / 896.0)).reshape(1, 3)
    ss = np.array([1.07], np.float32)
    bg = np.array([0.3, 0.6, 0.9], np.float32)
    G = np.random.default_rng(5).normal(size=(3, H, W)).astype(np.float32)
    xyz_c, scl_c, W_lbs = c32(av.xyz_canon), c32(av.scales), c32(av.lbs_weights)
    rot_c = None if iso else c32(av.rotmat_canon)
    rest, inv_A, parents = c32(av.rest), c32(av.inv_A_t2cano), np.ascontiguousarray(av.parents, np.int32)
    z = lambda *s: np.zeros(s, np.float32)
    # ---- separate kernels
    A, Gm = z(1, J, 4, 4), z(1, J, 12)
    xyz, rotq, sc = z(1, N, 3), z(1, N, 4), z(1, N, 3)
    assert L.sgs_pose_lbs_fwd(p(pose), p(rest), p(parents), p(inv_A), 1, N, J, p(A), p(Gm), p(xyz_c), p(W_lbs), p(rot_c),
                              p(scl_c), p(ss), p(transl), p(xyz), p(rotq), p(sc), None) == 0
    tc = torch.from_numpy
    A_o = lo.pose_to_A(tc(pose), tc(rest), av.parents, tc(inv_A))
    xo, qo, so, _ = lo.deform(A_o, tc(xyz_c), tc(W_lbs), tc(scl_c), None if iso else tc(rot_c), tc(ss).reshape(1, 1), tc(transl))
    assert np.abs(xyz - xo.numpy()).max() < 2e-5 and np.abs(sc - so.numpy()).max() < 1e-6
    scene = dict(means3D=xyz[0], rotations=rotq[0], scales=sc[0], opacity=av.opacity, shs=av.shs, view=view)
    sep = Frame(L, scene, bg, D)
    gs = sep.backward(G)
    d_xyz, d_scl, d_rot = z(N, 3), z(N, 3), (None if iso else z(N, 3, 3))
    d_A, d_tr, d_ss, d_pose = z(1, J, 4, 4), z(1, 3), z(1), z(1, J, 3)
    assert L.sgs_lbs_bwd(1, N, J, p(A), p(xyz_c), p(W_lbs), p(rot_c), p(scl_c), p(ss), p(transl), None, None, None,
                         p(gs["means3D"]), p(gs["rotations"]), p(gs["scales"]), None, p(d_xyz), p(d_rot), p(d_scl), p(d_A),
                         p(d_ss), p(d_tr), None) == 0
    assert L.sgs_pose_to_A_bwd(p(pose), p(rest), p(parents), p(inv_A), p(Gm), p(d_A), 1, J, p(d_pose), None) == 0
    # ---- fused kernels: packed weights, one per-Gaussian kernel each way
    K = 4
    while True:
        nb = int(L.sgs_lbs_packed_bytes(N, K))
        wq, iq, nnz = buf(nb).view(np.float32), buf(max(nb // 4, 16)).view(np.int32), np.zeros(1, np.int32)
        assert L.sgs_lbs_pack_weights(N, J, p(W_lbs), K, p(wq), p(iq), p(nnz), None) == 0
        if int(nnz[0]) <= K:
            break
        K = (int(nnz[0]) + 3) // 4 * 4
    f = dict(A=z(1, J, 4, 4), G=z(1, J, 12), xyz=z(1, N, 3), rotq=z(1, N, 4), sc=z(1, N, 3), d_xyz=z(N, 3), d_scl=z(N, 3),
             d_rot=None if iso else z(N, 3, 3), d_A=z(1, J, 4, 4), d_tr=z(1, 3), d_pose=z(1, J, 3))
    d = _lib.DeformArgs()
    d.N, d.J, d.K, d.rot6d = N, J, K, 0
    for name, arr in (("pose", pose), ("rest", rest), ("parents", parents), ("inv_A_t2cano", inv_A), ("xyz_canon", xyz_c),
                      ("scales", scl_c), ("rot_canon", rot_c), ("wq", wq), ("iq", iq), ("smpl_scale", ss), ("transl", transl),
                      ("A", f["A"]), ("G", f["G"]), ("xyz", f["xyz"]), ("rotq", f["rotq"]), ("scales_out", f["sc"]),
                      ("d_xyz_canon", f["d_xyz"]), ("d_rot_canon", f["d_rot"]), ("d_scales", f["d_scl"]), ("d_A", f["d_A"]),
                      ("d_transl", f["d_tr"]), ("d_pose", f["d_pose"])):
        setattr(d, name, p(arr))
    L_cap = sep.L_cap
    s = [C.c_size_t() for _ in range(4)]
    assert L.sgs_raster_sizes(N, W, H, L_cap, *[C.byref(x) for x in s]) == 0
    geom, binning, img, acc = (buf(int(x.value)) for x in s)
    color, radii, cnt = z(3, H, W), np.zeros(N, np.int32), np.zeros(2, np.int32)
    opa, shs = c32(av.opacity), c32(av.shs)
    M = shs.shape[1]
    assert L.sgs_raster_clear(N, W, H, L_cap, p(binning), p(acc), None, 0, None) == 0
    rc = L.sgs_avatar_forward(C.byref(d), D, M, W, H, p(bg), p(opa), 1.0, p(sep.viewm), p(sep.proj), p(sep.campos), sep.tfx,
                              sep.tfy, p(shs), L_cap, p(geom), p(binning), p(img), p(color), p(radii), None, None, p(cnt), None,
                              FLAG_PRECLEARED, None)
    assert rc == 0, L.sgs_error_string(rc)
    assert int(cnt[1]) == 0 and int(cnt[0]) == int(sep.counters[0]) > N
    assert np.array_equal(f["xyz"], xyz) and np.array_equal(f["rotq"], rotq) and np.array_equal(f["sc"], sc)
    assert np.array_equal(color, sep.color) and np.array_equal(radii, sep.radii)
    g_m2, d_opa, d_shs = z(N, 3), z(N, 1), z(N, M, 3)
    # densification statistics ride in the same kernel (sings_hybrid.py:1013-1015, gs_trainer.py:487-490): accumulated
    accum, denom, maxr = np.full(N, 0.5, np.float32), np.full(N, 2.0, np.float32), np.full(N, 3.0, np.float32)
    rc = L.sgs_avatar_backward(C.byref(d), D, M, W, H, p(bg), 1.0, p(sep.viewm), p(sep.proj), p(sep.campos), sep.tfx, sep.tfy,
                               p(shs), p(radii), p(G), L_cap, p(geom), p(binning), p(img), p(acc), p(g_m2), p(d_opa), p(d_shs),
                               p(accum), p(denom), p(maxr), None, FLAG_PRECLEARED, None)
    assert rc == 0, L.sgs_error_string(rc)
    vis = radii > 0
    assert vis.any()
    np.testing.assert_allclose(accum, 0.5 + vis * np.sqrt(g_m2[:, 0] ** 2 + g_m2[:, 1] ** 2), rtol=2e-6)
    assert np.array_equal(denom, 2.0 + vis) and np.array_equal(maxr, np.where(vis, np.maximum(3.0, radii), 3.0).astype(np.float32))
    rel = lambda a, b: np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)
    pairs = [("d_xyz_canon", f["d_xyz"], d_xyz), ("d_scales", f["d_scl"], d_scl), ("d_pose", f["d_pose"], d_pose),
             ("d_transl", f["d_tr"], d_tr), ("means2D", g_m2, gs["means2D"]), ("d_opacity", d_opa, gs["opacities"]),
             ("d_shs", d_shs, gs["sh"])]
    if not iso:
        pairs.append(("d_rot_canon", f["d_rot"], d_rot))
    for name, a, b in pairs:
        assert rel(a, b) < 2e-5, (name, rel(a, b))


def test_knn_and_mark_visible_through_the_c_abi(L):
    """sgs_knn_mean_dist (grid build on the library's radix sort + growing cell-block search) against brute
    force; sgs_mark_visible against the view-space depth test."""
    rng = np.random.default_rng(3)
    N, K = 700, 8
    xyz = np.ascontiguousarray(rng.random((N, 3)).astype(np.float32) * np.array([0.6, 1.7, 0.3], np.float32))
    scratch = buf(int(L.sgs_knn_scratch_bytes(N)))
    mean, idx, d2 = np.full(N, np.nan, np.float32), np.full((N, K), -5, np.int32), np.full((N, K), np.nan, np.float32)
    assert L.sgs_knn_mean_dist(N, p(xyz), K, p(scratch), scratch.nbytes, p(mean), p(idx), p(d2), None) == 0
    D = ((xyz[:, None, :].astype(np.float64) - xyz[None, :, :]) ** 2).sum(-1)
    np.fill_diagonal(D, np.inf)
    order = np.argsort(D, axis=1, kind="stable")[:, :K]
    ref_d2 = np.take_along_axis(D, order, 1)
    assert np.abs(d2 - ref_d2).max() <= 1e-6 * ref_d2.max()
    assert np.array_equal(np.sort(idx, 1), np.sort(order, 1))
    assert np.abs(mean - np.sqrt(ref_d2).mean(1)).max() <= 2e-6 * np.sqrt(ref_d2).mean(1).max()
    # the scale-edge loss of the reference (GaussiansEdgeLoss run with a brute-force knn_points, reg_golden_edge.npz)
    # from the library's neighbour distances: ((scale_i - mean edge length_i)^2).mean()
    zg = np.load(os.path.join(os.path.dirname(__file__), "golden", "reg_golden_edge.npz"))
    xg = np.ascontiguousarray(zg["xyz_canon"])
    Ng = xg.shape[0]
    scratch = buf(int(L.sgs_knn_scratch_bytes(Ng)))
    mg = np.full(Ng, np.nan, np.float32)
    assert L.sgs_knn_mean_dist(Ng, p(xg), 8, p(scratch), scratch.nbytes, p(mg), None, None, None) == 0
    loss = float(((zg["scales"][:, 0].astype(np.float64) - mg) ** 2).mean())
    assert abs(loss - float(zg["loss_f64"])) <= 1e-5 * float(zg["loss_f64"])
    sc = make_scene(N=300, H=32, W=32, seed=2)
    pts = np.ascontiguousarray(sc["means3D"].copy())
    pts[::3, 2] -= 20.0                                      # a third of the points behind the camera
    present = np.full(300, 7, np.uint8)
    view = c32(sc["view"].world_view_transform)
    assert L.sgs_mark_visible(300, p(pts), p(view), p(present), None) == 0
    vm = view.reshape(4, 4)
    z = pts @ vm[:3, 2] + vm[3, 2]
    assert np.array_equal(present.astype(bool), z > 0.2)


def test_edge_cases_through_the_c_abi(L):
    """No Gaussians, every Gaussian culled, and a pair list that does not fit its capacity (reported, nothing
    written out of bounds; rendered again with room it equals the oracle) -- on dirty scratch buffers."""
    bg = np.array([0.1, 0.5, 0.9], np.float32)
    sc = make_scene(N=400, H=40, W=56, seed=9, scale_range=(0.02, 0.05))
    # every Gaussian behind the camera
    hidden = dict(sc)
    hidden["means3D"] = np.ascontiguousarray(sc["means3D"] - np.array([0, 0, 40], np.float32))
    fr = Frame(L, hidden, bg, 3)
    assert int(fr.counters[0]) == 0 and not fr.radii.any()
    assert np.array_equal(fr.color, np.broadcast_to(bg[:, None, None], fr.color.shape)) and not fr.alpha.any()
    g = fr.backward(np.ones((3, 40, 56), np.float32))
    assert all(not np.asarray(v).any() for k, v in g.items() if k not in ("colors", "cov"))
    # no Gaussians at all
    empty = dict(sc)
    for k in ("means3D", "opacity", "shs", "scales", "rotations"):
        empty[k] = np.ascontiguousarray(sc[k][:0])
    fr0 = Frame(L, empty, bg, 3)
    assert np.array_equal(fr0.color, np.broadcast_to(bg[:, None, None], fr0.color.shape))
    # more Gaussians than a list entry can name (id in 24 bits): refused before anything is launched
    one = np.zeros(16, np.float32)
    rc = L.sgs_raster_forward(1 << 24, 0, 1, 16, 16, p(one), p(one), None, p(one), p(one), 1.0, p(one), None, p(one), p(one), p(one),
                              1.0, 1.0, p(one), 0, 1 << 16, p(one), p(one), p(one), p(one), p(one), None, None, None, None, 0, None)
    assert rc == -5 and b"capacity" in L.sgs_error_string(rc)
    # capacity too small: the needed count comes back with the overflow flag set
    st = ro.forward(oracle_camera(sc["view"]), sc["means3D"], sc["opacity"], bg, sh_degree=3, shs=sc["shs"],
                    scales=sc["scales"], rotations=sc["rotations"])
    assert st.num_rendered > 512
    small = Frame(L, sc, bg, 3, L_cap=512)
    assert int(small.counters[1]) == 1 and int(small.counters[0]) == st.num_rendered
    check_forward(Frame(L, sc, bg, 3, L_cap=int(st.num_rendered * 1.3) + 4096), st)


def test_precomputed_colours_and_joint_counts(L):
    """colors_precomp instead of SH, and the SMPL-H joint count (J = 52) through the fused deform kernel."""
    from oracle import lbs_oracle as lo
    from sings_b200 import synthetic as syn
    import torch
    sc = make_scene(N=500, H=48, W=48, seed=21)
    cols = np.random.default_rng(2).uniform(size=(500, 3)).astype(np.float32)
    bg = np.zeros(3, np.float32)
    st = ro.forward(oracle_camera(sc["view"]), sc["means3D"], sc["opacity"], bg, sh_degree=0, colors_precomp=cols,
                    scales=sc["scales"], rotations=sc["rotations"])
    check_forward(Frame(L, sc, bg, 0, colors=cols), st)
    N, J = 300, 52
    av = syn.make_avatar(N, J, seed=4)
    pose = c32(syn.random_pose(J, seed=6)).reshape(1, J, 3)
    transl = np.array([[0.1, -0.2, 9.0]], np.float32)
    A, Gm = np.zeros((1, J, 4, 4), np.float32), np.zeros((1, J, 12), np.float32)
    xyz, q, s_ = np.zeros((1, N, 3), np.float32), np.zeros((1, N, 4), np.float32), np.zeros((1, N, 3), np.float32)
    assert L.sgs_pose_lbs_fwd(p(pose), p(c32(av.rest)), p(np.ascontiguousarray(av.parents, np.int32)), p(c32(av.inv_A_t2cano)), 1, N, J,
                              p(A), p(Gm), p(c32(av.xyz_canon)), p(c32(av.lbs_weights)), p(c32(av.rotmat_canon)), p(c32(av.scales)),
                              None, p(transl), p(xyz), p(q), p(s_), None) == 0
    tc = torch.from_numpy
    A_o = lo.pose_to_A(tc(pose), tc(c32(av.rest)), av.parents, tc(c32(av.inv_A_t2cano)))
    xo, qo, so, _ = lo.deform(A_o, tc(c32(av.xyz_canon)), tc(c32(av.lbs_weights)), tc(c32(av.scales)), tc(c32(av.rotmat_canon)),
                              None, tc(transl))
    assert np.abs(A - A_o.numpy()).max() < 1e-5 and np.abs(xyz - xo.numpy()).max() < 2e-5
    assert np.abs(s_ - so.numpy()).max() < 1e-6
