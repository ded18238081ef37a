"""CPU: the rasterizer oracle (oracle/c/raster_oracle.c).  The upstream rasterizer is not in
the reference tree ("parity unpinned"), so the oracle is anchored on what IS citable:
the reference's SH evaluator and projection matrix (golden vectors generated from
spherical_harmonics.py:30-125 and graphics.py:65-85), a float64 torch-autograd restatement for
every gradient, analytic cases, and structural properties of keys / ranges / counts."""
import math
import os

import numpy as np
import pytest
import torch

from helpers import make_scene, oracle_camera, rel_err
from oracle import raster_oracle as ro
from oracle import raster_ref64 as r64
from sings_b200 import synthetic as syn

GOLD = os.path.join(os.path.dirname(__file__), "golden", "sh_golden.npz")


def test_sh_matches_reference_evaluator():
    g = np.load(GOLD)
    assert abs(float(g["C0"]) - 0.28209479177387814) < 1e-7 and abs(float(g["C1"]) - 0.4886025119029199) < 1e-7
    dirs, sh = g["dirs"], g["sh"]
    for D in range(4):
        ref = g[f"rgb_deg{D}"]                                   # reference eval_sh, float64
        mine = r64.sh_to_rgb(D, torch.from_numpy(sh), torch.from_numpy(dirs)).numpy()
        assert np.abs(mine - ref).max() < 2e-7      # reference constants are float32
        # C oracle: camera at the origin looking down +z, Gaussians placed along the directions
        pos = (dirs * 5.0).astype(np.float32)
        view = syn.make_view(64, 64, focal=20.0)     # wide field of view
        cam = oracle_camera(view)
        P = pos.shape[0]
        st = ro.forward(cam, pos, np.full((P, 1), 0.5, np.float32), np.zeros(3, np.float32),
                        shs=sh.astype(np.float32), scales=np.full((P, 3), 0.05, np.float32),
                        rotations=np.tile(np.array([1, 0, 0, 0], np.float32), (P, 1)), sh_degree=D)
        vis = st.radii > 0
        assert vis.sum() > 5
        want = np.maximum(ref + 0.5, 0.0)
        assert np.abs(st.rgb[vis] - want[vis]).max() < 2e-5
        assert np.array_equal(st.clamped[vis].astype(bool), (ref[vis] + 0.5) < 0)


def test_projection_matrix_matches_reference():
    g = np.load(GOLD)
    assert np.abs(syn.projection_matrix(0.01, 100.0, 0.4, 0.4) - g["proj_a"]).max() < 1e-6
    assert np.abs(syn.projection_matrix(0.01, 100.0, 0.2276, 0.3962) - g["proj_b"]).max() < 1e-6


def test_higher_msb_and_expneg():
    for n, want in [(4096, 13), (8160, 13), (16384, 15), (1024, 11), (16, 5), (1, 1)]:
        assert ro.higher_msb(n) == want
    xs = np.concatenate([np.linspace(-20, 0, 4001), [-79.9, -80.5, -1e-8, 0.0]])
    for x in xs:
        e, t = ro.expneg(float(np.float32(x))), math.exp(float(np.float32(x)))
        if x < -80:
            assert 0.0 < e < 2e-35          # clamped at exp(-80)
        else:
            assert abs(e - t) <= 2.5e-7 * t
    assert ro.expneg(0.0) == 1.0


@pytest.mark.parametrize("mode", ["sh3", "sh1", "colors", "cov3d"])
def test_backward_matches_float64_autograd(mode):
    sc = make_scene(N=600, H=64, W=80, seed=3, scale_range=(0.01, 0.05))
    cam = oracle_camera(sc["view"])
    bg = np.array([0.3, 0.5, 0.7], np.float32)
    kw = dict(scales=sc["scales"], rotations=sc["rotations"])
    D = 0
    rng = np.random.default_rng(0)
    if mode == "colors":
        kw["colors_precomp"] = rng.uniform(size=(600, 3)).astype(np.float32)
    else:
        kw["shs"] = sc["shs"]
        D = 3 if mode != "sh1" else 1
    if mode == "cov3d":
        st0 = ro.forward(cam, sc["means3D"], sc["opacity"], bg, sh_degree=D, **kw)
        kw = dict(shs=sc["shs"], cov3D_precomp=np.ascontiguousarray(
            ro.forward(cam, sc["means3D"], sc["opacity"], bg, shs=sc["shs"], scales=sc["scales"],
                       rotations=sc["rotations"], sh_degree=D).cov3D))
        # cov3D of culled Gaussians is zero; give them something valid
        bad = kw["cov3D_precomp"].sum(1) == 0
        kw["cov3D_precomp"][bad] = np.array([1e-4, 0, 0, 1e-4, 0, 1e-4], np.float32)
    st = ro.forward(cam, sc["means3D"], sc["opacity"], bg, sh_degree=D, **kw)
    assert st.num_rendered > 500
    G = rng.normal(size=st.color.shape).astype(np.float32)
    gr = ro.backward(st, G)
    t64 = lambda a: torch.tensor(np.asarray(a, np.float64), requires_grad=True)
    m3, op = t64(sc["means3D"]), t64(sc["opacity"])
    m2 = torch.zeros(600, 3, dtype=torch.float64, requires_grad=True)
    tk = {k: t64(v) for k, v in kw.items()}
    img = r64.render(st, m3, op, means2D=m2, **tk)
    assert np.abs(img.detach().numpy() - st.color).max() < 1e-4          # fp32 forward vs fp64
    (img * torch.from_numpy(G.astype(np.float64))).sum().backward()
    tol = 1e-4
    assert rel_err(gr["means3D"], m3.grad.numpy()) < tol
    assert rel_err(gr["means2D"], m2.grad.numpy()) < tol
    assert rel_err(gr["opacities"], op.grad.numpy()) < tol
    names = {"shs": "sh", "colors_precomp": "colors_precomp", "scales": "scales",
             "rotations": "rotations", "cov3D_precomp": "cov3Ds_precomp"}
    for k, t in tk.items():
        assert rel_err(gr[names[k]], t.grad.numpy()) < tol, k


def test_scale_gradient_has_no_modifier_factor():
    """Upstream quirk kept on purpose: dL/dscale carries no scale_modifier factor,
    so at modifier m it equals the true gradient divided by m."""
    sc = make_scene(N=300, H=48, W=48, seed=5, scale_range=(0.01, 0.04))
    cam = oracle_camera(sc["view"])
    bg = np.zeros(3, np.float32)
    st = ro.forward(cam, sc["means3D"], sc["opacity"], bg, shs=sc["shs"], scales=sc["scales"],
                    rotations=sc["rotations"], sh_degree=0, scale_modifier=1.5)
    G = np.random.default_rng(1).normal(size=st.color.shape).astype(np.float32)
    gr = ro.backward(st, G)
    t64 = lambda a: torch.tensor(np.asarray(a, np.float64), requires_grad=True)
    s_ = t64(sc["scales"])
    img = r64.render(st, t64(sc["means3D"]), t64(sc["opacity"]), shs=t64(sc["shs"]), scales=s_,
                     rotations=t64(sc["rotations"]))
    (img * torch.from_numpy(G.astype(np.float64))).sum().backward()
    assert rel_err(gr["scales"] * 1.5, s_.grad.numpy()) < 1e-4


def test_structure_of_keys_ranges_counts():
    sc = make_scene(N=4000, H=96, W=144, seed=9)
    cam = oracle_camera(sc["view"])
    st = ro.forward(cam, sc["means3D"], sc["opacity"], np.ones(3, np.float32), shs=sc["shs"],
                    scales=sc["scales"], rotations=sc["rotations"], sh_degree=2)
    L = st.num_rendered
    assert L == int(st.tiles_touched.astype(np.int64).sum()) == int(st.offsets[-1])
    # stable sort == numpy stable argsort of the unsorted keys (unique answer of LSD radix)
    order = np.argsort(st.keys_unsorted, kind="stable")
    assert np.array_equal(st.keys, st.keys_unsorted[order])
    assert np.array_equal(st.point_list, st.vals_unsorted[order])
    # ranges partition [0, L) over the non-empty tiles, in tile order
    tiles = (st.keys >> np.uint64(32)).astype(np.int64)
    r = st.ranges.astype(np.int64)
    nz = np.where(r[:, 1] > r[:, 0])[0]
    assert r[nz, 0].min() == 0 and r[nz, 1].max() == L
    assert np.array_equal(np.sort(np.unique(tiles)), nz)
    for t in nz[:50]:
        assert np.all(tiles[r[t, 0]:r[t, 1]] == t)
    assert int((r[:, 1] - r[:, 0]).sum()) == L
    # depth bits ascending inside a tile; contributor counts bounded by the tile list length
    gx = cam.grid[0]
    H, W = cam.H, cam.W
    for t in nz[:50]:
        d = (st.keys[r[t, 0]:r[t, 1]] & np.uint64(0xFFFFFFFF)).astype(np.uint32).view(np.float32)
        assert np.all(np.diff(d) >= 0)
        ty, tx = divmod(int(t), gx)
        blk = st.n_contrib[ty * 16:min(H, ty * 16 + 16), tx * 16:min(W, tx * 16 + 16)]
        assert blk.max() <= r[t, 1] - r[t, 0]


def test_analytic_single_gaussian_and_empty():
    view = syn.make_view(64, 64)
    cam = oracle_camera(view)
    bg = np.array([0.2, 0.3, 0.4], np.float32)
    # empty scene: background passes through
    st = ro.forward(cam, np.zeros((0, 3), np.float32), np.zeros((0, 1), np.float32), bg,
                    colors_precomp=np.zeros((0, 3), np.float32), scales=np.zeros((0, 3), np.float32),
                    rotations=np.zeros((0, 4), np.float32))
    assert st.num_rendered == 0 and np.allclose(st.color, bg[:, None, None])
    # one isotropic Gaussian on the optical axis: pixel centre (31.5, 31.5)
    z, s, o = 4.0, 0.05, 0.7
    col = np.array([[0.9, 0.1, 0.5]], np.float32)
    st = ro.forward(cam, np.array([[0, 0, z]], np.float32), np.array([[o]], np.float32), bg,
                    colors_precomp=col, scales=np.full((1, 3), s, np.float32),
                    rotations=np.array([[1, 0, 0, 0]], np.float32))
    f = 64 / (2 * view.tanfovx)
    var = (f * s / z) ** 2 + 0.3
    assert abs(st.xy[0, 0] - 31.5) < 1e-4 and abs(st.xy[0, 1] - 31.5) < 1e-4
    assert st.radii[0] == math.ceil(3 * math.sqrt(var))
    a = o * math.exp(-0.5 * (0.5 ** 2 + 0.5 ** 2) / var)          # pixel (31,31) is 0.5 px off-centre
    want = a * col[0] + (1 - a) * bg
    assert np.abs(st.color[:, 31, 31] - want).max() < 1e-5
    assert st.n_contrib[31, 31] == 1 and abs(st.alpha[31, 31] - a) < 1e-6
    assert abs(st.depth[31, 31] - a * z) < 1e-5
    # behind the near plane (z <= 0.2): culled
    st = ro.forward(cam, np.array([[0, 0, 0.2]], np.float32), np.array([[o]], np.float32), bg,
                    colors_precomp=col, scales=np.full((1, 3), s, np.float32),
                    rotations=np.array([[1, 0, 0, 0]], np.float32))
    assert st.radii[0] == 0 and st.num_rendered == 0
    assert not ro.mark_visible(np.array([[0, 0, 0.2]], np.float32), cam.view)[0]
    assert ro.mark_visible(np.array([[0, 0, 0.21]], np.float32), cam.view)[0]


def test_two_overlapping_gaussians_blend_front_to_back():
    view = syn.make_view(32, 32)
    cam = oracle_camera(view)
    bg = np.zeros(3, np.float32)
    pos = np.array([[0, 0, 6.0], [0, 0, 3.0]], np.float32)      # second one is in front
    col = np.array([[1, 0, 0], [0, 1, 0]], np.float32)
    op = np.array([[0.6], [0.5]], np.float32)
    st = ro.forward(cam, pos, op, bg, colors_precomp=col, scales=np.full((2, 3), 0.2, np.float32),
                    rotations=np.tile(np.array([1, 0, 0, 0], np.float32), (2, 1)))
    f = 32 / (2 * view.tanfovx)
    a_front = 0.5 * math.exp(-0.5 * 0.5 / ((f * 0.2 / 3.0) ** 2 + 0.3))
    a_back = 0.6 * math.exp(-0.5 * 0.5 / ((f * 0.2 / 6.0) ** 2 + 0.3))
    want = np.array([a_back * (1 - a_front), a_front, 0.0])
    assert np.abs(st.color[:, 15, 15] - want).max() < 1e-5
    assert st.n_contrib[15, 15] == 2
    assert list(st.point_list[st.ranges[0, 0]:st.ranges[0, 0] + 2]) == [1, 0]
