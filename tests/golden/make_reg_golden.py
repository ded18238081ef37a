"""Generate golden vectors for the non-image loss terms by running the REFERENCE's own classes on CPU.

Run in the build container only (needs /root/reference):
    python tests/golden/make_reg_golden.py
Writes tests/golden/reg_golden_<case>.npz.  Executed reference code:
  L2Norm, GaussiansEdgeLoss, RegionLaplacianLoss_v2 (reset_laplacians, forward, forward_hands), pcd_laplacian_smoothing
                                      /root/reference/sings/rec/losses/loss_items.py:15-54, 57-90, 93-190, 205-214
  parse_weights                       /root/reference/sings/rec/utils/body_model/smpl_parsing.py:38-44
                                      (reads /root/reference/data/human_models/smpl_parsing/*.json)
loss_items.py imports pytorch3d.ops at module level; pytorch3d is not installed (and not vendored:
install_all.sh:21 pulls its default branch).  Stand-ins registered for it:
  laplacian   the published algorithm of pytorch3d/ops/laplacian_matrices.py::laplacian as a torch sparse
              COO tensor (values computed in float32 as there; cast to verts.dtype so that the float64
              runs, which give the gradient truth, can multiply it)
  knn_points  brute force: exact K nearest neighbours by squared distance, ascending (reached by GaussiansEdgeLoss)
Everything the reference's classes do around that function -- region selection, renumbering, weights,
means, the calls' autograd -- is the reference's own code.
"""
import collections
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

POSITION_REGIONS_W = {'head-neck': 0.5, 'spine': 0.75, 'leftUpArm': 1., 'rightUpArm': 1., 'leftDownArm': 1.,
                      'rightDownArm': 1., 'leftHand': 1.5, 'rightHand': 1.5, 'hips': 1., 'leftUpLeg': 1.,
                      'rightUpLeg': 1., 'leftDownLeg': 1., 'rightDownLeg': 1., 'leftFoot': 0.75,
                      'rightFoot': 0.75}                       # cfgs/train/beta/human_complex.yaml:141-142
COLOR_REGIONS_W = {'head-neck': 0., 'spine': 0., 'leftUpArm': 0., 'rightUpArm': 0., 'leftDownArm': 1.,
                   'rightDownArm': 1., 'leftHand': 1., 'rightHand': 1., 'hips': 0., 'leftUpLeg': 0.,
                   'rightUpLeg': 0., 'leftDownLeg': 0., 'rightDownLeg': 0., 'leftFoot': 0.,
                   'rightFoot': 0.}                            # human_complex.yaml:137-138
L2_CFG = dict(lambda_xyz_offsets=0.001, lambda_scales_diff=0.005, max_scale_threshold=0.005, lambda_max_scale=0.01,
              min_opacity_threshold=0.2, lambda_min_opacity=0.001)          # human_complex.yaml:148-154


def pytorch3d_laplacian(verts, edges):
    V = verts.shape[0]
    e0, e1 = edges.unbind(1)
    idx01 = torch.stack([e0, e1], dim=1)
    idx10 = torch.stack([e1, e0], dim=1)
    idx = torch.cat([idx01, idx10], dim=0).t()
    ones = torch.ones(idx.shape[1], dtype=torch.float32)
    A = torch.sparse_coo_tensor(idx, ones, (V, V))
    deg = torch.sparse.sum(A, dim=1).to_dense()
    deg0 = deg[e0]
    deg0 = torch.where(deg0 > 0.0, 1.0 / deg0, deg0)
    deg1 = deg[e1]
    deg1 = torch.where(deg1 > 0.0, 1.0 / deg1, deg1)
    val = torch.cat([deg0, deg1])
    L = torch.sparse_coo_tensor(idx, val, (V, V))
    idx = torch.arange(V)
    idx = torch.stack([idx, idx], dim=0)
    ones = torch.ones(idx.shape[1], dtype=torch.float32)
    L = L - torch.sparse_coo_tensor(idx, ones, (V, V))
    return L.coalesce().to(verts.dtype)


_KNN = collections.namedtuple("KNN", "dists idx knn")          # pytorch3d returns this named 3-tuple


def brute_knn_points(p1, p2, K):
    d = torch.cdist(p1[0].double(), p2[0].double()) ** 2
    dists, idx = torch.topk(d, K, dim=1, largest=False, sorted=True)
    return _KNN(dists=dists[None].to(p1.dtype), idx=idx[None], knn=None)


def load_reference_loss_items():
    for name in ("pytorch3d", "pytorch3d.ops", "sings", "sings.rec", "sings.rec.utils", "sings.rec.utils.body_model"):
        sys.modules.setdefault(name, types.ModuleType(name))
    ops = sys.modules["pytorch3d.ops"]
    ops.knn_points = brute_knn_points
    ops.laplacian = pytorch3d_laplacian
    ops.cot_laplacian = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError("cot_laplacian"))
    ops.norm_laplacian = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError("norm_laplacian"))
    cwd = os.getcwd()
    os.chdir(REF)                                   # smpl_parsing.py opens its json tables relative to the repo root
    try:
        spec = importlib.util.spec_from_file_location("sings.rec.utils.body_model.smpl_parsing",
                                                      f"{REF}/sings/rec/utils/body_model/smpl_parsing.py")
        sp = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(sp)
        sys.modules["sings.rec.utils.body_model.smpl_parsing"] = sp
        spec = importlib.util.spec_from_file_location("ref_loss_items", f"{REF}/sings/rec/losses/loss_items.py")
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        os.chdir(cwd)
    return mod


def body_like_graph(V, R, seed, orphan=False):
    """V points on a wobbly tube cut into R label bands with ragged borders; unique undirected edges to a few
    nearest points (some cross a border and are dropped by the region selection).  orphan=True relabels
    two vertices so that they keep no edge inside their region (the renumbering quirk of
    reset_laplacians, see sings_b200/regularizers.py::region_laplacian)."""
    g = torch.Generator().manual_seed(seed)
    t = torch.sort(torch.rand(V, generator=g)).values
    ang = torch.rand(V, generator=g) * 6.2831853
    verts = torch.stack([0.1 * torch.cos(ang), 1.7 * t, 0.1 * torch.sin(ang)], 1) + 0.004 * torch.randn(V, 3, generator=g)
    labels = torch.clamp(((t + 0.01 * torch.randn(V, generator=g)) * R).floor().long(), 0, R - 1)
    d = torch.cdist(verts, verts)
    nbr = torch.topk(d, 5, dim=1, largest=False).indices[:, 1:]
    e = torch.stack([torch.arange(V).unsqueeze(1).expand(-1, 4).reshape(-1), nbr.reshape(-1)], 1)
    e = torch.sort(e, dim=1).values
    e = torch.unique(e, dim=0)
    if orphan:
        # a vertex deep inside band 3 relabelled 9, one inside band 11 relabelled 2: no same-label neighbour
        for src_band, new in ((3, 9), (11, 2)):
            cand = torch.nonzero((labels == src_band) & (torch.abs(t * R - src_band - 0.5) < 0.2)).reshape(-1)
            labels[cand[0]] = new
    return verts, e, labels


def main():
    M = load_reference_loss_items()

    # ---------------- region Laplacians
    for name, V, seed, orphan in (("region_a", 420, 11, False), ("region_b", 333, 12, True)):
        verts, edges, labels = body_like_graph(V, 15, seed, orphan)
        g = torch.Generator().manual_seed(seed + 100)
        xyz = verts + 0.01 * torch.randn(V, 3, generator=g)               # 'xyz_anchor_canon'
        shs = 0.5 * torch.randn(V, 16, 3, generator=g)                    # colours = shs[:, 0]
        out = dict(verts=verts.numpy(), edges=edges.numpy(), labels=labels.numpy(), xyz=xyz.numpy(), shs=shs.numpy())
        for dt, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
            pos = M.RegionLaplacianLoss_v2(verts=verts.to(dt), edges=edges, vertex_labels=labels,
                                           region_weights=POSITION_REGIONS_W)       # gs_trainer.py:177-184
            col = M.RegionLaplacianLoss_v2(verts=verts.to(dt), edges=edges, vertex_labels=labels,
                                           region_weights=COLOR_REGIONS_W)          # gs_trainer.py:187-192
            x = xyz.to(dt).clone().requires_grad_(True)
            s = shs.to(dt).clone().requires_grad_(True)
            l_pos = pos(x)                                                # gs_trainer.py:372
            l_col = col(s[:, 0])                                          # gs_trainer.py:373
            l_hand = pos.forward_hands(x)                                 # gs_trainer.py:395
            g_pos, = torch.autograd.grad(l_pos, x, retain_graph=True)
            g_hand, = torch.autograd.grad(l_hand, x)
            g_col, = torch.autograd.grad(l_col, s)
            out.update({f"loss_pos_{tag}": float(l_pos), f"loss_col_{tag}": float(l_col), f"loss_hand_{tag}": float(l_hand),
                        f"grad_pos_{tag}": g_pos.numpy(), f"grad_hand_{tag}": g_hand.numpy(), f"grad_col_{tag}": g_col.numpy()})
            if tag == "f32":
                # the reference's per-region operators, scattered to global vertex numbering the way forward()
                # applies them: row / column k of region i acts on the k-th vertex of x[labels == i]
                D = torch.zeros(V, V)
                for i, (L, part) in enumerate(zip(pos.laplacians, pos.vertex_partitions)):
                    ids = torch.nonzero(part).reshape(-1)
                    Ld = L.to_dense()
                    D[ids.unsqueeze(1), ids.unsqueeze(0)] = Ld
                out["L_dense"] = D.numpy()
                out["weights_pos"] = np.asarray(pos.weights, np.float64)
                out["weights_col"] = np.asarray(col.weights, np.float64)
        np.savez_compressed(os.path.join(HERE, f"reg_golden_{name}.npz"), **out)
        print(name, {k: out[k] for k in out if k.startswith("loss_")})

    # ---------------- point-cloud Laplacian smoothing (edges given: directed K-NN lists with repeats both ways)
    N, K = 300, 6
    g = torch.Generator().manual_seed(21)
    pts = torch.rand(N, 3, generator=g)
    d = torch.cdist(pts, pts)
    nbr = torch.topk(d, K + 1, dim=1, largest=False).indices[:, 1:]
    edges = torch.cat([torch.arange(N).unsqueeze(1).repeat(1, K).reshape(-1, 1), nbr.reshape(-1, 1)], dim=1)   # = build_edges
    pts[7] = (pts[nbr[7]].mean(0))                                        # a row whose Laplacian is ~0
    out = dict(pts=pts.numpy(), edges=edges.numpy())
    for dt, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
        x = pts.to(dt).clone().requires_grad_(True)
        l = M.pcd_laplacian_smoothing(x, edges)
        gx, = torch.autograd.grad(l, x)
        out.update({f"loss_{tag}": float(l), f"grad_{tag}": gx.numpy()})
    np.savez_compressed(os.path.join(HERE, "reg_golden_pcd.npz"), **out)
    print("pcd", out["loss_f32"], out["loss_f64"])

    # ---------------- scale-edge loss (GaussiansEdgeLoss, loss_items.py:57-90; gs_trainer.py:194, 367) with the brute-force
    # knn_points stand-in: everything the class does around the neighbour search is the reference's own code
    N = 400
    g = torch.Generator().manual_seed(41)
    pts = torch.rand(N, 3, generator=g) * torch.tensor([0.6, 1.7, 0.3])
    sc = (0.002 + 0.01 * torch.rand(N, 1, generator=g)).repeat(1, 3)
    out = dict(xyz_canon=pts.numpy(), scales=sc.numpy())
    edge = M.GaussiansEdgeLoss()                                         # K = 9 (the point itself + 8 neighbours)
    for dt, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
        x = pts.to(dt).clone().requires_grad_(True)
        s = sc.to(dt).clone().requires_grad_(True)
        l = edge({"xyz_canon": x, "scales": s})
        gs, gx = torch.autograd.grad(l, [s, x], allow_unused=True)
        assert gx is None or float(gx.abs().max()) == 0.0                # the edge lengths are detached (:79)
        out.update({f"loss_{tag}": float(l), f"grad_scales_{tag}": gs.numpy()})
    np.savez_compressed(os.path.join(HERE, "reg_golden_edge.npz"), **out)
    print("edge", out["loss_f32"], out["loss_f64"])

    # ---------------- L2Norm
    cases = [dict(name="l2_a", N=500, cfg=L2_CFG, opacity=True, seed=31, big=True),
             dict(name="l2_b", N=257, cfg=dict(), opacity=False, seed=32, big=True),          # constructor defaults, no opacity key
             dict(name="l2_c", N=64, cfg=L2_CFG, opacity=True, seed=33, big=False)]           # nothing above / below the thresholds
    for c in cases:
        g = torch.Generator().manual_seed(c["seed"])
        N = c["N"]
        off = 0.01 * torch.randn(N, 3, generator=g)
        if c["big"]:
            sc = torch.exp(torch.log(torch.tensor(0.002)) + torch.rand(N, 1, generator=g) * 1.8).repeat(1, 3)
            op = torch.sigmoid(1.5 * torch.randn(N, 1, generator=g))
        else:
            sc = (0.001 + 0.003 * torch.rand(N, 1, generator=g)).repeat(1, 3)
            op = 0.3 + 0.6 * torch.rand(N, 1, generator=g)
        out = dict(xyz_offsets=off.numpy(), scales=sc.numpy(), opacity=op.numpy(), has_opacity=c["opacity"],
                   cfg=np.asarray([c["cfg"].get(k, d) for k, d in (("lambda_xyz_offsets", 0.005), ("lambda_scales_diff", 0.005),
                                                                   ("lambda_max_scale", 0.001), ("max_scale_threshold", 0.008),
                                                                   ("lambda_min_opacity", 0.0001), ("min_opacity_threshold", 0.2))],
                                  np.float64))
        mod = M.L2Norm(**c["cfg"])
        for dt, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
            o = off.to(dt).clone().requires_grad_(True)
            s = sc.to(dt).clone().requires_grad_(True)
            p = op.to(dt).clone().requires_grad_(True)
            d_in = {"xyz_offsets": o, "scales": s}
            if c["opacity"]:
                d_in["opacity"] = p
            l = mod(d_in)
            grads = torch.autograd.grad(l, [o, s] + ([p] if c["opacity"] else []), allow_unused=True)
            out.update({f"loss_{tag}": float(l), f"grad_off_{tag}": grads[0].numpy(), f"grad_scales_{tag}": grads[1].numpy()})
            if c["opacity"]:
                out[f"grad_opacity_{tag}"] = (torch.zeros_like(p) if grads[2] is None else grads[2]).numpy()
        np.savez_compressed(os.path.join(HERE, f"reg_golden_{c['name']}.npz"), **out)
        print(c["name"], out["loss_f32"], out["loss_f64"])


if __name__ == "__main__":
    main()
