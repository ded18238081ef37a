"""First GPU contact: run every kernel once against the oracle and print the differences
(diagnostic script; the real assertions live in tests/)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from helpers import make_scene, oracle_camera, raster_settings, rel_err, inspect_state
from oracle import raster_oracle as ro, lbs_oracle as lo
from diff_gaussian_rasterization import GaussianRasterizer
from sings_b200 import rasterizer as R, deform as Dm, _lib

print("lib version", _lib.lib().sgs_version(), torch.cuda.get_device_name(0))
dev = "cuda"
for (N, H, W, D) in [(3000, 128, 160, 3), (20000, 256, 256, 0), (1000, 70, 45, 2)]:
    sc = make_scene(N=N, H=H, W=W, seed=N)
    cam = oracle_camera(sc["view"])
    bg = np.array([0.2, 0.4, 0.6], np.float32)
    st = ro.forward(cam, sc["means3D"], sc["opacity"], bg, shs=sc["shs"], scales=sc["scales"], rotations=sc["rotations"], sh_degree=D)
    t = lambda a, rg=False: torch.tensor(a, device=dev, requires_grad=rg)
    m3, op, shs, s_, r_ = t(sc["means3D"], True), t(sc["opacity"], True), t(sc["shs"], True), t(sc["scales"], True), t(sc["rotations"], True)
    m2 = torch.zeros_like(m3, requires_grad=True)
    rs = raster_settings(sc["view"], bg, D, debug=True)
    rast = GaussianRasterizer(rs)
    color, radii, alpha, depth = rast.forward_aux(m3, m2, op, shs=shs, scales=s_, rotations=r_)
    torch.cuda.synchronize()
    fn = color.grad_fn
    saved = fn.saved_tensors
    geom, binning, img = saved[7], saved[8], saved[9]
    ins = inspect_state((geom, binning, img), N, W, H, fn.L_cap)
    print(f"--- N={N} {W}x{H} D={D}: L oracle {st.num_rendered} gpu {ins['num_rendered']} ovf {ins['overflow']}")
    print(" radii equal", np.array_equal(radii.cpu().numpy(), st.radii))
    L = st.num_rendered
    if ins["num_rendered"] == L:
        print(" keys_unsorted equal", np.array_equal(ins["keys_unsorted"], st.keys_unsorted), "vals", np.array_equal(ins["vals_unsorted"], st.vals_unsorted))
        print(" keys sorted equal", np.array_equal(ins["keys"], st.keys), "point_list", np.array_equal(ins["point_list"], st.point_list))
        print(" ranges equal", np.array_equal(ins["ranges"], st.ranges))
    print(" n_contrib equal", np.array_equal(ins["n_contrib"], st.n_contrib), "mismatch", int((ins["n_contrib"] != st.n_contrib).sum()))
    print(" final_T maxabs", np.abs(ins["final_T"] - st.final_T).max(), " bit-equal", np.array_equal(ins["final_T"], st.final_T))
    print(" color maxabs", np.abs(color.detach().cpu().numpy() - st.color).max(), "bit-equal", np.array_equal(color.detach().cpu().numpy(), st.color))
    print(" alpha maxabs", np.abs(alpha.cpu().numpy() - st.alpha).max(), "depth maxabs", np.abs(depth.cpu().numpy() - st.depth).max())
    G = np.random.default_rng(1).normal(size=st.color.shape).astype(np.float32)
    (color * t(G)).sum().backward()
    torch.cuda.synchronize()
    gr = ro.backward(st, G)
    for name, g in [("means3D", m3.grad), ("means2D", m2.grad), ("opacities", op.grad), ("sh", shs.grad), ("scales", s_.grad), ("rotations", r_.grad)]:
        print(f" grad {name}: rel {rel_err(g.cpu().numpy(), gr[name]):.3e}")

# ---- LBS ----
import glob
for f in sorted(glob.glob(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "lbs_golden_*_f32.npz"))):
    g = {k: torch.from_numpy(v) for k, v in np.load(f).items()}
    c = lambda k, rg=False: g[k].to(dev).requires_grad_(rg) if k in g else None
    iso = bool(g["isotropic"])
    pose = c("pose", True)
    A = Dm.pose_to_A(pose, c("rest"), g["parents"], c("inv_A_t2cano"))
    print(os.path.basename(f), "A maxabs", (A.detach().cpu() - g["A_cano2pose"]).abs().max().item())
    ext = (c("ext_trans"), c("ext_rotmat"), c("ext_scale")) if "ext_trans" in g else None
    xyz_c, sc_c, rc, ss, tr = c("xyz_canon", True), c("scales", True), (None if iso else c("rotmat_canon", True)), c("smpl_scale", True), c("transl", True)
    xyz, q, sco = Dm.deform_gaussians(A, xyz_c, c("lbs_weights"), rc, sc_c, ss, tr, ext)
    print("  xyz", (xyz.detach().cpu() - g["xyz"]).abs().max().item(), "q", (q.detach().cpu() - g["rotq"]).abs().max().item(), "sc", (sco.detach().cpu() - g["scales_out"]).abs().max().item())
    loss = (xyz * c("gx")).sum() + (q * c("gq")).sum() + (sco * c("gs")).sum()
    loss.backward()
    print("  d_pose", rel_err(pose.grad.cpu(), g["d_pose"]), "d_xyz", rel_err(xyz_c.grad.cpu(), g["d_xyz_canon"]), "d_scales", rel_err(sc_c.grad.cpu(), g["d_scales"]),
          "d_rot", (rel_err(rc.grad.cpu(), g["d_rotmat_canon"]) if rc is not None else None), "d_ss", rel_err(ss.grad.cpu(), g["d_smpl_scale"]), "d_tr", rel_err(tr.grad.cpu(), g["d_transl"]))

# ---- standalone sort ----
n = 1_000_003
keys = torch.randint(0, 2**45, (n,), device=dev, dtype=torch.int64)
vals = torch.arange(n, device=dev, dtype=torch.int32)
kt, vt = torch.empty_like(keys), torch.empty_like(vals)
sb = _lib.lib().sgs_sort_scratch_bytes(n)
scratch = torch.empty(sb, device=dev, dtype=torch.uint8)
import ctypes
flag = ctypes.c_int(0)
k0 = keys.clone()
_lib.check(_lib.lib().sgs_sort_pairs_u64(keys.data_ptr(), vals.data_ptr(), kt.data_ptr(), vt.data_ptr(), scratch.data_ptr(), sb, n, 45, ctypes.byref(flag), torch.cuda.current_stream().cuda_stream))
torch.cuda.synchronize()
rk, rv = (kt, vt) if flag.value else (keys, vals)
sk, si = torch.sort(k0, stable=True)
print("sort keys equal", torch.equal(rk, sk), "vals equal", torch.equal(rv.long(), si))
print("DONE")
