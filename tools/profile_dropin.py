"""Host-side profile of the drop-in autograd path (diagnostic): where does the Python time go?"""
import cProfile, pstats, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer
from sings_b200 import deform, rasterizer as R, synthetic as syn

dev = torch.device("cuda", 0)
N, H, W, J, D = 200_000, 1024, 1024, 24, 3
av = syn.make_avatar(N, J, seed=0)
t = lambda a: torch.as_tensor(a, device=dev)
p = dict(xyz=t(av.xyz_canon).requires_grad_(True), rot=t(av.rotmat_canon).requires_grad_(True),
         scales=t(av.scales).requires_grad_(True), opacity=t(av.opacity).requires_grad_(True),
         shs=t(av.shs).requires_grad_(True), W=t(av.lbs_weights), rest=t(av.rest),
         parents=torch.from_numpy(av.parents).to(device=dev, dtype=torch.int32), inv_A=t(av.inv_A_t2cano))
pose0, tr0 = t(syn.random_pose(J, seed=2)), t(syn.default_transl(H))
v = syn.make_view(H, W)
vm, pm, cp, bg = t(v.world_view_transform), t(v.full_proj_transform), t(v.camera_center), t(np.ones(3, np.float32))
G = torch.randn(3, H, W, device=dev)


def step():
    pose = pose0.detach().requires_grad_(True)
    transl = tr0.detach().requires_grad_(True)
    A = deform.pose_to_A(pose, p["rest"], p["parents"], p["inv_A"])
    xyz, rotq, sc = deform.deform_gaussians(A, p["xyz"], p["W"], p["rot"], p["scales"], None, transl)
    rs = GaussianRasterizationSettings(image_height=H, image_width=W, tanfovx=v.tanfovx, tanfovy=v.tanfovy, bg=bg,
                                       scale_modifier=1.0, viewmatrix=vm, projmatrix=pm, sh_degree=D, campos=cp,
                                       prefiltered=False, debug=False)
    means2D = torch.zeros_like(xyz, requires_grad=True)
    img, radii = GaussianRasterizer(rs)(means3D=xyz, means2D=means2D, shs=p["shs"], opacities=p["opacity"], scales=sc, rotations=rotq)
    loss = (img * G).sum()
    for q in ("xyz", "rot", "scales", "opacity", "shs"):
        p[q].grad = None
    loss.backward()


for _ in range(5):
    step()
R.set_async(True)
for _ in range(20):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(200):
    step()
t_host = time.perf_counter() - t0
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
print(f"host time per step {t_host / 200 * 1e3:.3f} ms, with device drain {t_all / 200 * 1e3:.3f} ms")
pr = cProfile.Profile()
pr.enable()
for _ in range(200):
    step()
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
