"""GPU: drop-in integration.  The reference's renderer shim
(/root/reference/sings/rec/renderer/gs_renderer_single.py:12-107) is restated here call for
call (the GPU box has no /root/reference) and driven with the same `data` / `human_gs_out`
dicts SinGS builds; plus AvatarStep (the fused fast path) against the autograd path."""
import math

import numpy as np
import pytest
import torch

from helpers import make_scene, oracle_camera, rel_err
from oracle import raster_oracle as ro

pytestmark = pytest.mark.gpu


def reference_style_render(means3D, feats, opacity, scales, rotations, data, scaling_modifier=1.0,
                           bg_color=None, active_sh_degree=0):
    """The body of the reference's render(): same statements, same keyword names."""
    from diff_gaussian_rasterization import GaussianRasterizationSettings, GaussianRasterizer
    if bg_color is None:
        bg_color = torch.zeros(3, dtype=torch.float32, device="cuda")
    screenspace_points = torch.zeros_like(means3D, dtype=means3D.dtype, requires_grad=True, device="cuda") + 0
    try:
        screenspace_points.retain_grad()
    except Exception:
        pass
    means2D = screenspace_points
    tanfovx = math.tan(data['fovx'] * 0.5)
    tanfovy = math.tan(data['fovy'] * 0.5)
    shs, rgb = None, None
    if len(feats.shape) == 2:
        rgb = feats
    else:
        shs = feats
    raster_settings = GaussianRasterizationSettings(
        image_height=int(data['image_height']), image_width=int(data['image_width']), tanfovx=tanfovx,
        tanfovy=tanfovy, bg=bg_color, scale_modifier=scaling_modifier, viewmatrix=data['world_view_transform'],
        projmatrix=data['full_proj_transform'], sh_degree=active_sh_degree, campos=data['camera_center'],
        prefiltered=False, debug=False)
    rasterizer = GaussianRasterizer(raster_settings=raster_settings)
    rendered_image, radii = rasterizer(means3D=means3D, means2D=means2D, shs=shs, opacities=opacity,
                                       scales=scales, rotations=rotations, colors_precomp=rgb)
    rendered_image = torch.clamp(rendered_image, 0.0, 1.0)
    return {"render": rendered_image, "viewspace_points": screenspace_points,
            "visibility_filter": radii > 0, "radii": radii}


def test_renderer_shim_flow_and_densification_stats():
    sc = make_scene(N=4000, H=160, W=96, seed=12)
    view = sc["view"]
    dev = "cuda"
    t = lambda a, rg=False: torch.tensor(np.ascontiguousarray(a), device=dev, requires_grad=rg)
    data = dict(fovx=view.fovx, fovy=view.fovy, image_height=view.image_height, image_width=view.image_width,
                world_view_transform=t(view.world_view_transform), full_proj_transform=t(view.full_proj_transform),
                camera_center=t(view.camera_center))
    gs = dict(xyz=t(sc["means3D"], True), shs=t(sc["shs"], True), opacity=t(sc["opacity"], True),
              scales=t(sc["scales"], True), rotq=t(sc["rotations"], True), active_sh_degree=2)
    bg = torch.rand(3, device=dev)
    pkg = reference_style_render(gs["xyz"], gs["shs"], gs["opacity"], gs["scales"], gs["rotq"], data,
                                 bg_color=bg, active_sh_degree=gs["active_sh_degree"])
    st = ro.forward(oracle_camera(view), sc["means3D"], sc["opacity"], bg.cpu().numpy(), shs=sc["shs"],
                    scales=sc["scales"], rotations=sc["rotations"], sh_degree=2)
    assert np.abs(pkg["render"].detach().cpu().numpy() - np.clip(st.color, 0, 1)).max() <= 1e-4
    assert np.array_equal(pkg["visibility_filter"].cpu().numpy(), st.radii > 0)
    G = torch.randn(3, view.image_height, view.image_width, device=dev)
    (pkg["render"] * G).sum().backward()
    clip_mask = ((st.color > 0) & (st.color < 1)).astype(np.float32)
    gr = ro.backward(st, G.cpu().numpy() * clip_mask)
    vs = pkg["viewspace_points"]
    assert vs.grad is not None and rel_err(vs.grad.cpu().numpy(), gr["means2D"]) < 1e-3
    # densification statistics (sings_hybrid.py:1013-1015, gs_trainer.py:487-490) via the fused kernel
    from sings_b200 import _lib
    N = 4000
    accum = torch.zeros(N, device=dev); denom = torch.zeros(N, device=dev); maxr = torch.full((N,), 3.0, device=dev)
    _lib.check(_lib.lib().sgs_densify_stats(N, vs.grad.data_ptr(), pkg["radii"].data_ptr(), accum.data_ptr(),
                                            denom.data_ptr(), maxr.data_ptr(),
                                            torch.cuda.current_stream().cuda_stream))
    vis = pkg["visibility_filter"]
    exp_acc = torch.zeros(N, device=dev)
    exp_acc[vis] += torch.norm(vs.grad[vis, :2], dim=-1)
    exp_max = torch.full((N,), 3.0, device=dev)
    exp_max[vis] = torch.max(exp_max[vis], pkg["radii"][vis].float())
    assert torch.allclose(accum, exp_acc, rtol=1e-6, atol=1e-12)
    assert torch.equal(denom, vis.float()) and torch.equal(maxr, exp_max)


def test_avatar_step_matches_autograd_path():
    """AvatarStep (preallocated, sync-free C-ABI sequence) == deform + rasterizer autograd."""
    from diff_gaussian_rasterization import GaussianRasterizer
    from helpers import raster_settings
    from sings_b200 import deform
    from sings_b200.step import AvatarStep, FrameInputs
    sc = make_scene(N=6000, H=128, W=128, seed=15)
    av, view = sc["avatar"], sc["view"]
    dev = "cuda"
    t = lambda a, rg=False: torch.tensor(np.ascontiguousarray(a), device=dev, requires_grad=rg)
    bg = np.array([1.0, 1.0, 1.0], np.float32)
    step = AvatarStep(t(av.xyz_canon), t(av.rotmat_canon), t(av.scales), t(av.opacity), t(av.shs),
                      t(av.lbs_weights), t(av.rest), torch.from_numpy(av.parents), t(av.inv_A_t2cano),
                      128, 128, 3, timing=True)
    fr = FrameInputs(pose=t(sc["pose"]), transl=t(sc["transl"]), viewmatrix=t(view.world_view_transform),
                     projmatrix=t(view.full_proj_transform), campos=t(view.camera_center), bg=t(bg),
                     tanfovx=view.tanfovx, tanfovy=view.tanfovy)
    G = torch.randn(3, 128, 128, device=dev)
    img = step.forward(fr).clone()
    step.backward(G)
    torch.cuda.synchronize()
    assert step.check_capacity() > 0
    ms = step.stage_ms()
    assert ms["total"] > 0 and ms["blend_fwd"] > 0
    # autograd path
    pose = t(sc["pose"], True)
    xyz_c, rot_c, sc_c = t(av.xyz_canon, True), t(av.rotmat_canon, True), t(av.scales, True)
    opa, shs, tr = t(av.opacity, True), t(av.shs, True), t(sc["transl"], True)
    A = deform.pose_to_A(pose, t(av.rest), torch.from_numpy(av.parents), t(av.inv_A_t2cano))
    xyz, q, s = deform.deform_gaussians(A, xyz_c, t(av.lbs_weights), rot_c, sc_c, None, tr)
    m2 = torch.zeros_like(xyz, requires_grad=True)
    img2, radii = GaussianRasterizer(raster_settings(view, bg, 3))(means3D=xyz, means2D=m2, shs=shs,
                                                                   opacities=opa, scales=s, rotations=q)
    (img2 * G).sum().backward()
    assert torch.equal(img, img2.detach())
    assert torch.equal(step.radii, radii)
    for a, b, name in [(step.d_xyz_canon, xyz_c.grad, "xyz"), (step.d_scales, sc_c.grad, "scales"),
                       (step.d_rot_canon, rot_c.grad, "rot"), (step.d_opacity, opa.grad, "opacity"),
                       (step.d_shs, shs.grad, "shs"), (step.d_pose[0], pose.grad, "pose"),
                       (step.d_transl[0], tr.grad, "transl")]:
        assert rel_err(a.cpu().numpy(), b.cpu().numpy()) < 1e-4, name
    vis = radii > 0
    assert torch.equal(step.denom, vis.float())
    assert torch.allclose(step.grad_accum[vis], torch.norm(m2.grad[vis, :2], dim=-1), rtol=1e-4, atol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("J,iso,smooth", [(24, False, 0), (52, True, 2), (52, False, 1)])
def test_fused_deform_kernels_equal_the_separate_ones(J, iso, smooth, monkeypatch):
    """sgs_avatar_forward / _backward (LBS fused into the rasterizer's per-Gaussian kernels, packed
    skinning weights with K = 4 .. 16 slots) against the separate kernels on the same inputs:
    same image bits, same radii, gradients to fp32 re-association."""
    from sings_b200 import synthetic as syn
    from sings_b200.step import AvatarStep, FrameInputs
    H, W, N = 144, 112, 5000
    av = syn.make_avatar(N, J, seed=31, isotropic=iso, smooth_weights=smooth)
    view = syn.make_view(H, W)
    t = lambda a: torch.tensor(np.ascontiguousarray(a), device="cuda")
    fr = FrameInputs(pose=t(syn.random_pose(J, seed=33)), transl=t(syn.default_transl(H, focal=5000.0 * H / 896.0)),
                     viewmatrix=t(view.world_view_transform), projmatrix=t(view.full_proj_transform),
                     campos=t(view.camera_center), bg=t(np.array([0.3, 0.6, 0.9], np.float32)),
                     tanfovx=view.tanfovx, tanfovy=view.tanfovy, smpl_scale=t(np.array([1.07], np.float32)))
    G = torch.randn(3, H, W, device="cuda", generator=torch.Generator("cuda").manual_seed(5))
    out = {}
    for mode in ("fused", "separate"):
        if mode == "separate":
            monkeypatch.setenv("SGS_NO_FUSE", "1")
        st = AvatarStep(t(av.xyz_canon), None if iso else t(av.rotmat_canon), t(av.scales), t(av.opacity), t(av.shs),
                        t(av.lbs_weights), t(av.rest), torch.from_numpy(av.parents), t(av.inv_A_t2cano), H, W, 3)
        assert (st.K > 0) == (mode == "fused")
        if mode == "fused":
            nnz = int((av.lbs_weights != 0).sum(1).max())
            assert st.K >= nnz and st.K - nnz < 4 and (smooth == 0 or st.K > 4)
        img = st.forward(fr).clone()
        st.backward(G)
        torch.cuda.synchronize()
        assert st.check_capacity() > 0
        out[mode] = dict(img=img, radii=st.radii.clone(), xyz=st.xyz.clone(), q=st.rotq.clone(), bucket=st.bucket.clone(),
                         d_pose=st.d_pose.clone(), d_transl=st.d_transl.clone(), m2=st.g_means2D.clone())
    for name in ("fused",):
        a, b = out[name], out["separate"]
        assert torch.equal(a["xyz"], b["xyz"]) and torch.equal(a["q"], b["q"])
        assert torch.equal(a["img"], b["img"]) and torch.equal(a["radii"], b["radii"])
        for k in ("bucket", "d_pose", "d_transl", "m2"):
            assert rel_err(a[k].cpu().numpy(), b[k].cpu().numpy()) < 2e-5, (name, k)


def _avatar_step(sc, H, W, D=3, timing=False):
    from sings_b200.step import AvatarStep
    av = sc["avatar"]
    t = lambda a: torch.tensor(np.ascontiguousarray(a), device="cuda")
    return AvatarStep(t(av.xyz_canon), t(av.rotmat_canon), t(av.scales), t(av.opacity), t(av.shs),
                      t(av.lbs_weights), t(av.rest), torch.from_numpy(av.parents), t(av.inv_A_t2cano),
                      H, W, D, timing=timing)


def _frame(sc, pose=None, transl=None, bg=(1.0, 1.0, 1.0)):
    from sings_b200.step import FrameInputs
    view = sc["view"]
    t = lambda a: torch.tensor(np.ascontiguousarray(a, np.float32), device="cuda")
    return FrameInputs(pose=t(sc["pose"] if pose is None else pose), transl=t(sc["transl"] if transl is None else transl),
                       viewmatrix=t(view.world_view_transform), projmatrix=t(view.full_proj_transform),
                       campos=t(view.camera_center), bg=t(np.asarray(bg, np.float32)),
                       tanfovx=view.tanfovx, tanfovy=view.tanfovy)


def test_cuda_graph_replay_equals_eager_launches():
    """AvatarStep.capture(): the frame as one CUDA graph gives the same bits as eager launches,
    follows in-place input updates, and keeps the stage events usable."""
    sc = make_scene(N=5000, H=144, W=112, seed=21)
    step = _avatar_step(sc, 144, 112, timing=True)
    fr = _frame(sc)
    G = torch.randn(3, 144, 112, device="cuda")
    img = step.forward(fr).clone()
    step.backward(G)
    torch.cuda.synchronize()
    ref = [x.clone() for x in (step.d_xyz_canon, step.d_shs, step.d_pose, step.d_opacity)]
    step.reset_stats()
    replay = step.capture(fr, G, loss_weight=G)
    step.reset_stats()
    replay()
    torch.cuda.synchronize()
    assert step.check_capacity() > 0
    assert torch.equal(step.color, img)
    # parameter gradients come from floating-point atomics: equal up to summation order
    for a, b in zip((step.d_xyz_canon, step.d_shs, step.d_pose, step.d_opacity), ref):
        assert rel_err(a.cpu().numpy(), b.cpu().numpy()) < 1e-5
    assert abs(float(step.loss) - float((img * G).sum())) <= 1e-3 * abs(float((img * G).sum())) + 1e-3
    assert step.stage_ms()["blend_bwd"] > 0
    # new pose written in place -> the replay renders the new frame
    from sings_b200 import synthetic as syn
    pose2 = syn.random_pose(24, seed=99)
    fr.pose.copy_(torch.from_numpy(pose2).cuda())
    replay()
    torch.cuda.synchronize()
    img_graph = step.color.clone()
    step.forward(fr)
    torch.cuda.synchronize()
    assert torch.equal(step.color, img_graph) and not torch.equal(img_graph, img)


def test_animation_frames_1080p_sharded_forward_only():
    """BASELINE.json configs[3] shape (scaled): 1920x1080 frames (68 tile rows, the last one
    ragged), forward only, contiguous frame ranges per rank, every frame bit-equal to the oracle."""
    from oracle import lbs_oracle as lo
    from sings_b200 import synthetic as syn
    from sings_b200.animate import render_frames
    H, W, F = 1080, 1920, 5
    sc = make_scene(N=20_000, H=H, W=W, seed=31, scale_range=(0.003, 0.02))
    av = sc["avatar"]
    step = _avatar_step(sc, H, W, D=3)
    poses = [syn.random_pose(24, seed=200 + f) for f in range(F)]
    frames = [_frame(sc, pose=p, bg=(0.0, 0.0, 0.0)) for p in poses]
    got = {}
    for rank in range(2):                      # two "ranks" rendered one after the other
        lo_, hi_, imgs = render_frames(step, frames, rank=rank, world=2, clamp=False)
        for f in range(lo_, hi_):
            got[f] = imgs[f - lo_].cpu().numpy()
    assert sorted(got) == list(range(F))
    tc = torch.from_numpy
    for f in (0, F - 1):
        A = lo.pose_to_A(tc(poses[f])[None], tc(av.rest), av.parents, tc(av.inv_A_t2cano))
        xyz, q, s, _ = lo.deform(A, tc(av.xyz_canon), tc(av.lbs_weights), tc(av.scales), tc(av.rotmat_canon),
                                 None, tc(sc["transl"])[None])
        # the rasterizer oracle is fed with the CUDA deform of the same frame (identical inputs
        # at the boundary); the deform itself is checked against the oracle to 1e-5
        step.forward(frames[f])
        torch.cuda.synchronize()
        assert np.abs(step.xyz[0].cpu().numpy() - xyz[0].numpy()).max() < 1e-5
        st = ro.forward(oracle_camera(sc["view"]), step.xyz[0].cpu().numpy(), av.opacity, np.zeros(3, np.float32),
                        shs=av.shs, scales=step.sc[0].cpu().numpy(), rotations=step.rotq[0].cpu().numpy(), sh_degree=3)
        assert np.array_equal(got[f], st.color)


def test_animation_two_lanes_equal_one_lane():
    """render_frames over two AvatarSteps (two streams, frames dealt round-robin) returns the same bits as over
    one, also when a pair list overflows in the middle of the sequence (the block is redone)."""
    from sings_b200 import synthetic as syn
    from sings_b200.animate import render_frames
    from sings_b200.step import AvatarStep
    H, W, F = 272, 400, 9
    sc = make_scene(N=6000, H=H, W=W, seed=41, scale_range=(0.004, 0.03))
    frames = [_frame(sc, pose=syn.random_pose(24, seed=300 + f), bg=(0.2, 0.4, 0.6)) for f in range(F)]
    one = _avatar_step(sc, H, W, D=3)
    _, _, ref = render_frames(one, frames, clamp=True)
    ref = ref.clone()
    lanes = [_avatar_step(sc, H, W, D=3), _avatar_step(sc, H, W, D=3)]
    lo_, hi_, got = render_frames(lanes, frames, clamp=True)
    assert (lo_, hi_) == (0, F) and torch.equal(got, ref)
    assert not lanes[0].forward_only and not lanes[1].forward_only
    # sharded: rank 1 of 2 gets the second half
    lo_, hi_, got = render_frames(lanes, frames, rank=1, world=2)
    assert torch.equal(got, ref[lo_:hi_])


def test_animation_lanes_redo_a_block_after_overflow():
    """Huge Gaussians: every frame needs more pairs than the initial capacity of an AvatarStep; one lane or two,
    the capacities grow once, the block is rendered again, the images are the same."""
    from sings_b200 import synthetic as syn
    from sings_b200.animate import render_frames
    H, W, F = 272, 400, 5
    sc = make_scene(N=3000, H=H, W=W, seed=42, scale_range=(0.2, 0.4))
    frames = [_frame(sc, pose=syn.random_pose(24, seed=400 + f), bg=(0.0, 0.0, 0.0)) for f in range(F)]
    one = _avatar_step(sc, H, W, D=3)
    cap0 = one.L_cap
    _, _, ref = render_frames(one, frames)
    ref = ref.clone()
    assert one.L_cap > cap0                                  # the single lane overflowed and recovered
    lanes = [_avatar_step(sc, H, W, D=3), _avatar_step(sc, H, W, D=3)]
    _, _, got = render_frames(lanes, frames)
    assert min(s.L_cap for s in lanes) > cap0
    assert torch.equal(got, ref)


def test_config_c1_neutral_pose_sh0_512():
    """BASELINE.json configs[0]: 50k Gaussians, neutral pose, SH degree 0, one 512x512 view --
    LBS + forward splat through the C ABI against the CPU reference path."""
    from oracle import lbs_oracle as lo
    sc = make_scene(N=50_000, H=512, W=512, seed=41, scale_range=(0.003, 0.015))
    av = sc["avatar"]
    step = _avatar_step(sc, 512, 512, D=0)
    fr = _frame(sc, pose=np.zeros((24, 3), np.float32))
    img = step.forward(fr).clone()
    torch.cuda.synchronize()
    assert step.check_capacity() > 0
    tc = torch.from_numpy
    A = lo.pose_to_A(torch.zeros(1, 24, 3), tc(av.rest), av.parents, tc(av.inv_A_t2cano))
    xyz, q, s, _ = lo.deform(A, tc(av.xyz_canon), tc(av.lbs_weights), tc(av.scales), tc(av.rotmat_canon),
                             None, tc(sc["transl"])[None])
    assert np.abs(step.xyz[0].cpu().numpy() - xyz[0].numpy()).max() < 1e-5
    assert np.abs(step.rotq[0].cpu().numpy() - q[0].numpy()).max() < 1e-5
    st = ro.forward(oracle_camera(sc["view"]), step.xyz[0].cpu().numpy(), av.opacity, np.ones(3, np.float32),
                    shs=av.shs, scales=step.sc[0].cpu().numpy(), rotations=step.rotq[0].cpu().numpy(), sh_degree=0)
    assert np.array_equal(step.radii.cpu().numpy(), st.radii)
    assert np.array_equal(img.cpu().numpy(), st.color)


def test_config_c3_turnaround_views_gradient_accumulation():
    """BASELINE.json configs[2]: human_complex-style step, 300k Gaussians, a batch of 8 synthetic
    turn-around views at 1024^2.  One GPU plays the 8 ranks in turn (the multi-rank exchange is
    covered by tests/test_dp_gloo.py): per view AvatarStep forward+backward, the per-view
    buckets summed and the statistics folded by GradExchange exactly as bench.py does under
    torchrun.  Checked against the oracle chain (C rasterizer forward/backward + float64 autograd
    of the LBS restatement): every image bit for bit, the batch gradient of every canonical
    parameter within 1e-3, the densification statistics of the batch."""
    from oracle import lbs_oracle as lo
    from sings_b200 import dp
    from sings_b200 import synthetic as syn
    N, H, W, V = 300_000, 1024, 1024, 8
    sc = make_scene(N=N, H=H, W=W, seed=21, scale_range=(0.002, 0.010))
    av = sc["avatar"]
    step = _avatar_step(sc, H, W, 3)
    ex = dp.GradExchange(N, step.n_param_grads, "cuda", defer_max=True)
    bg = np.ones(3, np.float32)
    tz = float(sc["transl"][2])
    views = [syn.make_view(H, W, yaw=2.0 * math.pi * k / V, centre=(0.0, 0.0, tz)) for k in range(V)]
    rng = np.random.default_rng(33)
    total = torch.zeros(step.n_param_grads, device="cuda")
    d_pose = []
    # oracle accumulators
    t64 = lambda a: torch.from_numpy(np.ascontiguousarray(a)).double()
    xyz_c, rot_c, scl_c = t64(av.xyz_canon).requires_grad_(True), t64(av.rotmat_canon).requires_grad_(True), t64(av.scales).requires_grad_(True)
    A64 = t64(sc["A"])[None]
    xyz64, q64, s64, _ = lo.deform(A64, xyz_c, t64(av.lbs_weights), scl_c, rot_c, None, t64(sc["transl"])[None])
    o_sh, o_op = np.zeros_like(av.shs, dtype=np.float64), np.zeros_like(av.opacity, dtype=np.float64)
    o_accum, o_denom, o_maxr = np.zeros(N), np.zeros(N), np.zeros(N)
    g_xyz, g_q, g_s = np.zeros((N, 3)), np.zeros((N, 4)), np.zeros((N, 3))
    for k, view in enumerate(views):
        G = rng.normal(size=(3, H, W)).astype(np.float32)
        sck = dict(sc, view=view)
        img = step.forward(_frame(sck)).clone()
        step.backward(torch.from_numpy(G).cuda())
        total += step.bucket[:step.n_param_grads]
        d_pose.append(step.d_pose.clone())
        ex.exchange(step.bucket, step.max_radii2D, async_op=True, reset_step=True)()
        torch.cuda.synchronize()
        assert step.check_capacity() > 0
        # the oracle rasterizes what the CUDA deformer produced (LBS parity is tolerance-based,
        # tests/test_gpu_lbs.py), so the images can be compared bit for bit
        npy = lambda x: np.ascontiguousarray(x[0].cpu().numpy())
        st = ro.forward(oracle_camera(view), npy(step.xyz), sc["opacity"], bg, sh_degree=3, shs=sc["shs"],
                        scales=npy(step.sc), rotations=npy(step.rotq))
        assert np.array_equal(img.cpu().numpy(), st.color), f"view {k}: image must be bit-exact"
        assert np.array_equal(step.radii.cpu().numpy(), st.radii)
        gr = ro.backward(st, G)
        g_xyz += gr["means3D"]; g_q += gr["rotations"]; g_s += gr["scales"]
        o_sh += gr["sh"]; o_op += gr["opacities"]
        vis = st.radii > 0
        o_accum[vis] += np.linalg.norm(gr["means2D"][vis, :2].astype(np.float64), axis=1)
        o_denom[vis] += 1.0
        o_maxr = np.maximum(o_maxr, np.where(vis, st.radii, 0).astype(np.float64))
    ((xyz64[0] * t64(g_xyz)).sum() + (q64[0] * t64(g_q)).sum() + (s64[0] * t64(g_s)).sum()).backward()
    n3, n9, M = 3 * N, 9 * N, av.shs.shape[1]
    got = total.cpu().numpy().astype(np.float64)
    parts = [("xyz_canon", got[:n3], xyz_c.grad.numpy().ravel()),
             ("scales", got[n3:2 * n3], scl_c.grad.numpy().ravel()),
             ("rot_canon", got[2 * n3:2 * n3 + n9], rot_c.grad.numpy().ravel()),
             ("opacity", got[2 * n3 + n9:2 * n3 + n9 + N], o_op.ravel()),
             ("shs", got[2 * n3 + n9 + N:2 * n3 + n9 + N + 3 * M * N], o_sh.ravel())]
    for name, a, b in parts:
        assert rel_err(a, b) < 1e-3, f"batch gradient {name}: {rel_err(a, b)}"
    assert np.array_equal(ex.denom.cpu().numpy(), o_denom.astype(np.float32))
    assert np.array_equal(ex.sync_max().cpu().numpy(), o_maxr.astype(np.float32))
    assert rel_err(ex.xyz_gradient_accum.cpu().numpy(), o_accum) < 1e-3
    assert float(step.bucket[step.n_param_grads:].abs().max()) == 0.0 and float(step.max_radii2D.abs().max()) == 0.0
    assert all(torch.isfinite(p).all() for p in d_pose)


def test_forward_only_frames_render_the_same_image_and_refuse_a_backward():
    """A forward-only AvatarStep (animation / evaluation) leaves nothing for the backward blend: same
    image, same n_contrib / final_T, and backward() refuses instead of reading lists that were never
    written."""
    import numpy as np
    import torch
    from helpers import make_scene
    from sings_b200._lib import SgsError
    from sings_b200.step import AvatarStep, FrameInputs
    dev = torch.device("cuda", 0)
    sc = make_scene(N=6000, H=200, W=176, seed=21)
    av, view = sc["avatar"], sc["view"]
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), device=dev)
    fr = FrameInputs(pose=t(sc["pose"]), transl=t(sc["transl"]), viewmatrix=t(view.world_view_transform),
                     projmatrix=t(view.full_proj_transform), campos=t(view.camera_center), bg=t(np.array([0.2, 0.3, 0.4], np.float32)),
                     tanfovx=view.tanfovx, tanfovy=view.tanfovy)
    mk = lambda: AvatarStep(t(av.xyz_canon), t(av.rotmat_canon), t(av.scales), t(av.opacity), t(av.shs), t(av.lbs_weights),
                            t(av.rest), torch.from_numpy(av.parents), t(av.inv_A_t2cano), view.image_height, view.image_width, 3)
    a, b = mk(), mk()
    b.forward_only = True
    ia, ib = a.forward(fr).clone(), b.forward(fr).clone()
    torch.cuda.synchronize()
    assert torch.equal(ia, ib)
    sa, sb = a.image_state(), b.image_state()
    assert torch.equal(sa["n_contrib"], sb["n_contrib"]) and torch.equal(sa["final_T"], sb["final_T"])
    with pytest.raises(SgsError):
        b.backward(torch.ones_like(ib))


def test_avatar_renderer_matches_the_three_function_path():
    """sings_b200.fused.AvatarRenderer (one autograd Function over the fused kernels) gives the image
    and every gradient of the reference-shaped path (pose_to_A + deform_gaussians + GaussianRasterizer)."""
    import numpy as np
    import torch
    from helpers import assert_grad_close, make_scene, raster_settings
    from diff_gaussian_rasterization import GaussianRasterizer
    from sings_b200 import deform
    from sings_b200.fused import AvatarRenderer
    dev = torch.device("cuda", 0)
    sc = make_scene(N=7000, H=208, W=160, seed=31)
    av, view = sc["avatar"], sc["view"]
    t = lambda a, rg=False: torch.tensor(np.ascontiguousarray(a), device=dev, requires_grad=rg)
    bg = np.array([0.3, 0.1, 0.6], np.float32)
    G = t(np.random.default_rng(1).normal(size=(3, view.image_height, view.image_width)).astype(np.float32))

    def leaves():
        return dict(pose=t(sc["pose"], True), transl=t(sc["transl"], True), xyz=t(av.xyz_canon, True),
                    rot=t(av.rotmat_canon, True), scl=t(av.scales, True), opa=t(av.opacity, True), shs=t(av.shs, True))
    # reference-shaped path
    a = leaves()
    A = deform.pose_to_A(a["pose"], t(av.rest), torch.from_numpy(av.parents), t(av.inv_A_t2cano))
    xyz, rotq, scl = deform.deform_gaussians(A, a["xyz"], t(av.lbs_weights), a["rot"], a["scl"], None, a["transl"])
    m2 = torch.zeros_like(xyz, requires_grad=True)
    img_a, radii_a = GaussianRasterizer(raster_settings(view, bg, 3))(means3D=xyz, means2D=m2, shs=a["shs"], opacities=a["opa"],
                                                                        scales=scl, rotations=rotq)
    (img_a * G).sum().backward()
    # one call
    b = leaves()
    r = AvatarRenderer(b["xyz"], b["rot"], b["scl"], b["opa"], b["shs"], t(av.lbs_weights), t(av.rest),
                       torch.from_numpy(av.parents), t(av.inv_A_t2cano), view.image_height, view.image_width, 3)
    img_b, radii_b = r(b["pose"], b["transl"], t(view.world_view_transform), t(view.full_proj_transform),
                       t(view.camera_center), t(bg), view.tanfovx, view.tanfovy)
    (img_b * G).sum().backward()
    torch.cuda.synchronize()
    assert torch.equal(img_a, img_b) and torch.equal(radii_a, radii_b)
    for k in a:
        assert_grad_close(b[k].grad.cpu().numpy(), a[k].grad.cpu().numpy(), k)
    # a second frame reuses the buffers; no-grad call renders the same image
    with torch.no_grad():
        img_c, _ = r(b["pose"], b["transl"], t(view.world_view_transform), t(view.full_proj_transform),
                     t(view.camera_center), t(bg), view.tanfovx, view.tanfovy)
    assert torch.equal(img_c, img_b)
