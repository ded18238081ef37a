"""CPU: the SOURCE of lbs.cu -- pose -> A, the stand-alone LBS forward / backward (TMA bulk loads and stores
stood in by memcpy, the mbarrier by a counter), the 6D-rotation conversions -- executed under the SIMT
emulation of tests/cuda_emu against the golden vectors produced by the reference's own lbs_extra /
rotations / batch_rigid_transform (tests/golden/make_lbs_golden.py, make_rot6d_golden.py).  Same bars as
tests/test_gpu_lbs.py: values 1e-5, gradients 1e-3 against the reference's float64 autograd.  Test
infrastructure only -- the product has no CPU path."""
import ctypes as C
import glob
import os
import shutil

import numpy as np
import pytest

from helpers import assert_grad_close

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "lbs_golden_*_f32.npz")))
VAL_TOL, GRAD_TOL = 1e-5, 1e-3
vp, i32 = C.c_void_p, C.c_int

EXPORTS = r'''
extern "C" {
int emu_pose_to_A(const float* pose, const float* rest, const int* parents, const float* inv_A, int B, int J, float* A, float* G) {
    return sgs::launch_pose_to_A(pose, rest, parents, inv_A, B, J, A, G, nullptr);
}
int emu_pose_to_A_bwd(const float* pose, const float* rest, const int* parents, const float* inv_A, const float* G,
                      const float* dA, int B, int J, float* d_pose) {
    return sgs::launch_pose_to_A_bwd(pose, rest, parents, inv_A, G, dA, B, J, d_pose, nullptr);
}
static sgs::LbsArgs fill(int B, int N, int J, const float* A, const float* xyz, const float* W, const float* rot,
                         const float* scales, const float* ss, const float* tr, const float* et, const float* er,
                         const float* es, int rot6d) {      // = fill_lbs of api.cu
    sgs::LbsArgs a;
    a.B = B; a.N = N; a.J = J; a.A = A; a.xyz = xyz; a.W = W; a.rot = rot; a.scales = scales; a.smpl_scale = ss;
    a.transl = tr; a.ext_trans = et; a.ext_rot = er; a.ext_scale = es; a.rot6d = rot6d;
    return a;
}
int emu_lbs_fwd(int B, int N, int J, const float* A, const float* xyz, const float* W, const float* rot, const float* scales,
                const float* ss, const float* tr, const float* et, const float* er, const float* es, int rot6d,
                float* xyz_o, float* rotq_o, float* sc_o, float* T_o) {
    sgs::LbsOut o{xyz_o, rotq_o, sc_o, T_o};
    return sgs::launch_lbs_fwd(fill(B, N, J, A, xyz, W, rot, scales, ss, tr, et, er, es, rot6d), o, nullptr);
}
int emu_lbs_bwd(int B, int N, int J, const float* A, const float* xyz, const float* W, const float* rot, const float* scales,
                const float* ss, const float* tr, const float* et, const float* er, const float* es, int rot6d,
                const float* g_xyz, const float* g_rotq, const float* g_sc, const float* g_T, float* d_xyz, float* d_rot,
                float* d_sc, float* d_A, float* d_ss, float* d_tr) {
    sgs::LbsGrads g{g_xyz, g_rotq, g_sc, g_T, d_xyz, d_rot, d_sc, d_A, d_ss, d_tr};
    return sgs::launch_lbs_bwd(fill(B, N, J, A, xyz, W, rot, scales, ss, tr, et, er, es, rot6d), g, nullptr);
}
int emu_rot6d(const float* d6, int n, int mode, float* out) { return sgs::launch_rot6d_convert(d6, n, mode, out, nullptr); }
int emu_rot6d_bwd(const float* d6, const float* g, int n, int mode, float* g6) { return sgs::launch_rot6d_convert_bwd(d6, g, n, mode, g6, nullptr); }
}
'''


@pytest.fixture(scope="module")
def lbs():
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    from cuda_emu import LBS_REWRITES, build
    L = build(os.path.join(ROOT, "sings_b200", "csrc", "lbs.cu"), EXPORTS, rewrites=LBS_REWRITES)
    L.emu_pose_to_A.argtypes = [vp, vp, vp, vp, i32, i32, vp, vp]
    L.emu_pose_to_A_bwd.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32, vp]
    L.emu_lbs_fwd.argtypes = [i32, i32, i32] + [vp] * 10 + [i32] + [vp] * 4
    L.emu_lbs_bwd.argtypes = [i32, i32, i32] + [vp] * 10 + [i32] + [vp] * 10
    L.emu_rot6d.argtypes = [vp, i32, i32, vp]
    L.emu_rot6d_bwd.argtypes = [vp, vp, i32, i32, vp]
    return L


def p(a):
    return None if a is None else a.ctypes.data


def c32(a):
    return None if a is None else np.ascontiguousarray(a, np.float32)


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(q)[11:-8] for q in GOLD])
def test_lbs_kernels_against_reference_golden(lbs, path):
    g, g64 = np.load(path), np.load(path.replace("_f32", "_f64"))
    get = lambda k: c32(g[k]) if k in g.files else None
    pose, rest, inv_A = get("pose"), get("rest"), get("inv_A_t2cano")
    parents = np.ascontiguousarray(g["parents"], np.int32)
    B, J = pose.shape[0], pose.shape[1]
    xyz, W, scales = get("xyz_canon"), get("lbs_weights"), get("scales")
    N = xyz.shape[0]
    rot = None if bool(g["isotropic"]) else get("rotmat_canon")
    ss, tr = get("smpl_scale"), get("transl")
    et, er, es = get("ext_trans"), get("ext_rotmat"), get("ext_scale")
    nanf = lambda *s: np.full(s, np.nan, np.float32)
    # forward: pose -> A -> deformed means / quaternions / scales (+ T)
    A, G = nanf(B, J, 4, 4), nanf(B, J, 12)
    assert lbs.emu_pose_to_A(p(pose), p(rest), p(parents), p(inv_A), B, J, p(A), p(G)) == 0
    assert np.abs(A - g["A_cano2pose"]).max() < VAL_TOL
    xo, qo, so, To = nanf(B, N, 3), nanf(B, N, 4), nanf(B, N, 3), nanf(B, N, 4, 4)
    assert lbs.emu_lbs_fwd(B, N, J, p(A), p(xyz), p(W), p(rot), p(scales), p(ss), p(tr), p(et), p(er), p(es), 0,
                           p(xo), p(qo), p(so), p(To)) == 0
    assert np.abs(xo - g["xyz"]).max() < VAL_TOL and np.abs(qo - g["rotq"]).max() < VAL_TOL
    assert np.abs(so - g["scales_out"]).max() < VAL_TOL and np.abs(To - g["T"]).max() < VAL_TOL
    # backward of loss = <xyz, gx> + <q, gq> + <scales, gs>, down to the pose
    d_xyz, d_sc = nanf(N, 3), nanf(N, 3)
    d_rot = None if rot is None else nanf(N, 3, 3)
    d_A = np.zeros((B, J, 4, 4), np.float32)
    d_ss = None if ss is None else np.zeros(B, np.float32)
    d_tr = None if tr is None else np.zeros((B, 3), np.float32)
    assert lbs.emu_lbs_bwd(B, N, J, p(A), p(xyz), p(W), p(rot), p(scales), p(ss), p(tr), p(et), p(er), p(es), 0,
                           p(get("gx")), p(get("gq")), p(get("gs")), None, p(d_xyz), p(d_rot), p(d_sc), p(d_A), p(d_ss), p(d_tr)) == 0
    d_pose = nanf(B, J, 3)
    assert lbs.emu_pose_to_A_bwd(p(pose), p(rest), p(parents), p(inv_A), p(G), p(d_A), B, J, p(d_pose)) == 0
    checks = [("d_pose", d_pose), ("d_xyz_canon", d_xyz), ("d_scales", d_sc)]
    if ss is not None:
        checks.append(("d_smpl_scale", d_ss.reshape(g64["d_smpl_scale"].shape)))
    if tr is not None:
        checks.append(("d_transl", d_tr))
    if rot is not None:
        checks.append(("d_rotmat_canon", d_rot))
    for name, got in checks:
        assert_grad_close(got, g64[name], name, tol=GRAD_TOL)


def test_rot6d_kernels_against_reference_golden(lbs):
    base = os.path.join(ROOT, "tests", "golden", "rot6d_golden_")
    g, g64 = np.load(base + "f32.npz"), np.load(base + "f64.npz")
    d6 = c32(g["d6"])
    n = d6.shape[0]
    R, aa = np.full((n, 9), np.nan, np.float32), np.full((n, 3), np.nan, np.float32)
    assert lbs.emu_rot6d(p(d6), n, 0, p(R)) == 0 and lbs.emu_rot6d(p(d6), n, 1, p(aa)) == 0
    assert np.abs(R.reshape(n, 3, 3) - g["R"]).max() < VAL_TOL
    assert np.abs(aa - g["aa"]).max() < 2e-4                 # near pi the axis-angle is ill-conditioned in float32
    well = np.linalg.norm(g64["aa"], axis=-1) < 2.8
    assert np.abs(aa[well] - g64["aa"][well]).max() < 2e-5
    dR, daa = np.full((n, 6), np.nan, np.float32), np.full((n, 6), np.nan, np.float32)
    assert lbs.emu_rot6d_bwd(p(d6), p(c32(g["gR"]).reshape(n, 9)), n, 0, p(dR)) == 0
    assert lbs.emu_rot6d_bwd(p(d6), p(c32(g["gaa"])), n, 1, p(daa)) == 0
    rel = lambda a, b: np.abs(a - b).max() / np.abs(b).max()
    assert rel(dR, g64["d_d6_from_R"]) < GRAD_TOL and rel(daa[well], g64["d_d6_from_aa"][well]) < GRAD_TOL


def test_lbs_from_rot6d_equals_the_matrix_path(lbs):
    """The stored 6D canonical rotation straight into the LBS kernels: same values as converting first."""
    g = np.load(GOLD[0])
    rng = np.random.default_rng(5)
    A = c32(g["A_cano2pose"])
    B, J = A.shape[0], A.shape[1]
    xyz, W, scales = c32(g["xyz_canon"]), c32(g["lbs_weights"]), c32(g["scales"])
    N = xyz.shape[0]
    d6 = rng.standard_normal((N, 6)).astype(np.float32)
    R = np.full((N, 9), np.nan, np.float32)
    assert lbs.emu_rot6d(p(d6), N, 0, p(R)) == 0
    outs = []
    for rot, flag in ((R, 0), (d6, 1)):
        xo, qo, so = (np.full((B, N, k), np.nan, np.float32) for k in (3, 4, 3))
        assert lbs.emu_lbs_fwd(B, N, J, p(A), p(xyz), p(W), p(rot), p(scales), None, None, None, None, None, flag,
                               p(xo), p(qo), p(so), None) == 0
        outs.append((xo, qo, so))
    for a, b in zip(*outs):
        np.testing.assert_array_equal(a, b)
