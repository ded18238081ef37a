"""CPU: the SOURCE of the small streaming kernels -- densify.cu (densification statistics, data-parallel
fold, frame clear) and frame_out.cu (8-bit frame conversion) -- executed under the SIMT emulation of
tests/cuda_emu against the reference's expressions.  Complements the GPU tests (test_gpu_dropin.py,
test_gpu_dp.py, test_gpu_output.py); test infrastructure only, the product has no CPU path."""
import ctypes as C
import os
import shutil

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "sings_b200", "csrc")
vp, i32, sz = C.c_void_p, C.c_int, C.c_size_t


def p(a):
    return None if a is None else a.ctypes.data


@pytest.fixture(scope="module")
def densify():
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    from cuda_emu import build
    L = build(os.path.join(CSRC, "densify.cu"), r'''
extern "C" {
int emu_densify_stats(int P, const float* g, const int* radii, float* accum, float* denom, float* max_radii) {
    return sgs::launch_densify_stats(P, g, radii, accum, denom, max_radii, nullptr);
}
int emu_fold_stats(int P, float* sa, float* sd, float* sm, float* accum, float* denom, float* max_radii) {
    return sgs::launch_fold_stats(P, sa, sd, sm, accum, denom, max_radii, nullptr);
}
int emu_clear3(void* a, size_t na, void* b, size_t nb, void* c, size_t nc) { return sgs::launch_clear3(a, na, b, nb, c, nc, nullptr); }
}
''')
    L.emu_densify_stats.argtypes = [i32, vp, vp, vp, vp, vp]
    L.emu_fold_stats.argtypes = [i32, vp, vp, vp, vp, vp, vp]
    L.emu_clear3.argtypes = [vp, sz, vp, sz, vp, sz]
    return L


@pytest.fixture(scope="module")
def frame_out():
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    from cuda_emu import build
    L = build(os.path.join(CSRC, "frame_out.cu"), r'''
extern "C" int emu_frame_to_u8(const float* img, int H, int W, int bgr, unsigned char* out) {
    return sgs::launch_frame_to_u8(img, H, W, bgr, out, nullptr);
}
''')
    L.emu_frame_to_u8.argtypes = [vp, i32, i32, i32, vp]
    return L


def test_densification_statistics(densify):
    """sings_hybrid.py:1013-1015 + gs_trainer.py:487-490 over two views, then the data-parallel fold."""
    rng = np.random.default_rng(0)
    P = 1000
    accum, denom, maxr = np.zeros(P, np.float32), np.zeros(P, np.float32), np.zeros(P, np.float32)
    ra, rd, rm = accum.copy(), denom.copy(), maxr.copy()
    for _ in range(2):
        g = rng.standard_normal((P, 3)).astype(np.float32)
        radii = rng.integers(-1, 40, P).astype(np.int32)
        assert densify.emu_densify_stats(P, p(g), p(radii), p(accum), p(denom), p(maxr)) == 0
        vis = radii > 0
        ra[vis] += np.sqrt(g[vis, 0] * g[vis, 0] + g[vis, 1] * g[vis, 1])      # torch.norm(grad[vis, :2], dim=-1) in binary32
        rd[vis] += 1
        rm[vis] = np.maximum(rm[vis], radii[vis].astype(np.float32))
    np.testing.assert_array_equal(denom, rd)
    np.testing.assert_array_equal(maxr, rm)
    np.testing.assert_allclose(accum, ra, rtol=2e-7)
    tot_a, tot_d, tot_m = np.ones(P, np.float32), np.full(P, 2, np.float32), np.full(P, 7.5, np.float32)
    ea, ed, em = tot_a + accum, tot_d + denom, np.maximum(tot_m, maxr)
    assert densify.emu_fold_stats(P, p(accum), p(denom), p(maxr), p(tot_a), p(tot_d), p(tot_m)) == 0
    np.testing.assert_array_equal(tot_a, ea); np.testing.assert_array_equal(tot_d, ed); np.testing.assert_array_equal(tot_m, em)
    assert not accum.any() and not denom.any() and not maxr.any()              # step buffers cleared for the next view


def test_frame_clear_touches_exactly_its_regions(densify):
    buf = np.full(3 * 4096 + 64, 0xAB, np.uint8)
    base = buf.ctypes.data
    off = (-base) % 16
    a, na = base + off, 500 * 16
    b, nb = a + na + 16, 515                      # ragged tail handled bytewise
    c, nc = b + 528, 0
    assert densify.emu_clear3(a, na, b, nb, c, nc) == 0
    ref = np.full_like(buf, 0xAB)
    ref[off:off + na] = 0
    ref[off + na + 16:off + na + 16 + nb] = 0
    np.testing.assert_array_equal(buf, ref)


@pytest.mark.parametrize("H,W", [(48, 64), (37, 53), (1, 3)])
@pytest.mark.parametrize("bgr", [0, 1])
def test_frame_to_uint8_bit_exact(frame_out, H, W, bgr):
    """gs_trainer.py:716-718: (image.clamp(0, 1).permute(1, 2, 0).numpy() * 255).astype('uint8') [+ RGB -> BGR]."""
    rng = np.random.default_rng(H * W + bgr)
    img = (rng.random((3, H, W), np.float32) * 1.4 - 0.2).astype(np.float32)
    img[:, 0, 0] = [0.0, 1.0, 0.999999]
    out = np.full((H, W, 3), 0x55, np.uint8)
    assert frame_out.emu_frame_to_u8(p(img), H, W, bgr, p(out)) == 0
    ref = (np.clip(img, 0, 1).transpose(1, 2, 0) * np.float32(255)).astype("uint8")
    if bgr:
        ref = ref[:, :, ::-1]
    np.testing.assert_array_equal(out, ref)
