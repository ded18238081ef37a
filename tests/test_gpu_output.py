"""Animation output path: device-side 8-bit conversion (bit-exact against the reference's numpy
expression, gs_trainer.py:716-718) and the asynchronous frame writer."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _reference_u8(image: torch.Tensor, bgr: bool) -> np.ndarray:
    img_np = (image.detach().cpu().clamp(0, 1).permute(1, 2, 0).numpy() * 255).astype('uint8')     # gs_trainer.py:717
    return img_np[..., ::-1].copy() if bgr else img_np                                             # :718 cv2.COLOR_RGB2BGR


@pytest.mark.parametrize("H,W", [(1080, 1920), (67, 45), (896, 512)])
def test_frame_to_uint8_bit_exact(H, W):
    from sings_b200.animate import frame_to_uint8
    g = torch.Generator("cuda").manual_seed(H)
    img = torch.rand(3, H, W, device="cuda", generator=g) * 1.4 - 0.2          # values below 0 and above 1 too
    img[0, 0, :5] = torch.tensor([0.0, 1.0, 0.999999, 1.0 / 255, 254.9999 / 255], device="cuda")
    for bgr in (False, True):
        out = frame_to_uint8(img, bgr=bgr)
        assert out.shape == (H, W, 3) and out.dtype == torch.uint8
        assert np.array_equal(out.cpu().numpy(), _reference_u8(img, bgr))


def test_frame_writer_overlaps_and_delivers_every_frame(tmp_path):
    import cv2
    from sings_b200.animate import FrameWriter
    H, W, F = 120, 160, 20
    frames = [torch.rand(3, H, W, device="cuda", generator=torch.Generator("cuda").manual_seed(i)) for i in range(F)]
    got = {}
    w = FrameWriter(None, H, W, "cuda", depth=3, workers=2, encode=False, sink=lambda n, a: got.__setitem__(n, a.copy()))
    for i, f in enumerate(frames):
        w.submit(f, f"{i:05d}")
    assert w.close() == F and len(got) == F
    for i, f in enumerate(frames):
        assert np.array_equal(got[f"{i:05d}"], _reference_u8(f, False))
    # through the encoder: PNG is lossless, so the file holds exactly the reference's BGR bytes
    w = FrameWriter(str(tmp_path), H, W, "cuda", depth=4, workers=2, ext="png")
    for i, f in enumerate(frames[:6]):
        w.submit(f, f"{i:05d}")
    assert w.close() == 6
    for i, f in enumerate(frames[:6]):
        assert np.array_equal(cv2.imread(os.path.join(str(tmp_path), f"{i:05d}.png")), _reference_u8(f, True))
