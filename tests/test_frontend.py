"""Front end of the frame loop: YUV 4:2:0 -> padded RGB, MATLAB-compatible bicubic EL -> BL resize, PSNR.

CPU: the oracle restatement (oracle/frontend_oracle.py) against the fixture generated from the unmodified reference
(tools/make_golden_frontend.py, bit-identical), and the host-side resize plan / padding rule of lssvc_b200/frontend.py.
GPU: the CUDA kernels through the C-ABI against the oracle and the fixture — bit-exact for the colour conversion and the x2
resize (same operation order, no FMA contraction), at the fixture sizes and at 1080p."""
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "frontend.npz")


def _sub(tag):
    return 1 if tag == "a" else 4


@pytest.mark.parametrize("tag", ["a", "b"])
def test_oracle_matches_reference_fixture(tag):
    from oracle import frontend_oracle as fo
    g = np.load(GOLD)
    y8, uv8 = g[f"{tag}_y8"], g[f"{tag}_uv8"]
    H, W = y8.shape
    p = g[f"{tag}_pad"]
    y, uv = fo.read_yuv420_frame(y8.tobytes() + uv8.tobytes(), H, W)
    rgb = fo.ycbcr420_to_rgb(y, uv)
    s = _sub(tag)
    assert np.array_equal(rgb[:, ::s, ::s], g[f"{tag}_rgb"])
    x_el = fo.pad_el(rgb, tuple(int(v) for v in p[0:4]))
    assert tuple(x_el.shape[2:]) == (int(p[8]), int(p[9]))
    bl = fo.base_layer(x_el, (int(p[10]), int(p[11])))
    assert np.array_equal(bl.numpy()[:, :, ::s, ::s], g[f"{tag}_bl"])
    x15 = fo.imresize_cubic(x_el, (x_el.shape[2] * 2 // 3, x_el.shape[3] * 2 // 3))
    assert np.array_equal(x15.numpy()[:, :, ::3, ::3], g[f"{tag}_x15"])
    crop = x_el[:, :, :40, :48].contiguous()
    assert fo.psnr(crop, torch.from_numpy(g[f"{tag}_noisy"])) == float(g[f"{tag}_psnr"])


def test_padding_rule_and_resize_plan():
    from lssvc_b200 import frontend as fe
    pad = fe.get_interlayer_padding(1080, 1920, 2)        # BASELINE config 2: EL 1080p -> 1152x1920, BL 540x960 -> 576x960
    assert pad["HR_padded_size"] == (1152, 1920) and pad["LR_padded_size"] == (576, 960)
    assert pad["P_HR"] == (0, 0, 0, 72) and pad["P_LR"] == (0, 0, 0, 36) and pad["LR_size"] == (540, 960)
    pad = fe.get_interlayer_padding(320, 512, 2)          # config 1
    assert pad["HR_padded_size"] == (384, 512) and pad["LR_padded_size"] == (192, 256)
    pad = fe.get_interlayer_padding(2160, 3840, 2)        # config 5
    assert pad["HR_padded_size"] == (2176, 3840) and pad["LR_padded_size"] == (1088, 1920)
    # x2 down-sampling: 10 taps (4 / 0.5 + 2), weights sum to 1, border samples reflected with the edge used twice
    w, taps, K = fe.resize_plan(128, 64)
    assert K == 10 and w.shape == (64, 10) and taps.shape == (64, 10)
    assert torch.allclose(w.sum(1), torch.ones(64), atol=1e-6)
    assert taps[0].tolist() == [3, 2, 1, 0, 0, 1, 2, 3, 4, 5] and taps[-1].tolist() == [122, 123, 124, 125, 126, 127, 127, 126, 125, 124]
    # plan-based evaluation on the host == oracle (the kernel evaluates exactly this sum)
    from oracle import frontend_oracle as fo
    x = torch.rand(1, 2, 128, 96, generator=torch.Generator().manual_seed(0))
    ref = fo.imresize_cubic(x, (64, 48))
    wv, tv, _ = fe.resize_plan(128, 64)
    wh, th, _ = fe.resize_plan(96, 48)
    rows = torch.zeros(1, 2, 64, 96)
    for k in range(wv.shape[1]):
        rows = rows + x[:, :, tv[:, k].long(), :] * wv[:, k].view(1, 1, -1, 1)
    out = torch.zeros(1, 2, 64, 48)
    for k in range(wh.shape[1]):
        out = out + rows[:, :, :, th[:, k].long()] * wh[:, k].view(1, 1, 1, -1)
    assert (out - ref).abs().max().item() < 2e-7


def test_frontend_requires_cuda():
    from lssvc_b200 import _lib, frontend as fe
    with pytest.raises(_lib.LssvcError):
        fe.imresize(torch.zeros(1, 3, 8, 8), sizes=(4, 4))


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["a", "b"])
def test_cuda_front_end_against_fixture_and_oracle(tag, cuda_device):
    from lssvc_b200 import frontend as fe
    from oracle import frontend_oracle as fo
    g = np.load(GOLD)
    y8, uv8 = g[f"{tag}_y8"], g[f"{tag}_uv8"]
    H, W = y8.shape
    front = fe.FrontEnd(H, W, 2, cuda_device)
    x_el = front.rgb_from_yuv420(torch.from_numpy(y8).to(cuda_device), torch.from_numpy(uv8).to(cuda_device))
    y, uv = fo.read_yuv420_frame(y8.tobytes() + uv8.tobytes(), H, W)
    ref_el = fo.pad_el(fo.ycbcr420_to_rgb(y, uv), front.padding["P_HR"])
    s = _sub(tag)
    assert np.array_equal(x_el.cpu().numpy()[0, :, :H:s, :W:s], g[f"{tag}_rgb"]), "colour conversion differs from the reference fixture"
    assert torch.equal(x_el.cpu(), ref_el), "padded EL frame differs from the oracle"
    x_bl = front.base_layer(x_el)
    ref_bl = fo.base_layer(ref_el, front.bl_size)
    d = (x_bl.cpu() - ref_bl).abs().max().item()
    print(f"{tag}: BL max|d| vs oracle {d:.3e}")
    assert d <= 1.2e-7                                    # one fp32 ulp at 1.0: the summation order inside torch's sum(dim=1)
    assert np.abs(x_bl.cpu().numpy()[:, :, ::s, ::s] - g[f"{tag}_bl"]).max() <= 1.2e-7
    x15 = fe.imresize(x_el, sizes=(x_el.shape[2] * 2 // 3, x_el.shape[3] * 2 // 3))
    assert np.abs(x15.cpu().numpy()[:, :, ::3, ::3] - g[f"{tag}_x15"]).max() <= 2.4e-7
    crop = x_el[:, :, :40, :48].contiguous()
    noisy = torch.from_numpy(g[f"{tag}_noisy"]).to(cuda_device)
    assert abs(fe.psnr(crop, noisy) - float(g[f"{tag}_psnr"])) < 1e-4


@pytest.mark.gpu
def test_cuda_front_end_1080p(cuda_device):
    """BASELINE config 2 sizes: EL 1080x1920 -> 1152x1920 padded, BL 576x960, against the oracle on the host."""
    from lssvc_b200 import frontend as fe
    from oracle import frontend_oracle as fo
    H, W = 1080, 1920
    rng = np.random.default_rng(5)
    yy, xx = np.mgrid[0:H, 0:W]
    y8 = np.clip(120 + 80 * np.sin(yy / 37.0) * np.cos(xx / 53.0) + rng.normal(0, 12, (H, W)), 0, 255).astype(np.uint8)
    uv8 = np.clip(128 + 60 * np.cos(np.mgrid[0:2, 0:H // 2, 0:W // 2].sum(0) / 41.0) + rng.normal(0, 6, (2, H // 2, W // 2)), 0, 255).astype(np.uint8)
    front = fe.FrontEnd(H, W, 2, cuda_device)
    assert front.el_size == (1152, 1920) and front.bl_size == (576, 960)
    x_el = front.rgb_from_yuv420(torch.from_numpy(y8).to(cuda_device), torch.from_numpy(uv8).to(cuda_device))
    y, uv = fo.read_yuv420_frame(y8.tobytes() + uv8.tobytes(), H, W)
    ref_el = fo.pad_el(fo.ycbcr420_to_rgb(y, uv), front.padding["P_HR"])
    assert torch.equal(x_el.cpu(), ref_el)
    assert float(x_el[:, :, H:, :].abs().max()) == 0.0
    x_bl = front.base_layer(x_el)
    ref_bl = fo.base_layer(ref_el, front.bl_size)
    d = (x_bl.cpu() - ref_bl).abs().max().item()
    print(f"1080p: BL max|d| vs oracle {d:.3e}, equal {100 * (x_bl.cpu() == ref_bl).float().mean().item():.3f} %")
    assert d <= 1.2e-7
    assert abs(fe.psnr(x_bl, ref_bl.to(cuda_device).roll(1, 3)) - fo.psnr(x_bl.cpu(), ref_bl.roll(1, 3))) < 1e-3


@pytest.mark.parametrize("shape, sizes", [((96, 64), (64, 48)), ((64, 64), (96, 80)), ((60, 90), (20, 31)), ((128, 128), (64, 64))])
def test_resize_plans_reproduce_the_oracle(shape, sizes):
    """Down-sampling with antialiasing (x2/3, x1/3, x1/2) and up-sampling (x1.5, x1.25): evaluating the host-built weights and
    reflect-resolved taps the way the kernel does (taps in order, multiply then add) gives the oracle's imresize."""
    from lssvc_b200 import frontend as fe
    from oracle import frontend_oracle as fo
    (h, w), (ho, wo) = shape, sizes
    x = torch.rand(1, 2, h, w, generator=torch.Generator().manual_seed(1))
    ref = fo.imresize_cubic(x, (ho, wo))
    wv, tv, _ = fe.resize_plan(h, ho)
    wh, th, _ = fe.resize_plan(w, wo)
    rows = torch.zeros(1, 2, ho, w)
    for k in range(wv.shape[1]):
        rows = rows + x[:, :, tv[:, k].long(), :] * wv[:, k].view(1, 1, -1, 1)
    out = torch.zeros(1, 2, ho, wo)
    for k in range(wh.shape[1]):
        out = out + rows[:, :, :, th[:, k].long()] * wh[:, k].view(1, 1, 1, -1)
    assert (out - ref).abs().max().item() <= 1.2e-7
