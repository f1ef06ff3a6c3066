"""Golden fixture of the front end (tests/golden/frontend.npz) from the UNMODIFIED reference (build container only):
src/utils/functional.py ycbcr420_to_rgb, src/utils/core.py imresize, src/utils/common.py get_interlayer_padding and
test.py's PSNR are imported from /root/reference, evaluated on a seeded random 8-bit YUV 4:2:0 frame, and the oracle
restatement (oracle/frontend_oracle.py) is required to agree bit for bit before anything is written."""
import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import frontend_oracle as fo  # noqa: E402

REF = "/root/reference"


def load(path, name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, path))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def main():
    core = load("src/utils/core.py", "ref_core")
    functional = load("src/utils/functional.py", "ref_functional")
    common = load("src/utils/common.py", "ref_common")
    out = {}
    for tag, (H, W) in (("a", (72, 104)), ("b", (160, 256))):
        rng = np.random.default_rng(7 if tag == "a" else 11)
        # smooth-ish content plus noise so that the clip and both chroma gradients are exercised
        yy, xx = np.mgrid[0:H, 0:W]
        y8 = np.clip(128 + 90 * np.sin(yy / 9.0) * np.cos(xx / 13.0) + rng.normal(0, 25, (H, W)), 0, 255).astype(np.uint8)
        uv8 = rng.integers(0, 256, (2, H // 2, W // 2), dtype=np.uint8)
        pad = common.get_interlayer_padding(H, W, 2)
        assert pad == __import__("lssvc_b200.frontend", fromlist=["x"]).get_interlayer_padding(H, W, 2)
        y, uv = y8[None].astype(np.float32) / 255, uv8.astype(np.float32) / 255
        rgb_ref = functional.ycbcr420_to_rgb(y, uv)
        rgb_orc = fo.ycbcr420_to_rgb(y, uv)
        assert rgb_ref.dtype == np.float32 and np.array_equal(rgb_ref, rgb_orc), "oracle ycbcr420_to_rgb != reference"
        x_el = torch.nn.functional.pad(torch.from_numpy(rgb_ref).type(torch.FloatTensor).unsqueeze(0), pad["P_HR"], mode="constant", value=0)
        assert torch.equal(x_el, fo.pad_el(rgb_orc, pad["P_HR"]))
        bl_ref = core.imresize(x_el, sizes=pad["LR_padded_size"], kernel="cubic").clamp_(0, 1)
        bl_orc = fo.base_layer(x_el, pad["LR_padded_size"])
        assert torch.equal(bl_ref, bl_orc), f"oracle imresize != reference (max diff {(bl_ref - bl_orc).abs().max()})"
        # a non-integer ratio as well (x1.5 of recommend_test_config.json)
        sz15 = (x_el.shape[2] * 2 // 3, x_el.shape[3] * 2 // 3)
        r15_ref = core.imresize(x_el, sizes=sz15, kernel="cubic")
        assert torch.equal(r15_ref, fo.imresize_cubic(x_el, sz15))
        crop = x_el[:, :, :40, :48].contiguous()
        noisy = (crop + 0.01 * torch.randn(crop.shape, generator=torch.Generator().manual_seed(3))).clamp(0, 1)
        mse = torch.mean((crop - noisy) ** 2)
        psnr_ref = (10 * torch.log10(1.0 / mse)).item()          # test.py:115-118 (a nested function of test.py: restated)
        assert psnr_ref == fo.psnr(crop, noisy)
        sub = (slice(None), slice(None), slice(None, None, 1 if tag == "a" else 4), slice(None, None, 1 if tag == "a" else 4))
        out.update({f"{tag}_y8": y8, f"{tag}_uv8": uv8, f"{tag}_rgb": rgb_ref[sub[1:]].copy(), f"{tag}_bl": bl_ref.numpy()[sub].copy(),
                    f"{tag}_x15": r15_ref.numpy()[:, :, ::3, ::3].copy(), f"{tag}_noisy": noisy.numpy(), f"{tag}_psnr": np.float64(psnr_ref),
                    f"{tag}_pad": np.array(pad["P_HR"] + pad["P_LR"] + pad["HR_padded_size"] + pad["LR_padded_size"])})
        print(tag, (H, W), "->", tuple(x_el.shape), tuple(bl_ref.shape), "psnr", psnr_ref)
    path = os.path.join(ROOT, "tests", "golden", "frontend.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
