"""ORACLE — test infrastructure only (tests/ and tools/make_golden_frontend.py).

CPU restatement of the host-side front end of the reference's frame loop (test.py:183-199, 253-254): YUV 4:2:0 -> RGB,
zero padding, MATLAB-compatible bicubic EL -> BL resize, PSNR.  numpy / scipy / torch in the reference's operation order;
each function cites the reference file:line it follows (paths under /root/reference).  scipy.ndimage.zoom is the
reference's own third-party dependency (functional.py:4,49) and is called as the reference calls it.

Pinning: tools/make_golden_frontend.py imports the unmodified reference functions in the build container, requires
bit-identical outputs from this restatement and commits tests/golden/frontend.npz, which tests/test_frontend.py re-checks
wherever the reference is absent (the GPU box).  Nothing under lssvc_b200/ imports this module.
"""
import math

import numpy as np
import scipy.ndimage
import torch


def read_yuv420_frame(buf, height, width):
    """video_reader.py:139-155: bytes of one frame -> y [1, H, W], uv [2, H/2, W/2] float32 in [0, 1]."""
    y = np.frombuffer(buf[:height * width], dtype=np.uint8).reshape(1, height, width)
    uv = np.frombuffer(buf[height * width:height * width * 3 // 2], dtype=np.uint8).reshape(2, height // 2, width // 2)
    return y.astype(np.float32) / 255, uv.astype(np.float32) / 255


def ycbcr420_to_rgb(y, uv, order=1):
    """functional.py:42-58 (BT.709, K_r = 0.2126, K_b = 0.0722; functional.py:10-13)."""
    uv = scipy.ndimage.zoom(uv, (1, 2, 2), order=order)
    cb, cr = uv[0:1], uv[1:2]
    Kr, Kg, Kb = 0.2126, 0.7152, 0.0722
    r = y + (2 - 2 * Kr) * (cr - 0.5)
    b = y + (2 - 2 * Kb) * (cb - 0.5)
    g = (y - Kr * r - Kb * b) / Kg
    return np.clip(np.concatenate((r, g, b), axis=0), 0., 1.)


def pad_el(rgb, p_size):
    """test.py:189-197: np_image_to_tensor + F.pad(rgb, (left, right, top, bottom), mode='constant', value=0)."""
    x = torch.from_numpy(rgb).type(torch.FloatTensor).unsqueeze(0)
    return torch.nn.functional.pad(x, p_size, mode="constant", value=0)


def _cubic_contribution(x, a=-0.5):
    """core.py:40-55"""
    ax = x.abs()
    ax2 = ax * ax
    ax3 = ax * ax2
    cont_01 = ((a + 2) * ax3 - (a + 3) * ax2 + 1) * ax.le(1).to(dtype=x.dtype)
    cont_12 = ((a * ax3) - (5 * a * ax2) + (8 * a * ax) - (4 * a)) * torch.logical_and(ax.gt(1), ax.le(2)).to(dtype=x.dtype)
    return cont_01 + cont_12


def _reflect_pad(x, dim, pad_pre, pad_post):
    """core.py:97-129: MATLAB-style reflection, border samples used twice ([a, b, c, d] -> [a, a, b, c, d, d])."""
    n = x.size(dim)
    pre = x.narrow(dim, 0, pad_pre).flip(dim) if pad_pre else None
    post = x.narrow(dim, n - pad_post, pad_post).flip(dim) if pad_post else None
    return torch.cat([t for t in (pre, x, post) if t is not None], dim)


def _resize_1d(x, dim, size, scale):
    """core.py:268-337 for kernel='cubic', antialiasing=True, padding_type='reflect'.  x: [N, 1, H, W]."""
    if scale == 1:
        return x
    kernel_size = 4
    if scale < 1:
        antialiasing_factor = scale
        kernel_size = math.ceil(kernel_size / antialiasing_factor)
    else:
        antialiasing_factor = 1
    kernel_size += 2
    pos = torch.linspace(0, size - 1, steps=size, dtype=x.dtype)
    pos = (pos + 0.5) / scale - 0.5
    base = pos.floor() - (kernel_size // 2) + 1
    dist = pos - base
    buffer_pos = dist.new_zeros(kernel_size, len(dist))            # get_weight, core.py:172-193
    for idx, buffer_sub in enumerate(buffer_pos):
        buffer_sub.copy_(dist - idx)
    buffer_pos *= antialiasing_factor
    weight = _cubic_contribution(buffer_pos)
    weight /= weight.sum(dim=0, keepdim=True)
    base = base.long()                                             # get_padding, core.py:148-169
    r_min, r_max = int(base.min()), int(base.max()) + kernel_size - 1
    pad_pre = -r_min if r_min <= 0 else 0
    base = base + pad_pre
    pad_post = r_max - x.size(dim) + 1 if r_max >= x.size(dim) else 0
    x_pad = _reflect_pad(x, dim, pad_pre, pad_post)
    # unfold + "subsampling first" (core.py:196-211, 322-331): sample[:, k, i] = x_pad[base[i] + k]
    taps = base[None, :] + torch.arange(kernel_size)[:, None]     # [K, size]
    if dim == 2 or dim == -2:
        sample = x_pad[:, 0][:, taps, :]                           # [N, K, size, W]
        weight = weight.view(1, kernel_size, size, 1)
    else:
        sample = x_pad[:, 0][:, :, taps].permute(0, 2, 1, 3).contiguous()   # [N, K, H, size], the layout the reference sums over
        weight = weight.view(1, kernel_size, 1, size)
    x = sample * weight
    return x.sum(dim=1, keepdim=True)


def imresize_cubic(x, sizes):
    """core.py:364-432 with sizes=(H, W), kernel='cubic': rows (dim -2) first, then columns (core.py:417-418)."""
    b, c, h, w = x.shape
    x = x.reshape(-1, 1, h, w).float()
    x = _resize_1d(x, -2, sizes[0], sizes[0] / h)
    x = _resize_1d(x, -1, sizes[1], sizes[1] / w)
    return x.view(b, c, sizes[0], sizes[1])


def base_layer(x_el_padded, sizes):
    """test.py:199"""
    return imresize_cubic(x_el_padded, sizes).clamp_(0, 1)


def psnr(img1, img2):
    """test.py:115-118"""
    mse = torch.mean((img1 - img2) ** 2)
    return (10 * torch.log10(1.0 / mse)).item()
