"""Front end of the reference's frame loop on the GPU: the per-frame work test.py does on the host before a
frame reaches the models and after it leaves them.

    reference (test.py)                                              here
    ------------------------------------------------------------    -------------------------------------------------
    get_interlayer_padding(H, W, ratio)       common.py:48-86        get_interlayer_padding (same dict)
    YUVReader.read_one_frame -> ycbcr420_to_rgb -> F.pad             FrontEnd.rgb_from_yuv420(y_u8, uv_u8)
        video_reader.py:139-155, functional.py:42-58, test.py:185-197
    imresize(x, sizes=..., kernel='cubic').clamp_(0, 1)              imresize(x, sizes=...) / FrontEnd.base_layer(x)
        core.py:364-432, test.py:199
    PSNR(a, b)                                test.py:115-118        psnr(a, b)

Tensors are NCHW fp32 on the CUDA device, as in test.py.  The kernels are lssvc_yuv420_to_rgb / lssvc_resample_1d /
lssvc_sse_flat (csrc/frontend.cu); there is no CPU fallback."""
import math

import torch

from . import _lib
from .ops import _ptr, _stream


def round_to_even(x):
    """common.py:40-45"""
    tmp = int(x)
    return tmp + 1 if tmp % 2 != 0 else tmp


def get_interlayer_padding(H_HR, W_HR, ratio):
    """common.py:48-86: EL padded so that H, W are multiples of 64 and of 64 * ratio; padding on the right / bottom."""
    def padded(n):
        i = 0
        while True:
            p = 64 + 32 * i
            t = (n + p - 1) // p * p
            if t % 64 == 0 and t % (64 * ratio) == 0:
                return t
            i += 1
    new_H, new_W = padded(H_HR), padded(W_HR)
    H_LR, W_LR = round_to_even(H_HR / ratio), round_to_even(W_HR / ratio)
    new_H_LR, new_W_LR = int(new_H / ratio), int(new_W / ratio)
    return {"P_LR": (0, new_W_LR - W_LR, 0, new_H_LR - H_LR), "P_HR": (0, new_W - W_HR, 0, new_H - H_HR),
            "LR_padded_size": (new_H_LR, new_W_LR), "HR_padded_size": (new_H, new_W), "LR_size": (H_LR, W_LR),
            "HR_size": (H_HR, W_HR)}


def _cubic(x, a=-0.5):
    """core.py:40-55"""
    ax = x.abs()
    ax2 = ax * ax
    ax3 = ax * ax2
    c01 = ((a + 2) * ax3 - (a + 3) * ax2 + 1) * ax.le(1).to(x.dtype)
    c12 = ((a * ax3) - (5 * a * ax2) + (8 * a * ax) - (4 * a)) * torch.logical_and(ax.gt(1), ax.le(2)).to(x.dtype)
    return c01 + c12


def resize_plan(in_size, out_size, antialiasing=True):
    """Weights and source indices of resize_1d (core.py:268-337) for one axis: (w [out][K] fp32, taps [out][K] int32, K).
    The arithmetic is the reference's, in fp32 on the host: pos = (i + .5) / scale - .5, base = floor(pos) - K // 2 + 1,
    w_k = cubic((pos - base - k) * af) normalised over k; the reflect padding of core.py:97-129 (border samples used twice)
    is resolved into indices: j < 0 -> -j - 1, j >= n -> 2n - 1 - j."""
    scale = out_size / in_size
    K = 4
    if antialiasing and scale < 1:
        af = scale
        K = math.ceil(K / af)
    else:
        af = 1
    K += 2
    pos = torch.linspace(0, out_size - 1, steps=out_size, dtype=torch.float32)
    pos = (pos + 0.5) / scale - 0.5
    base = pos.floor() - (K // 2) + 1
    dist = pos - base
    buf = torch.stack([dist - k for k in range(K)], 0)
    buf *= af
    w = _cubic(buf)
    w /= w.sum(dim=0, keepdim=True)
    j = base.long()[None, :] + torch.arange(K)[:, None]
    j = torch.where(j < 0, -j - 1, j)
    j = torch.where(j >= in_size, 2 * in_size - 1 - j, j)
    assert int(j.min()) >= 0 and int(j.max()) < in_size, "resize_plan: kernel wider than the reflected image"
    return w.t().contiguous(), j.t().to(torch.int32).contiguous(), K


_PLANS = {}


def _plan(in_size, out_size, device):
    key = (in_size, out_size, str(device))
    if key not in _PLANS:
        w, taps, K = resize_plan(in_size, out_size)
        _PLANS[key] = (w.to(device), taps.to(device), K)
    return _PLANS[key]


def _require_cuda(t, what):
    if not t.is_cuda:
        raise _lib.LssvcError(f"{what}: lssvc_b200.frontend runs on a CUDA (sm_100a) device only — there is no CPU fallback")


def imresize(x, scale=None, sizes=None, kernel="cubic", antialiasing=True, clamp=False):
    """core.py:364-432 for kernel='cubic', padding_type='reflect': rows first, then columns.  x: [B, C, H, W] (or fewer
    dims) fp32 CUDA tensor.  clamp=True fuses the .clamp_(0, 1) test.py:199 applies to the result."""
    if kernel != "cubic":
        raise NotImplementedError("imresize: only the bicubic kernel of the coding path (test.py:199)")
    if (scale is None) == (sizes is None):
        raise ValueError("One of scale or sizes must be specified!" if scale is None else "Please specify scale or sizes to avoid conflict!")
    if not antialiasing:
        raise NotImplementedError("imresize: antialiasing=False is not on the coding path")
    _require_cuda(x, "imresize")
    shape = x.shape
    h, w = shape[-2], shape[-1]
    if sizes is None:
        sizes = (math.ceil(h * scale), math.ceil(w * scale))
    planes = x.reshape(-1, h, w).to(torch.float32).contiguous()
    C = planes.shape[0]
    lib = _lib.load()
    cur, ch, cw = planes, h, w
    for dim, n_out in ((0, sizes[0]), (1, sizes[1])):
        n_in = ch if dim == 0 else cw
        if n_out == n_in:          # "Identity case" core.py:289-290
            continue
        wts, taps, K = _plan(n_in, n_out, x.device)
        last = dim == 1 or sizes[1] == cw
        out = torch.empty((C, n_out, cw) if dim == 0 else (C, ch, n_out), dtype=torch.float32, device=x.device)
        _lib.check(lib.lssvc_resample_1d(_ptr(cur), C, ch, cw, dim, _ptr(wts), _ptr(taps), K, n_out, _ptr(out),
                                         1 if (clamp and last) else 0, _stream()), "resample_1d")
        cur = out
        if dim == 0:
            ch = n_out
        else:
            cw = n_out
    if clamp and cur is planes:
        cur = planes.clamp(0, 1)
    return cur.reshape(*shape[:-2], ch, cw)


def psnr(a, b):
    """test.py:115-118: 10 log10(1 / mean((a - b)^2)); returns a Python float (one device->host read)."""
    _require_cuda(a, "psnr")
    assert a.shape == b.shape
    a, b = a.to(torch.float32).contiguous(), b.to(torch.float32).contiguous()
    out = torch.zeros(1, dtype=torch.float64, device=a.device)
    _lib.check(_lib.load().lssvc_sse_flat(_ptr(a), _ptr(b), a.numel(), _ptr(out), _stream()), "sse_flat")
    mse = out.item() / a.numel()
    return 10.0 * math.log10(1.0 / mse) if mse > 0 else float("inf")


class FrontEnd:
    """Per-sequence front end: sizes and padding of get_interlayer_padding, YUV 4:2:0 -> padded RGB, EL -> BL resize."""

    def __init__(self, height, width, ratio=2, device="cuda"):
        self.height, self.width, self.ratio = height, width, ratio
        self.device = torch.device(device)
        self.padding = get_interlayer_padding(height, width, ratio)
        self.el_size = self.padding["HR_padded_size"]
        self.bl_size = self.padding["LR_padded_size"]

    def rgb_from_yuv420(self, y, uv):
        """y: uint8 [H, W] (or [1, H, W]), uv: uint8 [2, H/2, W/2], CUDA tensors (e.g. non_blocking copies of the pinned file
        buffer) -> x_EL_padded [1, 3, Hp, Wp] fp32 (test.py:185-197)."""
        _require_cuda(y, "rgb_from_yuv420")
        H, W = self.height, self.width
        assert y.dtype == torch.uint8 and uv.dtype == torch.uint8 and y.numel() == H * W and uv.numel() == H * W // 2
        Hp, Wp = self.el_size
        out = torch.empty(1, 3, Hp, Wp, dtype=torch.float32, device=y.device)
        _lib.check(_lib.load().lssvc_yuv420_to_rgb(_ptr(y.contiguous()), _ptr(uv.contiguous()), H, W, _ptr(out), Hp, Wp, _stream()),
                   "yuv420_to_rgb")
        return out

    def base_layer(self, x_el_padded):
        """imresize(x_EL_padded, sizes=LR_padded_size, kernel='cubic').clamp_(0, 1)   test.py:199"""
        return imresize(x_el_padded, sizes=self.bl_size, clamp=True)
