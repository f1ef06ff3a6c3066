"""Deterministic synthetic inputs: frames (smooth seeded field under global motion + noise) and the padded
two-layer geometry of the reference (src/utils/common.py:48-86).  There is no dataset or checkpoint access,
so tests and bench.py feed both this implementation and the oracle from here."""
import math

import torch
import torch.nn.functional as F


def round_to_even(x):
    t = int(x)
    return t + 1 if t % 2 else t


def interlayer_padding(H_HR, W_HR, ratio=2.0):
    """get_interlayer_padding (common.py:48-86): EL padded so H, W are multiples of 64 and of 64*ratio."""
    def pad_dim(v):
        i = 0
        while True:
            p = 64 + 32 * i
            t = (v + p - 1) // p * p
            if t % 64 == 0 and t % (64 * ratio) == 0:
                return t
            i += 1
    H, W = pad_dim(H_HR), pad_dim(W_HR)
    return {"HR_padded_size": (H, W), "LR_padded_size": (int(H / ratio), int(W / ratio)),
            "HR_size": (H_HR, W_HR), "LR_size": (round_to_even(H_HR / ratio), round_to_even(W_HR / ratio))}


def _smooth_field(C, H, W, gen, octaves=4):
    out = torch.zeros(1, C, H, W)
    amp = 1.0
    for o in range(octaves):
        h, w = max(2, H >> (octaves + 1 - o)), max(2, W >> (octaves + 1 - o))
        out += amp * F.interpolate(torch.randn(1, C, h, w, generator=gen), size=(H, W), mode="bicubic",
                                   align_corners=False)
        amp *= 0.5
    out = out - out.amin()
    return out / out.amax().clamp_min(1e-6)


def make_sequence(H, W, n_frames, seed=0, ratio=2.0, noise=0.003):
    """Returns a list of (x_bl [1,3,H/ratio,W/ratio], x_el [1,3,H,W]) fp32 CPU tensors in [0,1].
    H, W are the PADDED enhancement-layer size."""
    gen = torch.Generator().manual_seed(1000 * seed + 17)
    margin = 4 * n_frames + 8
    canvas = _smooth_field(3, H + 2 * margin, W + 2 * margin, gen)
    vx = float(torch.rand(1, generator=gen)) * 3.0 - 1.5
    vy = float(torch.rand(1, generator=gen)) * 2.0 - 1.0
    ys = torch.arange(H, dtype=torch.float32)
    xs = torch.arange(W, dtype=torch.float32)
    frames = []
    Hc, Wc = canvas.shape[2], canvas.shape[3]
    for t in range(n_frames):
        gx = (xs + margin + vx * t) / (Wc - 1) * 2 - 1
        gy = (ys + margin + vy * t) / (Hc - 1) * 2 - 1
        grid = torch.stack(torch.meshgrid(gy, gx, indexing="ij")[::-1], dim=-1).unsqueeze(0)
        x_el = F.grid_sample(canvas, grid, mode="bilinear", padding_mode="border", align_corners=True)
        x_el = (x_el + noise * torch.randn(x_el.shape, generator=gen)).clamp_(0, 1)
        x_bl = F.interpolate(x_el, size=(int(H / ratio), int(W / ratio)), mode="bicubic", align_corners=False,
                             antialias=True).clamp_(0, 1)
        frames.append((x_bl.contiguous(), x_el.contiguous()))
    return frames
