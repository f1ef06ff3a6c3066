"""GOP-wise sharding of a coding job over ranks.

An I-frame rebuilds the decoded-picture buffer from scratch (test.py:219-223), so every (sequence, GOP) pair is an
independent work unit; inside a unit the P-chain is strictly serial.  The reference spreads *sequences* over a
process pool (test.py:644-656, 739-743); here the finer (sequence, GOP) units are dealt round-robin to one process
per GPU.  There is no data-path exchange: the only collective is the gather of the per-frame rate/distortion rows at
the end of the run.
"""
from collections import namedtuple

import torch

Unit = namedtuple("Unit", "seq gop first_frame n_frames")
STAT_COLUMNS = ("seq", "frame", "is_intra", "bits_bl", "bits_el", "sse_bl", "sse_el")


def work_units(n_seq, n_frames, gop_size):
    """All (sequence, GOP) units of a job, in (seq, gop) order.  The last GOP of a sequence may be short."""
    assert n_seq >= 0 and n_frames >= 0 and gop_size >= 1
    units = []
    for s in range(n_seq):
        for g, first in enumerate(range(0, n_frames, gop_size)):
            units.append(Unit(s, g, first, min(gop_size, n_frames - first)))
    return units


def shard(units, world_size, rank):
    """Static round-robin deal: unit i goes to rank i % world_size (all units cost the same: 1 I + (gop-1) P)."""
    assert 0 <= rank < world_size
    return units[rank::world_size]


def frames_of(unit):
    """(frame_idx, is_intra) of every frame of a unit, in coding order (test.py:219: frame_idx % gop_size == 0)."""
    return [(unit.first_frame + i, i == 0) for i in range(unit.n_frames)]


def gather_stats(rows, dist=None, device="cpu"):
    """rows: this rank's list of STAT_COLUMNS tuples.  Returns the [total, 7] float64 table of all ranks sorted by
    (seq, frame), identical on every rank.  dist: an initialised torch.distributed module, or None for one rank."""
    t = torch.tensor(rows, dtype=torch.float64, device=device).reshape(-1, len(STAT_COLUMNS))
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        world = dist.get_world_size()
        n = torch.tensor([t.shape[0]], dtype=torch.int64, device=device)
        counts = [torch.zeros_like(n) for _ in range(world)]
        dist.all_gather(counts, n)
        counts = [int(c.item()) for c in counts]
        cap = max(counts) if counts else 0
        padded = torch.zeros(cap, len(STAT_COLUMNS), dtype=torch.float64, device=device)
        padded[: t.shape[0]] = t
        parts = [torch.empty_like(padded) for _ in range(world)]
        dist.all_gather(parts, padded)
        t = torch.cat([p[:c] for p, c in zip(parts, counts)], 0)
    if t.shape[0]:
        order = torch.argsort(t[:, 0] * 1e9 + t[:, 1], stable=True)
        t = t[order]
    return t


def summarize(table, pixels_el, pixels_bl):
    """Mean bpp per layer and PSNR per layer from a gathered table (test.py:253-263 semantics, RGB MSE)."""
    if table.shape[0] == 0:
        return {"frames": 0}
    out = {"frames": int(table.shape[0]), "bpp_bl": float(table[:, 3].mean()) / pixels_bl,
           "bpp_el": float(table[:, 4].mean()) / pixels_el}
    for name, col, px in (("psnr_bl", 5, pixels_bl), ("psnr_el", 6, pixels_el)):
        mse = (table[:, col] / (3.0 * px)).clamp_min(1e-12)
        out[name] = float((10.0 * torch.log10(1.0 / mse)).mean())
    return out
