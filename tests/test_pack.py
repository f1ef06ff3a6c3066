"""Host-side weight packing of the split-fp16 engine (lssvc_b200/ops.py): hi/lo split, power-of-two scaling and the
accumulator-truncation compensation — CPU only."""
import math

import torch

from lssvc_b200 import ops


def _unpack(pc):
    packed, cin16, acc_scale = pc.weight_h2()
    w = (packed[:, 0].float() + packed[:, 1].float()) * acc_scale      # [taps][n_pad][cin16]
    return w, cin16


def test_acc_comp_formula():
    assert float(ops.acc_comp(0)) == 1.0
    for steps in (4, 36, 196):
        assert math.isclose(float(ops.acc_comp(steps)), 1.0 + (0.264 * steps + 0.6) * 2.0 ** -24, rel_tol=0, abs_tol=6e-8)
    t = ops.acc_comp(torch.tensor([0, 36]))
    assert float(t[0]) == 1.0 and float(t[1]) > 1.0


def test_split_weights_reconstruct_the_compensated_weights():
    g = torch.Generator().manual_seed(0)
    w = torch.randn(8, 20, 3, 3, generator=g)
    pc = ops.PackedConv(w, None, src_channels=[(20, 24)])
    got, cin16 = _unpack(pc)
    assert cin16 == 32 and got.shape == (9, 16, 32)
    # 20 real channels = two 16-channel slices with data, 9 taps -> 18 accumulation steps for every real output channel
    comp = float(ops.acc_comp(18))
    ref = w.permute(2, 3, 0, 1).reshape(9, 8, 20) * comp
    assert (got[:, :8, :20] - ref).abs().max().item() < 2.0 ** -21 * w.abs().max().item()      # hi + lo keeps ~22 bits
    assert got[:, 8:].abs().max().item() == 0 and got[:, :, 20:].abs().max().item() == 0       # channel padding stays zero
    # the same layer fed by quantised symbols is left alone
    got_exact, _ = _unpack(ops.PackedConv(w, None, src_channels=[(20, 24)], exact_in=True))
    ref = w.permute(2, 3, 0, 1).reshape(9, 8, 20)
    assert (got_exact[:, :8, :20] - ref).abs().max().item() < 2.0 ** -21 * w.abs().max().item()
    ratio = (got[:, :8, :20] / got_exact[:, :8, :20])[ref.abs() > 0.5]
    assert (ratio - comp).abs().max().item() < 1e-6


def test_empty_taps_do_not_count_as_accumulation_steps():
    """Sub-pixel-decomposed ConvTranspose2d: an output phase only uses 1, 2 or 4 of the 9 taps; all-zero (tap, slice) blocks
    add exact zeros to the accumulator and must not be compensated for."""
    g = torch.Generator().manual_seed(1)
    w = torch.zeros(4, 32, 3, 3)
    w[0, :, 1, 1] = torch.randn(32, generator=g)                     # 1 tap  x 2 slices = 2 steps
    w[1, :, 1, 1:] = torch.randn(32, 2, generator=g)                 # 2 taps x 2 slices = 4 steps
    w[2, :16, 1:, 1:] = torch.randn(16, 2, 2, generator=g)           # 4 taps x 1 slice  = 4 steps
    w[3] = torch.randn(32, 3, 3, generator=g)                        # 9 taps x 2 slices = 18 steps
    got, _ = _unpack(ops.PackedConv(w, None))
    base, _ = _unpack(ops.PackedConv(w, None, exact_in=True))
    for ch, steps in ((0, 2), (1, 4), (2, 4), (3, 18)):
        m = base[:, ch].abs() > 0.5
        ratio = (got[:, ch][m] / base[:, ch][m]).mean().item()
        assert abs(ratio - float(ops.acc_comp(steps))) < 2e-7, (ch, steps, ratio)


def test_laplace_pair_tile_interleaves_scale_and_mean_per_channel_tile():
    """Entropy epilogues over several channel tiles: the channel tile conv_hs picks, and the pack-time permutation that
    puts a latent channel's scale and mean into the same tile ([scale of Ct channels | their means] per tile of 2 Ct)."""
    assert ops.hs_channel_tile(128) == 128 and ops.hs_channel_tile(192) == 96 and ops.hs_channel_tile(256) == 128
    assert ops.hs_channel_tile(1152) == 128 and ops.hs_channel_tile(144) == 48
    assert ops.laplace_pair_tile(128) == 0                      # one tile: natural (scale | mean) order
    assert ops.laplace_pair_tile(192) == 96 and ops.laplace_pair_tile(256) == 128
    assert ops.laplace_pair_tile(192, engine="simt") == 0       # nothing to interleave for an engine without the epilogue
    assert ops.laplace_pair_tile(144) == 0                      # tile 48: its halves are not 16-channel aligned -> not fused
    C = 96
    w = torch.arange(2 * C, dtype=torch.float32).reshape(2 * C, 1, 1, 1).repeat(1, 16, 1, 1)      # every weight of channel c is c
    b = torch.arange(2 * C, dtype=torch.float32)
    pc = ops.PackedConv(w, b, pad=0, exact_in=True, pair_tile=96)
    order = pc.bias[:2 * C].to(torch.int64).tolist()
    for t in range(2):
        tile = order[96 * t:96 * (t + 1)]
        assert tile[:48] == list(range(48 * t, 48 * t + 48))                   # scales of latent channels 48 t ..
        assert tile[48:] == list(range(C + 48 * t, C + 48 * t + 48))           # ... and their means
    assert torch.equal(pc.weight[0, :, 0].to(torch.int64), pc.bias.to(torch.int64))   # weights permuted with the bias
    plain = ops.PackedConv(w, b, pad=0, exact_in=True)
    assert plain.bias[:2 * C].tolist() == list(range(2 * C)) and plain.pair_tile == 0
