"""The oracle restatement against the golden vectors generated from the unmodified reference
(tools/make_golden.py; there reference == oracle bit for bit).  On another CPU model the MKL-DNN kernels may round
differently, hence small tolerances instead of exact equality."""
import os

import pytest
import torch

from lssvc_b200 import nets, synth
from oracle import lssvc_oracle as orc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def golden():
    return torch.load(os.path.join(GOLD, "forward_128.pt"))


def _sub(t, step=8):
    return t[:, :, ::step, ::step]


def test_oracle_reproduces_reference_golden(golden):
    torch.set_num_threads(8)
    H, W, seed = golden["H"], golden["W"], golden["seed"]
    sd_i = nets.ParamBag(nets.intra_ss_spec(), seed=seed, gains=nets.model_gains("I")).state_dict()
    sd_p = nets.ParamBag(nets.lssvc_spec(), seed=seed + 1, gains=nets.model_gains("P")).state_dict()
    frames = synth.make_sequence(H, W, 3, seed=seed)
    dpb = None
    with torch.no_grad():
        for t, (x_bl, x_el) in enumerate(frames):
            g = golden["frames"][t]
            if t == 0:
                o = orc.intra_ss(sd_i, x_bl, x_el, (H, W))
                x_bl_hat, x_el_hat, feat = o["x_hat_bl"], o["x_hat_el"], o["feature_el"]
                sym_el = torch.round(o["y"] - o["means"])
                idx_el = orc.build_indexes_image(o["scales"])
                dpb = {"ref_frame_bl": x_bl_hat, "ref_frame_el": x_el_hat, "ref_feature_bl": None, "ref_feature_el": feat}
            else:
                o = orc.lssvc(sd_p, x_bl, x_el, dpb, (H, W), 2.0)
                dpb = o["dpb"]
                x_bl_hat, x_el_hat, feat = dpb["ref_frame_bl"], dpb["ref_frame_el"], dpb["ref_feature_el"]
                sym_el = o["four_part"]["y_q"]
                idx_el = orc.build_indexes_video(o["four_part"]["scales_hat"])
                assert (_sub(o["mv_hat"]) - g["mv_hat"]).abs().max() < 1e-3
                assert (o["mv_y_q"] == g["sym_mv"].float()).float().mean() > 0.999
            assert abs(o["bit_bl"] - g["bit_bl"]) / g["bit_bl"] < 1e-4
            assert abs(o["bit_el"] - g["bit_el"]) / g["bit_el"] < 1e-4
            assert (_sub(x_bl_hat, 4) - g["x_hat_bl"]).abs().max() < 1e-4
            assert (_sub(x_el_hat) - g["x_hat_el"]).abs().max() < 1e-4
            assert (_sub(feat) - g["feature_el"]).abs().max() < 1e-3
            assert (sym_el == g["sym_el"].float()).float().mean() > 0.999
            assert (idx_el == g["index_el"].int()).float().mean() > 0.999
            # the symbols are not degenerate (H1): enough non-zeros, many CDF rows in use
            assert (sym_el != 0).float().mean() > 0.2
            assert idx_el.unique().numel() >= 20
            dpb["ref_frame_bl"] = dpb["ref_frame_bl"].clamp_(0, 1)
            dpb["ref_frame_el"] = dpb["ref_frame_el"].clamp_(0, 1)


def test_four_part_masks_partition_the_tensor():
    masks = orc._masks(6, 10, torch.float32)
    assert torch.equal(sum(masks), torch.ones(1, 1, 6, 10))
    for step in range(4):
        assert sorted(orc.MASK_ORDER[step]) == [0, 1, 2, 3]
    for k in range(4):   # every quarter visits every checkerboard phase exactly once
        assert sorted(orc.MASK_ORDER[s][k] for s in range(4)) == [0, 1, 2, 3]


def test_interlayer_padding_matches_survey_sizes():
    # SURVEY.md §8 size table ([probe] of common.py:48-86)
    assert synth.interlayer_padding(320, 512)["HR_padded_size"] == (384, 512)
    assert synth.interlayer_padding(320, 512)["LR_padded_size"] == (192, 256)
    assert synth.interlayer_padding(1080, 1920)["HR_padded_size"] == (1152, 1920)
    assert synth.interlayer_padding(1080, 1920)["LR_padded_size"] == (576, 960)
    assert synth.interlayer_padding(2160, 3840)["HR_padded_size"] == (2176, 3840)
    assert synth.interlayer_padding(2160, 3840)["LR_padded_size"] == (1088, 1920)


def test_oracle_pad_size_reproduces_reference_golden():
    """Non-zero pad_size (get_depadded_feature, LSSVC_net.py:271-282 / 453-456, IntraSS.py:124-147): the base layer is coded on a
    frame with 64 more padding pixels than the enhancement layer's size / 2 and cropped by pad_size = (0, -64, 0, -64) before the
    inter-layer resamplers.  Fixture: tools/make_golden_padsize.py (reference == oracle bit for bit there)."""
    g = torch.load(os.path.join(GOLD, "padsize_128.pt"))
    torch.set_num_threads(8)
    H, W, PAD = g["H"], g["W"], tuple(g["pad_size"])
    sd_i = nets.ParamBag(nets.intra_ss_spec(), seed=0, gains=nets.model_gains("I")).state_dict()
    sd_p = nets.ParamBag(nets.lssvc_spec(), seed=1, gains=nets.model_gains("P")).state_dict()
    frames = [(torch.nn.functional.pad(b, (0, 64, 0, 64), mode="replicate"), e) for b, e in synth.make_sequence(H, W, 2, seed=g["seed"])]
    with torch.no_grad():
        o = orc.intra_ss(sd_i, *frames[0], (H, W), PAD)
        f = g["frames"][0]
        assert tuple(o["x_hat_bl"].shape[2:]) == (128, 128) and tuple(o["x_hat_el"].shape[2:]) == (H, W)
        for k in ("bit_bl", "bit_el"):
            assert abs(o[k] - f[k]) / f[k] < 1e-4
        assert (_sub(o["x_hat_bl"], 4) - f["x_hat_bl"]).abs().max() < 1e-4
        assert (_sub(o["x_hat_el"]) - f["x_hat_el"]).abs().max() < 1e-4
        assert (_sub(o["feature_el"]) - f["feature_el"]).abs().max() < 1e-3
        # a different pad_size must give a different enhancement layer (the argument is not ignored) ...
        o0 = orc.intra_ss(sd_i, frames[0][0][:, :, :64, :64], frames[0][1], (H, W))
        assert (o0["x_hat_el"] - o["x_hat_el"]).abs().max() > 1e-3
        dpb = {"ref_frame_bl": o["x_hat_bl"].clamp(0, 1), "ref_frame_el": o["x_hat_el"].clamp(0, 1), "ref_feature_bl": None,
               "ref_feature_el": o["feature_el"]}
        o = orc.lssvc(sd_p, *frames[1], dpb, (H, W), 2.0, PAD)
        f = g["frames"][1]
        for k in ("bit_bl", "bit_el"):
            assert abs(o[k] - f[k]) / f[k] < 1e-4
        assert (_sub(o["dpb"]["ref_frame_bl"], 4) - f["ref_frame_bl"]).abs().max() < 1e-4
        assert (_sub(o["dpb"]["ref_frame_el"]) - f["ref_frame_el"]).abs().max() < 1e-4
        assert (_sub(o["dpb"]["ref_feature_el"]) - f["ref_feature_el"]).abs().max() < 1e-3
        assert (_sub(o["mv_hat"]) - f["mv_hat"]).abs().max() < 1e-3
        assert (o["four_part"]["y_q"] == f["sym_el"].float()).float().mean() > 0.999
