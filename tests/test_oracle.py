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
