"""Non-integer layer ratio (x1_5 of recommend_test_config.json / test.py:27-33,693): I-frame + first P-frame at
EL 384x384 / BL 256x256 (the padding rule of common.py:48-86 makes H, W multiples of 64 and of 64 * 1.5 = 96), against the
oracle, teacher-forced on its symbols, with the north-star tolerances.  Exercises the resamplers (MvResampler,
TextureResampler, LayerPriorResampler, lssvc_modules.py:339-428 / layers.py:258-285) away from the exact x2 case."""
import pytest
import torch

pytestmark = pytest.mark.gpu

H = W = 384
RATIO = 1.5
HB, WB = 256, 256


def _run(net, call, force):
    try:
        net._debug, net._force, net._force_flips = {}, force, {}
        r = call()
        flips = dict(net._force_flips)
    finally:
        net._debug = net._force = None
    return r, flips


def _check(name, got, ref, tol):
    d = (got.cpu() - ref).abs().max().item()
    print(f"  {name:14s} max|d| {d:.3e}")
    assert d < tol, f"{name}: max|d| {d:.3e} >= {tol}"


def test_ratio_1_5_against_oracle(cuda_device):
    from lssvc_b200 import IntraSS, LSSVC_extend, frontend, synth
    from oracle import lssvc_oracle as orc
    pad = frontend.get_interlayer_padding(H, W, RATIO)
    assert pad["HR_padded_size"] == (H, W) and pad["LR_padded_size"] == (HB, WB)
    torch.set_num_threads(8)
    net_i, net_p = IntraSS(seed=0), LSSVC_extend(seed=1)
    sd_i = {k: v.clone() for k, v in net_i.state_dict().items()}
    sd_p = {k: v.clone() for k, v in net_p.state_dict().items()}
    net_i.to(cuda_device)
    net_p.to(cuda_device)
    for n in (net_i, net_p):
        n.set_scale_information(RATIO, (H, W), (0, 0, 0, 0))
    frames = synth.make_sequence(H, W, 2, seed=2, ratio=RATIO)
    assert tuple(frames[0][0].shape[2:]) == (HB, WB)
    dev = cuda_device
    x_bl, x_el = frames[0]
    with torch.no_grad():
        o = orc.intra_ss(sd_i, x_bl, x_el, (H, W))
    q_ref = {"bl_z_hat": o["bl"]["z_hat"], "bl_y_q": torch.round(o["bl"]["y"] - o["bl"]["means"]), "z_hat": o["z_hat"],
             "y_q": torch.round(o["y"] - o["means"])}
    r, flips = _run(net_i, lambda: net_i.encode_decode(x_bl.to(dev), x_el.to(dev), None, None, HB, WB, H, W), q_ref)
    total = sum(v.numel() for v in q_ref.values())
    print(f"x1.5 I-frame: bits {r['bit_bl']:.1f}/{r['bit_el']:.1f} oracle {o['bit_bl']:.1f}/{o['bit_el']:.1f}; "
          f"{sum(flips.values())} of {total} symbols differ {flips}")
    assert sum(flips.values()) <= 1e-4 * total
    _check("x_hat_bl", r["x_hat_bl"], o["x_hat_bl"], 1e-3)
    _check("x_hat_el", r["x_hat_el"], o["x_hat_el"], 1e-3)
    for k in ("bit_bl", "bit_el"):
        assert abs(r[k] - o[k]) / o[k] < 1e-3

    dpb = {"ref_frame_bl": o["x_hat_bl"].clamp(0, 1), "ref_frame_el": o["x_hat_el"].clamp(0, 1), "ref_feature_bl": None,
           "ref_feature_el": o["feature_el"]}
    x_bl, x_el = frames[1]
    with torch.no_grad():
        o = orc.lssvc(sd_p, x_bl, x_el, dpb, (H, W), RATIO)
    q_ref = {"bl_mv_z_hat": o["bl"]["mv_z_hat"], "bl_mv_y_q": o["bl"]["mv_y_q"], "bl_z_hat": o["bl"]["z_hat"],
             "bl_y_q": o["bl"]["y_q"], "mv_z_hat": o["mv_z_hat"], "mv_y_q": o["mv_y_q"], "z_hat": o["z_hat"],
             "y_q": o["four_part"]["y_q"]}
    dpb_dev = {k: (None if v is None else v.to(dev)) for k, v in dpb.items()}
    r, flips = _run(net_p, lambda: net_p.encode_decode(x_bl.to(dev), x_el.to(dev), dpb_dev, None, None, W, H, WB, HB), q_ref)
    total = sum(v.numel() for v in q_ref.values())
    print(f"x1.5 P-frame: bits {r['bit_bl']:.1f}/{r['bit_el']:.1f} oracle {o['bit_bl']:.1f}/{o['bit_el']:.1f}; "
          f"{sum(flips.values())} of {total} symbols differ {flips}")
    assert sum(flips.values()) <= 1e-4 * total
    for k in ("mv_hat", "warp_frame"):
        _check(k, r[k], o[k], 1e-3)
    for k in ("ref_frame_bl", "ref_frame_el"):
        _check(k, r["dpb"][k], o["dpb"][k], 1e-3)
    for k in ("bit_bl", "bit_el"):
        assert abs(r[k] - o[k]) / o[k] < 1e-3


def test_pad_size_depadding_against_oracle(cuda_device):
    """Non-zero pad_size (get_depadded_feature, LSSVC_net.py:271-282 / 453-456, IntraSS.py:124-147): the base layer is coded on a
    frame with 64 more padding pixels (right / bottom) than the enhancement layer's size / ratio and its tensors are cropped by
    pad_size = (0, -64, 0, -64) (/ 16 for the latent) before the inter-layer resamplers.  Same contract as above against the
    oracle (== the reference bit for bit with this pad_size: tools/check_oracle_vs_reference.py 3 2 128 64), I + P frame."""
    from lssvc_b200 import IntraSS, LSSVC_extend, synth
    from oracle import lssvc_oracle as orc
    He = We = 128
    PAD = (0, -64, 0, -64)
    torch.set_num_threads(8)
    net_i, net_p = IntraSS(seed=0), LSSVC_extend(seed=1)
    sd_i = {k: v.clone() for k, v in net_i.state_dict().items()}
    sd_p = {k: v.clone() for k, v in net_p.state_dict().items()}
    net_i.to(cuda_device)
    net_p.to(cuda_device)
    for n in (net_i, net_p):
        n.set_scale_information(2.0, (He, We), PAD)
    frames = [(torch.nn.functional.pad(b, (0, 64, 0, 64), mode="replicate"), e) for b, e in synth.make_sequence(He, We, 2, seed=3)]
    dev = cuda_device
    x_bl, x_el = frames[0]
    assert tuple(x_bl.shape[2:]) == (128, 128)
    with torch.no_grad():
        o = orc.intra_ss(sd_i, x_bl, x_el, (He, We), PAD)
    q_ref = {"bl_z_hat": o["bl"]["z_hat"], "bl_y_q": torch.round(o["bl"]["y"] - o["bl"]["means"]), "z_hat": o["z_hat"],
             "y_q": torch.round(o["y"] - o["means"])}
    r, flips = _run(net_i, lambda: net_i.encode_decode(x_bl.to(dev), x_el.to(dev), None, None, 128, 128, He, We), q_ref)
    total = sum(v.numel() for v in q_ref.values())
    print(f"pad_size I-frame: bits {r['bit_bl']:.1f}/{r['bit_el']:.1f} oracle {o['bit_bl']:.1f}/{o['bit_el']:.1f}; "
          f"{sum(flips.values())} of {total} symbols differ {flips}")
    assert sum(flips.values()) <= max(1, 1e-4 * total)
    assert tuple(r["x_hat_bl"].shape[2:]) == (128, 128) and tuple(r["x_hat_el"].shape[2:]) == (He, We)
    _check("x_hat_bl", r["x_hat_bl"], o["x_hat_bl"], 1e-3)
    _check("x_hat_el", r["x_hat_el"], o["x_hat_el"], 1e-3)
    for k in ("bit_bl", "bit_el"):
        assert abs(r[k] - o[k]) / o[k] < 1e-3

    dpb = {"ref_frame_bl": o["x_hat_bl"].clamp(0, 1), "ref_frame_el": o["x_hat_el"].clamp(0, 1), "ref_feature_bl": None,
           "ref_feature_el": o["feature_el"]}
    x_bl, x_el = frames[1]
    with torch.no_grad():
        o = orc.lssvc(sd_p, x_bl, x_el, dpb, (He, We), 2.0, PAD)
    q_ref = {"bl_mv_z_hat": o["bl"]["mv_z_hat"], "bl_mv_y_q": o["bl"]["mv_y_q"], "bl_z_hat": o["bl"]["z_hat"],
             "bl_y_q": o["bl"]["y_q"], "mv_z_hat": o["mv_z_hat"], "mv_y_q": o["mv_y_q"], "z_hat": o["z_hat"],
             "y_q": o["four_part"]["y_q"]}
    dpb_dev = {k: (None if v is None else v.to(dev)) for k, v in dpb.items()}
    r, flips = _run(net_p, lambda: net_p.encode_decode(x_bl.to(dev), x_el.to(dev), dpb_dev, None, None, We, He, 128, 128), q_ref)
    total = sum(v.numel() for v in q_ref.values())
    print(f"pad_size P-frame: bits {r['bit_bl']:.1f}/{r['bit_el']:.1f} oracle {o['bit_bl']:.1f}/{o['bit_el']:.1f}; "
          f"{sum(flips.values())} of {total} symbols differ {flips}")
    assert sum(flips.values()) <= max(1, 1e-4 * total)
    for k in ("mv_hat", "warp_frame"):
        _check(k, r[k], o[k], 1e-3)
    for k in ("ref_frame_bl", "ref_frame_el"):
        _check(k, r["dpb"][k], o["dpb"][k], 1e-3)
    for k in ("bit_bl", "bit_el"):
        assert abs(r[k] - o[k]) / o[k] < 1e-3
