"""BASELINE config 1: IntraSS I-frame + LSSVC two-layer P-frames, one 12-frame GOP at 512x320 (BL 256x160), i.e. the padded
sizes 384x512 / 192x256 the reference runs (common.py:48-86), estimate mode.  Every frame is coded from the ORACLE's DPB and
teacher-forced on the oracle's symbols, so the twelve frames are twelve independent checks of the north-star tolerances
(symbols >= 99.99 % equal over the GOP, reconstructions within 1e-3, per-layer bits within 0.1 %), including the P-after-P
feature path (48-channel EL reference feature) over a full GOP of drift in the reference data."""
import pytest
import torch

pytestmark = pytest.mark.gpu

GOP = 12


def test_cfg1_gop_against_oracle(cuda_device):
    from lssvc_b200 import IntraSS, LSSVC_extend, frontend, synth
    from oracle import lssvc_oracle as orc
    pad = frontend.get_interlayer_padding(320, 512, 2)
    H, W = pad["HR_padded_size"]
    assert (H, W) == (384, 512) and pad["LR_padded_size"] == (192, 256)
    dev = cuda_device
    torch.set_num_threads(8)
    net_i, net_p = IntraSS(seed=0), LSSVC_extend(seed=1)
    sd_i = {k: v.clone() for k, v in net_i.state_dict().items()}
    sd_p = {k: v.clone() for k, v in net_p.state_dict().items()}
    net_i.to(dev)
    net_p.to(dev)
    for n in (net_i, net_p):
        n.set_scale_information(2.0, (H, W), (0, 0, 0, 0))
    frames = synth.make_sequence(H, W, GOP, seed=5)
    bad = total = 0
    worst = {"recon": 0.0, "bits": 0.0}
    dpb = None
    for t, (x_bl, x_el) in enumerate(frames):
        with torch.no_grad():
            if t == 0:
                o = orc.intra_ss(sd_i, x_bl, x_el, (H, W))
                q_ref = {"bl_z_hat": o["bl"]["z_hat"], "bl_y_q": torch.round(o["bl"]["y"] - o["bl"]["means"]), "z_hat": o["z_hat"],
                         "y_q": torch.round(o["y"] - o["means"])}
                net, call = net_i, (lambda: net_i.encode_decode(x_bl.to(dev), x_el.to(dev), None, None, H // 2, W // 2, H, W))
            else:
                o = orc.lssvc(sd_p, x_bl, x_el, dpb, (H, W), 2.0)
                q_ref = {"bl_mv_z_hat": o["bl"]["mv_z_hat"], "bl_mv_y_q": o["bl"]["mv_y_q"], "bl_z_hat": o["bl"]["z_hat"],
                         "bl_y_q": o["bl"]["y_q"], "mv_z_hat": o["mv_z_hat"], "mv_y_q": o["mv_y_q"], "z_hat": o["z_hat"],
                         "y_q": o["four_part"]["y_q"]}
                dpb_dev = {k: (None if v is None else v.to(dev)) for k, v in dpb.items()}
                net, call = net_p, (lambda: net_p.encode_decode(x_bl.to(dev), x_el.to(dev), dpb_dev, None, None, W, H, W // 2, H // 2))
        try:
            net._debug, net._force, net._force_flips = {}, q_ref, {}
            r = call()
            flips = dict(net._force_flips)
        finally:
            net._debug = net._force = None
        if t == 0:
            recs = [(r["x_hat_bl"], o["x_hat_bl"]), (r["x_hat_el"], o["x_hat_el"])]
            dpb = {"ref_frame_bl": o["x_hat_bl"].clamp(0, 1), "ref_frame_el": o["x_hat_el"].clamp(0, 1), "ref_feature_bl": None,
                   "ref_feature_el": o["feature_el"]}
        else:
            recs = [(r["dpb"]["ref_frame_bl"], o["dpb"]["ref_frame_bl"]), (r["dpb"]["ref_frame_el"], o["dpb"]["ref_frame_el"])]
            dpb = dict(o["dpb"])
            dpb["ref_frame_bl"] = dpb["ref_frame_bl"].clamp(0, 1)          # test.py:249-250
            dpb["ref_frame_el"] = dpb["ref_frame_el"].clamp(0, 1)
        d = max((a.cpu() - b).abs().max().item() for a, b in recs)
        rb = max(abs(r[k] - o[k]) / o[k] for k in ("bit_bl", "bit_el"))
        n_bad, n = sum(flips.values()), sum(v.numel() for v in q_ref.values())
        print(f"frame {t:2d} ({'I' if t == 0 else 'P'}): bits {r['bit_bl']:.0f}/{r['bit_el']:.0f} (rel {rb:.1e}), recon max|d| {d:.2e}, "
              f"{n_bad} of {n} symbols differ")
        assert d < 1e-3 and rb < 1e-3
        bad += n_bad
        total += n
        worst["recon"], worst["bits"] = max(worst["recon"], d), max(worst["bits"], rb)
    print(f"GOP of {GOP}: {bad} of {total} symbols differ ({100 * bad / total:.4f} %), worst recon {worst['recon']:.2e}, worst bits {worst['bits']:.1e}")
    assert bad <= 1e-4 * total


def test_cfg1_ip32_free_running_drift(cuda_device):
    """SURVEY H2 / BASELINE config 3's GOP length at config-1 size: ONE IP32 GOP (1 I + 31 P, 384x512 / 192x256) coded
    FREE-RUNNING — every implementation carries its own DPB, nothing is teacher-forced — by the oracle, by the default
    tensor-core engine and by the fp32 CUDA-core engine.

    Two fp32-accurate implementations of this recurrent codec cannot stay bit-identical: a latent within float noise of a
    rounding boundary flips a symbol (<= 0.01 % per frame is the contract, 0-2 of 200 k here), and with RANDOM-INIT synthetic
    weights a single flipped symbol moves the reconstruction by O(0.1) locally (the synthesis transform is not trained to be
    smooth) and the next frames inherit it through the DPB.  What the chain has to show is that the divergence SATURATES
    instead of compounding, that it is a property of the algorithm under these weights and not of the split-fp16 arithmetic
    (the fp32 engine, whose per-layer error is 10x smaller, diverges from the oracle just as much), and that the rate stays
    put.  Recorded per frame: max|d| and rms of the reconstructions against the oracle chain, PSNR(cuda, oracle), relative
    difference of the bits.  Asserted over all 32 frames: rms <= 0.05 (PSNR(cuda, oracle) >= 26 dB) with the last 16 frames
    no worse than the first 16 (+ 10 %), per-layer bits within 2 % per frame and 0.5 % per GOP, and the default engine no
    further from the oracle than 1.5 x the fp32 engine.  (First measurement, both engines alike: the I-frame already carries
    two flipped symbols = 5e-2 max|d| locally; rms settles at 2.0e-2 = 34 dB from frame 5 on; bits within 1.1 % per frame,
    0.09 % per GOP.)"""
    import math
    import os
    from lssvc_b200 import IntraSS, LSSVC_extend, ops, synth
    from oracle import lssvc_oracle as orc
    H, W, N = 384, 512, 32
    dev = cuda_device
    torch.set_num_threads(os.cpu_count() or 8)
    net_i, net_p = IntraSS(seed=0), LSSVC_extend(seed=1)
    sd_i = {k: v.clone() for k, v in net_i.state_dict().items()}
    sd_p = {k: v.clone() for k, v in net_p.state_dict().items()}
    net_i.to(dev)
    net_p.to(dev)
    for n in (net_i, net_p):
        n.set_scale_information(2.0, (H, W), (0, 0, 0, 0))
    frames = synth.make_sequence(H, W, N, seed=6)
    engines = [ops.default_engine(), "simt"]
    dpb_o, dpb_g = None, {e: None for e in engines}
    rows = {e: [] for e in engines}
    tot = {e: {"o_bl": 0.0, "o_el": 0.0, "g_bl": 0.0, "g_el": 0.0} for e in engines}
    for t, (x_bl, x_el) in enumerate(frames):
        with torch.no_grad():
            if t == 0:
                o = orc.intra_ss(sd_i, x_bl, x_el, (H, W))
                dpb_o = {"ref_frame_bl": o["x_hat_bl"], "ref_frame_el": o["x_hat_el"], "ref_feature_bl": None, "ref_feature_el": o["feature_el"]}
            else:
                o = orc.lssvc(sd_p, x_bl, x_el, dpb_o, (H, W), 2.0)
                dpb_o = dict(o["dpb"])
        dpb_o["ref_frame_bl"] = dpb_o["ref_frame_bl"].clamp_(0, 1)          # test.py:249-250
        dpb_o["ref_frame_el"] = dpb_o["ref_frame_el"].clamp_(0, 1)
        for e in engines:
            prev = ops.set_engine(e)
            try:
                if t == 0:
                    g = net_i.encode_decode(x_bl.to(dev), x_el.to(dev), None, None, H // 2, W // 2, H, W)
                    d = {"ref_frame_bl": g["x_hat_bl"], "ref_frame_el": g["x_hat_el"], "ref_feature_bl": None, "ref_feature_el": g["feature_el"]}
                else:
                    g = net_p.encode_decode(x_bl.to(dev), x_el.to(dev), dpb_g[e], None, None, W, H, W // 2, H // 2)
                    d = g["dpb"]
            finally:
                ops.set_engine(prev)
            d["ref_frame_bl"].clamp_(0, 1)
            d["ref_frame_el"].clamp_(0, 1)
            dpb_g[e] = d
            e_el = d["ref_frame_el"].cpu() - dpb_o["ref_frame_el"]
            e_bl = d["ref_frame_bl"].cpu() - dpb_o["ref_frame_bl"]
            rms = max(e_el.pow(2).mean().sqrt().item(), e_bl.pow(2).mean().sqrt().item())
            mx = max(e_el.abs().max().item(), e_bl.abs().max().item())
            rb = max(abs(g[k] - o[k]) / o[k] for k in ("bit_bl", "bit_el"))
            for k in ("bl", "el"):
                tot[e]["o_" + k] += o["bit_" + k]
                tot[e]["g_" + k] += g["bit_" + k]
            rows[e].append((t, mx, rms, rb))
        a, b = rows[engines[0]][-1], rows[engines[1]][-1]
        psnr = lambda r: 10 * math.log10(1.0 / max(r * r, 1e-20))
        print(f"frame {t:2d} ({'I' if t == 0 else 'P'}): {engines[0]}: max|d| {a[1]:.2e} rms {a[2]:.2e} PSNR(cuda, oracle) {psnr(a[2]):5.1f} dB bits rel {a[3]:.1e}"
              f"   |   simt: max|d| {b[1]:.2e} rms {b[2]:.2e} PSNR {psnr(b[2]):5.1f} dB bits rel {b[3]:.1e}")
    summary = {}
    for e in engines:
        r = rows[e]
        gop_bits = max(abs(tot[e]["g_" + k] - tot[e]["o_" + k]) / tot[e]["o_" + k] for k in ("bl", "el"))
        summary[e] = {"worst_rms": max(x[2] for x in r), "first16": sum(x[2] for x in r[:16]) / 16, "last16": sum(x[2] for x in r[16:]) / 16,
                      "worst_bits": max(x[3] for x in r), "gop_bits": gop_bits, "i_frame": r[0][1]}
        print(f"IP32 free-running, {e}: worst rms {summary[e]['worst_rms']:.2e}, mean rms frames 0-15 {summary[e]['first16']:.2e} / 16-31 "
              f"{summary[e]['last16']:.2e}, worst per-frame bits {summary[e]['worst_bits']:.1e}, GOP bits {gop_bits:.1e}, I-frame max|d| {r[0][1]:.1e}")
    for e in engines:
        s_ = summary[e]
        assert s_["worst_rms"] <= 0.05 and s_["last16"] <= 1.1 * s_["first16"] + 1e-3, "the drift must saturate, not compound"
        assert s_["worst_bits"] < 2e-2 and s_["gop_bits"] < 5e-3
    if summary["simt"]["last16"] > 1e-3:        # the fp32 engine left the oracle's trajectory too (the expected case)
        assert summary[engines[0]]["last16"] <= 1.5 * summary["simt"]["last16"] + 1e-3, \
            "the split-fp16 engine drifts further from the oracle than the fp32 engine does"
