"""Measure the per-conv weight multipliers of the synthetic-weight recipe (lssvc_b200/synth_gains.json).

One pass of the oracle over a synthetic I + P + P sequence; every conv / transposed conv is rescaled the first
time it runs so that its (bias-free) output has the target std its spec entry declares, and the forward pass
continues with the rescaled weights (layer-sequential unit-variance initialisation).  The multipliers are
relative to N(0, 1/fan_in) and are what lssvc_b200.nets.init_tensor applies for ANY seed.
"""
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lssvc_b200 import nets, synth  # noqa: E402
from oracle import lssvc_oracle as orc  # noqa: E402


def main(H=128, W=128, seed=0):
    spec_i, spec_p = nets.intra_ss_spec(), nets.lssvc_spec()
    sd_i = nets.ParamBag(spec_i, seed=seed, gains={"__raw__": 1.0}).state_dict()
    sd_p = nets.ParamBag(spec_p, seed=seed + 1, gains={"__raw__": 1.0}).state_dict()
    frames = synth.make_sequence(H, W, 4, seed=seed)
    factors = {}
    raw_gain = {}
    for spec in (spec_i, spec_p):
        for k, e in spec.items():
            if e["kind"] in ("conv_w", "deconv_w"):
                raw_gain[k[:-len(".weight")]] = min(e["gain"], 1.0)
    targets = {}
    for tag, spec in (("I.", spec_i), ("P.", spec_p)):
        for k, e in spec.items():
            if e["kind"] in ("conv_w", "deconv_w"):
                targets[tag + k[:-len(".weight")]] = min(e["gain"], 1.0)
    state = {"tag": "I."}
    o_conv, o_deconv = orc.conv, orc.deconv

    def calibrate(p, y0):
        name = p.prefix[:-1]
        key = state["tag"] + name
        if key in factors:
            return
        std = y0.std().item()
        f = targets[key] / max(std, 1e-12)
        p.sd[p.prefix + "weight"].mul_(f)
        factors[key] = f * raw_gain[name]

    def c_conv(p, x, stride=1, padding=None):
        w = p["weight"]
        pad = w.shape[-1] // 2 if padding is None else padding
        calibrate(p, F.conv2d(x, w, None, stride=stride, padding=pad))
        return o_conv(p, x, stride, padding)

    def c_deconv(p, x, stride):
        calibrate(p, F.conv_transpose2d(x, p["weight"], None, stride=stride, padding=1, output_padding=stride - 1))
        return o_deconv(p, x, stride)

    orc.conv, orc.deconv = c_conv, c_deconv
    try:
        with torch.no_grad():
            x_bl, x_el = frames[0]
            o = orc.intra_ss(sd_i, x_bl, x_el, (H, W))
            dpb = {"ref_frame_bl": o["x_hat_bl"].clamp(0, 1), "ref_frame_el": o["x_hat_el"].clamp(0, 1),
                   "ref_feature_bl": None, "ref_feature_el": o["feature_el"]}
            state["tag"] = "P."
            for t in (1, 2):
                x_bl, x_el = frames[t]
                o = orc.lssvc(sd_p, x_bl, x_el, dpb, (H, W), 2.0)
                dpb = o["dpb"]
                dpb["ref_frame_bl"] = dpb["ref_frame_bl"].clamp(0, 1)
                dpb["ref_frame_el"] = dpb["ref_frame_el"].clamp(0, 1)
    finally:
        orc.conv, orc.deconv = o_conv, o_deconv
    # the two models never share a key except through identical sub-module names; keep them apart by model
    out = {"I": {k[2:]: round(v, 6) for k, v in factors.items() if k.startswith("I.")},
           "P": {k[2:]: round(v, 6) for k, v in factors.items() if k.startswith("P.")}}
    path = os.path.join(ROOT, "lssvc_b200", "synth_gains.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)
    print(f"wrote {path}: {len(out['I'])} + {len(out['P'])} convs; "
          f"factor range [{min(factors.values()):.3g}, {max(factors.values()):.3g}]")


if __name__ == "__main__":
    main()
