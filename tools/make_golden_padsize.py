"""Golden fixture for the non-zero pad_size path (get_depadded_feature, LSSVC_net.py:271-282 / IntraSS.py:124-135), from the
UNMODIFIED reference (build container only): I + 1 P frame at EL 128x128 with the base layer coded on a 128x128 frame (64 px of
extra padding right / bottom) and pad_size = (0, -64, 0, -64).  The oracle is required to agree with the reference exactly
before anything is written.  -> tests/golden/padsize_128.pt"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ref_harness  # noqa: E402
from lssvc_b200 import nets, synth  # noqa: E402
from oracle import lssvc_oracle as orc  # noqa: E402


def sub(t, step=8):
    return t[:, :, ::step, ::step].contiguous().clone()


def main():
    torch.manual_seed(0)
    H = W = 128
    PAD, seed = (0, -64, 0, -64), 3
    IntraSS, LSSVC_extend = ref_harness.import_reference()
    sd_i = nets.ParamBag(nets.intra_ss_spec(), seed=0, gains=nets.model_gains("I")).state_dict()
    sd_p = nets.ParamBag(nets.lssvc_spec(), seed=1, gains=nets.model_gains("P")).state_dict()
    ref_i = IntraSS.from_state_dict(dict(sd_i)).eval()
    ref_p = LSSVC_extend().eval()
    ref_p.load_dict(dict(sd_p))
    frames = [(torch.nn.functional.pad(b, (0, 64, 0, 64), mode="replicate"), e) for b, e in synth.make_sequence(H, W, 2, seed=seed)]
    out = {"H": H, "W": W, "seed": seed, "pad_size": PAD, "frames": []}
    with torch.no_grad():
        for m in (ref_i, ref_p):
            m.set_scale_information(2.0, (H, W), PAD)
        x_bl, x_el = frames[0]
        r = ref_i.encode_decode(x_bl, x_el, None, None, 128, 128, H, W)
        o = orc.intra_ss(sd_i, x_bl, x_el, (H, W), PAD)
        for k in ("x_hat_bl", "x_hat_el", "feature_el"):
            assert torch.equal(r[k], o[k]), k
        assert r["bit_bl"] == o["bit_bl"] and r["bit_el"] == o["bit_el"]
        out["frames"].append({"bit_bl": r["bit_bl"], "bit_el": r["bit_el"], "x_hat_bl": sub(r["x_hat_bl"], 4), "x_hat_el": sub(r["x_hat_el"]),
                              "feature_el": sub(r["feature_el"])})
        dpb = {"ref_frame_bl": r["x_hat_bl"].clamp(0, 1), "ref_frame_el": r["x_hat_el"].clamp(0, 1), "ref_feature_bl": None,
               "ref_feature_el": r["feature_el"]}
        x_bl, x_el = frames[1]
        r = ref_p.encode_decode(x_bl, x_el, dict(dpb), None, None, W, H, 128, 128)
        o = orc.lssvc(sd_p, x_bl, x_el, dict(dpb), (H, W), 2.0, PAD)
        for k in r["dpb"]:
            assert torch.equal(r["dpb"][k], o["dpb"][k]), k
        assert torch.equal(r["mv_hat"], o["mv_hat"]) and r["bit_bl"] == o["bit_bl"] and r["bit_el"] == o["bit_el"]
        out["frames"].append({"bit_bl": r["bit_bl"], "bit_el": r["bit_el"], "ref_frame_bl": sub(r["dpb"]["ref_frame_bl"], 4),
                              "ref_frame_el": sub(r["dpb"]["ref_frame_el"]), "ref_feature_el": sub(r["dpb"]["ref_feature_el"]),
                              "mv_hat": sub(r["mv_hat"]), "sym_el": o["four_part"]["y_q"].to(torch.int8)})
    path = os.path.join(ROOT, "tests", "golden", "padsize_128.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes; bits", [(f["bit_bl"], f["bit_el"]) for f in out["frames"]])


if __name__ == "__main__":
    main()
