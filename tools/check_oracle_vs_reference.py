"""Run the unmodified reference and the oracle restatement on identical seeded weights/frames (CPU) and
report the maximum deviation of every output (expected: 0.0, same torch ops in the same order)."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ref_harness  # noqa: E402
from lssvc_b200 import nets, synth  # noqa: E402
from oracle import lssvc_oracle as orc  # noqa: E402


def main(H=128, W=128, n_frames=4, seed=0, ratio=2.0, bl_pad=0):
    """bl_pad > 0: the base layer is coded on a frame padded by bl_pad more pixels (right / bottom) than the enhancement layer's
    size / ratio, and the models de-pad its tensors with pad_size = (0, -bl_pad, 0, -bl_pad) (get_depadded_feature)."""
    torch.manual_seed(0)
    IntraSS, LSSVC_extend = ref_harness.import_reference()
    sd_i = nets.ParamBag(nets.intra_ss_spec(), seed=seed, gains=nets.model_gains("I")).state_dict()
    sd_p = nets.ParamBag(nets.lssvc_spec(), seed=seed + 1, gains=nets.model_gains("P")).state_dict()
    ref_i = IntraSS.from_state_dict(dict(sd_i)).eval()
    ref_p = LSSVC_extend().eval()
    ref_p.load_dict(dict(sd_p))
    frames = synth.make_sequence(H, W, n_frames, seed=seed, ratio=ratio)
    pad_size = (0, -bl_pad, 0, -bl_pad)
    if bl_pad:
        frames = [(torch.nn.functional.pad(b, (0, bl_pad, 0, bl_pad), mode="replicate"), e) for b, e in frames]
    worst = 0.0
    with torch.no_grad():
        dpb_r = dpb_o = None
        for t, (x_bl, x_el) in enumerate(frames):
            ref_i.set_scale_information(ratio, (H, W), pad_size)
            ref_p.set_scale_information(ratio, (H, W), pad_size)
            t0 = time.time()
            if t == 0:
                r = ref_i.encode_decode(x_bl, x_el, None, None, x_bl.shape[2], x_bl.shape[3], H, W)
                o = orc.intra_ss(sd_i, x_bl, x_el, (H, W), pad_size)
                pairs = [("x_hat_bl", r["x_hat_bl"], o["x_hat_bl"]), ("x_hat_el", r["x_hat_el"], o["x_hat_el"]),
                         ("feature_el", r["feature_el"], o["feature_el"])]
                dpb_r = {"ref_frame_bl": r["x_hat_bl"], "ref_frame_el": r["x_hat_el"], "ref_feature_bl": None,
                         "ref_feature_el": r["feature_el"]}
                dpb_o = {"ref_frame_bl": o["x_hat_bl"], "ref_frame_el": o["x_hat_el"], "ref_feature_bl": None,
                         "ref_feature_el": o["feature_el"]}
            else:
                r = ref_p.encode_decode(x_bl, x_el, dpb_r, None, None, W, H, x_bl.shape[3], x_bl.shape[2])
                o = orc.lssvc(sd_p, x_bl, x_el, dpb_o, (H, W), ratio, pad_size)
                pairs = [(k, r["dpb"][k], o["dpb"][k]) for k in r["dpb"]] + [("mv_hat", r["mv_hat"], o["mv_hat"]),
                                                                             ("warp_frame", r["warp_frame"], o["warp_frame"])]
                dpb_r, dpb_o = r["dpb"], o["dpb"]
            for d in (dpb_r, dpb_o):
                d["ref_frame_bl"] = d["ref_frame_bl"].clamp_(0, 1)
                d["ref_frame_el"] = d["ref_frame_el"].clamp_(0, 1)
            devs = {k: (a - b).abs().max().item() for k, a, b in pairs}
            worst = max([worst] + list(devs.values()) + [abs(r["bit_bl"] - o["bit_bl"]), abs(r["bit_el"] - o["bit_el"])])
            print(f"frame {t}: bits ref ({r['bit_bl']:.1f}, {r['bit_el']:.1f}) oracle ({o['bit_bl']:.1f}, {o['bit_el']:.1f}) "
                  f"max dev {max(devs.values()):.3e}  [{time.time() - t0:.1f}s]")
            if t > 0:
                fp = o["four_part"]
                nz = (fp["y_q"] != 0).float().mean().item()
                rows = orc.build_indexes_video(fp["scales_hat"]).unique().numel()
                print(f"   EL y_q nonzero {nz:.3f}, |y_q| max {fp['y_q'].abs().max().item():.0f}, scale rows {rows}, "
                      f"mv_y_q nonzero {(o['mv_y_q'] != 0).float().mean().item():.3f}, z_hat nz {(o['z_hat'] != 0).float().mean().item():.3f}, "
                      f"BL y_q nz {(o['bl']['y_q'] != 0).float().mean().item():.3f}, mv_hat absmax {o['mv_hat'].abs().max().item():.2f}, "
                      f"feat_el absmax {o['dpb']['ref_feature_el'].abs().max().item():.2f} feat_bl absmax {o['dpb']['ref_feature_bl'].abs().max().item():.2f} "
                      f"recon_el range [{o['dpb']['ref_frame_el'].min().item():.2f},{o['dpb']['ref_frame_el'].max().item():.2f}]")
            else:
                print(f"   I: EL y sym nz {(torch.round(o['y'] - o['means']) != 0).float().mean().item():.3f}, rows "
                      f"{orc.build_indexes_image(o['scales']).unique().numel()}, BL y nz "
                      f"{(torch.round(o['bl']['y'] - o['bl']['means']) != 0).float().mean().item():.3f}, x_hat_el range "
                      f"[{o['x_hat_el'].min().item():.2f},{o['x_hat_el'].max().item():.2f}] feat absmax {o['feature_el'].abs().max().item():.2f}")
    print("WORST DEVIATION", worst)
    return worst


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    if len(sys.argv) > 2:          # e.g. "3 1.5 192": the x1_5 ratio of recommend_test_config.json at EL 192x192 / BL 128x128
        # "3 2 128 64": BL coded at 128x128 for an EL of 128x128 (64 px of extra BL padding, pad_size = (0, -64, 0, -64))
        main(H=int(sys.argv[3]), W=int(sys.argv[3]), n_frames=n, ratio=float(sys.argv[2]),
             bl_pad=int(sys.argv[4]) if len(sys.argv) > 4 else 0)
    else:
        main(n_frames=n)
