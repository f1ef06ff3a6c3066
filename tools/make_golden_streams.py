"""Reference-produced bitstreams: tests/golden/streams_128.pt (build container only; /root/reference is not on the GPU box).

The UNMODIFIED reference codes I + 2 P frames at EL 128x128 / BL 64x64 in `--write_stream 1` mode exactly as test.py does
(test.py:212-250, 561-564): `update(force=True)`, `IntraSS.encode_decode(..., bin_path_bl, bin_path_el, ...)`
(IntraSS.py:245-302 -> IntraNoAR.compress/decompress priors.py:420-452, IntraSS.compress/decompress :304-336) and
`LSSVC_extend.encode_decode(..., output_path_bl, output_path_el, ...)` (LSSVC_net_extend.py:144-191 ->
DMCExtend.compress/decompress dmc_net_extend.py:55-147, LSSVC_extend.compress/decompress :24-142), on the seeded synthetic
weights, with the reference's OWN C++ coder (oracle/_ref, compiled from /root/reference/src/cpp by oracle/Makefile).

The image path imports `RansEncoder` / `RansDecoder.decode_with_indexes`, which src/cpp does not define (SURVEY §8c "Gap"):
they are supplied as two-line adapters over the reference's own BufferedRansEncoder / RansDecoder (encode + flush,
set_stream + decode_stream).  `torch.cuda.synchronize` is a no-op here (no GPU in the build container).

Before anything is written the script REQUIRES, frame by frame, that
  * the stream-mode DPB of the reference (the decoder's reconstruction) equals the oracle's estimate-mode DPB bit for bit
    (after the caller's clamp of test.py:249-250), so the GPU tests can rebuild every DPB from the oracle instead of
    shipping ~10 MB of feature maps;
  * the symbols / CDF indices the oracle exposes, pushed through the oracle-side composition below, reproduce the
    reference's files byte for byte (this is the composition tests/test_streams.py re-checks against the PRODUCT coder).

Fixture: per frame the two files (bytes), bits, sha256 of the reconstructions."""
import hashlib
import os
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import make_golden  # noqa: E402
import ref_harness  # noqa: E402
from lssvc_b200 import nets, synth  # noqa: E402
from oracle import lssvc_oracle as orc  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
H = W = 128
SEED = 0


def install_image_coder(rans):
    """RansEncoder / decode_with_indexes of the prebuilt cpython-36 module, expressed on the reference's own C++ classes."""
    class RansEncoder:
        def encode_with_indexes(self, symbols, indexes, cdfs, sizes, offsets):
            e = rans.BufferedRansEncoder()
            e.encode_with_indexes(symbols, indexes, cdfs, sizes, offsets)
            return e.flush()

    base = rans.RansDecoder

    class RansDecoder:
        def __init__(self):
            self._d = base()

        def set_stream(self, s):
            self._d.set_stream(s)

        def decode_stream(self, *a):
            return self._d.decode_stream(*a)

        def decode_with_indexes(self, s, indexes, cdfs, sizes, offsets):
            self._d.set_stream(s)
            return list(self._d.decode_stream(indexes, cdfs, sizes, offsets))

    rans.RansEncoder = RansEncoder
    rans.RansDecoder = RansDecoder


def sha(t):
    return hashlib.sha256(np.ascontiguousarray(t.detach().cpu().numpy()).tobytes()).hexdigest()


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    IntraSS, LSSVC_extend = ref_harness.import_reference()
    mods = make_golden.load_ref_native()
    install_image_coder(mods["MLCodec_rans"])
    torch.cuda.synchronize = lambda *a, **k: None

    sd_i = nets.ParamBag(nets.intra_ss_spec(), seed=SEED, gains=nets.model_gains("I")).state_dict()
    sd_p = nets.ParamBag(nets.lssvc_spec(), seed=SEED + 1, gains=nets.model_gains("P")).state_dict()
    ref_i = IntraSS.from_state_dict(dict(sd_i)).eval()
    ref_p = LSSVC_extend().eval()
    ref_p.load_dict(dict(sd_p))
    ref_i.update(force=True)
    ref_p.update(force=True)
    frames = synth.make_sequence(H, W, 3, seed=SEED)

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import stream_compose  # the oracle-side composition shared with tests/test_streams.py

    rec = {"H": H, "W": W, "seed": SEED, "frames": []}
    tmp = tempfile.mkdtemp()
    with torch.no_grad():
        dpb_r = dpb_o = None
        for t, (x_bl, x_el) in enumerate(frames):
            for m in (ref_i, ref_p):
                m.set_scale_information(2.0, (H, W), (0, 0, 0, 0))
            p_bl, p_el = os.path.join(tmp, f"{t}_bl.bin"), os.path.join(tmp, f"{t}_el.bin")
            if t == 0:
                r = ref_i.encode_decode(x_bl, x_el, p_bl, p_el, H // 2, W // 2, H, W)
                o = orc.intra_ss(sd_i, x_bl, x_el, (H, W))
                assert torch.equal(r["x_hat_bl"], o["x_hat_bl"]) and torch.equal(r["x_hat_el"], o["x_hat_el"])
                assert torch.equal(r["feature_el"], o["feature_el"])
                dpb_r = {"ref_frame_bl": r["x_hat_bl"], "ref_frame_el": r["x_hat_el"], "ref_feature_bl": None,
                         "ref_feature_el": r["feature_el"]}
                dpb_o = {"ref_frame_bl": o["x_hat_bl"], "ref_frame_el": o["x_hat_el"], "ref_feature_bl": None,
                         "ref_feature_el": o["feature_el"]}
                composed = stream_compose.intra_files(orc, o, ref_i.state_dict(), H, W, coder="oracle")
                fr = {"type": "I", "x_hat_bl": sha(o["x_hat_bl"]), "x_hat_el": sha(o["x_hat_el"])}
            else:
                r = ref_p.encode_decode(x_bl, x_el, dict(dpb_r), p_bl, p_el, W, H, W // 2, H // 2)
                o = orc.lssvc(sd_p, x_bl, x_el, dpb_o, (H, W), 2.0)
                # the reference's stream-mode DPB (decoder output; BL recon clamped by the decoder) == oracle estimate mode
                assert torch.equal(r["dpb"]["ref_frame_bl"], o["dpb"]["ref_frame_bl"].clamp(0, 1))
                for k in ("ref_feature_bl", "ref_frame_el", "ref_feature_el"):
                    assert torch.equal(r["dpb"][k], o["dpb"][k]), k
                assert torch.equal(r["mv_hat"], o["mv_hat"])
                dpb_r, dpb_o = r["dpb"], o["dpb"]
                composed = stream_compose.inter_files(orc, o, sd_p, coder="oracle")
                fr = {"type": "P", "x_hat_bl": sha(o["dpb"]["ref_frame_bl"].clamp(0, 1)), "x_hat_el": sha(o["dpb"]["ref_frame_el"])}
            fr["file_bl"], fr["file_el"] = open(p_bl, "rb").read(), open(p_el, "rb").read()
            assert composed[0] == fr["file_bl"], f"frame {t}: BL composition from the oracle's symbols differs from the reference file"
            assert composed[1] == fr["file_el"], f"frame {t}: EL composition from the oracle's symbols differs from the reference file"
            assert r["bit_bl"] == 8 * len(fr["file_bl"]) and r["bit_el"] == 8 * len(fr["file_el"])
            fr["bit_bl"], fr["bit_el"] = r["bit_bl"], r["bit_el"]
            fr["est_bl"], fr["est_el"] = o["bit_bl"], o["bit_el"]
            rec["frames"].append(fr)
            for d in (dpb_r, dpb_o):
                d["ref_frame_bl"] = d["ref_frame_bl"].clamp_(0, 1)
                d["ref_frame_el"] = d["ref_frame_el"].clamp_(0, 1)
            print(f"frame {t} ({fr['type']}): files {len(fr['file_bl'])} + {len(fr['file_el'])} B "
                  f"(estimated {o['bit_bl'] / 8:.0f} + {o['bit_el'] / 8:.0f} B); reference stream DPB == oracle; "
                  f"composition from oracle symbols == reference files")
    torch.save(rec, os.path.join(GOLD, "streams_128.pt"))
    print("streams_128.pt:", os.path.getsize(os.path.join(GOLD, "streams_128.pt")), "bytes")


if __name__ == "__main__":
    main()
