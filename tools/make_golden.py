"""Generate the golden fixtures under tests/golden/ from the UNMODIFIED reference (run in the build container only;
/root/reference does not exist on the GPU box).

For each fixture the reference and the oracle restatement are both evaluated and required to agree exactly before
anything is written, which is what pins the oracle:

  state_dict_layout.json   keys / shapes / dtypes of IntraSS and LSSVC_extend                (nets.py mirror)
  forward_128.pt           I + 2 P frames at EL 128x128: bits, sub-sampled reconstructions, symbols, scales
  entropy_tables.json      sha256 of the CDF tables built by the reference's update() methods  (entropy.py mirror)
  rans_vectors.npz         symbols/indexes and the byte strings the reference's own C++ coder produces
"""
import hashlib
import importlib.util
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ref_harness  # noqa: E402
from lssvc_b200 import nets, synth  # noqa: E402
from oracle import lssvc_oracle as orc  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
REF_SO = os.path.join(ROOT, "oracle", "_ref")


def load_ref_native():
    """Register the reference's own C++ modules (compiled into oracle/_ref by oracle/Makefile) under the names the
    reference's relative imports resolve to."""
    mods = {}
    for name in ("MLCodec_CXX", "MLCodec_rans"):
        path = [os.path.join(REF_SO, f) for f in os.listdir(REF_SO) if f.startswith(name)][0]
        spec = importlib.util.spec_from_file_location(name, path)
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        sys.modules["src.entropy_models." + name] = m
        mods[name] = m
    return mods


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def sub(t, step=8):
    return t[:, :, ::step, ::step].contiguous().clone()


def golden_layout(IntraSS, LSSVC_extend):
    out = {}
    for tag, model in (("IntraSS", IntraSS()), ("LSSVC_extend", LSSVC_extend())):
        out[tag] = [[k, list(v.shape), str(v.dtype)] for k, v in model.state_dict().items()]
    with open(os.path.join(GOLD, "state_dict_layout.json"), "w") as f:
        json.dump(out, f)
    print("layout:", {k: len(v) for k, v in out.items()})


def golden_forward(IntraSS, LSSVC_extend, H=128, W=128, seed=0):
    sd_i = nets.ParamBag(nets.intra_ss_spec(), seed=seed, gains=nets.model_gains("I")).state_dict()
    sd_p = nets.ParamBag(nets.lssvc_spec(), seed=seed + 1, gains=nets.model_gains("P")).state_dict()
    ref_i = IntraSS.from_state_dict(dict(sd_i)).eval()
    ref_p = LSSVC_extend().eval()
    ref_p.load_dict(dict(sd_p))
    frames = synth.make_sequence(H, W, 3, seed=seed)
    rec = {"H": H, "W": W, "seed": seed, "frames": []}
    with torch.no_grad():
        dpb_r = dpb_o = None
        for t, (x_bl, x_el) in enumerate(frames):
            ref_i.set_scale_information(2.0, (H, W), (0, 0, 0, 0))
            ref_p.set_scale_information(2.0, (H, W), (0, 0, 0, 0))
            if t == 0:
                r = ref_i.encode_decode(x_bl, x_el, None, None, x_bl.shape[2], x_bl.shape[3], H, W)
                o = orc.intra_ss(sd_i, x_bl, x_el, (H, W))
                for k in ("x_hat_bl", "x_hat_el", "feature_el"):
                    assert torch.equal(r[k], o[k]), k
                dpb_r = {"ref_frame_bl": r["x_hat_bl"], "ref_frame_el": r["x_hat_el"], "ref_feature_bl": None,
                         "ref_feature_el": r["feature_el"]}
                dpb_o = {"ref_frame_bl": o["x_hat_bl"], "ref_frame_el": o["x_hat_el"], "ref_feature_bl": None,
                         "ref_feature_el": o["feature_el"]}
                fr = {"type": "I", "x_hat_bl": sub(o["x_hat_bl"], 4), "x_hat_el": sub(o["x_hat_el"]),
                      "feature_el": sub(o["feature_el"]),
                      "sym_el": torch.round(o["y"] - o["means"]).to(torch.int16),
                      "index_el": orc.build_indexes_image(o["scales"]).to(torch.uint8),
                      "sym_bl": torch.round(o["bl"]["y"] - o["bl"]["means"]).to(torch.int16)}
            else:
                r = ref_p.encode_decode(x_bl, x_el, dpb_r, None, None, W, H, x_bl.shape[3], x_bl.shape[2])
                o = orc.lssvc(sd_p, x_bl, x_el, dpb_o, (H, W), 2.0)
                for k in r["dpb"]:
                    assert torch.equal(r["dpb"][k], o["dpb"][k]), k
                assert torch.equal(r["mv_hat"], o["mv_hat"]) and torch.equal(r["warp_frame"], o["warp_frame"])
                dpb_r, dpb_o = r["dpb"], o["dpb"]
                fp = o["four_part"]
                fr = {"type": "P", "x_hat_bl": sub(o["dpb"]["ref_frame_bl"], 4), "x_hat_el": sub(o["dpb"]["ref_frame_el"]),
                      "feature_el": sub(o["dpb"]["ref_feature_el"]), "mv_hat": sub(o["mv_hat"]),
                      "sym_el": fp["y_q"].to(torch.int16), "index_el": orc.build_indexes_video(fp["scales_hat"]).to(torch.uint8),
                      "sym_mv": o["mv_y_q"].to(torch.int16), "sym_z": o["z_hat"].to(torch.int16),
                      "sym_bl": o["bl"]["y_q"].to(torch.int16)}
            assert r["bit_bl"] == o["bit_bl"] and r["bit_el"] == o["bit_el"], (r["bit_bl"], o["bit_bl"])
            fr["bit_bl"], fr["bit_el"] = o["bit_bl"], o["bit_el"]
            rec["frames"].append(fr)
            for d in (dpb_r, dpb_o):
                d["ref_frame_bl"] = d["ref_frame_bl"].clamp_(0, 1)
                d["ref_frame_el"] = d["ref_frame_el"].clamp_(0, 1)
            print(f"frame {t} ({fr['type']}): bits {fr['bit_bl']:.2f} / {fr['bit_el']:.2f}  reference == oracle")
    torch.save(rec, os.path.join(GOLD, "forward_128.pt"))
    print("forward_128.pt:", os.path.getsize(os.path.join(GOLD, "forward_128.pt")) // 1024, "KiB")


def golden_tables_and_rans(LSSVC_extend, IntraSS, seed=0):
    from src.entropy_models.video_entropy_models import BitEstimator, EntropyCoder, GaussianEncoder
    ec = EntropyCoder()
    ge = GaussianEncoder()
    ge.update(force=True, entropy_coder=ec)
    cdf, sizes, offs = ge.cdf_helper.get_cdf_info_list()
    tables = {"laplace": {"cdf": sha(np.asarray(cdf, np.int32)), "sizes": sha(np.asarray(sizes, np.int32)),
                          "offsets": sha(np.asarray(offs, np.int32)), "shape": list(np.asarray(cdf).shape)}}
    torch.manual_seed(seed)
    be = BitEstimator(64)
    with torch.no_grad():
        for f in (be.f1, be.f2, be.f3, be.f4):
            f.h.normal_(-0.5, 0.3)
            f.b.normal_(0, 0.3)
            if f.a is not None:
                f.a.normal_(0, 0.3)
    be.update(force=True, entropy_coder=ec)
    bcdf, bsizes, boffs = be.cdf_helper.get_cdf_info_list()
    tables["bitparm"] = {"cdf": sha(np.asarray(bcdf, np.int32)), "sizes": sha(np.asarray(bsizes, np.int32)),
                         "offsets": sha(np.asarray(boffs, np.int32)), "shape": list(np.asarray(bcdf).shape)}
    bit_sd = {k: v.clone() for k, v in be.state_dict().items()}
    # image-side tables: GaussianConditional and EntropyBottleneck of a default IntraSS
    # (their update() needs RansEncoder, absent from src/cpp: give the proxy a stand-in, tables do not depend on it)
    import src.entropy_models.img_entropy_models as iem
    iem._EntropyCoder = lambda: object()
    torch.manual_seed(seed)
    net = IntraSS()
    net.update(force=True)
    gc = net.gaussian_conditional
    eb = net.entropy_bottleneck
    tables["gaussian"] = {"cdf": sha(gc._quantized_cdf.numpy().astype(np.int32)), "sizes": sha(gc._cdf_length.numpy().astype(np.int32)),
                          "offsets": sha(gc._offset.numpy().astype(np.int32)), "shape": list(gc._quantized_cdf.shape)}
    tables["eb"] = {"cdf": sha(eb._quantized_cdf.numpy().astype(np.int32)), "sizes": sha(eb._cdf_length.numpy().astype(np.int32)),
                    "offsets": sha(eb._offset.numpy().astype(np.int32)), "shape": list(eb._quantized_cdf.shape)}
    eb_sd = {k: v.clone() for k, v in eb.state_dict().items() if not k.startswith("_q") and not k.startswith("_o") and not k.startswith("_c")}
    with open(os.path.join(GOLD, "entropy_tables.json"), "w") as f:
        json.dump(tables, f, indent=1)
    torch.save({"bitparm_state": bit_sd, "eb_state": eb_sd}, os.path.join(GOLD, "entropy_params.pt"))
    print("tables:", {k: v["shape"] for k, v in tables.items()})

    # rANS known-answer vectors from the reference's own coder
    rng = np.random.default_rng(seed)
    vec = {}
    ncdf = np.asarray(cdf, np.int32)
    for name, n, scale in (("small", 257, 1.5), ("bypass", 4000, 30.0), ("big", 60000, 3.0)):
        idx = rng.integers(0, 256, size=n).astype(np.int32)
        symv = np.round(rng.laplace(0, scale, size=n)).astype(np.int32)
        if name == "bypass":
            symv[::97] = rng.integers(-70000, 70000, size=symv[::97].size)
        ec.reset_encoder()
        ec.encode_with_indexes(symv.tolist(), idx.tolist(), cdf, sizes, offs)
        if name == "big":   # two pushes into one stream, like the per-frame P stream
            ec.encode_with_indexes(symv[:1000].tolist(), idx[:1000].tolist(), cdf, sizes, offs)
        stream = ec.flush_encoder()
        ec.set_stream(stream)
        dec = np.asarray(ec.decoder.decode_stream(idx.tolist(), cdf, sizes, offs), np.int32)
        assert np.array_equal(dec, symv), name
        vec[name + "_sym"], vec[name + "_idx"] = symv, idx
        vec[name + "_bytes"] = np.frombuffer(stream, dtype=np.uint8)
        print(f"rans vector {name}: {n} symbols -> {len(stream)} bytes")
    pm = rng.random(40).astype(np.float32)
    pm /= pm.sum()
    pm[5:9] = 1e-9
    vec["pmf"] = pm
    vec["pmf_cdf"] = np.asarray(sys.modules["src.entropy_models.MLCodec_CXX"].pmf_to_quantized_cdf(pm.tolist(), 16), np.uint32)
    np.savez_compressed(os.path.join(GOLD, "rans_vectors.npz"), **vec)


def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.manual_seed(0)
    torch.set_num_threads(8)
    IntraSS, LSSVC_extend = ref_harness.import_reference()
    load_ref_native()
    golden_layout(IntraSS, LSSVC_extend)
    golden_tables_and_rans(LSSVC_extend, IntraSS)
    golden_forward(IntraSS, LSSVC_extend)


if __name__ == "__main__":
    main()
