"""Bitstream compatibility with the REFERENCE ITSELF: tests/golden/streams_128.pt holds the files the
unmodified reference wrote for I + 2 P frames at EL 128x128 in `--write_stream 1` mode with its own C++ coder
(tools/make_golden_streams.py, which also proved that the reference's stream-mode DPB equals the oracle's estimate-mode DPB
bit for bit, so the DPBs are rebuilt here from the oracle).

CPU (every box):
  * the product's host code (table builders, rANS coder, container) fed with the oracle's symbols and scales writes the
    reference's files byte for byte, and decodes them back to the oracle's symbols  -> pins symbol order, CDF-row rule,
    tables, coder and container against reference-produced bytes;
  * the importable MLCodec_rans / MLCodec_CXX shims behave like the reference's pybind11 modules, and — where /root/reference
    exists — the reference's own EntropyCoder / GaussianEncoder / BitEstimator / EntropyBottleneck classes run on them and
    produce the reference's bytes.
GPU:
  * the CUDA ENCODER on the fp32 engine (symbols and indices exact) writes the reference's files byte for byte through the
    public `encode_decode(..., bin paths)` API;
  * the CUDA DECODER (`decompress` of both layers) reads the REFERENCE's strings and reproduces the reference decoder's
    reconstruction within 1e-3 — default tensor-core engine and fp32 engine."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import stream_compose  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden", "streams_128.pt")


@pytest.fixture(scope="module")
def gold():
    return torch.load(GOLD, weights_only=False)


@pytest.fixture(scope="module")
def coded(gold):
    """The oracle's I + 2 P frames on the fixture's frames and weights (CPU, a few seconds)."""
    from lssvc_b200 import nets, synth
    from oracle import lssvc_oracle as orc
    torch.set_num_threads(8)
    H, W, seed = gold["H"], gold["W"], gold["seed"]
    sd_i = nets.ParamBag(nets.intra_ss_spec(), seed=seed, gains=nets.model_gains("I")).state_dict()
    sd_p = nets.ParamBag(nets.lssvc_spec(), seed=seed + 1, gains=nets.model_gains("P")).state_dict()
    frames = synth.make_sequence(H, W, 3, seed=seed)
    outs, dpbs = [], []
    with torch.no_grad():
        o = orc.intra_ss(sd_i, frames[0][0], frames[0][1], (H, W))
        outs.append(o)
        dpb = {"ref_frame_bl": o["x_hat_bl"].clamp(0, 1), "ref_frame_el": o["x_hat_el"].clamp(0, 1), "ref_feature_bl": None,
               "ref_feature_el": o["feature_el"]}
        for t in (1, 2):
            dpbs.append(dpb)
            o = orc.lssvc(sd_p, frames[t][0], frames[t][1], dpb, (H, W), 2.0)
            outs.append(o)
            dpb = dict(o["dpb"])
            dpb["ref_frame_bl"] = dpb["ref_frame_bl"].clamp(0, 1)
            dpb["ref_frame_el"] = dpb["ref_frame_el"].clamp(0, 1)
    return {"orc": orc, "sd_i": sd_i, "sd_p": sd_p, "frames": frames, "o": outs, "dpb": dpbs, "H": H, "W": W}


# ---------------------------------------------------------------------------------------------------------------------
# CPU
# ---------------------------------------------------------------------------------------------------------------------
def test_product_coder_writes_the_reference_files(gold, coded):
    c = coded
    files = [stream_compose.intra_files(c["orc"], c["o"][0], c["sd_i"], c["H"], c["W"])]
    tv = stream_compose.video_tables(c["sd_p"])
    files += [stream_compose.inter_files(c["orc"], c["o"][t], c["sd_p"], tables=tv) for t in (1, 2)]
    for t, (f_bl, f_el) in enumerate(files):
        g = gold["frames"][t]
        assert f_bl == g["file_bl"], f"frame {t}: BL file differs from the reference's ({len(f_bl)} vs {len(g['file_bl'])} B)"
        assert f_el == g["file_el"], f"frame {t}: EL file differs from the reference's ({len(f_el)} vs {len(g['file_el'])} B)"
        assert 8 * len(f_bl) == g["bit_bl"] and 8 * len(f_el) == g["bit_el"]
        print(f"frame {t} ({g['type']}): {len(f_bl)} + {len(f_el)} B == reference files")


def test_product_decoder_reads_the_reference_files(gold, coded, tmp_path):
    """Decoder direction on the host: the reference's strings, the oracle's CDF rows -> the oracle's symbols, in the
    reference's decode order (dmc_net_extend.py:106-135, LSSVC_net_extend.py:104-263, priors.py:437-452, IntraSS.py:316-336)."""
    from lssvc_b200 import entropy as E
    from lssvc_b200 import stream
    c, orc = coded, coded["orc"]
    flat = stream_compose._flat
    rows = stream_compose._channel_rows
    # ---- P-frames
    tv = stream_compose.video_tables(c["sd_p"])
    for t in (1, 2):
        o, g = c["o"][t], gold["frames"][t]
        for layer, data in (("bl", g["file_bl"]), ("el", g["file_el"])):
            path = tmp_path / f"{t}{layer}.bin"
            path.write_bytes(data)
            dec = E.RansDecoder()
            dec.set_stream(stream.decode_p(str(path)))
            src = o["bl"] if layer == "bl" else o
            steps = [(src["mv_z_hat"], rows(src["mv_z_hat"]), tv[layer + "_mv_z"]),
                     (src["mv_y_q"], flat(orc.build_indexes_video(src["mv_scales"])), tv["laplace"]),
                     (src["z_hat"], rows(src["z_hat"]), tv[layer + "_z"])]
            if layer == "bl":
                steps.append((src["y_q"], flat(orc.build_indexes_video(src["scales"])), tv["laplace"]))
            else:
                steps += [(q, flat(orc.build_indexes_video(s)), tv["laplace"])
                          for q, s in zip(o["four_part"]["y_q_w"], o["four_part"]["scales_w"])]
            for want, idx, table in steps:
                got = dec.decode_stream(idx, table)
                assert np.array_equal(got, flat(want)), f"frame {t} {layer}: decoded symbols differ"
    # ---- I-frame
    ti = stream_compose.image_tables(c["sd_i"])
    o, g = c["o"][0], gold["frames"][0]
    for layer, data, src, prefix, ztab in (("bl", g["file_bl"], o["bl"], "base_layer_model.entropy_bottleneck.", ti["bl_z"]),
                                           ("el", g["file_el"], o, "entropy_bottleneck.", ti["el_z"])):
        path = tmp_path / f"0{layer}.bin"
        path.write_bytes(data)
        h, w, y_string, z_string = stream.decode_i(str(path))
        assert (h, w) == ((c["H"] // 2, c["W"] // 2) if layer == "bl" else (c["H"], c["W"]))
        assert stream.get_downsampled_shape(h, w, 64) == tuple(src["z"].shape[-2:])
        med = c["sd_i"][prefix + "quantiles"].detach().float()[:, 0, 1].view(1, -1, 1, 1)
        dec = E.RansDecoder()
        dec.set_stream(z_string)
        assert np.array_equal(dec.decode_stream(rows(src["z"]), ztab), flat(torch.round(src["z"] - med)))
        dec.set_stream(y_string)
        got = dec.decode_stream(flat(orc.build_indexes_image(src["scales"])), ti["gaussian"])
        assert np.array_equal(got, flat(torch.round(src["y"] - src["means"])))


def test_mlcodec_shims_have_the_reference_surface():
    """Names and call signatures of rans_interface.cpp:246-261 / ops.cpp:84-91 (+ the image path's RansEncoder /
    decode_with_indexes), list arguments as the reference passes them, round trip incl. bypass symbols."""
    from lssvc_b200 import MLCodec_CXX, MLCodec_rans, compat
    from lssvc_b200 import entropy as E
    assert compat.install("some_pkg.entropy_models") == (MLCodec_rans, MLCodec_CXX)
    assert sys.modules["some_pkg.entropy_models.MLCodec_rans"] is MLCodec_rans
    lap = E.laplace_table()
    cdfs, sizes, offs = lap.cdf.tolist(), lap.sizes.tolist(), lap.offsets.tolist()
    rng = np.random.default_rng(5)
    idx = rng.integers(0, 256, size=3000).astype(np.int32)
    sym = np.round(rng.laplace(0, 4.0, size=3000)).astype(np.int32)
    sym[::53] = rng.integers(-70000, 70000, size=sym[::53].size)
    enc = MLCodec_rans.BufferedRansEncoder()
    assert enc.encode_with_indexes(sym.tolist(), idx.tolist(), cdfs, sizes, offs) is None
    enc.encode_with_indexes(sym[:100].tolist(), idx[:100].tolist(), cdfs, sizes, offs)
    s = enc.flush()
    assert isinstance(s, bytes) and len(s) % 4 == 0
    dec = MLCodec_rans.RansDecoder()
    dec.set_stream(s)
    a = dec.decode_stream(idx.tolist(), cdfs, sizes, offs)
    b = dec.decode_stream(idx[:100].tolist(), cdfs, sizes, offs)
    assert np.array_equal(np.asarray(a), sym) and np.array_equal(np.asarray(b), sym[:100])
    enc.encode_with_indexes([1, 2], [0, 0], cdfs, sizes, offs)
    enc.reset()
    assert len(enc.flush()) == 8                                   # nothing buffered after reset(): the bare state
    one = MLCodec_rans.RansEncoder().encode_with_indexes(sym.tolist(), idx.tolist(), cdfs, sizes, offs)
    assert MLCodec_rans.RansDecoder().decode_with_indexes(one, idx.tolist(), cdfs, sizes, offs) == sym.tolist()
    v = np.load(os.path.join(ROOT, "tests", "golden", "rans_vectors.npz"))
    assert MLCodec_CXX.pmf_to_quantized_cdf(v["pmf"].tolist(), 16) == v["pmf_cdf"].tolist()
    # argument validation the reference leaves to asserts: a row outside the table must not be dereferenced
    from lssvc_b200._lib import LssvcError
    with pytest.raises(LssvcError):
        MLCodec_rans.RansEncoder().encode_with_indexes([0], [256], cdfs, sizes, offs)
    with pytest.raises(LssvcError):
        dec.set_stream(s)
        dec.decode_stream([-1], cdfs, sizes, offs)


@pytest.mark.skipif(not os.path.isdir("/root/reference/src/entropy_models"), reason="reference sources not on this box")
def test_reference_entropy_models_run_on_the_shims(gold, coded):
    """The reference's OWN classes (EntropyCoder video_entropy_models.py:8-61, GaussianEncoder :247-336, BitEstimator
    :150-244, EntropyBottleneck / GaussianConditional img_entropy_models.py) imported from /root/reference with
    lssvc_b200.compat.install() in place of its binaries: update() builds the tables through MLCodec_CXX, encode / flush
    through MLCodec_rans, and the P-frame EL string + the I-frame BL strings come out byte-identical to the golden files."""
    import importlib
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import ref_harness
    from lssvc_b200 import compat
    saved = {k: sys.modules.get(k) for k in ("src.entropy_models.MLCodec_rans", "src.entropy_models.MLCodec_CXX")}
    try:
        ref_harness.import_reference()
        compat.install()
        vem = importlib.import_module("src.entropy_models.video_entropy_models")
        iem = importlib.import_module("src.entropy_models.img_entropy_models")
        c, o = coded, coded["o"][1]
        ec = vem.EntropyCoder()
        ge = vem.GaussianEncoder()
        ge.update(force=True, entropy_coder=ec)
        bes = {}
        for tag, ch in (("bit_estimator_z_mv.", 64), ("bit_estimator_z.", 128)):
            be = vem.BitEstimator(ch)
            be.load_state_dict({k[len(tag):]: v for k, v in c["sd_p"].items() if k.startswith(tag)})
            be.update(force=True, entropy_coder=ec)
            bes[tag] = be
        ec.reset_encoder()
        bes["bit_estimator_z_mv."].encode(o["mv_z_hat"])
        ge.encode(o["mv_y_q"], o["mv_scales"])
        bes["bit_estimator_z."].encode(o["z_hat"])
        for q, s in zip(o["four_part"]["y_q_w"], o["four_part"]["scales_w"]):
            ge.encode(q, s)
        string = ec.flush_encoder()
        assert gold["frames"][1]["file_el"][4:] == string, "reference EntropyCoder on the shim: EL string differs from the golden file"
        ec.set_stream(string)
        assert torch.equal(bes["bit_estimator_z_mv."].decode_stream(o["mv_z_hat"].shape[-2:]), o["mv_z_hat"])
        assert torch.equal(ge.decode_stream(o["mv_scales"]), o["mv_y_q"])
        # image path: EntropyBottleneck + GaussianConditional of the BL I-frame codec
        oi = coded["o"][0]["bl"]
        eb = iem.EntropyBottleneck(192)
        p = "base_layer_model.entropy_bottleneck."
        sd = {k[len(p):]: v for k, v in c["sd_i"].items() if k.startswith(p)}
        for k in ("_offset", "_quantized_cdf", "_cdf_length"):
            sd.pop(k, None)
        eb.load_state_dict(sd, strict=False)
        eb.update(force=True)
        gc = iem.GaussianConditional()
        gc.update()
        z_strings = eb.compress(oi["z"])
        y_strings = gc.compress(oi["y"], gc.build_indexes(oi["scales"]), means=oi["means"])
        f = gold["frames"][0]["file_bl"]
        assert f[16:16 + len(y_strings[0])] == y_strings[0] and f[16 + len(y_strings[0]):] == z_strings[0]
        z_hat = eb.decompress(z_strings, oi["z"].shape[-2:])
        assert torch.equal(z_hat, oi["z_hat"])
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


# ---------------------------------------------------------------------------------------------------------------------
# GPU
# ---------------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def nets_gpu(cuda_device, coded):
    from lssvc_b200 import IntraSS, LSSVC_extend
    net_i, net_p = IntraSS(seed=0), LSSVC_extend(seed=1)
    net_i.to(cuda_device)
    net_p.to(cuda_device)
    for n in (net_i, net_p):
        n.set_scale_information(2.0, (coded["H"], coded["W"]), (0, 0, 0, 0))
        n.update(force=True)
    return net_i, net_p


def _dev(d, dev):
    return {k: (None if v is None else v.to(dev)) for k, v in d.items()}


@pytest.mark.gpu
def test_cuda_encoder_writes_the_reference_files(gold, coded, nets_gpu, cuda_device, tmp_path):
    """encode_decode(..., bin paths) on the fp32 CUDA-core engine (its symbols and CDF rows equal the oracle's), every frame
    coded from the oracle's DPB as test.py would hand it over: the files on disk are the reference's, byte for byte."""
    from lssvc_b200 import ops
    net_i, net_p = nets_gpu
    c, dev = coded, cuda_device
    H, W = c["H"], c["W"]
    prev = ops.set_engine("simt")
    try:
        for t in range(3):
            x_bl, x_el = (x.to(dev) for x in c["frames"][t])
            p_bl, p_el = str(tmp_path / f"{t}_bl.bin"), str(tmp_path / f"{t}_el.bin")
            if t == 0:
                r = net_i.encode_decode(x_bl, x_el, p_bl, p_el, H // 2, W // 2, H, W)
            else:
                r = net_p.encode_decode(x_bl, x_el, _dev(c["dpb"][t - 1], dev), p_bl, p_el, W, H, W // 2, H // 2)
            g = gold["frames"][t]
            for layer, path in (("bl", p_bl), ("el", p_el)):
                data = open(path, "rb").read()
                assert data == g["file_" + layer], (f"frame {t} {layer}: {len(data)} B written, reference file has "
                                                    f"{len(g['file_' + layer])} B" + ("" if len(data) != len(g["file_" + layer]) else " (same size, different bytes)"))
                assert r["bit_" + layer] == g["bit_" + layer]
            print(f"frame {t} ({g['type']}): CUDA encoder files == reference files ({g['bit_bl'] // 8} + {g['bit_el'] // 8} B)")
    finally:
        ops.set_engine(prev)


@pytest.mark.gpu
@pytest.mark.parametrize("engine", ["default", "simt"])
def test_cuda_decoder_reads_the_reference_streams(gold, coded, nets_gpu, cuda_device, tmp_path, engine):
    """decompress() of every layer given ONLY the reference's string and the DPB: reconstructions within 1e-3 of the reference
    decoder's (== the oracle's, proven at fixture time), latents y_hat within 1e-3 as well (a wrong CDF row anywhere would
    desynchronise the rANS decoder and show up as garbage or LSSVC_ERR_STREAM)."""
    from lssvc_b200 import ops, stream
    net_i, net_p = nets_gpu
    c, dev = coded, cuda_device
    H, W = c["H"], c["W"]
    prev = ops.set_engine(ops.default_engine() if engine == "default" else engine)

    def close(name, got, ref, tol=1e-3):
        d = (got.cpu() - ref).abs().max().item()
        print(f"  {name:18s} max|d| {d:.2e}")
        assert d < tol, f"{name}: {d:.3e}"

    try:
        # ---- I-frame (priors.py:437-452, IntraSS.py:316-336)
        g, o = gold["frames"][0], c["o"][0]
        (tmp_path / "i_bl.bin").write_bytes(g["file_bl"])
        (tmp_path / "i_el.bin").write_bytes(g["file_el"])
        h, w, ys, zs = stream.decode_i(str(tmp_path / "i_bl.bin"))
        dec_bl = net_i.base_layer_model.decompress([[ys], [zs]], stream.get_downsampled_shape(h, w, 64))
        close("I x_hat_bl", dec_bl["x_hat"], o["x_hat_bl"])
        close("I y_hat_bl", dec_bl["y_hat"], o["bl"]["y_hat"])
        h, w, ys, zs = stream.decode_i(str(tmp_path / "i_el.bin"))
        dec = net_i.decompress([[ys], [zs]], {"x_hat_bl": dec_bl["x_hat"], "y_hat_bl": dec_bl["y_hat"]},
                               stream.get_downsampled_shape(h, w, 64))
        close("I x_hat_el", dec["x_hat"], o["x_hat_el"])
        close("I feature_el", dec["feature"], o["feature_el"], 5e-3)
        # ---- P-frames (dmc_net_extend.py:106-147, LSSVC_net_extend.py:88-142, 200-263)
        for t in (1, 2):
            g, o = gold["frames"][t], c["o"][t]
            dpb = _dev(c["dpb"][t - 1], dev)
            (tmp_path / f"{t}_bl.bin").write_bytes(g["file_bl"])
            (tmp_path / f"{t}_el.bin").write_bytes(g["file_el"])
            bl = net_p.base_layer_model.decompress(stream.decode_p(str(tmp_path / f"{t}_bl.bin")), H // 2, W // 2, dpb)["dpb"]
            close(f"P{t} ref_frame_bl", bl["ref_frame_bl"], o["dpb"]["ref_frame_bl"].clamp(0, 1))
            close(f"P{t} y_hat_bl", bl["y_hat_bl"], o["bl"]["y_hat"])
            close(f"P{t} mv_hat_bl", bl["mv_hat_bl"], o["bl"]["mv_hat"])
            dpb["texture"], dpb["y_hat_bl"], dpb["mv_hat_bl"] = bl["ref_feature_bl"], bl["y_hat_bl"], bl["mv_hat_bl"]
            el = net_p.decompress(stream.decode_p(str(tmp_path / f"{t}_el.bin")), H, W, dpb)["dpb"]
            close(f"P{t} ref_frame_el", el["ref_frame_el"], o["dpb"]["ref_frame_el"])
            close(f"P{t} ref_feature_el", el["ref_feature_el"], o["dpb"]["ref_feature_el"], 5e-3)
    finally:
        ops.set_engine(prev)
