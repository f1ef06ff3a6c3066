"""Composes the per-frame bitstream FILES of the reference's `--write_stream 1` flow from the ORACLE's tensors (symbols and
scales of one coded frame) with the PRODUCT's host code: table builders (lssvc_b200/entropy.py), rANS coder (C-ABI,
csrc/rans.cpp) and container (lssvc_b200/stream.py).  tools/make_golden_streams.py requires the result to equal the files the
unmodified reference wrote; tests/test_streams.py re-checks it on every box (no GPU needed), which pins the symbol order,
the CDF-row rule, the tables, the coder and the container against reference-produced bytes.

Stream composition followed (file:line under /root/reference/src):
  P, per layer, ONE string:  BL  [mv_z | mv_y | z | y]                         models/dmc_net_extend.py:92-98
                             EL  [mv_z | mv_y | z | y_w0 | y_w1 | y_w2 | y_w3]  models/LSSVC_net_extend.py:66-74
      symbols NCHW-flattened (`x.reshape(-1).int().tolist()`, entropy_models/video_entropy_models.py:232-236, 315-319);
      CDF row = channel for the factorised z (:225-230), build_indexes(scales) for y (:309-313);
      file = `>I` length + string                                              utils/stream_helper.py:85-89
  I, per layer, TWO strings: y = GaussianConditional.compress(y, build_indexes(scales), means), z = EntropyBottleneck.compress(z)
      (symbols = round(x - means|medians), entropy_models/img_entropy_models.py:286-313, 563-566, 687-691;
      models/priors.py:420-435, models/IntraSS.py:304-314); file = `>4I` (H, W, len y, len z) + y + z   stream_helper.py:61-68"""
import os
import tempfile

import numpy as np
import torch

from lssvc_b200 import entropy as E
from lssvc_b200 import stream


def _flat(t):
    return t.detach().reshape(-1).to(torch.int32).numpy()


def _channel_rows(t):
    _, C, H, W = t.shape
    return np.repeat(np.arange(C, dtype=np.int32), H * W)


def _encode(parts):
    enc = E.RansEncoder()
    for sym, idx, table in parts:
        enc.encode_with_indexes(sym, idx, table)
    return enc.flush()


def _file(write, *args):
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "f.bin")
        write(*args, path)
        with open(path, "rb") as f:
            return f.read()


def video_tables(sd_p):
    g = lambda p: E.bitparm_table(E.bitparm_coef([sd_p[f"{p}f{i}.h"] for i in (1, 2, 3, 4)], [sd_p[f"{p}f{i}.b"] for i in (1, 2, 3, 4)],
                                                 [sd_p[f"{p}f{i}.a"] for i in (1, 2, 3)]))
    return {"laplace": E.laplace_table(), "el_z": g("bit_estimator_z."), "el_mv_z": g("bit_estimator_z_mv."),
            "bl_z": g("base_layer_model.bit_estimator_z."), "bl_mv_z": g("base_layer_model.bit_estimator_z_mv.")}


def image_tables(sd_i):
    def eb(p):
        return E.eb_table([sd_i[f"{p}_matrices.{i}"] for i in range(5)], [sd_i[f"{p}_biases.{i}"] for i in range(5)],
                          [sd_i[f"{p}_factors.{i}"] for i in range(4)], sd_i[f"{p}quantiles"])
    return {"gaussian": E.gaussian_table(), "bl_z": eb("base_layer_model.entropy_bottleneck."), "el_z": eb("entropy_bottleneck.")}


def inter_strings(orc, o, sd_p, tables=None):
    """(BL string, EL string) of one P-frame from the oracle's result dict `o` (oracle.lssvc_oracle.lssvc)."""
    t = tables or video_tables(sd_p)
    bl, fp = o["bl"], o["four_part"]
    lap = t["laplace"]
    s_bl = _encode([(_flat(bl["mv_z_hat"]), _channel_rows(bl["mv_z_hat"]), t["bl_mv_z"]),
                    (_flat(bl["mv_y_q"]), _flat(orc.build_indexes_video(bl["mv_scales"])), lap),
                    (_flat(bl["z_hat"]), _channel_rows(bl["z_hat"]), t["bl_z"]),
                    (_flat(bl["y_q"]), _flat(orc.build_indexes_video(bl["scales"])), lap)])
    parts = [(_flat(o["mv_z_hat"]), _channel_rows(o["mv_z_hat"]), t["el_mv_z"]),
             (_flat(o["mv_y_q"]), _flat(orc.build_indexes_video(o["mv_scales"])), lap),
             (_flat(o["z_hat"]), _channel_rows(o["z_hat"]), t["el_z"])]
    parts += [(_flat(q), _flat(orc.build_indexes_video(s)), lap) for q, s in zip(fp["y_q_w"], fp["scales_w"])]
    return s_bl, _encode(parts)


def inter_files(orc, o, sd_p, coder=None, tables=None):
    s_bl, s_el = inter_strings(orc, o, sd_p, tables)
    return _file(stream.encode_p, s_bl), _file(stream.encode_p, s_el)


def intra_strings(orc, o, sd_i, tables=None):
    """((y, z) strings of the BL, (y, z) strings of the EL) of one I-frame from oracle.lssvc_oracle.intra_ss's result."""
    t = tables or image_tables(sd_i)
    out = []
    for layer, prefix, ztab in ((o["bl"], "base_layer_model.entropy_bottleneck.", t["bl_z"]), (o, "entropy_bottleneck.", t["el_z"])):
        med = sd_i[prefix + "quantiles"].detach().float()[:, 0, 1].view(1, -1, 1, 1)
        z_sym = torch.round(layer["z"] - med)
        y_sym = torch.round(layer["y"] - layer["means"])
        y_string = _encode([(_flat(y_sym), _flat(orc.build_indexes_image(layer["scales"])), t["gaussian"])])
        z_string = _encode([(_flat(z_sym), _channel_rows(z_sym), ztab)])
        out.append((y_string, z_string))
    return out


def intra_files(orc, o, sd_i, H, W, coder=None, tables=None):
    (yb, zb), (ye, ze) = intra_strings(orc, o, sd_i, tables)
    return _file(stream.encode_i, H // 2, W // 2, yb, zb), _file(stream.encode_i, H, W, ye, ze)
