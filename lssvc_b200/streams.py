"""--write_stream 1 orchestration: real rANS bitstreams in the reference's container and symbol order.

Encoder side = the CUDA forward pass with the symbol / CDF-index dumps switched on (the entropy kernels write int32
NCHW buffers next to their NHWC outputs), then the host rANS coder (csrc/rans.cpp, byte-identical to src/cpp/rans).

  P-frame, per layer (LSSVC_net_extend.py:66-74, dmc_net_extend.py:91-96): ONE string
        [mv_z | mv_y | z | y]  (EL: y as the four parts y_w0..y_w3 of the 4-step prior), file = >I length + string
  I-frame, per layer (IntraSS.py:251-274, priors.py:420-435): y-string and z-string, file = >4I (H, W, len y, len z)

This module is the SINGLE-PASS variant (model.single_pass_streams = True): one encoder pass codes both layers, the streams
are decoded with the CDF rows of the dumped indices and required to return exactly the coded symbols, and the DPB comes
from the encoder-side pass.  The default path (codec.py) runs the genuine decoder — synthesis networks on the decoded
symbols, progressive four-part decode — and takes the DPB from it as the reference does (LSSVC_net_extend.py:168-178);
both paths write byte-identical files and bit-identical reconstructions (tests/test_parity_gpu.py).
"""
import time

import numpy as np
import torch

from . import entropy, stream

from .codec import _Dump as SymbolDump     # model._write hook: int32 NCHW device buffers for symbols / CDF rows
from .codec import _encode


def _verify(string, parts, what):
    dec = entropy.RansDecoder()
    dec.set_stream(string)
    for k, (sym, idx, table) in enumerate(parts):
        got = dec.decode_stream(idx, table)
        if not np.array_equal(got, np.asarray(sym, dtype=np.int32).reshape(-1)):
            raise RuntimeError(f"{what}: part {k} of the stream does not decode to the coded symbols")


_POOL = None


def _pool():
    global _POOL
    if _POOL is None:
        from concurrent.futures import ThreadPoolExecutor
        _POOL = ThreadPoolExecutor(max_workers=2, thread_name_prefix="lssvc-rans")   # one worker per layer string
    return _POOL


def _layer_parts(dump, tag, t):
    """[(symbols, CDF rows, table)] of one layer's string in push order (dmc_net_extend.py:92-98, LSSVC_net_extend.py:66-74)."""
    h = dump.host
    parts = [(h(f"{tag}_mv_z"), dump.channel_index(f"{tag}_mv_z"), t[f"{tag}_mv_z"]), (h(f"{tag}_mv_y"), h(f"{tag}_mv_y_idx"), t["laplace"]),
             (h(f"{tag}_z"), dump.channel_index(f"{tag}_z"), t[f"{tag}_z"])]
    if tag == "bl":
        parts.append((h("bl_y"), h("bl_y_idx"), t["laplace"]))
    else:
        parts += [(h(f"el_y{k}"), h(f"el_y{k}_idx"), t["laplace"]) for k in range(4)]
    return parts


def _encode_job(dump, tag, tables, path):
    """Worker thread: waits for the layer's symbols to land in pinned memory (a CUDA event: the GPU is still running the
    layer's synthesis networks), codes the string and writes the file.  The C coder releases the GIL."""
    t0 = time.time()
    parts = _layer_parts(dump, tag, tables)
    t1 = time.time()
    stream.encode_p(_encode(parts), path)
    return parts, stream.filesize(path) * 8, time.time() - t1, t1 - t0


def _verify_job(path, parts, what):
    t0 = time.time()
    _verify(stream.decode_p(path), parts, what)
    return time.time() - t0


def drain(model):
    """Waits for the background stream verifications of earlier frames; raises what they raised."""
    checks, model.__dict__["_stream_checks"] = model.__dict__.get("_stream_checks", []), []
    times = [c.result() for c in checks]
    if times:
        model.__dict__["stream_check_seconds"] = times
    return times


def inter_encode_decode(model, x_bl, x_el, dpb, output_path_bl, output_path_el, pic_width, pic_height, pic_width_bl,
                        pic_height_bl):
    """LSSVC_extend.encode_decode_extend (LSSVC_net_extend.py:144-191) + DMCExtend.encode_decode_extend, single pass and
    overlapped: ONE forward pass codes both layers; the symbols of a layer leave for pinned host memory as soon
    as its last entropy kernel has been issued (codec._Dump.layer_done) and a worker thread per layer codes its string while
    the GPU is still running the rest of the frame; the decode-and-compare verification of the two strings runs in the
    background, overlapped with the NEXT frame's forward pass (its failure is raised by the next call, or by drain())."""
    model.update()
    t = model._tables
    futures = {}

    def on_layer(tag):
        if tag in futures:         # the frame is being coded again on the fp32 engine (models._recode_fp32): same dump, new symbols
            futures.pop(tag).result()
        futures[tag] = _pool().submit(_encode_job, dump, tag, t, output_path_bl if tag == "bl" else output_path_el)

    dump = SymbolDump(model.device, owner=model, on_layer=on_layer)
    torch.cuda.synchronize(model.device)
    t0 = time.time()
    r = model.forward_one_frame(x_bl, x_el, dpb["ref_frame_bl"], dpb["ref_frame_el"], dpb["ref_feature_bl"],
                                dpb["ref_feature_el"], _dpb=dpb, _write=dump)
    t1 = time.time()
    bl_parts, r["bit_bl"], enc_bl, _ = futures["bl"].result()
    el_parts, r["bit_el"], enc_el, _ = futures["el"].result()
    t2 = time.time()
    drain(model)                                   # frame t - 1's verification (ran under this frame's forward pass)
    model.__dict__["_stream_checks"] = [_pool().submit(_verify_job, output_path_bl, bl_parts, "base layer"),
                                        _pool().submit(_verify_job, output_path_el, el_parts, "enhancement layer")]
    # the decoder of the reference's base layer clamps its reconstruction (dmc_net_extend.py:138); this path takes the
    # DPB from the encoder-side pass, so the clamp is applied here to hand out the same tensor as codec.bl_decompress
    r["dpb"]["ref_frame_bl"].clamp_(0, 1)
    # ONE forward pass codes both layers; the rANS passes ran under it.  encoding_time_EL = the whole call (forward pass +
    # what was left of the string coding when it returned), encoding_time_BL = the BL string's own coding time;
    # decoding (verification) is off the critical path: model.stream_check_seconds holds the previous frame's figures
    r["encoding_time_BL"], r["encoding_time_EL"] = enc_bl, t2 - t0
    r["decoding_time_BL"], r["decoding_time_EL"] = 0.0, 0.0
    return r


def intra_encode_decode(model, x_bl, x_el, bin_path_bl, bin_path_el, pic_height_bl, pic_width_bl, pic_height_el,
                        pic_width_el):
    """IntraSS.encode_decode with bitstreams (IntraSS.py:245-302)."""
    model.update()
    t = model._tables
    dump = SymbolDump(model.device)
    r = model.forward(x_bl, x_el, _write=dump)
    torch.cuda.synchronize(model.device)
    h = dump.host
    for tag, path, ph, pw in (("bl", bin_path_bl, pic_height_bl, pic_width_bl), ("el", bin_path_el, pic_height_el, pic_width_el)):
        y_part = (h(f"{tag}_y"), h(f"{tag}_y_idx"), t["gaussian"])
        z_part = (h(f"{tag}_z"), dump.channel_index(f"{tag}_z"), t[f"{tag}_z"])
        stream.encode_i(ph, pw, _encode([y_part]), _encode([z_part]), path)
        _, _, y_string, z_string = stream.decode_i(path)
        _verify(y_string, [y_part], f"{tag} y")
        _verify(z_string, [z_part], f"{tag} z")
        r[f"bit_{tag}"] = stream.filesize(path) * 8
    return r
