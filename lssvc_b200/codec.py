"""Real bitstream coding of P-frames: compress / decompress of the base layer (DMCExtend.compress / decompress,
dmc_net_extend.py:55-147) and of the enhancement layer (LSSVC_extend.compress / decompress / decompress_four_part_prior,
LSSVC_net_extend.py:24-142, 200-263), on the CUDA kernels.

The encoder is the forward pass with the symbol / CDF-index dumps switched on (one pass, no host round trip before the
end), then the host rANS coder.  The decoder is a genuine decoder: it only sees the string and the DPB; synthesis
networks run on the GPU, and at every point where the reference calls `decode_stream` the CDF indices of the scales just
computed go to the host (int32, NCHW order), the rANS decoder returns the symbols, and they come back as an NHWC view
(`symbols_to_view`, `four_part_dec_step`).  Same kernels, same order => the decoder's reconstruction is bit-identical
to the encoder's (tests/test_parity_gpu.py::test_bitstream_round_trip).

Stream order per layer (one string, one flush): BL [mv_z | mv_y | z | y], EL [mv_z | mv_y | z | y_w0 | y_w1 | y_w2 | y_w3].
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import entropy, ops, stream
from .ops import View


class _Dump:
    """model._write hook: int32 device buffers for symbols / CDF indices (NCHW order) next to the NHWC outputs.

    With `owner` (a model) the buffers of a layer travel to PINNED host memory as soon as the layer's last entropy kernel has
    been issued (layer_done, called by the models): the copy runs on a side stream behind an event, so the host coder can
    start while the GPU is still running the layer's synthesis networks.  Pinned buffers are cached on the
    owner, two sets used alternately so that a background job of the previous frame never sees them overwritten."""

    def __init__(self, device, owner=None, on_layer=None):
        self.device = device
        self.bufs = {}
        self.shapes = {}
        self.owner = owner
        self.on_layer = on_layer
        self.pinned = {}
        self.done = {}
        if owner is not None:
            st = owner.__dict__.setdefault("_dump_state", {"stream": torch.cuda.Stream(device=device), "sets": ({}, {}), "turn": 0})
            st["turn"] ^= 1
            self._copy_stream, self._cache = st["stream"], st["sets"][st["turn"]]

    def buf(self, name, view, C=None):
        C = view.real if C is None else C
        t = torch.empty(C * view.H * view.W, dtype=torch.int32, device=self.device)
        self.bufs[name] = t
        self.shapes[name] = (C, view.H, view.W)
        return t

    def layer_done(self, tag):
        """Every dump of layer `tag` ('bl' / 'el') has been issued on the current stream."""
        if self.owner is None:
            return
        ev = torch.cuda.Event()
        ev.record()
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(ev)
            for name, t in self.bufs.items():
                if name.startswith(tag + "_"):
                    p = self._cache.get(name)
                    if p is None or p.numel() != t.numel():
                        p = self._cache[name] = torch.empty(t.numel(), dtype=torch.int32).pin_memory()
                    p.copy_(t, non_blocking=True)
                    self.pinned[name] = p
            done = torch.cuda.Event()
            done.record()
        for name in self.pinned:
            if name.startswith(tag + "_"):
                self.done[name] = done
        if self.on_layer is not None:
            self.on_layer(tag)

    def host(self, name):
        p = self.pinned.get(name)
        if p is not None:
            self.done[name].synchronize()
            return p.numpy()
        return self.bufs[name].cpu().numpy()

    def channel_index(self, name):
        C, H, W = self.shapes[name]
        return _channel_index(C, H, W)


def _channel_index(C, H, W):
    """BitEstimator.build_indexes (video_entropy_models.py:225-230): the CDF row is the channel."""
    return np.repeat(np.arange(C, dtype=np.int32), H * W)


def _encode(parts):
    enc = entropy.RansEncoder()
    for sym, idx, table in parts:
        enc.encode_with_indexes(sym, idx, table)
    return enc.flush()


def _bits_scratch(model):
    from .models import _Bits
    return _Bits(model.device)


def _range_guarded(fn):
    """Stream-mode entry points: after the call the range flag of the split-fp16 kernels (csrc/range.cu) is fetched; if an
    operand left the fp16 range the whole call is repeated on the fp32 engine (models._recode_fp32)."""
    import functools

    @functools.wraps(fn)
    def inner(model, *a, **k):
        from .models import _recode_fp32
        out = fn(model, *a, **k)
        if ops.default_engine() != "h2":
            return out
        bits = _bits_scratch(model)
        bits.fetch_range()
        bits.read()
        if bits.out_of_range:
            return _recode_fp32(model, fn.__name__, lambda: fn(model, *a, **k))
        return out
    return inner


def _to_view(model, t, image=False):
    """NCHW tensor (or an NHWC View handed over natively) -> NHWC view."""
    if isinstance(t, View) or t is None:
        return t
    return model.image_view(t) if image else model.feature_view(t)


def _rows(model, name, idx):
    """CDF rows (host int32) about to drive a decode_stream call.  Test hook (tests/test_streams.py): when
    model._force_rows holds reference rows under `name`, the rows that differ are counted in model._row_flips and replaced —
    a scale within float noise of a row threshold otherwise desynchronises the decoder of a FOREIGN stream for good (the
    reference has the same property between its CPU and GPU runs).  Never active outside tests."""
    f = getattr(model, "_force_rows", None)
    if f and name in f:
        ref = np.ascontiguousarray(f[name], dtype=np.int32).reshape(-1)
        model._row_flips[name] = int((ref != idx).sum())
        return ref
    return idx


class _Reader:
    """Decoder-side glue: rANS decoder + device<->host hops of the symbol / index tensors."""

    def __init__(self, model, string):
        self.m = model
        self.dec = entropy.RansDecoder()
        self.dec.set_stream(string)

    def _view(self, sym, C, H, W, add=None):
        t = torch.from_numpy(np.ascontiguousarray(sym, dtype=np.int32)).to(self.m.device)
        out = self.m.new(H, W, C)
        ops.symbols_to_view(t, None if add is None else add.exact(), out.exact())
        return out

    def factorized(self, table, H, W):
        """BitEstimator.decode_stream(size) (video_entropy_models.py:232-241): z_hat = the decoded integers."""
        C = int(table.cdf.shape[0])
        return self._view(self.dec.decode_stream(_channel_index(C, H, W), table), C, H, W)

    def laplace(self, params, table, name):
        """GaussianEncoder.decode_stream(scales) + means (LSSVC_net_extend.py:120-123): params = (scales | means)."""
        C = params.real // 2
        scale, mean = params.slice(0, C), params.slice(C, 2 * C)
        idx = torch.empty(C * params.H * params.W, dtype=torch.int32, device=self.m.device)
        ops.scale_index(scale, idx, self.m._thr())
        sym = self.dec.decode_stream(_rows(self.m, name, idx.cpu().numpy()), table)
        return self._view(sym, C, params.H, params.W, add=mean)

    def four_part(self, common, table):
        """decompress_four_part_prior (LSSVC_net_extend.py:200-263)."""
        m = self.m
        C = common.real // 2
        y_hat = View.alloc(common.H, common.W, C, m.device, zero=True)
        n = (C // 4) * common.H * common.W
        prm = common
        for step in range(4):
            idx = torch.empty(n, dtype=torch.int32, device=m.device)
            ops.four_part_index(prm, step, idx, m._thr())
            rows = _rows(m, f"el_y{step}", idx.cpu().numpy())
            sym = torch.from_numpy(self.dec.decode_stream(rows, table)).to(m.device)
            ops.four_part_dec_step(sym, prm, step, y_hat)
            if step < 3:
                prm = m._spatial_prior(step + 1, y_hat, common)
        return y_hat


# ---- base layer -------------------------------------------------------------------------------------------------
@_range_guarded
def bl_compress(model, x, dpb):
    """DMCExtend.compress(x, dpb) -> {"string", "dpb": {ref_frame_bl, ref_feature_bl, y_hat_bl, mv_hat_bl}}."""
    model.update()
    t = model._tables
    dump = _Dump(model.device, owner=model)
    bl = model._base_layer(model.image_view(x), _to_view(model, dpb["ref_frame_bl"], image=True),
                           _to_view(model, dpb.get("ref_feature_bl")), _bits_scratch(model), dump)
    h = dump.host
    string = _encode([(h("bl_mv_z"), dump.channel_index("bl_mv_z"), t["bl_mv_z"]), (h("bl_mv_y"), h("bl_mv_y_idx"), t["laplace"]),
                      (h("bl_z"), dump.channel_index("bl_z"), t["bl_z"]), (h("bl_y"), h("bl_y_idx"), t["laplace"])])
    return {"string": string, "dpb": _bl_dpb(bl, clamp=False)}


@_range_guarded
def bl_decompress(model, string, height, width, dpb):
    """DMCExtend.decompress(string, height, width, dpb): the reconstruction is clamped to [0, 1] (dmc_net_extend.py:138)."""
    model.update()
    t = model._tables
    p = "base_layer_model."
    rd = _Reader(model, string)
    ref = _to_view(model, dpb["ref_frame_bl"], image=True)
    ref_feature = _to_view(model, dpb.get("ref_feature_bl"))
    zh, zw = stream.get_downsampled_shape(height, width, 64)
    mv_z_hat = rd.factorized(t["bl_mv_z"], zh, zw)
    mv_y_hat = rd.laplace(model._bl_mv_params(p, mv_z_hat), t["laplace"], "bl_mv_y")
    mv_hat = model._bl_mv_decode(p, mv_y_hat)
    c1, c2, c3 = model._bl_contexts(p, ref, ref_feature, mv_hat)
    z_hat = rd.factorized(t["bl_z"], zh, zw)
    y_hat = rd.laplace(model._bl_res_params(p, z_hat, c1, c2, c3), t["laplace"], "bl_y")
    rec_feat = model._res_decoder_gdn(p + "res_decoder", y_hat, c2, c3, intra=False)
    feature, recon = model._recon_generation(p + "recon_generation_net", rec_feat, c1)
    return {"dpb": _bl_dpb({"recon": recon, "feature": feature, "y_hat": y_hat, "mv_hat": mv_hat}, clamp=True)}


def _bl_dpb(bl, clamp):
    rec = bl["recon"].to_nchw()
    if clamp:
        rec.clamp_(0, 1)
    return {"ref_frame_bl": rec, "ref_feature_bl": bl["feature"].to_nchw(), "y_hat_bl": bl["y_hat"].to_nchw(),
            "mv_hat_bl": bl["mv_hat"].to_nchw()}


def bl_encode_decode_extend(model, x, dpb, output_path=None, pic_width=None, pic_height=None):
    """DMCExtend.encode_decode_extend (dmc_net_extend.py:149-173): compress, write, read back, decompress."""
    import time
    torch.cuda.synchronize(model.device)
    t0 = time.time()
    encoded = bl_compress(model, x, dpb)
    stream.encode_p(encoded["string"], output_path)
    bits = stream.filesize(output_path) * 8
    torch.cuda.synchronize(model.device)
    t1 = time.time()
    decoded = bl_decompress(model, stream.decode_p(output_path), pic_height, pic_width, dpb)
    torch.cuda.synchronize(model.device)
    t2 = time.time()
    return {"dpb": decoded["dpb"], "bit": bits, "encoding_time": t1 - t0, "decoding_time": t2 - t1}


# ---- enhancement layer ------------------------------------------------------------------------------------------
@_range_guarded
def el_compress(model, x, dpb):
    """LSSVC_extend.compress(x, dpb): dpb carries the decoded base layer as 'texture', 'y_hat_bl', 'mv_hat_bl'."""
    model.update()
    t = model._tables
    dump = _Dump(model.device, owner=model)
    el = model._el_layer(model.image_view(x), _to_view(model, dpb["ref_frame_el"], image=True),
                         _to_view(model, dpb.get("ref_feature_el")), _to_view(model, dpb["texture"]),
                         _to_view(model, dpb["y_hat_bl"]), _to_view(model, dpb["mv_hat_bl"]), _bits_scratch(model), dump)
    h = dump.host
    parts = [(h("el_mv_z"), dump.channel_index("el_mv_z"), t["el_mv_z"]), (h("el_mv_y"), h("el_mv_y_idx"), t["laplace"]),
             (h("el_z"), dump.channel_index("el_z"), t["el_z"])]
    parts += [(h(f"el_y{k}"), h(f"el_y{k}_idx"), t["laplace"]) for k in range(4)]
    return {"string": _encode(parts),
            "dpb": {"ref_frame_el": el["recon"].to_nchw(), "ref_feature_el": el["feature"].to_nchw(),
                    "warp_frame": el["warp_frame"].to_nchw(), "mv_hat": el["mv_hat"].to_nchw()}}


@_range_guarded
def el_decompress(model, string, height, width, dpb):
    """LSSVC_extend.decompress(string, height, width, dpb) -> {"dpb": {ref_frame_el, ref_feature_el}}."""
    model.update()
    t = model._tables
    rd = _Reader(model, string)
    ref = _to_view(model, dpb["ref_frame_el"], image=True)
    ref_feature = _to_view(model, dpb.get("ref_feature_el"))
    # (de-padded base-layer tensors: LSSVC_net_extend.py:97-99)
    mv_ctx_prior, mv_ctx = model._mv_contexts(model.depad(_to_view(model, dpb["mv_hat_bl"])))
    zh, zw = stream.get_downsampled_shape(height, width, 64)
    mv_z_hat = rd.factorized(t["el_mv_z"], zh, zw)
    mv_y_hat = rd.laplace(model._mv_params(mv_z_hat, mv_ctx_prior), t["laplace"], "el_mv_y")
    mv_hat = model._mv_decode(mv_y_hat, mv_ctx)
    c1, c2, c3, _ = model._hybrid_contexts(model.depad(_to_view(model, dpb["texture"])), mv_hat, ref, ref_feature)
    z_hat = rd.factorized(t["el_z"], zh, zw)
    params = model._res_params(z_hat, c3, model.depad(_to_view(model, dpb["y_hat_bl"]), 16))
    y_hat = rd.four_part(params, t["laplace"])
    feature, recon = model._res_decode(y_hat, c1, c2, c3)
    return {"dpb": {"ref_frame_el": recon.to_nchw(), "ref_feature_el": feature.to_nchw()}}


def encode_decode_extend(model, x_bl, x_el, dpb, output_path_bl=None, output_path_el=None, pic_width=None, pic_height=None,
                         pic_width_bl=None, pic_height_bl=None):
    """LSSVC_extend.encode_decode_extend (LSSVC_net_extend.py:144-191): every layer is compressed, written, read back and
    DECODED; the DPB of the next frame is the decoder's output, as in the reference."""
    import time
    bl = bl_encode_decode_extend(model, x_bl, dpb, output_path_bl, pic_width_bl, pic_height_bl)
    layer = bl["dpb"]
    dpb = dict(dpb)
    dpb["texture"], dpb["y_hat_bl"], dpb["mv_hat_bl"] = layer["ref_feature_bl"], layer["y_hat_bl"], layer["mv_hat_bl"]
    torch.cuda.synchronize(model.device)
    t0 = time.time()
    encoded = el_compress(model, x_el, dpb)
    stream.encode_p(encoded["string"], output_path_el)
    bits = stream.filesize(output_path_el) * 8
    torch.cuda.synchronize(model.device)
    t1 = time.time()
    decoded = el_decompress(model, stream.decode_p(output_path_el), pic_height, pic_width, dpb)
    torch.cuda.synchronize(model.device)
    t2 = time.time()
    out = {"ref_frame_bl": layer["ref_frame_bl"], "ref_feature_bl": layer["ref_feature_bl"],
           "ref_frame_el": decoded["dpb"]["ref_frame_el"], "ref_feature_el": decoded["dpb"]["ref_feature_el"]}
    return {"dpb": out, "bit_bl": bl["bit"], "bit_el": bits, "encoding_time_EL": t1 - t0, "decoding_time_EL": t2 - t1,
            "encoding_time_BL": bl["encoding_time"], "decoding_time_BL": bl["decoding_time"],
            "mv_hat": encoded["dpb"]["mv_hat"], "warp_frame": encoded["dpb"]["warp_frame"]}


# =================================================================================================================
# I-frames: IntraNoAR (base layer, priors.py:368-452) and IntraSS (IntraSS.py:239-336)
# =================================================================================================================
def _thr_img(model):
    return model.cached("thr_img", lambda: entropy.image_scale_thresholds().to(model.device))


def _eb_encode(model, z, prefix, table):
    """EntropyBottleneck.compress (img_entropy_models.py:286-313 with medians): z view -> (string, z_hat view)."""
    dump = _Dump(model.device)
    z_hat = model.new(z.H, z.W, z.real)
    ops.eb_quant(z, model._eb_coef(prefix), z_hat, None, sym=dump.buf("z", z))
    return _encode([(dump.host("z"), dump.channel_index("z"), table)]), z_hat


def _eb_decode(model, string, prefix, table, H, W):
    """EntropyBottleneck.decompress(strings, size): z_hat = decoded integers + medians."""
    C = int(table.cdf.shape[0])
    dec = entropy.RansDecoder()
    dec.set_stream(string)
    sym = dec.decode_stream(_channel_index(C, H, W), table).reshape(C, H * W)
    med = model._eb_coef(prefix)[:, 58].float().cpu().numpy()
    z = (sym.astype(np.float32) + med[:, None]).reshape(1, C, H, W)
    return model.feature_view(torch.from_numpy(z).to(model.device))


def _gaussian_encode(model, y, prm, table):
    """GaussianConditional.compress(y, build_indexes(scales), means) -> (string, y_hat view); prm = (scales | means)."""
    C = y.real
    dump = _Dump(model.device)
    y_hat = model.new(y.H, y.W, C)
    ops.gaussian_quant(y, prm.slice(C, 2 * C), prm.slice(0, C), y_hat, None, sym=dump.buf("y", y), index=dump.buf("y_idx", y),
                       thresholds=_thr_img(model))
    return _encode([(dump.host("y"), dump.host("y_idx"), table)]), y_hat


def _gaussian_decode(model, string, prm, table, name):
    C = prm.real // 2
    idx = torch.empty(C * prm.H * prm.W, dtype=torch.int32, device=model.device)
    ops.scale_index(prm.slice(0, C), idx, _thr_img(model))
    dec = entropy.RansDecoder()
    dec.set_stream(string)
    sym = torch.from_numpy(dec.decode_stream(_rows(model, name, idx.cpu().numpy()), table)).to(model.device)
    y_hat = model.new(prm.H, prm.W, C)
    ops.symbols_to_view(sym, prm.slice(C, 2 * C), y_hat.exact())
    return y_hat


_BL_EB = "base_layer_model.entropy_bottleneck."


@_range_guarded
def intra_bl_get_y_z(model, x):
    """IntraNoAR.get_y_z(x)."""
    y, z = model._bl_analysis(model.image_view(x))
    return y.to_nchw(), z.to_nchw()


@_range_guarded
def intra_bl_compress(model, x, y, z):
    """IntraNoAR.compress(x, y, z) (priors.py:420-435) -> {"strings": [[y], [z]], "shape"}."""
    model.update()
    t = model._tables
    z_string, z_hat = _eb_encode(model, _to_view(model, z), _BL_EB, t["bl_z"])
    y_string, _ = _gaussian_encode(model, _to_view(model, y), model._bl_params(z_hat), t["gaussian"])
    return {"strings": [[y_string], [z_string]], "shape": tuple(z.shape[-2:])}


@_range_guarded
def intra_bl_get_y_hat_recon(model, y, z):
    """IntraNoAR.get_y_hat_recon(y, z): the encoder-side reconstruction {x_hat, y_hat, z_hat}."""
    model.update()
    t = model._tables
    _, z_hat = _eb_encode(model, _to_view(model, z), _BL_EB, t["bl_z"])
    _, y_hat = _gaussian_encode(model, _to_view(model, y), model._bl_params(z_hat), t["gaussian"])
    return {"x_hat": model._bl_synthesis(y_hat).to_nchw(), "y_hat": y_hat.to_nchw(), "z_hat": z_hat.to_nchw()}


@_range_guarded
def intra_bl_decompress(model, strings, shape):
    """IntraNoAR.decompress(strings, shape) (priors.py:437-452) -> {"x_hat", "y_hat"}."""
    model.update()
    t = model._tables
    z_hat = _eb_decode(model, strings[1][0], _BL_EB, t["bl_z"], int(shape[0]), int(shape[1]))
    y_hat = _gaussian_decode(model, strings[0][0], model._bl_params(z_hat), t["gaussian"], "bl_y")
    return {"x_hat": model._bl_synthesis(y_hat).to_nchw(), "y_hat": y_hat.to_nchw()}


@_range_guarded
def intra_get_y_z_ctx(model, x_hat_bl, x_el):
    """IntraSS.get_y_z_ctx (IntraSS.py:239-243)."""
    c1, c2, c3 = model._context_mining(model.image_view(x_hat_bl))
    y, z = model._el_analysis(model.image_view(x_el), c1, c2, c3)
    return y.to_nchw(), z.to_nchw(), (c1.to_nchw(), c2.to_nchw(), c3.to_nchw())


@_range_guarded
def intra_compress(model, y=None, z=None, ctx3=None, y_hat_bl=None):
    """IntraSS.compress(y, z, ctx3, y_hat_bl) (IntraSS.py:304-314)."""
    model.update()
    t = model._tables
    z_string, z_hat = _eb_encode(model, _to_view(model, z), "entropy_bottleneck.", t["el_z"])
    prm = model._el_params(z_hat, _to_view(model, y_hat_bl), _to_view(model, ctx3))
    y_string, _ = _gaussian_encode(model, _to_view(model, y), prm, t["gaussian"])
    return {"strings": [[y_string], [z_string]], "shape": tuple(z.shape[-2:])}


@_range_guarded
def intra_decompress(model, strings, DPB_layer, shape):
    """IntraSS.decompress(strings, DPB_layer, shape) (IntraSS.py:316-336) -> {"x_hat", "feature"}."""
    model.update()
    t = model._tables
    c1, c2, c3 = model._context_mining(model.image_view(DPB_layer["x_hat_bl"]))
    z_hat = _eb_decode(model, strings[1][0], "entropy_bottleneck.", t["el_z"], int(shape[0]), int(shape[1]))
    prm = model._el_params(z_hat, _to_view(model, DPB_layer["y_hat_bl"]), c3)
    y_hat = _gaussian_decode(model, strings[0][0], prm, t["gaussian"], "el_y")
    res_hat = model._res_decoder_gdn("g_s", y_hat, c2, c3, intra=True)
    feature, x_hat = model._recon_generation("recon_net", res_hat, c1)
    return {"x_hat": x_hat.to_nchw(), "feature": feature.to_nchw()}


def intra_encode_decode(model, x_bl, x_el, bin_path_bl, bin_path_el, pic_height_bl, pic_width_bl, pic_height_el, pic_width_el):
    """IntraSS.encode_decode with bitstreams (IntraSS.py:245-302): encode BL and EL, then decode both from the files; the
    returned reconstructions are the DECODER's."""
    # ---- encode
    y_bl, z_bl = intra_bl_get_y_z(model, x_bl)
    comp = intra_bl_compress(model, None, y_bl, z_bl)
    stream.encode_i(pic_height_bl, pic_width_bl, comp["strings"][0][0], comp["strings"][1][0], bin_path_bl)
    bit_bl = stream.filesize(bin_path_bl) * 8
    enc = intra_bl_get_y_hat_recon(model, y_bl, z_bl)
    dp = lambda t, p=1: t if not any(int(a / p) for a in model.pad_size) else F.pad(t, [int(a / p) for a in model.pad_size])
    # (x_hat_bl_depadded / y_hat_bl_depadded: IntraSS.py:262-263, 285-286)
    y_el, z_el, ctx = intra_get_y_z_ctx(model, dp(enc["x_hat"]), x_el)
    comp = intra_compress(model, y=y_el, z=z_el, ctx3=ctx[2], y_hat_bl=dp(enc["y_hat"], 16))
    stream.encode_i(pic_height_el, pic_width_el, comp["strings"][0][0], comp["strings"][1][0], bin_path_el)
    bit_el = stream.filesize(bin_path_el) * 8
    # ---- decode
    h, w, y_string, z_string = stream.decode_i(bin_path_bl)
    dec_bl = intra_bl_decompress(model, [[y_string], [z_string]], stream.get_downsampled_shape(h, w, 64))
    h, w, y_string, z_string = stream.decode_i(bin_path_el)
    dec = intra_decompress(model, [[y_string], [z_string]], {"x_hat_bl": dp(dec_bl["x_hat"]), "y_hat_bl": dp(dec_bl["y_hat"], 16)},
                           stream.get_downsampled_shape(h, w, 64))
    return {"bit_bl": bit_bl, "bit_el": bit_el, "x_hat_bl": dec_bl["x_hat"], "x_hat_el": dec["x_hat"], "feature_el": dec["feature"]}
