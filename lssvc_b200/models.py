"""Host-side mirror of the reference model API over the sm_100a kernels.

  IntraSS        src/models/IntraSS.py:74        two-layer I-frame codec (BL IntraNoAR + EL contextual AE)
  LSSVC          src/models/LSSVC_net.py:12      two-layer P-frame codec (BL DMC + EL), estimate mode
  LSSVC_extend   src/models/LSSVC_net_extend.py:8  + real bitstreams (compress / decompress / update)

Same constructors, `from_state_dict` / `load_dict`, `set_scale_information`, `forward` / `encode_decode` signatures
and state_dict layout; tensors in and out are NCHW fp32 torch tensors (a batch > 1 is coded item by item, estimate mode
only, as in the reference).  Everything between the input conversion and the output conversion runs in the kernels of
liblssvc_b200.so; there is no PyTorch compute path.
"""
import math
import os

import torch

from . import _lib, entropy, ops
from .engine import Engine
from .nets import intra_ss_spec, lssvc_spec
from .ops import View
from .stream import decode_i, decode_p, encode_i, encode_p, filesize, get_downsampled_shape


# Engine of the I-frame hyper-decoders (h_s of both layers: eight convolutions at 1/64 .. 1/16 resolution, < 0.1 % of the
# frame's FLOPs).  Their output is the mean every latent is quantised against, so their rounding error turns directly into
# symbol flips: on the fp32 CUDA cores the 1080p I-frame differs from the oracle in 67-69 of 1.30 M symbols (what the
# all-fp32 engine gets: 64-68), on the tensor-core engine in 83-97 (tools/fullsize_parity.py, gpurun_out/fullsize4.log).
# LSSVC_PRIOR_ENGINE= (empty) selects the default engine for A/B.
PRIOR_ENGINE = os.environ.get("LSSVC_PRIOR_ENGINE", "simt") or None


def _strip_module_prefix(sd):
    return {(k[7:] if k[:7] == "module." else k): v for k, v in sd.items()}


def _dbg(model, **views):
    """Test hook: when model._debug is a dict, keep named intermediate views."""
    d = getattr(model, "_debug", None)
    if d is not None:
        d.update(views)


def _force(model, key, out, mean=None, mask=None, q_view=None):
    """Test hook (symbol teacher-forcing): when model._force holds reference symbols under `key`, count how many of
    the symbols just produced differ and overwrite `out` (= symbols + mean) with the reference symbols, so that
    everything downstream can be compared with the oracle even when a value sitting on a rounding boundary flipped.
    Host-side torch indexing on the strided NHWC views; never active outside tests."""
    f = getattr(model, "_force", None)
    if not f or key not in f:
        return
    ref = f[key].to(model.device)[0].permute(1, 2, 0)
    dst = out.exact().as_tensor()
    if q_view is not None:
        # progressive coding (four-part prior): out = symbols + means on `mask`, symbols kept in q_view
        q = q_view.exact().as_tensor()
        delta = torch.where(mask & (q != ref), ref - q, torch.zeros_like(q))
        model._force_flips[key] = model._force_flips.get(key, 0) + int((delta != 0).sum().item())
        dst.add_(delta)
        q.add_(delta)
        return
    m = mean.as_tensor() if mean is not None else None
    got = torch.round(dst - m) if m is not None else dst
    model._force_flips[key] = int((got != ref).sum().item())
    dst.copy_(ref + m if m is not None else ref)


def _merge_batch(results):
    """Per-item results of a batch -> one result: tensors concatenated along the batch dimension, bit counts and times summed
    (the reference sums its likelihoods over the whole batch: LSSVC_net.py:154-167, IntraSS.py:160), nested dicts merged likewise;
    the NHWC hand-over of the DPB ("_native") is per frame and is dropped."""
    out = {}
    for k, v in results[0].items():
        if k.startswith("_"):
            continue
        vs = [r[k] for r in results]
        if isinstance(v, torch.Tensor):
            out[k] = torch.cat(vs, dim=0)
        elif isinstance(v, dict):
            out[k] = _merge_batch(vs)
        elif isinstance(v, (int, float)):
            out[k] = sum(vs)
        else:
            out[k] = v
    return out


def _item(t, b):
    return None if t is None else t[b:b + 1]


class _Bits:
    """Device-side double accumulators for the per-layer bit counts of one frame."""

    def __init__(self, device, t=None):
        # [bits_bl, bits_el, range flag of the split-fp16 kernels (csrc/range.cu)]
        self.t = torch.zeros(3, dtype=torch.float64, device=device) if t is None else t
        self.out_of_range = False

    def ptr(self, layer):
        return self.t[layer:layer + 1]

    def fetch_range(self):
        """Stream-ordered: moves (and clears) the library's range flag into slot 2; call after the frame's last kernel."""
        ops.range_flag_fetch(self.t[2:3])

    def read(self):
        bl, el, flag = self.t.tolist()   # the one device->host sync of a frame
        self.out_of_range = flag != 0.0
        return bl, el


def _recode_fp32(model, what, call):
    """An operand of the split-fp16 tensor-core kernels left the fp16 range while `what` was coded (its outputs are NaN):
    code it again on the fp32 CUDA-core engine.  Counted in ops.RANGE_FALLBACKS; warned about once per model."""
    if ops.default_engine() != "h2":
        raise _lib.LssvcError(f"{what}: range flag raised although the split-fp16 engine is not in use")
    ops.RANGE_FALLBACKS += 1
    if not getattr(model, "_range_warned", False):
        import warnings
        warnings.warn(f"lssvc_b200: {what}: an activation reached the fp16 limit of the split-fp16 tensor-core engine "
                      "(|x| >= 65520, or a GDN input beyond 255.9); the frame is re-coded on the fp32 CUDA-core engine "
                      "(about 10x slower). Set LSSVC_CONV_ENGINE=simt to keep this checkpoint off the split-fp16 engine altogether.")
        model._range_warned = True
    prev = ops.set_engine("simt")
    try:
        return call()
    finally:
        ops.set_engine(prev)


# =============================================================================================================
# I-frame
# =============================================================================================================
class IntraSS(Engine):
    """IntraSS(channel_BL=192, channel_N=64, channel_M=96, base_layer_model_path=None)   IntraSS.py:74-117"""

    def __init__(self, channel_BL=192, channel_N=64, channel_M=96, base_layer_model_path=None, **kwargs):
        super().__init__(intra_ss_spec(channel_BL, channel_N, channel_M), "I", seed=kwargs.pop("seed", None))
        self.N, self.M, self.channel_BL = int(channel_N), int(channel_M), int(channel_BL)
        self._tables = None
        # the reference's `model.base_layer_model` is an IntraNoAR with get_y_z / compress / decompress / get_y_hat_recon
        # (priors.py:390-452): the parameter container of the same name gets those entry points
        api = _IntraBaseLayerAPI(self)
        for name in ("get_y_z", "compress", "decompress", "get_y_hat_recon", "update"):
            object.__setattr__(self.base_layer_model, name, getattr(api, name))
        if base_layer_model_path is not None:
            self.load_bl_pretrain(base_layer_model_path)

    # ---- checkpoint conventions (IntraSS.py:174-220) ------------------------------------------------------------
    def load_state_dict(self, state_dict, strict=True):
        sd = dict(state_dict)
        sd.pop("gaussian_conditional.scale_table", None)
        # CDF buffers are sized dynamically in the reference (update_registered_buffers): accept any size
        for k in list(sd.keys()):
            if k.endswith(("._quantized_cdf", "._offset", "._cdf_length")):
                mod_name, buf = k.rsplit(".", 1)
                mod = self
                for part in mod_name.split("."):
                    mod = mod._modules[part]
                mod._buffers[buf] = torch.zeros(sd[k].shape, dtype=torch.int32, device=mod._buffers[buf].device)
        return super().load_state_dict(sd, strict=strict)

    @classmethod
    def from_state_dict(cls, state_dict, base_layer_model_path=None):
        sd = _strip_module_prefix(state_dict)
        if base_layer_model_path is not None:
            ck = torch.load(base_layer_model_path, map_location=torch.device("cpu"))
            ck = ck.get("state_dict", ck)
            for k in ck:
                sd["base_layer_model." + k] = ck[k]
        n_bl = sd["base_layer_model.g_s.0.conv1.weight"].size(0)
        net = cls(n_bl)
        net.load_state_dict(sd)
        return net

    def load_bl_pretrain(self, path):
        ck = torch.load(path, map_location=torch.device("cpu"))
        ck = ck.get("state_dict", ck)
        sd = self.state_dict()
        for k, v in ck.items():
            sd["base_layer_model." + k] = v
        self.load_state_dict(sd)

    # ---- building blocks ----------------------------------------------------------------------------------------
    def _rb(self, name, x):
        # ResidualBlock (layers.py:122-145): lrelu(conv2(lrelu(conv1 x))) + x
        t = self.conv(name + ".conv1", x, act=0.01)
        return self.conv(name + ".conv2", t, act=0.01, res1=x)

    def _rb_stride(self, name, x):
        # ResidualBlockWithStride (layers.py:60-91): gdn(conv2(lrelu(conv1 s2 x))) + conv1x1 s2 (x)
        t = self.conv(name + ".conv1", x, stride=2, act=0.01)
        t = self.conv(name + ".conv2", t)
        idn = self.conv(name + ".downsample", x, stride=2, pad=0)
        return self.gdn(name + ".gdn", t, intra=True, res1=idn)

    def _rb_up(self, name, x):
        # ResidualBlockUpsample (layers.py:94-119): igdn(conv(lrelu(subpel x))) + subpel'(x)
        t = self.conv(name + ".subpel_conv.0", x, ps=True, act=0.01)
        t = self.conv(name + ".conv", t)
        idn = self.conv(name + ".upsample.0", x, ps=True)
        return self.gdn(name + ".igdn", t, inverse=True, intra=True, res1=idn)

    def _eb_coef(self, prefix):
        def build():
            g = lambda n: self.tensor(prefix + n)
            return entropy.eb_coef([g(f"_matrices.{i}") for i in range(5)], [g(f"_biases.{i}") for i in range(5)],
                                   [g(f"_factors.{i}") for i in range(4)], g("quantiles")).to(self.device)
        return self.cached(("eb", prefix), build)

    def _bl_analysis(self, x):
        p = "base_layer_model."
        y = x
        for i in range(6):
            y = self._rb_stride(f"{p}g_a.{i}", y) if i % 2 == 0 else self._rb(f"{p}g_a.{i}", y)
        y = self.conv(p + "g_a.6", y, stride=2)
        z = self.conv(p + "h_a.0", y, act=0.01)
        z = self.conv(p + "h_a.2", z, act=0.01)
        z = self.conv(p + "h_a.4", z, stride=2, act=0.01)
        z = self.conv(p + "h_a.6", z, act=0.01)
        z = self.conv(p + "h_a.8", z, stride=2)
        return y, z

    def _bl_params(self, z_hat):
        p = "base_layer_model."
        e = PRIOR_ENGINE
        g = self.conv(p + "h_s.0", z_hat, act=0.01, exact_in=True, engine=e)
        g = self.conv(p + "h_s.2.0", g, ps=True, act=0.01, engine=e)
        g = self.conv(p + "h_s.4", g, act=0.01, engine=e)
        g = self.conv(p + "h_s.6.0", g, ps=True, act=0.01, engine=e)
        return self.conv(p + "h_s.8", g, engine=e)

    def _bl_synthesis(self, y_hat):
        p = "base_layer_model."
        x = y_hat
        for i in range(7):
            x = self._rb(f"{p}g_s.{i}", x) if i % 2 == 0 else self._rb_up(f"{p}g_s.{i}", x)
        return self.conv(p + "g_s.7.0", x, ps=True)

    def _context_mining(self, x_hat_bl):
        """multi_scale_context_mining (IntraSS.py:119-122)."""
        H, W = self.shape_hr
        t = self.seq2("texture_resampler.conv_adaptor", x_hat_bl)
        t = self.resize(t, H, W)
        t1, t2, t3 = self.extractor3("texture_extractor", t)
        return self.fusion3("context_fusion_net", t1, t2, t3)

    def _el_analysis(self, x_el, c1, c2, c3):
        """ResEncoder (layers.py:342-367) + h_a (IntraSS.py:87-93)."""
        y = self._res_encoder_gdn("g_a", x_el, c1, c2, c3, intra=True)
        z = self.conv("h_a.0", y, act=0.01)
        z = self.conv("h_a.2", z, stride=2, act=0.01)
        z = self.conv("h_a.4", z, stride=2)
        return y, z

    def _res_encoder_gdn(self, name, x, c1, c2, c3, intra):
        f = self.conv(name + ".conv1", [x, c1], stride=2)
        cat = View.alloc(f.H, f.W, f.real + c2.real, self.device)
        self.gdn(name + ".gdn1", f, intra=intra, out=cat.slice(0, f.real))
        self.copy(c2, cat.slice(f.real, cat.C))
        f = self.res_block(name + ".res1", cat, slope=0.1, start_from_relu=False, end_with_relu=True)
        f = self.conv(name + ".conv2", f, stride=2)
        cat = View.alloc(f.H, f.W, f.real + c3.real, self.device)
        self.gdn(name + ".gdn2", f, intra=intra, out=cat.slice(0, f.real))
        self.copy(c3, cat.slice(f.real, cat.C))
        f = self.res_block(name + ".res2", cat, slope=0.1, start_from_relu=False, end_with_relu=True)
        f = self.conv(name + ".conv3", f, stride=2)
        f = self.gdn(name + ".gdn3", f, intra=intra)
        return self.conv(name + ".conv4", f, stride=2)

    def _res_decoder_gdn(self, name, y_hat, c2, c3, intra):
        """ResDecoder (layers.py:370-395, dmc_net.py:93-118)."""
        f = self.conv(name + ".up1.0", y_hat, ps=True)
        f = self.gdn(name + ".gdn1", f, inverse=True, intra=intra)
        f = self.conv(name + ".up2.0", f, ps=True)
        cat = View.alloc(f.H, f.W, f.real + c3.real, self.device)
        self.gdn(name + ".gdn2", f, inverse=True, intra=intra, out=cat.slice(0, f.real))
        self.copy(c3, cat.slice(f.real, cat.C))
        f = self.res_block(name + ".res1", cat, slope=0.1, start_from_relu=False, end_with_relu=True)
        f = self.conv(name + ".up3.0", f, ps=True)
        cat = View.alloc(f.H, f.W, f.real + c2.real, self.device)
        self.gdn(name + ".gdn3", f, inverse=True, intra=intra, out=cat.slice(0, f.real))
        self.copy(c2, cat.slice(f.real, cat.C))
        f = self.res_block(name + ".res2", cat, slope=0.1, start_from_relu=False, end_with_relu=True)
        return self.conv(name + ".up4.0", f, ps=True)

    def _recon_generation(self, name, first, second):
        """ReconGeneration called as (res_feature, context1): cat order (first, second)   layers.py:398-411."""
        t, ta = self.conv(name + ".feature_conv.0", [first, second], act_copy=0.01)
        t, ta = self.res_block(name + ".feature_conv.1", t, x_act=ta, act_copy=0.01)
        feature = self.res_block(name + ".feature_conv.2", t, x_act=ta)
        return feature, self.conv(name + ".recon_conv", feature)

    def _el_params(self, z_hat, y_hat_bl, c3):
        H, W = self.shape_hr
        e = PRIOR_ENGINE
        hyper = self.conv("h_s.0.0", z_hat, ps=True, act=0.01, exact_in=True, engine=e)
        hyper = self.conv("h_s.2.0", hyper, ps=True, act=0.01, engine=e)
        hyper = self.conv("h_s.4", hyper, engine=e)
        lp = self.seq2("layer_prior_resampler.conv_adaptor", y_hat_bl)
        lp = self.resize(lp, H // 16, W // 16)
        cp = self.conv("prior_fusion_net.context_parameters.0", c3, stride=2, act=0.1)
        cp = self.conv("prior_fusion_net.context_parameters.2", cp, stride=2)
        prm = self.conv("prior_fusion_net.params_net.0", [hyper, lp, cp], act=0.01)
        prm = self.conv("prior_fusion_net.params_net.2", prm, act=0.01)
        return self.conv("prior_fusion_net.params_net.4", prm)

    # ---- forward ------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, x_bl, x_el, train_with_recon=False, _native=False, _write=None, _async=False):
        """IntraSS.forward (IntraSS.py:137-172).  _async (runner.py): no host synchronisation — the bit counts stay on the
        device (result["_bits"]) and "bit_bl" / "bit_el" are None until the caller reads them."""
        self._require_cuda()
        if x_bl.dim() == 4 and x_bl.shape[0] > 1:
            # batch > 1 (estimate mode only, as in the reference): the items are independent frames, coded one after the other
            if _write is not None or _async:
                raise ValueError("IntraSS: bitstream / asynchronous coding takes one frame at a time (batch 1)")
            return _merge_batch([self.forward(x_bl[b:b + 1], x_el[b:b + 1], train_with_recon, _native)
                                 for b in range(x_bl.shape[0])])
        bits = _Bits(self.device)
        xb, xe = self.image_view(x_bl), self.image_view(x_el)
        # ---- base layer: IntraNoAR.get_layer_information (priors.py:368-388)
        y_bl, z_bl = self._bl_analysis(xb)
        z_hat_bl = self.new(z_bl.H, z_bl.W, z_bl.real)
        w = _write
        ops.eb_quant(z_bl, self._eb_coef("base_layer_model.entropy_bottleneck."), z_hat_bl, bits.ptr(0),
                     sym=w.buf("bl_z", z_bl) if w else None)
        _force(self, "bl_z_hat", z_hat_bl)
        prm = self._bl_params(z_hat_bl)
        C = y_bl.real
        y_hat_bl = self.new(y_bl.H, y_bl.W, C)
        thr = self.cached("thr_img", lambda: entropy.image_scale_thresholds().to(self.device))
        ops.gaussian_quant(y_bl, prm.slice(C, 2 * C), prm.slice(0, C), y_hat_bl, bits.ptr(0),
                           sym=w.buf("bl_y", y_bl) if w else None, index=w.buf("bl_y_idx", y_bl) if w else None,
                           thresholds=thr if w else None)
        _force(self, "bl_y_q", y_hat_bl, mean=prm.slice(C, 2 * C))
        _dbg(self, params_bl=prm)
        x_hat_bl = self._bl_synthesis(y_hat_bl)
        # ---- enhancement layer (the base-layer tensors it reads are de-padded first: IntraSS.py:145-147)
        c1, c2, c3 = self._context_mining(self.depad(x_hat_bl))
        y, z = self._el_analysis(xe, c1, c2, c3)
        z_hat = self.new(z.H, z.W, z.real)
        ops.eb_quant(z, self._eb_coef("entropy_bottleneck."), z_hat, bits.ptr(1), sym=w.buf("el_z", z) if w else None)
        _force(self, "z_hat", z_hat)
        prm = self._el_params(z_hat, self.depad(y_hat_bl, 16), c3)
        C = y.real
        y_hat = self.new(y.H, y.W, C)
        ops.gaussian_quant(y, prm.slice(C, 2 * C), prm.slice(0, C), y_hat, bits.ptr(1),
                           sym=w.buf("el_y", y) if w else None, index=w.buf("el_y_idx", y) if w else None,
                           thresholds=thr if w else None)
        _force(self, "y_q", y_hat, mean=prm.slice(C, 2 * C))
        res_hat = self._res_decoder_gdn("g_s", y_hat, c2, c3, intra=True)
        feature, x_hat = self._recon_generation("recon_net", res_hat, c1)
        _dbg(self, y_bl=y_bl, y_hat_bl=y_hat_bl, z_hat_bl=z_hat_bl, y=y, y_hat=y_hat, z_hat=z_hat, params_el=prm,
             c1=c1, c2=c2, c3=c3)
        bits.fetch_range()
        bit_bl, bit_el = (None, None) if _async else bits.read()
        if bits.out_of_range:
            return _recode_fp32(self, "IntraSS.forward", lambda: self.forward(x_bl, x_el, train_with_recon, _native, _write))
        result = {"bit_bl": bit_bl, "bit_el": bit_el, "x_hat_bl": x_hat_bl.to_nchw(), "x_hat_el": x_hat.to_nchw(),
                  "feature_el": feature.to_nchw_shared()}
        if _async:
            result["_bits"] = bits
        if _native:
            result["_native"] = {"x_hat_bl": x_hat_bl, "x_hat_el": x_hat, "feature_el": feature, "y_hat_bl": y_hat_bl,
                                 "y": y, "z_hat": z_hat, "params_el": prm, "y_bl": y_bl, "ctx": (c1, c2, c3)}
        return result

    def encode_decode(self, x_bl, x_el, bin_path_bl, bin_path_el, pic_height_bl, pic_width_bl, pic_height_el,
                      pic_width_el):
        """IntraSS.encode_decode (IntraSS.py:245-302)."""
        if bin_path_bl is None:
            return self.forward(x_bl, x_el)
        self._require_cuda()
        if self.single_pass_streams:
            from .streams import intra_encode_decode
            return intra_encode_decode(self, x_bl, x_el, bin_path_bl, bin_path_el, pic_height_bl, pic_width_bl,
                                       pic_height_el, pic_width_el)
        from . import codec
        with torch.no_grad():
            return codec.intra_encode_decode(self, x_bl, x_el, bin_path_bl, bin_path_el, pic_height_bl, pic_width_bl,
                                             pic_height_el, pic_width_el)

    # ---- stream-mode entry points of the reference (IntraSS.py:239-243, 304-336); real coder + decoder in codec.py ----
    single_pass_streams = False     # True: one forward pass + stream verification (streams.py) instead of encode + decode

    @torch.no_grad()
    def get_y_z_ctx(self, x_bl, x_el):
        from . import codec
        self._require_cuda()
        return codec.intra_get_y_z_ctx(self, x_bl, x_el)

    @torch.no_grad()
    def compress(self, y=None, z=None, ctx3=None, y_hat_bl=None):
        from . import codec
        self._require_cuda()
        return codec.intra_compress(self, y=y, z=z, ctx3=ctx3, y_hat_bl=y_hat_bl)

    @torch.no_grad()
    def decompress(self, strings, DPB_layer, shape):
        from . import codec
        self._require_cuda()
        return codec.intra_decompress(self, strings, DPB_layer, shape)

    def update(self, force=False):
        """Build the CDF tables of both layers (IntraSS.py:234-237)."""
        if self._tables is not None and not force:
            return
        g = lambda p, n: self.tensor(p + n)
        t = {"gaussian": entropy.gaussian_table()}
        for tag, p in (("bl_z", "base_layer_model.entropy_bottleneck."), ("el_z", "entropy_bottleneck.")):
            t[tag] = entropy.eb_table([g(p, f"_matrices.{i}") for i in range(5)], [g(p, f"_biases.{i}") for i in range(5)],
                                      [g(p, f"_factors.{i}") for i in range(4)], g(p, "quantiles"))
        self._tables = t


class _IntraBaseLayerAPI:
    """IntraNoAR's stream-mode calls (priors.py:390-452) bound to an IntraSS model."""

    def __init__(self, model):
        self._m = model

    def update(self, force=False):
        self._m.update(force=force)

    @torch.no_grad()
    def get_y_z(self, x):
        from . import codec
        self._m._require_cuda()
        return codec.intra_bl_get_y_z(self._m, x)

    @torch.no_grad()
    def compress(self, x, y, z):
        from . import codec
        self._m._require_cuda()
        return codec.intra_bl_compress(self._m, x, y, z)

    @torch.no_grad()
    def decompress(self, strings, shape):
        from . import codec
        self._m._require_cuda()
        return codec.intra_bl_decompress(self._m, strings, shape)

    @torch.no_grad()
    def get_y_hat_recon(self, y, z):
        from . import codec
        self._m._require_cuda()
        return codec.intra_bl_get_y_hat_recon(self._m, y, z)


# =============================================================================================================
# P-frame
# =============================================================================================================
class LSSVC(Engine):
    """LSSVC(bl_model_path=None, mv_pretrain_path=None, win_size=11)   LSSVC_net.py:12-139"""

    def __init__(self, bl_model_path=None, mv_pretrain_path=None, win_size=11, seed=None):
        super().__init__(lssvc_spec(), "P", seed=seed)
        self.version = "Final"
        self.channel_N, self.channel_mv = 64, 64
        self._tables = None

    def load_dict(self, pretrained_dict, strict=True):
        """LSSVC.load_dict (LSSVC_net.py:141-149)."""
        self.load_state_dict(_strip_module_prefix(pretrained_dict), strict=strict)

    # shared with IntraSS (same reference blocks with the inter-frame GDN)
    _res_encoder_gdn = IntraSS._res_encoder_gdn
    _res_decoder_gdn = IntraSS._res_decoder_gdn
    _recon_generation = IntraSS._recon_generation

    def _bitparm_coef(self, prefix):
        def build():
            g = lambda n: self.tensor(prefix + n)
            return entropy.bitparm_coef([g(f"f{i}.h") for i in (1, 2, 3, 4)], [g(f"f{i}.b") for i in (1, 2, 3, 4)],
                                        [g(f"f{i}.a") for i in (1, 2, 3)]).to(self.device)
        return self.cached(("bitparm", prefix), build)

    def _thr(self):
        return self.cached("thr_vid", lambda: entropy.video_scale_thresholds().to(self.device))

    @staticmethod
    def _laplace(y, y_hat, bits, w, key, thr):
        """Laplace entropy epilogue (ops.conv `entropy`) for latent y coded against the (scale | mean) the convolution emits."""
        return {"mode": "laplace", "y": y.exact(), "y_hat": y_hat.exact(), "bits": bits, "sym": w.buf(key, y) if w else None,
                "index": w.buf(key + "_idx", y) if w else None, "thresholds": thr}

    def _prior_encoder(self, name, y, coef, bits, w, key):
        """Hyper-encoder + the factorised quantiser of z (BitEstimator prior): the last convolution rounds z, counts its bits and
        dumps its symbols in its own epilogue (LSSVC_EPI_BITPARM; stand-alone lssvc_bitparm_quant when it cannot).  Returns z_hat."""
        t = self.conv(name + ".0", y, act=0.01)
        t = self.conv(name + ".2", t, stride=2, act=0.01)
        wt = self.tensor(name + ".4.weight")
        k = wt.shape[-1]
        z_hat = self.new((t.H + 2 * (k // 2) - k) // 2 + 1, (t.W + 2 * (k // 2) - k) // 2 + 1, wt.shape[0])
        ent = {"mode": "bitparm", "coef": coef, "bits": bits, "sym": w.buf(key, z_hat) if w else None}
        return self.conv(name + ".4", t, stride=2, out=z_hat, entropy=ent)

    # ---- base layer: DMC.get_inter_layer_information (dmc_net.py:421-488) ------------------------------------------
    def _bl_motion(self, p, xb, ref):
        est_mv = self.spynet(p + "optic_flow", xb, ref)
        t = est_mv
        for base in (0, 4, 8):
            t = self.conv(f"{p}mv_encoder.{base}", t, stride=2)
            t = self.gdn(f"{p}mv_encoder.{base + 1}", t)
            # ResBlock(start_from_relu=False) then LeakyReLU(0.1): take the activated copy
            _, t = self.res_block(f"{p}mv_encoder.{base + 2}", t, start_from_relu=False, act_copy=0.1)
        return self.conv(p + "mv_encoder.12", t, stride=2)

    def _bl_mv_params(self, p, mv_z_hat, entropy=None):
        """entropy: Laplace epilogue of the convolution that emits (scale | mean) (encoder side; ops.conv)."""
        t = self.deconv_s2(p + "mv_prior_decoder.0", mv_z_hat, act=0.01, exact_in=True)
        t = self.deconv_s2(p + "mv_prior_decoder.2", t, act=0.01)
        return self.conv(p + "mv_prior_decoder.4", t, transposed=True, entropy=entropy)

    def _bl_mv_decode(self, p, mv_y_hat):
        t = self.deconv_s2(p + "mv_decoder.0", mv_y_hat, act=0.1)
        t = self.res_block(p + "mv_decoder.2", t, start_from_relu=False)
        t = self.gdn(p + "mv_decoder.3", t, inverse=True)
        t = self.gdn(p + "mv_decoder.5", self.deconv_s2(p + "mv_decoder.4", t), inverse=True)
        t = self.gdn(p + "mv_decoder.7", self.deconv_s2(p + "mv_decoder.6", t), inverse=True)
        return self.deconv_s2(p + "mv_decoder.8", t)

    def _bl_contexts(self, p, ref, ref_feature, mv_hat):
        """DMC.motion_compensation (dmc_net.py:359-368)."""
        mv2 = self.resize(mv_hat, mv_hat.H // 2, mv_hat.W // 2, scale=0.5)
        mv3 = self.resize(mv2, mv2.H // 2, mv2.W // 2, scale=0.5)
        if ref_feature is None:
            feat = self.conv(p + "feature_adaptor_I", ref)
        else:
            feat = self.conv(p + "feature_adaptor_P", ref_feature, pad=0)
        f1, f2, f3 = self.extractor3(p + "feature_extractor", feat)
        c1, c2, c3 = self.warp(f1, mv_hat), self.warp(f2, mv2), self.warp(f3, mv3)
        return self.fusion3(p + "context_fusion_net", c1, c2, c3)

    def _bl_res_params(self, p, z_hat, c1, c2, c3, entropy=None):
        t = self.deconv_s2(p + "res_prior_decoder.0", z_hat, act=0.01, exact_in=True)
        t = self.deconv_s2(p + "res_prior_decoder.2", t, act=0.01)
        hier = self.conv(p + "res_prior_decoder.4", t, transposed=True)
        tp = p + "temporal_prior_encoder"
        t = self.gdn(tp + ".gdn1", self.conv(tp + ".conv1", c1, stride=2))
        t = self.gdn(tp + ".gdn2", self.conv(tp + ".conv2", [t, c2], stride=2))
        t = self.gdn(tp + ".gdn3", self.conv(tp + ".conv3", [t, c3], stride=2))
        temporal = self.conv(tp + ".conv4", t, stride=2)
        g = self.conv(p + "res_entropy_parameter.0", [temporal, hier], act=0.01)
        g = self.conv(p + "res_entropy_parameter.2", g, act=0.01)
        return self.conv(p + "res_entropy_parameter.4", g, entropy=entropy)

    def _base_layer(self, xb, ref_frame, ref_feature, bits, w=None):
        p = "base_layer_model."
        thr = self._thr() if w else None
        mv_y = self._bl_motion(p, xb, ref_frame)
        mv_z_hat = self._prior_encoder(p + "mv_prior_encoder", mv_y, self._bitparm_coef(p + "bit_estimator_z_mv."), bits.ptr(0),
                                       w, "bl_mv_z")
        _force(self, "bl_mv_z_hat", mv_z_hat)
        C = mv_y.real
        mv_y_hat = self.new(mv_y.H, mv_y.W, C)
        mv_prm = self._bl_mv_params(p, mv_z_hat, entropy=self._laplace(mv_y, mv_y_hat, bits.ptr(0), w, "bl_mv_y", thr))
        _force(self, "bl_mv_y_q", mv_y_hat, mean=mv_prm.slice(C, 2 * C))
        mv_hat = self._bl_mv_decode(p, mv_y_hat)
        c1, c2, c3 = self._bl_contexts(p, ref_frame, ref_feature, mv_hat)
        y = self._res_encoder_gdn(p + "res_encoder", xb, c1, c2, c3, intra=False)
        z_hat = self._prior_encoder(p + "res_prior_encoder", y, self._bitparm_coef(p + "bit_estimator_z."), bits.ptr(0), w, "bl_z")
        _force(self, "bl_z_hat", z_hat)
        C = y.real
        y_hat = self.new(y.H, y.W, C)
        # (2 x 96 parameter channels = two channel tiles of 96: scale / mean interleaved per tile at pack time, ops.PackedConv)
        prm = self._bl_res_params(p, z_hat, c1, c2, c3, entropy=self._laplace(y, y_hat, bits.ptr(0), w, "bl_y", thr))
        _force(self, "bl_y_q", y_hat, mean=prm.slice(C, 2 * C))
        if w:
            w.layer_done("bl")      # every symbol of the BL string is on its way to the host while the synthesis runs
        rec_feat = self._res_decoder_gdn(p + "res_decoder", y_hat, c2, c3, intra=False)
        feature, recon = self._recon_generation(p + "recon_generation_net", rec_feat, c1)
        _dbg(self, bl_mv_y=mv_y, bl_mv_y_hat=mv_y_hat, bl_mv_prm=mv_prm, bl_mv_z_hat=mv_z_hat, bl_y=y, bl_y_hat=y_hat,
             bl_prm=prm, bl_z_hat=z_hat, bl_mv_hat=mv_hat, bl_c1=c1)
        return {"recon": recon, "feature": feature, "y_hat": y_hat, "mv_hat": mv_hat}

    # ---- enhancement layer -----------------------------------------------------------------------------------------
    def _resampler_tail(self, name, up):
        """conv2 (conv, lrelu, conv) -> 2 DepthConvBlocks, + skip   lssvc_modules.py:361-363, 394-396, 426-428"""
        u = self.seq2(name + ".conv2", up)
        r = self.depth_conv_block(name + ".feature_refine.0", u)
        return self.depth_conv_block(name + ".feature_refine.1", r, res2=u)

    def _mv_resampler(self, mv_bl):
        """MvResampler (lssvc_modules.py:339-365)."""
        H, W = self.shape_hr
        f = self.seq2("mv_resampler.conv1", mv_bl)
        f = self._resampler_tail("mv_resampler", self.resize(f, H, W))
        return self.conv("mv_resampler.recon_conv", f, out_scale=float(self.scale_factor))

    def _texture_resampler(self, texture_bl):
        """TextureResampler (lssvc_modules.py:368-397)."""
        H, W = self.shape_hr
        key = "base_layer_adaptor" if texture_bl.real == 64 else "enhance_layer_adaptor"
        f = self.conv(f"texture_resampler.conv_adaptor.{key}", texture_bl)
        f = self.seq2("texture_resampler.conv1", f)
        return self._resampler_tail("texture_resampler", self.resize(f, H, W))

    def _layer_prior_resampler(self, y_hat_bl):
        """LayerPriorResampler (lssvc_modules.py:400-429)."""
        H, W = self.shape_hr
        key = "base_layer_adaptor" if y_hat_bl.real == 96 else "enhance_layer_adaptor"
        f = self.conv(f"layer_prior_resampler.conv_adaptor.{key}", y_hat_bl)
        f = self.seq2("layer_prior_resampler.conv1", f)
        return self._resampler_tail("layer_prior_resampler", self.resize(f, H // 16, W // 16))

    def _mv_contexts(self, mv_bl_hat):
        mv_up = self._mv_resampler(mv_bl_hat)
        t = mv_up
        for i in (0, 2, 4):    # mv_ctx_prior_encoder (LSSVC_net.py:108-116)
            t = self.gdn(f"mv_ctx_prior_encoder.{i + 1}", self.conv(f"mv_ctx_prior_encoder.{i}", t, stride=2))
        mv_ctx_prior = self.conv("mv_ctx_prior_encoder.6", t, stride=2)
        t, ta = self.conv("mv_ctx_transform.transform.0", mv_up, stride=2, act_copy=0.01)   # MVContextTransformer
        mv_ctx = self.res_block("mv_ctx_transform.transform.1", t, x_act=ta)
        return mv_ctx_prior, mv_ctx

    def _mv_encode(self, mv, mv_ctx):
        """MVResEncoder (lssvc_modules.py:445-469) + mv_prior_encoder."""
        t = self.gdn("mv_encoder.encoder1.1", self.conv("mv_encoder.encoder1.0", mv, stride=2))
        _, t = self.res_block("mv_encoder.encoder1.2", t, start_from_relu=False, act_copy=0.1)
        t = self.gdn("mv_encoder.encoder2.1", self.conv("mv_encoder.encoder2.0", [t, mv_ctx], stride=2))
        _, t = self.res_block("mv_encoder.encoder2.2", t, start_from_relu=False, act_copy=0.1)
        t = self.gdn("mv_encoder.encoder2.5", self.conv("mv_encoder.encoder2.4", t, stride=2))
        _, t = self.res_block("mv_encoder.encoder2.6", t, start_from_relu=False, act_copy=0.1)
        return self.conv("mv_encoder.encoder2.8", t, stride=2)

    def _mv_params(self, mv_z_hat, mv_ctx_prior, entropy=None):
        h = self.conv("mv_prior_decoder.0.0", mv_z_hat, ps=True, act=0.01, exact_in=True)
        h = self.conv("mv_prior_decoder.2.0", h, ps=True, act=0.01)
        h = self.conv("mv_prior_decoder.4", h)
        g = self.conv("mv_prior_fusion.0", [h, mv_ctx_prior], act=0.01)
        g = self.conv("mv_prior_fusion.2", g, act=0.01)
        return self.conv("mv_prior_fusion.4", g, entropy=entropy)

    def _mv_decode(self, mv_y_hat, mv_ctx):
        """MVResDecoder (lssvc_modules.py:472-494)."""
        t = self.conv("mv_decoder.decoder1.0.0", mv_y_hat, ps=True, act=0.1)
        t = self.res_block("mv_decoder.decoder1.2", t, start_from_relu=False)
        t = self.gdn("mv_decoder.decoder1.3", t, inverse=True)
        t = self.gdn("mv_decoder.decoder1.5", self.conv("mv_decoder.decoder1.4.0", t, ps=True), inverse=True)
        t = self.gdn("mv_decoder.decoder1.7", self.conv("mv_decoder.decoder1.6.0", t, ps=True), inverse=True)
        t = self.conv("mv_decoder.decoder2.0", [t, mv_ctx], act=0.1)
        return self.conv("mv_decoder.decoder2.2.0", t, ps=True)

    def _temporal_contexts(self, ref, ref_feature, mv_hat):
        """LSSVC.motion_compensation (LSSVC_net.py:229-244)."""
        warp_frame = self.warp(ref, mv_hat)
        mv2 = self.resize(mv_hat, mv_hat.H // 2, mv_hat.W // 2, scale=0.5)
        mv3 = self.resize(mv2, mv2.H // 2, mv2.W // 2, scale=0.5)
        if ref_feature is None:
            f0 = self.conv("feature_adaptor_EL_I", ref)
        elif ref_feature.real == 64:
            f0 = self.conv("feature_adaptor_EL_first_P", ref_feature)
        else:
            f0 = self.conv("feature_adaptor_EL", ref_feature)
        rf1, rf2, rf3 = self.extractor3("feature_extractor", f0)
        c1_init = self.warp(rf1, mv_hat)
        # OffsetDiversity (lssvc_modules.py:75-112): conv_offset on cat(context1_init, warpframe, mv) at half res
        o = self.conv("align.conv_offset.0", [c1_init, warp_frame, mv_hat], stride=2, act=0.1)
        o = self.conv("align.conv_offset.2", o, act=0.1)
        o = self.conv("align.conv_offset.4", o)
        fw, fb = self.cached("align.fusion", lambda: (
            self.tensor("align.fusion.weight").detach().float().reshape(48, 6).contiguous().to(self.device),
            self.tensor("align.fusion.bias").detach().float().to(self.device)))
        c1 = self.new(rf1.H, rf1.W, rf1.real)
        ops.offset_diversity(rf1.exact(), o.exact(), mv_hat, fw, fb, 16, 2, 40.0, c1.exact())
        c2, c3 = self.warp(rf2, mv2), self.warp(rf3, mv3)
        return self.fusion3("context_fusion_net", c1, c2, c3), warp_frame

    def _hybrid_contexts(self, texture_bl, mv_hat, ref, ref_feature):
        """hybrid_temporal_layer_context_fusion (LSSVC_net.py:246-259)."""
        (t1, t2, t3), warp_frame = self._temporal_contexts(ref, ref_feature, mv_hat)
        texture = self._texture_resampler(texture_bl)
        s1, s2, s3 = self.extractor3("texture_extractor", texture)
        blended = []
        for i, (t, s) in enumerate(((t1, s1), (t2, s2), (t3, s3)), start=1):
            g = f"weight_map_generator.generator{i}"
            m, ma = self.conv(g + ".0", [t, s], act_copy=0.01)
            m = self.res_block(g + ".1", m, x_act=ma, end_with_relu=True)
            logits = self.conv(g + ".2", m)
            b = self.new(t.H, t.W, t.real)
            ops.softmax2_blend(logits.exact(), t.exact(), s.exact(), b.exact())
            blended.append(b)
        c1, c2, c3 = self.fusion3("context_fusion_net", *blended)
        return c1, c2, c3, warp_frame

    def _res_encode(self, xe, c1, c2, c3):
        """ResEncoder (lssvc_modules.py:235-254): ResBlocks start with LeakyReLU(0.1) on the concatenation."""
        f = self._cat_res_block("res_encoder.res1", c2, "res_encoder.conv1", [xe, c1], stride=2)
        f = self._cat_res_block("res_encoder.res2", c3, "res_encoder.conv2", f, stride=2)
        f = self.conv("res_encoder.conv3", f, stride=2)
        return self.conv("res_encoder.conv4", f, stride=2)

    def _cat_res_block(self, name, ctx, conv_name, conv_src, **conv_kw):
        """ResBlock on cat(conv(conv_src), ctx): the convolution writes its output straight into its slice of the
        concatenation buffer (the residual add of the block needs the concatenation as ONE view), the context is copied in."""
        cout = self.tensor(conv_name + ".weight").shape[0] // (4 if conv_kw.get("ps") else 1)
        assert cout % 4 == 0
        cat = View.alloc(ctx.H, ctx.W, cout + ctx.real, self.device)
        self.conv(conv_name, conv_src, out=cat.slice(0, cout), **conv_kw)
        self.copy(ctx, cat.slice(cout, cat.C))
        return self.res_block(name, cat, slope=0.1, start_from_relu=True, end_with_relu=True)

    def _res_params(self, z_hat, c3, y_bl_hat, entropy=None):
        """entropy: step 0 of the four-part prior, coded in the epilogue of the block's last convolution (encoder side)."""
        h = self.conv("res_prior_decoder.0", z_hat, act=0.01, exact_in=True)
        h = self.conv("res_prior_decoder.2.0", h, ps=True, act=0.01, pad=0)
        h = self.conv("res_prior_decoder.4", h, act=0.01)
        h = self.conv("res_prior_decoder.6.0", h, ps=True, act=0.01, pad=0)
        cat = View.alloc(h.H, h.W, 3 * 128, self.device)
        self.conv("res_prior_decoder.8", h, out=cat.slice(0, 128))
        t = self.conv("temporal_prior_encoder.0", c3, stride=2, act=0.1)
        self.conv("temporal_prior_encoder.2", t, stride=2, out=cat.slice(128, 256))
        self.copy(self._layer_prior_resampler(y_bl_hat), cat.slice(256, 384))
        # PriorFusion (lssvc_modules.py:432-442)
        t = self.depth_conv_block("prior_fusion_net.prior_fusion_conv.0", cat)
        return self.depth_conv_block("prior_fusion_net.prior_fusion_conv.1", t, entropy=entropy)

    def _spatial_prior(self, step, y_hat_so_far, common, entropy=None):
        """y_spatial_prior(y_spatial_prior_adaptor_k(cat(y_hat_so_far, common_params)))   LSSVC_net.py:372-404"""
        t = self.conv(f"y_spatial_prior_adaptor_{step}", [y_hat_so_far, common], pad=0)
        for i in range(3):
            t = self.depth_conv_block(f"y_spatial_prior.{i}", t, entropy=entropy if i == 2 else None)
        return t

    def _four_part(self, y, z_hat, c3, y_bl_hat, bits, w=None):
        """forward_four_part_prior (LSSVC_net.py:338-443) together with the networks that produce its parameters: step k is coded
        in the epilogue of the last convolution of the block that emits step k's (scale | mean) — the prior-fusion net for step 0,
        y_spatial_prior for steps 1-3 (LSSVC_EPI_FOURPART; lssvc_four_part_step after the block when that cannot be fused, e.g.
        with the debug outputs on).  Returns (y_hat, common parameters)."""
        C = y.real
        y_hat = self.new(y.H, y.W, C)
        thr = self._thr() if w else None
        dbg = getattr(self, "_debug", None)
        y_q = self.new(y.H, y.W, C) if dbg is not None else None
        s_hat = self.new(y.H, y.W, C) if dbg is not None else None
        _dbg(self, y_q=y_q, scales_hat=s_hat)
        common = None
        for step in range(4):
            ent = {"mode": "fourpart", "step": step, "y": y.exact(), "y_hat": y_hat.exact(), "y_q": y_q, "s_hat": s_hat,
                   "bits": bits.ptr(1), "sym": w.buf(f"el_y{step}", y, C // 4) if w else None,
                   "index": w.buf(f"el_y{step}_idx", y, C // 4) if w else None, "thresholds": thr}
            if step == 0:
                common = self._res_params(z_hat, c3, y_bl_hat, entropy=ent)
            else:
                self._spatial_prior(step, y_hat, common, entropy=ent)
            if y_q is not None and getattr(self, "_force", None) and "y_q" in self._force:
                # coded-so-far positions are the ones with a (strictly positive) scale recorded
                _force(self, "y_q", y_hat, mask=s_hat.exact().as_tensor() != 0, q_view=y_q)
        return y_hat, common

    def _res_decode(self, y_hat, c1, c2, c3):
        """ResDecoder (lssvc_modules.py:257-276) + ReconGeneration (:279-292, called as (recon_image_feature, context1))."""
        f = self.conv("res_decoder.up1.0", y_hat, ps=True)
        f = self._cat_res_block("res_decoder.res1", c3, "res_decoder.up2.0", f, ps=True)
        f = self._cat_res_block("res_decoder.res2", c2, "res_decoder.up3.0", f, ps=True)
        rec = self.conv("res_decoder.up4.0", f, ps=True)
        f = self.conv("recon_generation_net.first_conv", [rec, c1])
        f = self._unet("recon_generation_net.unet_1", f)
        feature = self._unet("recon_generation_net.unet_2", f)
        return feature, self.conv("recon_generation_net.recon_conv", feature)

    def _unet(self, name, x):
        """UNet (lssvc_modules.py:295-336)."""
        cat2 = View.alloc(x.H, x.W, 64, self.device)                 # (x1 | up2(d3))
        x1 = self.depth_conv_block(name + ".conv1", x, out=cat2.slice(0, 32))
        p = self.new(x.H // 2, x.W // 2, 32)
        ops.maxpool2(x1, p)
        cat3 = View.alloc(p.H, p.W, 128, self.device)                # (x2 | up3(x3))
        x2 = self.depth_conv_block(name + ".conv2", p, out=cat3.slice(0, 64))
        p = self.new(p.H // 2, p.W // 2, 64)
        ops.maxpool2(x2, p)
        x3 = self.depth_conv_block(name + ".conv3", p)
        for i in range(4):
            x3 = self.depth_conv_block(f"{name}.context_refine.{i}", x3)
        self.conv(name + ".up3.0", x3, ps=True, pad=0, out=cat3.slice(64, 128))
        d3 = self.depth_conv_block(name + ".up_conv3", cat3)
        self.conv(name + ".up2.0", d3, ps=True, pad=0, out=cat2.slice(32, 64))
        return self.depth_conv_block(name + ".up_conv2", cat2)

    # ---- forward ------------------------------------------------------------------------------------------------
    def _dpb_view(self, dpb, key, image, as_view_of=None):
        """NHWC view of a DPB entry.  as_view_of: a static view to convert into (graph replay) when no valid native
        view travels with the tensor."""
        t = dpb.get(key)
        if t is None:
            return None
        native = dpb.get("_native", {}).get(key)
        if native is not None and native[1] is t and native[2] == t._version:
            token = native[3] if len(native) > 3 else None
            if token is None or token[0]["gen"] == token[1]:    # a graph's static output is valid until its next replay
                return native[0]
        if as_view_of is not None:
            if t.device != self.device:
                t = t.to(self.device)
            return View.from_nchw(t.float(), out=as_view_of)
        return self.image_view(t) if image else self.feature_view(t)

    def _frame_core(self, xb, xe, rb, re, fb, fe, bits, w):
        """The kernel launches of one P-frame on NHWC views (no host synchronisation, no host-side state besides the
        weight caches): what a CUDA graph of the frame captures."""
        bl = self._base_layer(xb, rb, fb, bits, w)
        el = self._el_layer(xe, re, fe, bl["feature"], bl["y_hat"], bl["mv_hat"], bits, w)
        bits.fetch_range()
        return {"bl_recon": bl["recon"], "bl_feature": bl["feature"], "recon": el["recon"], "feature": el["feature"],
                "mv_hat": el["mv_hat"], "warp_frame": el["warp_frame"]}

    def _el_layer(self, xe, re, fe, texture_bl, y_hat_bl, mv_hat_bl, bits, w):
        """Enhancement layer of one P-frame given the decoded base layer (LSSVC_net.py:455-508, LSSVC_net_extend.py:24-86)."""
        # inter-layer processing: de-padded base-layer texture / motion / latent (LSSVC_net.py:453-456)
        texture_bl, mv_hat_bl, y_hat_bl = self.depad(texture_bl), self.depad(mv_hat_bl), self.depad(y_hat_bl, 16)
        # EL motion
        mv_ctx_prior, mv_ctx = self._mv_contexts(mv_hat_bl)
        mv = self.spynet("optic_flow", xe, re)
        mv_y = self._mv_encode(mv, mv_ctx)
        mv_z_hat = self._prior_encoder("mv_prior_encoder", mv_y, self._bitparm_coef("bit_estimator_z_mv."), bits.ptr(1), w, "el_mv_z")
        _force(self, "mv_z_hat", mv_z_hat)
        C = mv_y.real
        mv_y_hat = self.new(mv_y.H, mv_y.W, C)
        mv_prm = self._mv_params(mv_z_hat, mv_ctx_prior,
                                 entropy=self._laplace(mv_y, mv_y_hat, bits.ptr(1), w, "el_mv_y", self._thr() if w else None))
        _force(self, "mv_y_q", mv_y_hat, mean=mv_prm.slice(C, 2 * C))
        mv_hat = self._mv_decode(mv_y_hat, mv_ctx)
        # contexts, residual coding
        c1, c2, c3, warp_frame = self._hybrid_contexts(texture_bl, mv_hat, re, fe)
        y = self._res_encode(xe, c1, c2, c3)
        z_hat = self._prior_encoder("res_prior_encoder", y, self._bitparm_coef("bit_estimator_z."), bits.ptr(1), w, "el_z")
        _force(self, "z_hat", z_hat)
        y_hat, params = self._four_part(y, z_hat, c3, y_hat_bl, bits, w)
        if w:
            w.layer_done("el")
        feature, recon = self._res_decode(y_hat, c1, c2, c3)
        _dbg(self, mv=mv, mv_y=mv_y, mv_y_hat=mv_y_hat, mv_prm=mv_prm, mv_z_hat=mv_z_hat, z_hat=z_hat, y=y, y_hat=y_hat,
             params=params, c1=c1, c2=c2, c3=c3)
        return {"recon": recon, "feature": feature, "mv_hat": mv_hat, "warp_frame": warp_frame}

    def _frame_result(self, v, bit_bl, bit_el, token=None):
        """Views of one coded frame -> the reference's result dict (fresh NCHW tensors the caller may mutate)."""
        # the 64-channel feature maps leave as channels_last tensors sharing the NHWC buffers (no conversion pass) unless they
        # are a graph's static outputs, which the next replay overwrites
        feat = (lambda x: x.to_nchw()) if token is not None else (lambda x: x.to_nchw_shared())
        out = {"ref_frame_bl": v["bl_recon"].to_nchw(), "ref_feature_bl": feat(v["bl_feature"]),
               "ref_frame_el": v["recon"].to_nchw(), "ref_feature_el": feat(v["feature"])}
        # NHWC originals of the features: the next frame reads them directly when the caller passes the tensors back
        # untouched.  token: (graph entry, generation) when the views are a graph's static outputs, valid until its next replay.
        out["_native"] = {"ref_feature_bl": (v["bl_feature"], out["ref_feature_bl"], out["ref_feature_bl"]._version, token),
                          "ref_feature_el": (v["feature"], out["ref_feature_el"], out["ref_feature_el"]._version, token)}
        return {"dpb": out, "bit_bl": bit_bl, "bit_el": bit_el, "encoding_time_EL": 0.0, "decoding_time_EL": 0.0,
                "encoding_time_BL": 0.0, "decoding_time_BL": 0.0, "mv_hat": v["mv_hat"].to_nchw(),
                "warp_frame": v["warp_frame"].to_nchw()}

    @torch.no_grad()
    def forward_one_frame(self, x_bl, x_el, ref_frame_bl, ref_frame_el, ref_feature_bl, ref_feature_el, _dpb=None,
                          _write=None, _async=False):
        """LSSVC.forward_one_frame (LSSVC_net.py:445-528).  _async (runner.py): no host synchronisation — the bit counts
        stay on the device (result["_bits"]) and "bit_bl" / "bit_el" are None until the caller reads them."""
        self._require_cuda()
        dpb = _dpb if _dpb is not None else {"ref_frame_bl": ref_frame_bl, "ref_frame_el": ref_frame_el,
                                             "ref_feature_bl": ref_feature_bl, "ref_feature_el": ref_feature_el}
        if x_bl.dim() == 4 and x_bl.shape[0] > 1:
            # batch > 1 (estimate mode only, as in the reference): independent frames with their own DPB entries, one after the other
            if _write is not None or _async:
                raise ValueError("LSSVC: bitstream / asynchronous coding takes one frame at a time (batch 1)")
            keys = ("ref_frame_bl", "ref_frame_el", "ref_feature_bl", "ref_feature_el")
            return _merge_batch([self.forward_one_frame(x_bl[b:b + 1], x_el[b:b + 1], None, None, None, None,
                                                        _dpb={k: _item(dpb.get(k), b) for k in keys})
                                 for b in range(x_bl.shape[0])])
        if (self.use_graphs and _write is None and getattr(self, "_debug", None) is None and not getattr(self, "_force", None)
                and ops.TRACE is None and not _lib.DRY_RUN):
            return self._forward_graphed(x_bl, x_el, dpb, _async)
        bits = _Bits(self.device)
        xb, xe = self.image_view(x_bl), self.image_view(x_el)
        rb, re = self._dpb_view(dpb, "ref_frame_bl", True), self._dpb_view(dpb, "ref_frame_el", True)
        fb, fe = self._dpb_view(dpb, "ref_feature_bl", False), self._dpb_view(dpb, "ref_feature_el", False)
        v = self._frame_core(xb, xe, rb, re, fb, fe, bits, _write)
        return self._frame_done(v, bits, _async, redo=(x_bl, x_el, dpb, _write))

    def _frame_done(self, v, bits, _async, token=None, redo=None):
        if _async:
            r = self._frame_result(v, None, None, token=token)
            r["_bits"] = bits
            return r
        bit_bl, bit_el = bits.read()
        if bits.out_of_range:
            x_bl, x_el, dpb, w = redo

            def again():
                graphs, self.use_graphs = self.use_graphs, False
                try:
                    return self.forward_one_frame(x_bl, x_el, None, None, None, None, _dpb=dpb, _write=w)
                finally:
                    self.use_graphs = graphs
            return _recode_fp32(self, "LSSVC.forward_one_frame", again)
        return self._frame_result(v, bit_bl, bit_el, token=token)

    # ---- whole-frame CUDA graph ---------------------------------------------------------------------------------
    # A P-frame is ~530 kernel launches with static shapes: after one eager frame per input signature (which packs the
    # weights and warms the allocator) the launches are captured once and replayed; inputs are staged into static
    # buffers, results leave as fresh tensors, the only host<->device sync stays the read of the two bit counters.
    # `lane` (runner.py): coding lanes of one GPU share the model (weights, packed weights) but replay their own graphs —
    # a graph owns its static inputs / outputs, and lanes run concurrently on their own streams.
    lane = 0

    def _forward_graphed(self, x_bl, x_el, dpb, _async=False):
        fbt, fet = dpb.get("ref_feature_bl"), dpb.get("ref_feature_el")
        key = (self.lane, tuple(x_bl.shape), tuple(x_el.shape), None if fbt is None else tuple(fbt.shape),
               None if fet is None else tuple(fet.shape))
        g = self._graphs.get(key)
        if g is None or g == "warm":
            if g is None:
                self._graphs[key] = "warm"
                bits = _Bits(self.device)
                v = self._frame_core(self.image_view(x_bl), self.image_view(x_el), self._dpb_view(dpb, "ref_frame_bl", True),
                                     self._dpb_view(dpb, "ref_frame_el", True), self._dpb_view(dpb, "ref_feature_bl", False),
                                     self._dpb_view(dpb, "ref_feature_el", False), bits, None)
                return self._frame_done(v, bits, _async, redo=(x_bl, x_el, dpb, None))
            g = self._graphs[key] = self._capture_frame(x_bl, x_el, dpb)
        # ---- stage the inputs
        for name, t in (("x_bl", x_bl), ("x_el", x_el), ("ref_frame_bl", dpb["ref_frame_bl"]), ("ref_frame_el", dpb["ref_frame_el"])):
            g["in"][name].copy_(t, non_blocking=True)
        for name in ("ref_feature_bl", "ref_feature_el"):
            dst = g["in"][name]
            if dst is None:
                continue
            src = self._dpb_view(dpb, name, False, as_view_of=dst)
            if src is not dst:
                dst.buf.copy_(src.buf)
        g["gen"] += 1
        g["graph"].replay()
        _lib.load().lssvc_launch_count_add(g["launches"])
        # the graph zeroes its static counters at the start of every replay: an asynchronous caller gets a snapshot
        bits = _Bits(self.device, g["bits"].t.clone()) if _async else g["bits"]
        return self._frame_done(g["out"], bits, _async, token=(g, g["gen"]), redo=(x_bl, x_el, dpb, None))

    def _capture_frame(self, x_bl, x_el, dpb):
        dev = self.device
        new_like = lambda t: torch.empty(tuple(t.shape), dtype=torch.float32, device=dev)
        static = {"x_bl": new_like(x_bl), "x_el": new_like(x_el), "ref_frame_bl": new_like(dpb["ref_frame_bl"]),
                  "ref_frame_el": new_like(dpb["ref_frame_el"])}
        for name in ("ref_feature_bl", "ref_feature_el"):
            t = dpb.get(name)
            if t is None:
                static[name] = None
            else:
                C = t.shape[1]
                v = View.alloc(t.shape[2], t.shape[3], ops.round_up(C, 8), dev, zero=True)
                v.real = C
                static[name] = v
        bits = _Bits(dev)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        l0 = _lib.launch_count()
        # the frame graphs of one lane never run concurrently (the P-chain is serial): they share one memory pool
        pool = self._graph_pools.setdefault(self.lane, torch.cuda.graph_pool_handle())
        with torch.cuda.graph(graph, pool=pool):
            bits.t.zero_()
            v = self._frame_core(self.image_view(static["x_bl"]), self.image_view(static["x_el"]),
                                 self.image_view(static["ref_frame_bl"]), self.image_view(static["ref_frame_el"]),
                                 static["ref_feature_bl"], static["ref_feature_el"], bits, None)
        return {"graph": graph, "in": static, "out": v, "bits": bits, "gen": 0, "launches": _lib.launch_count() - l0}

    def encode_decode_extend(self, *args, **kwargs):
        raise NotImplementedError("bitstream writing needs LSSVC_extend")

    def encode_decode(self, x_bl, x_el, dpb, output_path_bl=None, output_path_el=None, pic_width=None, pic_height=None,
                      pic_width_bl=None, pic_height_bl=None):
        """LSSVC.encode_decode (LSSVC_net.py:172-185)."""
        if output_path_el is not None:
            return self.encode_decode_extend(x_bl, x_el, dpb, output_path_bl, output_path_el, pic_width, pic_height,
                                             pic_width_bl, pic_height_bl)
        return self.forward_one_frame(x_bl, x_el, dpb["ref_frame_bl"], dpb["ref_frame_el"], dpb["ref_feature_bl"],
                                      dpb["ref_feature_el"], _dpb=dpb)


class LSSVC_extend(LSSVC):
    """LSSVC_extend()   LSSVC_net_extend.py:8-22 — adds update / compress / decompress / encode_decode_extend."""

    def __init__(self, seed=None):
        super().__init__(seed=seed)
        # the reference's `model.base_layer_model` is a DMCExtend with its own compress / decompress / encode_decode_extend
        # / update (dmc_net_extend.py:49-173): the parameter container of the same name gets those entry points
        api = _BaseLayerAPI(self)
        for name in ("compress", "decompress", "encode_decode_extend", "update"):
            object.__setattr__(self.base_layer_model, name, getattr(api, name))

    def update(self, force=False):
        """LSSVC_extend.update + DMCExtend.update (LSSVC_net_extend.py:17-22, dmc_net_extend.py:49-53)."""
        if self._tables is not None and not force:
            return
        t = {"laplace": entropy.laplace_table()}
        for tag, p in (("el_z", "bit_estimator_z."), ("el_mv_z", "bit_estimator_z_mv."),
                       ("bl_z", "base_layer_model.bit_estimator_z."), ("bl_mv_z", "base_layer_model.bit_estimator_z_mv.")):
            t[tag] = entropy.bitparm_table(self._bitparm_coef(p).cpu())
        self._tables = t

    # ---- real bitstreams: codec.py -------------------------------------------------------------------------------
    single_pass_streams = False     # True: one encoder pass + stream verification (streams.py) instead of encode + decode

    @torch.no_grad()
    def compress(self, x, dpb):
        """LSSVC_extend.compress(x, dpb) (LSSVC_net_extend.py:24-86): EL of one P-frame -> {"string", "dpb"}."""
        from . import codec
        self._require_cuda()
        return codec.el_compress(self, x, dpb)

    @torch.no_grad()
    def decompress(self, string, height, width, dpb):
        """LSSVC_extend.decompress(string, height, width, dpb) (LSSVC_net_extend.py:88-142)."""
        from . import codec
        self._require_cuda()
        return codec.el_decompress(self, string, height, width, dpb)

    @torch.no_grad()
    def encode_decode_extend(self, x_bl, x_el, dpb, output_path_bl=None, output_path_el=None, pic_width=None,
                             pic_height=None, pic_width_bl=None, pic_height_bl=None):
        """LSSVC_extend.encode_decode_extend (LSSVC_net_extend.py:144-191): per layer compress -> file -> decompress; the
        next frame's DPB is what the DECODER reconstructed."""
        self._require_cuda()
        if self.single_pass_streams:
            from .streams import inter_encode_decode
            return inter_encode_decode(self, x_bl, x_el, dpb, output_path_bl, output_path_el, pic_width, pic_height,
                                       pic_width_bl, pic_height_bl)
        from . import codec
        return codec.encode_decode_extend(self, x_bl, x_el, dpb, output_path_bl, output_path_el, pic_width, pic_height,
                                          pic_width_bl, pic_height_bl)


class _BaseLayerAPI:
    """DMCExtend's public calls (dmc_net_extend.py:49-173) bound to an LSSVC model."""

    def __init__(self, model):
        self._m = model

    def update(self, force=False):
        self._m.update(force=force)

    @torch.no_grad()
    def compress(self, x, dpb):
        from . import codec
        self._m._require_cuda()
        return codec.bl_compress(self._m, x, dpb)

    @torch.no_grad()
    def decompress(self, string, height, width, dpb):
        from . import codec
        self._m._require_cuda()
        return codec.bl_decompress(self._m, string, height, width, dpb)

    @torch.no_grad()
    def encode_decode_extend(self, x, dpb, output_path=None, pic_width=None, pic_height=None):
        from . import codec
        self._m._require_cuda()
        return codec.bl_encode_decode_extend(self._m, x, dpb, output_path, pic_width, pic_height)
