"""Host-side operator layer: NHWC views over torch-owned device memory and thin wrappers that hand raw
pointers to the C-ABI kernels.  torch is used for allocation and stream handles only."""
import ctypes
import math
import os

import torch

from . import _lib
from ._lib import CConv, CView, byref

_NULL_VIEW = CView(None, 0, 0, 0, 0)


def _stream():
    if _lib.DRY_RUN and not torch.cuda.is_available():
        return None
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def round_up(x, m):
    return (x + m - 1) // m * m


class View:
    """A window of C channels of an NHWC fp32 buffer with `pitch` floats per pixel."""

    __slots__ = ("buf", "H", "W", "C", "pitch", "coff", "real")

    def __init__(self, buf, H, W, C, pitch, coff=0, real=None):
        self.buf, self.H, self.W, self.C, self.pitch, self.coff = buf, H, W, C, pitch, coff
        self.real = C if real is None else real   # channels that carry data; C - real trailing channels are zero

    @staticmethod
    def alloc(H, W, C, device, pitch=None, zero=False):
        pitch = pitch or round_up(C, 4)
        fn = torch.zeros if zero else torch.empty
        return View(fn(H * W * pitch, dtype=torch.float32, device=device), H, W, C, pitch)

    def slice(self, c0, c1):
        assert 0 <= c0 < c1 <= self.C, (c0, c1, self.C)
        return View(self.buf, self.H, self.W, c1 - c0, self.pitch, self.coff + c0)

    def widen(self, C):
        """Same window start, C channels wide, exposing zeroed pad channels to a conv; `real` is kept."""
        assert self.coff + C <= self.pitch and C >= self.real
        return View(self.buf, self.H, self.W, C, self.pitch, self.coff, real=self.real)

    def exact(self):
        """The window without its pad channels (what a producer kernel writes)."""
        return View(self.buf, self.H, self.W, self.real, self.pitch, self.coff)

    @staticmethod
    def alloc_padded(H, W, C, device, mult=8):
        """C data channels in a buffer padded (and zeroed) to a multiple of `mult` so tensor-core convs can read it."""
        Cp = round_up(C, mult)
        if Cp == C:
            return View.alloc(H, W, C, device)
        v = View.alloc(H, W, Cp, device, pitch=Cp, zero=True)
        v.real = C
        return v

    @property
    def device(self):
        return self.buf.device

    def c(self):
        return CView(self.buf.data_ptr() + 4 * self.coff, self.H, self.W, self.C, self.pitch)

    def as_tensor(self):
        """[H, W, C] strided torch view (tests / debugging)."""
        return self.buf.view(self.H, self.W, self.pitch)[:, :, self.coff:self.coff + self.C]

    def to_nchw(self):
        """[1, real, H, W] contiguous copy."""
        v = self.exact() if self.real != self.C else self
        out = torch.empty(1, v.C, v.H, v.W, dtype=torch.float32, device=self.device)
        lib = _lib.load()
        _lib.check(lib.lssvc_nhwc_to_nchw(byref(v.c()), _ptr(out), _stream()), "nhwc_to_nchw")
        return out

    def to_nchw_shared(self):
        """[1, real, H, W] tensor in torch's channels_last memory format that SHARES this view's buffer (no copy) when the
        view is a whole unpadded buffer; a contiguous copy otherwise.  For the wide feature maps of the DPB: the caller gets a
        regular tensor (in-place edits included), the next frame re-imports it without a pass (View.from_channels_last)."""
        if self.coff == 0 and self.pitch == self.C and self.real == self.C and self.buf.numel() == self.H * self.W * self.C:
            return self.buf.view(1, self.H, self.W, self.C).permute(0, 3, 1, 2)
        return self.to_nchw()

    @staticmethod
    def from_channels_last(t):
        """Zero-copy NHWC view of a [1, C, H, W] fp32 tensor stored channels_last (what to_nchw_shared hands out); None if
        the tensor is not of that kind."""
        if (t.dim() == 4 and t.shape[0] == 1 and t.dtype == torch.float32 and t.shape[1] % 8 == 0 and t.shape[1] > 1
                and t.stride() == (t.shape[1] * t.shape[2] * t.shape[3], 1, t.shape[3] * t.shape[1], t.shape[1])
                and t.data_ptr() % 16 == 0):
            _, C, H, W = t.shape
            return View(t.permute(0, 2, 3, 1).reshape(-1), H, W, C, C)
        return None

    @staticmethod
    def from_nchw(t, C_view=None, out=None):
        """[1, C, H, W] (or [C, H, W]) contiguous fp32 -> NHWC view with C_view >= C channels (rest zero)."""
        if t.dim() == 4:
            assert t.shape[0] == 1, "batch 1 only"
            t = t[0]
        t = t.contiguous()
        assert t.dtype == torch.float32 and (t.is_cuda or _lib.DRY_RUN)
        C, H, W = t.shape
        if out is None:
            out = View.alloc(H, W, C_view or C, t.device)
            out.real = C
        lib = _lib.load()
        _lib.check(lib.lssvc_nchw_to_nhwc(_ptr(t), C, byref(out.c()), _stream()), "nchw_to_nhwc")
        return out


class LazyAct(View):
    """lrelu(base, slope) that has NOT been written anywhere: the consuming convolution applies the LeakyReLU in its
    operand path (in_transform).  Handing it to a kernel directly is a wiring error, so c() refuses."""

    __slots__ = ("base", "slope")

    def __init__(self, base, slope):
        super().__init__(base.buf, base.H, base.W, base.C, base.pitch, base.coff, real=base.real)
        self.base, self.slope = base, float(slope)

    def c(self):
        raise RuntimeError("LazyAct view reached a kernel: it must be consumed by Engine.conv (or materialised)")

    def exact(self):
        raise RuntimeError("LazyAct view reached a kernel: it must be consumed by Engine.conv (or materialised)")


def _cv(v):
    return v.c() if v is not None else _NULL_VIEW


# The tensor core adds every K = 16 MMA into the fp32 TMEM accumulator with truncation (round towards zero), so a sum of
# `steps` MMAs comes out SMALLER in magnitude than the exact sum by a nearly deterministic factor: on the 15 layer shapes
# of tools/conv_bench.py (CONV_BENCH_BIAS=1, profiles/r1_accumulation_bias.txt) the mean signed relative error is
# -(0.264 * steps + 0.6) * 2^-24 for steps = 4 .. 196, and it — not the split-fp16 operands — was ~85 % of the rms error
# (3x3 64->64: 7.2e-7 against 4.3e-7 for fp32 CUDA cores), coherent from layer to layer.  Its expected value is undone at
# pack time: output channel c of the split-fp16 weights is scaled by 1 + (0.264 * steps_c + 0.6) * 2^-24, steps_c = the
# number of (tap, 16-channel slice) weight blocks of that channel that are not all zero (all-zero blocks — channel
# padding, the empty taps of a sub-pixel-decomposed ConvTranspose2d — add exact zeros and truncate nothing).
# LSSVC_ACC_COMP=0 disables it (A/B).
ACC_COMP = os.environ.get("LSSVC_ACC_COMP", "1") != "0"


def acc_comp(steps, coherent=False):
    """Per-output-channel weight factor compensating the accumulator truncation; steps: tensor or number of MMA steps.
    coherent: every product of the layer has the same sign (GDN's norm pool: gamma >= 0 times x^2 >= 0), so the accumulator
    grows monotonically and every truncation costs ~1/2 ulp of a value close to the final sum: measured -(1.43 T - 1.9) x 2^-24
    for T = 4 .. 12 accumulation steps (profiles/r2_accumulation_bias_coherent.txt) against -(0.264 T + 0.6) for zero-mean
    data."""
    steps = torch.as_tensor(steps, dtype=torch.float32)
    if not ACC_COMP:
        return torch.ones_like(steps)
    if coherent:
        return torch.where(steps > 0, 1.0 + (1.43 * steps - 1.9).clamp_min(0.0) * 2.0 ** -24, torch.ones_like(steps))
    return torch.where(steps > 0, 1.0 + (0.264 * steps + 0.6) * 2.0 ** -24, torch.ones_like(steps))


class PackedConv:
    """Weights of one Conv2d repacked for the kernels: [kh*kw][n_pad][cin_total] + bias[n_pad]."""

    __slots__ = ("weight", "bias", "kh", "kw", "stride", "pad", "cout", "n_pad", "cin_total", "src_c", "pixel_shuffle",
                 "_h2", "exact_in", "coherent", "pair_tile", "_host")

    def __init__(self, w, b, stride=1, pad=None, src_channels=None, pixel_shuffle=False, transposed=False, device=None,
                 exact_in=False, coherent=False, pair_tile=0):
        """w: [Cout, Cin, kh, kw] (Conv2d) or, with transposed=True, a stride-1 ConvTranspose2d weight
        [Cin, Cout, kh, kw] (turned into the equivalent flipped Conv2d).  src_channels: list of
        (real, view) channel counts per source when sources are padded; default one unpadded source.
        exact_in: the input holds quantised symbols (z_hat: small integers, x_lo = 0, hi*hi products that add without
        truncation), so the accumulator-truncation compensation (acc_comp) must stay off for this layer."""
        self.exact_in = exact_in
        self.coherent = coherent      # sign-coherent products (GDN norm pool): acc_comp(..., coherent=True)
        # pair_tile: the layer emits (scale | mean) for a Laplace entropy epilogue spanning several channel tiles of this width:
        # packed tile t = [scale of channels t*pair_tile/2 .. | their means] (lssvc_conv::ent_tile); usable ONLY with that epilogue
        self.pair_tile = pair_tile
        w = w.detach().to(torch.float32).cpu()
        if transposed:
            w = w.permute(1, 0, 2, 3).flip(2, 3)
        cout, cin, kh, kw = w.shape
        b = torch.zeros(cout) if b is None else b.detach().to(torch.float32).cpu()
        if src_channels is None:
            src_channels = [(cin, cin)]
        assert sum(r for r, _ in src_channels) == cin, (src_channels, cin)
        if pair_tile:
            assert not pixel_shuffle and pair_tile % 32 == 0 and cout % pair_tile == 0, (cout, pair_tile)
            C, Ct = cout // 2, pair_tile // 2
            perm = torch.tensor([(C if r >= Ct else 0) + t * Ct + (r % Ct) for t in range(cout // pair_tile) for r in range(pair_tile)])
            w, b = w[perm], b[perm]
        if pixel_shuffle:
            assert cout % 4 == 0
            cq = cout // 4
            # packed channel (2i+j)*cq + c  <-  reference channel 4c + 2i + j
            perm = torch.tensor([4 * c + s for s in range(4) for c in range(cq)])
            w, b = w[perm], b[perm]
        n_pad = round_up(cout, 16)
        cin_total = sum(v for _, v in src_channels)
        packed = torch.zeros(kh * kw, n_pad, cin_total)
        wt = w.permute(2, 3, 0, 1).reshape(kh * kw, cout, cin)
        ci = co = 0
        for real, view in src_channels:
            packed[:, :cout, co:co + real] = wt[:, :, ci:ci + real]
            ci += real
            co += view
        bias = torch.zeros(n_pad)
        bias[:cout] = b
        # the host copy stays until the split-fp16 image has been derived from it (weight_h2): packing a model used to issue
        # ~5 000 tiny device launches on first use; now it is host arithmetic + one upload per tensor
        self._host = packed.contiguous()
        self.weight = self._host.to(device)
        self.bias = bias.to(device)
        self.kh, self.kw, self.stride = kh, kw, stride
        self.pad = kh // 2 if pad is None else pad
        self.cout, self.n_pad, self.cin_total = cout, n_pad, cin_total
        self.src_c = [v for _, v in src_channels]
        self.pixel_shuffle = pixel_shuffle
        self._h2 = None

    def weight_h2(self):
        """Split-fp16 weights for lssvc_conv_hs: (fp16 [taps][2 (hi, lo)][n_pad][cin_pad16], cin_pad16, acc_scale).
        The weights are first scaled by the power of two that puts max|w| in [2^13, 2^14), so that w_lo is a normal
        fp16 for every weight that matters; acc_scale = 2^-shift undoes it on the fp32 accumulator (exact).
        Each source's channels start at a multiple of 16 (the MMA's K step); pad columns are zero."""
        if self._h2 is None:
            w = self._host if self._host is not None else self.weight.cpu()      # host arithmetic (IEEE: same bits as on the device)
            taps, n_pad, _ = w.shape
            m = float(w.abs().max())
            shift = 0 if m == 0.0 else 13 - int(math.floor(math.log2(m)))
            ws = w * (2.0 ** shift)
            cin16 = sum(round_up(c, 16) for c in self.src_c)
            # accumulation steps per output channel: (tap, 16-channel slice) blocks holding a non-zero weight
            nz = torch.zeros(taps, n_pad, cin16, dtype=torch.bool, device=w.device)
            ci = co = 0
            for c in self.src_c:
                nz[:, :, co:co + c] = w[:, :, ci:ci + c] != 0
                ci += c
                co += round_up(c, 16)
            steps = nz.reshape(taps, n_pad, cin16 // 16, 16).any(-1).sum(dim=(0, 2))
            if not self.exact_in:
                ws = ws * acc_comp(steps, self.coherent).to(w.device)[None, :, None]
            packed = torch.zeros(taps, 2, n_pad, cin16, dtype=torch.float16, device=w.device)
            ci = co = 0
            for c in self.src_c:
                blk = ws[:, :, ci:ci + c]
                hi = blk.to(torch.float16)
                packed[:, 0, :, co:co + c] = hi
                packed[:, 1, :, co:co + c] = (blk - hi.to(torch.float32)).to(torch.float16)
                ci += c
                co += round_up(c, 16)
            self._h2 = (packed.contiguous().to(self.weight.device), cin16, 2.0 ** -shift)
            self._host = None
        return self._h2


# Convolution engines:
#   "h2"   split-fp16 on tcgen05 (csrc/conv_hs.cu): fp32-level accuracy, the engine of the path
#   "simt" fp32 CUDA cores (csrc/conv_simt.cu): exact fp32 arithmetic — the on-device cross-check, the home of views the TMA
#          path cannot address, and what a frame is re-coded on when its activations leave the fp16 range (csrc/range.cu)
# (the earlier TF32 / 3xTF32 / TMEM-operand kernels live under tools/engines/ as study material; they are not built into
# the library)
ENGINES = {
    "simt": {"kernel": "conv_simt_kernel", "dtype": "f32", "peak_vs_bf16": 0.5,
             "note": "fp32 FMA on CUDA cores (no tensor cores); shown against the TF32 tensor peak for comparison"},
    "h2": {"kernel": "conv_hs_kernel", "dtype": "f16x2-split", "peak_vs_bf16": 1.0,
           "note": "tcgen05.mma kind::f16 on split-fp16 operands (x = x_hi + x_lo; 3 MMAs per algorithmic MAC, both operands "
                   "read from shared memory, taps = shifted descriptors on one converted halo tile, hi*hi and cross terms in "
                   "separate fp32 TMEM accumulators); FLOPs counted are algorithmic, so the kernel's own ceiling is peak/3"},
}
# Cout <= 4 heads leave the tensor-core engine for the register-blocked fp32 kernel (csrc/conv_head.cu) when the caller did not
# name an engine; LSSVC_NO_HEAD=1 keeps them on conv_hs (A/B)
HEAD_KERNEL = os.environ.get("LSSVC_NO_HEAD", "0") in ("", "0")
# Entropy epilogues: the convolution that produces (scale | mean) / z quantises the latent, counts the bits and dumps the symbols
# in its own epilogue (csrc/conv_hs.cu epilogue_entropy); LSSVC_NO_ENT_FUSE=1 runs the stand-alone entropy kernels instead (A/B)
ENT_FUSE = os.environ.get("LSSVC_NO_ENT_FUSE", "0") in ("", "0")
HEAD_MIN_PIXELS = 400_000     # below ~1/2 of 1080p per side the launch is latency-bound on either kernel (tools/head_bench.py)
_ENGINE = os.environ.get("LSSVC_CONV_ENGINE", "h2")
assert _ENGINE in ENGINES, _ENGINE


def default_engine():
    return _ENGINE


def set_engine(name):
    global _ENGINE
    assert name in ENGINES, name
    prev, _ENGINE = _ENGINE, name
    return prev


def engine_info(name):
    return dict(ENGINES[name])


def force_simt(flag):
    """Route every convolution through the fp32 CUDA-core kernel (on-device cross-check of tcgen05)."""
    global _ENGINE, _SAVED
    if flag:
        _SAVED = set_engine("simt")
    else:
        set_engine(_SAVED)


_SAVED = _ENGINE


# Optional launch trace (tools/join_launches.py): when TRACE is a list every conv appends its shape and engine.
TRACE = None
TRACE_NAME = None
# Convolutions the tensor-core engine could not take (a view that fails the TMA alignment rules, an unsupported stride /
# kernel size) and that ran on the fp32 CUDA-core kernel instead: a 10x slower launch nobody would otherwise notice.
# bench.py reports the count of its timed region (`roofline.simt_downgrades`).
SIMT_DOWNGRADES = 0
# Frames (or stream-mode layer calls) re-coded on the fp32 engine because an operand of the split-fp16 kernels left the fp16
# range (csrc/range.cu, models._recode_fp32).
RANGE_FALLBACKS = 0


def hs_channel_tile(n_pad):
    """The channel tile lssvc_conv_hs picks for n_pad packed output channels (csrc/conv_hs.cu: equal tiles of at most 128)."""
    n_tile = n_pad
    if n_tile > 128:
        n_tile = 128
        while n_tile >= 16 and n_pad % n_tile:
            n_tile -= 16
    return n_tile


def laplace_pair_tile(cout, engine=None):
    """pair_tile to pack a (scale | mean) convolution with so that its Laplace / four-part epilogue can span several channel tiles;
    0 when one tile holds all 2C channels (natural order) or when the epilogue will not be fused anyway."""
    if cout <= 128 or not ENT_FUSE or (engine or _ENGINE) not in ("h2", "hs") or cout % 16:
        return 0
    t = hs_channel_tile(cout)
    return t if t % 32 == 0 else 0


def _entropy_fusable(pc, out, act, res1, res2, out2, out_scale, epi, ent):
    """Can lssvc_conv_hs take this entropy epilogue?  (16-byte aligned views, LeakyReLU + one residual at most in front of it,
    scale / mean pairs inside a channel tile, no debug outputs)"""
    if not ENT_FUSE or res2 is not None or out2 is not None or out_scale != 1.0 or (res1 is not None and not view_aligned(res1)):
        return False
    if epi != _lib.EPI_PLAIN or pc.pixel_shuffle or pc.cout % 16 or pc.cout != pc.n_pad or not view_aligned(out):
        return False
    if ent["mode"] in ("laplace", "fourpart"):
        tile = hs_channel_tile(pc.n_pad)
        paired = pc.pair_tile == tile if pc.cout > 128 else pc.pair_tile in (0, tile)
        if ent["mode"] == "fourpart" and (pc.cout % 128 or ent.get("y_q") is not None or ent.get("s_hat") is not None):
            return False
        return tile % 32 == 0 and paired and view_aligned(ent["y"]) and view_aligned(ent["y_hat"])
    return ent["mode"] == "bitparm"


def entropy_standalone(ent, prm):
    """The stand-alone entropy kernel for an `entropy` request of ops.conv, on parameters `prm` = (scale | mean) already in
    memory (or, for "bitparm", prm = z and ent["z_hat"] the destination)."""
    if ent["mode"] == "bitparm":
        return bitparm_quant(prm, ent["coef"], ent["z_hat"], ent.get("bits"), sym=ent.get("sym"))
    C = prm.C // 2
    if ent["mode"] == "fourpart":
        return four_part_step(ent["y"], prm, ent["step"], ent["y_hat"], ent.get("y_q"), ent.get("s_hat"), ent.get("bits"),
                              sym=ent.get("sym"), index=ent.get("index"), thresholds=ent.get("thresholds"))
    return laplace_quant(ent["y"], prm.slice(C, 2 * C), prm.slice(0, C), None, ent["y_hat"], ent.get("bits"),
                         sym=ent.get("sym"), index=ent.get("index"), thresholds=ent.get("thresholds"))


def conv(pc, srcs, out, act=None, res1=None, res2=None, out2=None, slope2=0.0, out_scale=1.0,
         in_transform=_lib.IN_NONE, in_slope=0.0, epi=_lib.EPI_PLAIN, gdn_x=None, engine=None, entropy=None):
    """Run one packed convolution.  act: None or the LeakyReLU slope (0.0 = ReLU).
    engine: None (the default engine), 'h2' / 'hs' or 'simt'.
    entropy: the convolution produces entropy parameters and codes the latent in its epilogue (LSSVC_EPI_LAPLACE / _BITPARM):
      {"mode": "laplace", "y": View (C ch), "y_hat": View, "bits": tensor | None, "sym": int32 tensor | None,
       "index": int32 tensor | None, "thresholds": tensor | None}  — `out` receives (scale | mean) as usual;
      {"mode": "bitparm", "coef": tensor [C][11], "bits": ..., "sym": ...}  — `out` receives rint(conv output).
    When the tensor-core kernel cannot take it (other engine, several channel tiles, LSSVC_NO_ENT_FUSE=1) the convolution and the
    stand-alone entropy kernel run one after the other: same results bit for bit."""
    if isinstance(srcs, View):
        srcs = [srcs]
    assert len(srcs) == len(pc.src_c), (len(srcs), pc.src_c)
    if entropy is not None:
        eng = engine or _ENGINE
        fuse = eng in ("h2", "hs") and _entropy_fusable(pc, out, act, res1, res2, out2, out_scale, epi, entropy)
        fuse = (fuse and all(s.C % 4 == 0 and s.pitch % 4 == 0 and s.coff % 4 == 0 for s in srcs) and pc.stride in (1, 2)
                and pc.kh * pc.kw <= 49 and srcs[0].H % pc.stride == 0 and srcs[0].W % pc.stride == 0)   # conv_hs's own conditions
        assert fuse or not pc.pair_tile, "weights interleaved for a Laplace epilogue (pair_tile) cannot run without it"
        if not fuse:
            kw = dict(act=act, res1=res1, res2=res2, out2=out2, slope2=slope2, out_scale=out_scale, in_transform=in_transform,
                      in_slope=in_slope, epi=epi, gdn_x=gdn_x, engine=engine)
            if entropy["mode"] == "bitparm":
                z = View.alloc(out.H, out.W, out.C, out.device)
                conv(pc, srcs, z, **kw)
                entropy_standalone(dict(entropy, z_hat=out), z)
            else:
                conv(pc, srcs, out, **kw)
                entropy_standalone(entropy, out)
            return out
    d = CConv()
    d.n_src = len(srcs)
    for i, s in enumerate(srcs):
        assert s.C == pc.src_c[i], f"source {i} has {s.C} channels, packed for {pc.src_c[i]}"
        d.src[i] = s.c()
    d.weight, d.bias = pc.weight.data_ptr(), pc.bias.data_ptr()
    d.kh, d.kw, d.stride, d.pad = pc.kh, pc.kw, pc.stride, pc.pad
    d.cout, d.n_pad, d.cin_total = pc.cout, pc.n_pad, pc.cin_total
    d.in_transform, d.in_slope, d.epi = in_transform, in_slope, epi
    d.act = _lib.ACT_NONE if act is None else _lib.ACT_LRELU
    d.slope = 0.0 if act is None else float(act)
    d.out_scale = float(out_scale)
    d.pixel_shuffle = 1 if pc.pixel_shuffle else 0
    d.out = out.c()
    d.res1, d.res2, d.out2, d.gdn_x = _cv(res1), _cv(res2), _cv(out2), _cv(gdn_x)
    d.slope2 = float(slope2)
    if entropy is not None:                          # (fusable: checked above)
        g = lambda k: None if entropy.get(k) is None else entropy[k].data_ptr()
        d.ent_bits, d.ent_sym = g("bits"), g("sym")
        if entropy["mode"] in ("laplace", "fourpart"):
            d.epi = _lib.EPI_LAPLACE if entropy["mode"] == "laplace" else _lib.EPI_FOURPART
            d.ent_step = int(entropy.get("step") or 0)
            d.ent_y, d.ent_y_hat = entropy["y"].c(), entropy["y_hat"].c()
            d.ent_index, d.ent_thr = g("index"), g("thresholds")
            d.ent_n_thr = 0 if entropy.get("thresholds") is None else entropy["thresholds"].numel()
            d.ent_tile = pc.pair_tile
        else:
            d.epi = _lib.EPI_BITPARM
            d.ent_coef = entropy["coef"].data_ptr()
    lib = _lib.load()
    auto = engine is None
    engine = engine or _ENGINE
    if engine == "hs":
        engine = "h2"
    assert engine in ENGINES or engine == "head", engine
    if (engine == "h2" and auto and HEAD_KERNEL and pc.cout <= 4 and out.H * out.W >= HEAD_MIN_PIXELS
            and lib.lssvc_conv_head_supported(byref(d))):
        engine = "head"       # narrow heads (Cout 2..4): fp32 CUDA-core direct convolution, csrc/conv_head.cu
    if engine == "h2":
        ok = (all(s.C % 4 == 0 and s.pitch % 4 == 0 and (s.coff % 4 == 0) for s in srcs) and pc.kh * pc.kw <= 49
              and pc.stride in (1, 2) and srcs[0].H % pc.stride == 0 and srcs[0].W % pc.stride == 0)
        if not ok:
            global SIMT_DOWNGRADES
            SIMT_DOWNGRADES += 1
            engine = "simt"
    assert entropy is None or engine == "h2", engine
    if TRACE is not None:
        Ho, Wo = (out.H // 2, out.W // 2) if pc.pixel_shuffle else (out.H, out.W)
        TRACE.append({"name": TRACE_NAME, "engine": "hs" if engine == "h2" else engine, "k": pc.kh, "stride": pc.stride, "cin": pc.cin_total,
                      "src_c": list(pc.src_c), "cout": pc.cout, "Ho": Ho, "Wo": Wo, "ps": bool(pc.pixel_shuffle),
                      "extras": ("r" if res1 is not None else "") + ("s" if res2 is not None else "") + ("o" if out2 is not None else "")
                                + ("l" if in_transform == _lib.IN_LRELU else "") + ("g" if epi != _lib.EPI_PLAIN else "")
                                + ("e" if entropy is not None else ""),
                      "flops": 2.0 * Ho * Wo * pc.kh * pc.kw * sum(s.real for s in srcs) * pc.cout})
    if engine == "simt":
        _lib.check(lib.lssvc_conv_simt(byref(d), _stream()), "conv_simt")
    elif engine == "head":
        _lib.check(lib.lssvc_conv_head(byref(d), _stream()), "conv_head")
    else:
        wh, cin16, acc_scale = pc.weight_h2()
        d.precision = _lib.PREC_H2
        d.weight_h2, d.cin_pad16, d.acc_scale = wh.data_ptr(), cin16, acc_scale
        _lib.check(lib.lssvc_conv_hs(byref(d), _stream()), "conv_hs")
    return out


class PackedFfn:
    """Weights of one ConvFFN (two 1x1 convs, lssvc_modules.py:42-60) packed for lssvc_conv_ffn: split fp16 hi/lo of
    W * 2^shift in [chunk of 32 hidden channels][16-wide K slice][hi | lo][row][16] sub-tiles whose 32-byte rows carry
    the SWIZZLE_32B pattern (16-byte halves swapped when (row >> 2) & 1), so a plain bulk copy lands them UMMA-ready."""

    __slots__ = ("w1", "w2", "b1", "b2", "scale1", "scale2", "C", "hidden")

    @staticmethod
    def supported(C, hidden):
        return C % 16 == 0 and 16 <= C <= 64 and hidden % 64 == 0 and hidden >= 64 and 8 * C * hidden + 3 * 128 * C * 4 <= 224 * 1024

    def __init__(self, w1, b1, w2, b2, device):
        w1 = w1.detach().to(torch.float32).cpu().reshape(w1.shape[0], w1.shape[1])     # [hidden, C]
        w2 = w2.detach().to(torch.float32).cpu().reshape(w2.shape[0], w2.shape[1])     # [C, hidden]
        hidden, C = w1.shape
        assert w2.shape == (C, hidden) and PackedFfn.supported(C, hidden), (C, hidden)

        def scaled(w):
            m = float(w.abs().max())
            shift = 0 if m == 0.0 else 13 - int(math.floor(math.log2(m)))
            return w * (2.0 ** shift), 2.0 ** -shift

        def split(w):
            hi = w.to(torch.float16)
            return hi, (w - hi.to(torch.float32)).to(torch.float16)

        def swizzle(t):                     # t: [..., rows, 16] fp16 -> same shape, halves swapped on rows with (r >> 2) & 1
            rows = t.shape[-2]
            t = t.reshape(*t.shape[:-1], 2, 8)
            swap = ((torch.arange(rows) >> 2) & 1).bool()
            out = t.clone()
            out[..., swap, 0, :] = t[..., swap, 1, :]
            out[..., swap, 1, :] = t[..., swap, 0, :]
            return out.reshape(*out.shape[:-2], 16)

        w1s, self.scale1 = scaled(w1 * float(acc_comp(C // 16)))
        w2s, self.scale2 = scaled(w2 * float(acc_comp(hidden // 16)))
        n_chunks = hidden // 32
        # W1: rows = hidden channel of the chunk, K = input channel     -> [chunk][ks][hi|lo][32][16]
        h1, l1 = split(w1s)
        t1 = torch.stack([h1, l1], 0).reshape(2, n_chunks, 32, C // 16, 16).permute(1, 3, 0, 2, 4)
        # W2: rows = output channel, K = hidden channel of the chunk    -> [chunk][ks][hi|lo][C][16]
        h2, l2 = split(w2s)
        t2 = torch.stack([h2, l2], 0).reshape(2, C, n_chunks, 2, 16).permute(2, 3, 0, 1, 4)
        # rows of a sub-tile run over (hi|lo, row): the swizzle pattern follows the row index inside the sub-tile
        self.w1 = swizzle(t1.reshape(n_chunks, C // 16, 64, 16)).contiguous().to(device)
        self.w2 = swizzle(t2.reshape(n_chunks, 2, 2 * C, 16)).contiguous().to(device)
        self.b1 = b1.detach().to(torch.float32).contiguous().to(device)
        self.b2 = b2.detach().to(torch.float32).contiguous().to(device)
        self.C, self.hidden = C, hidden


def ffn(pf, x, out, slope1=0.1, slope2=0.1, res2=None):
    """out = x + lrelu(W2 . lrelu(W1 . x + b1, slope1) + b2, slope2) (+ res2), one fused kernel."""
    d = _lib.CFfn()
    d.inp, d.out, d.res2 = x.c(), out.c(), _cv(res2)
    d.hidden = pf.hidden
    d.w1, d.w2, d.b1, d.b2 = pf.w1.data_ptr(), pf.w2.data_ptr(), pf.b1.data_ptr(), pf.b2.data_ptr()
    d.scale1, d.scale2, d.slope1, d.slope2 = pf.scale1, pf.scale2, float(slope1), float(slope2)
    if TRACE is not None:
        TRACE.append({"name": TRACE_NAME, "engine": "ffn", "k": 1, "stride": 1, "cin": pf.C, "src_c": [pf.C], "cout": pf.C,
                      "hidden": pf.hidden, "Ho": out.H, "Wo": out.W, "ps": False, "flops": 4.0 * out.H * out.W * pf.C * pf.hidden})
    lib = _lib.load()
    _lib.check(lib.lssvc_conv_ffn(byref(d), _stream()), "conv_ffn")
    return out


class PackedPw:
    """Weights of one 1x1 conv (optionally with the depthwise 3x3 in front of it) packed for lssvc_conv_pw: split fp16
    hi/lo of W * 2^shift as [Cin/16][hi | lo][Cout][16] sub-tiles with the SWIZZLE_32B row pattern (see PackedFfn)."""

    __slots__ = ("w", "bias", "scale", "cin", "cout", "dw_w", "dw_b")

    @staticmethod
    def supported(cin, cout, dw=False):
        if not (cin % 16 == 0 and 16 <= cin <= 128 and cout % 16 == 0 and 16 <= cout <= 64 and 2 * cin + 4 * cout <= 512):
            return False
        # shared memory of conv_pw.cu with two input buffers (it takes three when they fit)
        slab_w = 32 if cin % 32 == 0 else 16
        rows = 18 * 10 if dw else 128
        in_bytes = (cin // slab_w) * round_up(rows * slab_w * 4, 1024)
        smem = round_up(round_up(4 * cin * cout, 1024) + (40 * cin if dw else 0), 1024) + 2 * in_bytes + 512 * cout + 1024
        return smem <= 226 * 1024

    def __init__(self, w, b, device, dw_w=None, dw_b=None):
        w = w.detach().to(torch.float32).cpu().reshape(w.shape[0], w.shape[1])          # [cout, cin]
        cout, cin = w.shape
        assert PackedPw.supported(cin, cout, dw_w is not None), (cin, cout)
        m = float(w.abs().max())
        shift = 0 if m == 0.0 else 13 - int(math.floor(math.log2(m)))
        ws = w * (2.0 ** shift) * float(acc_comp(cin // 16))
        hi = ws.to(torch.float16)
        lo = (ws - hi.to(torch.float32)).to(torch.float16)
        t = torch.stack([hi, lo], 0).reshape(2, cout, cin // 16, 16).permute(2, 0, 1, 3).reshape(cin // 16, 2 * cout, 2, 8)
        swap = ((torch.arange(2 * cout) >> 2) & 1).bool()
        sw = t.clone()
        sw[:, swap, 0, :] = t[:, swap, 1, :]
        sw[:, swap, 1, :] = t[:, swap, 0, :]
        self.w = sw.reshape(cin // 16, 2 * cout, 16).contiguous().to(device)
        self.bias = (torch.zeros(cout) if b is None else b.detach().to(torch.float32).cpu()).contiguous().to(device)
        self.scale = 2.0 ** -shift
        self.cin, self.cout = cin, cout
        if dw_w is not None:
            dw = dw_w.detach().to(torch.float32).cpu()
            assert dw.shape[0] == cin and dw.numel() == 9 * cin, dw.shape
            self.dw_w = dw.reshape(cin, 9).t().contiguous().to(device)                 # [9][cin], tap = 3*ky + kx
            self.dw_b = dw_b.detach().to(torch.float32).contiguous().to(device)
        else:
            self.dw_w = self.dw_b = None


def view_aligned(v):
    return v.pitch % 4 == 0 and v.coff % 4 == 0


def pw(pp, x, out, act=None, res1=None, res2=None, out_scale=1.0):
    """out = act(W . u + b) * out_scale (+ res1) (+ res2), u = x or dw3x3(x) + dw_b; one resident-weight kernel."""
    d = _lib.CPw()
    assert x.C == pp.cin and out.C == pp.cout, (x.C, out.C, pp.cin, pp.cout)
    d.inp, d.out, d.res1, d.res2 = x.c(), out.c(), _cv(res1), _cv(res2)
    d.w, d.bias = pp.w.data_ptr(), pp.bias.data_ptr()
    d.dw_weight = None if pp.dw_w is None else pp.dw_w.data_ptr()
    d.dw_bias = None if pp.dw_b is None else pp.dw_b.data_ptr()
    d.act = _lib.ACT_NONE if act is None else _lib.ACT_LRELU
    d.slope = 0.0 if act is None else float(act)
    d.out_scale, d.acc_scale = float(out_scale), pp.scale
    if TRACE is not None:
        TRACE.append({"name": TRACE_NAME, "engine": "pw", "k": 1, "stride": 1, "cin": pp.cin, "src_c": [pp.cin],
                      "cout": pp.cout, "Ho": out.H, "Wo": out.W, "ps": False, "dw": pp.dw_w is not None,
                      "flops": 2.0 * out.H * out.W * pp.cin * (pp.cout + (9 if pp.dw_w is not None else 0))})
    lib = _lib.load()
    _lib.check(lib.lssvc_conv_pw(byref(d), _stream()), "conv_pw")
    return out


def dwconv3x3(x, weight9c, bias, out):
    lib = _lib.load()
    _lib.check(lib.lssvc_dwconv3x3(byref(x.c()), _ptr(weight9c), _ptr(bias), byref(out.c()), _stream()), "dwconv3x3")
    return out


def deconv3x3_s2(x, weight, bias, out, act=None):
    lib = _lib.load()
    _lib.check(lib.lssvc_deconv3x3_s2(byref(x.c()), _ptr(weight), _ptr(bias), 0 if act is None else 1,
                                      0.0 if act is None else float(act), byref(out.c()), _stream()), "deconv3x3_s2")
    return out


def lrelu_copy(x, slope, out):
    lib = _lib.load()
    _lib.check(lib.lssvc_lrelu_copy(byref(x.c()), float(slope), byref(out.c()), _stream()), "lrelu_copy")
    return out


def softmax2_blend(logits, a, b, out):
    lib = _lib.load()
    _lib.check(lib.lssvc_softmax2_blend(byref(logits.c()), byref(a.c()), byref(b.c()), byref(out.c()), _stream()),
               "softmax2_blend")
    return out


def flow_warp(src, flow, out, flow_scale=1.0):
    lib = _lib.load()
    _lib.check(lib.lssvc_flow_warp(byref(src.c()), byref(flow.c()), float(flow_scale), byref(out.c()), _stream()),
               "flow_warp")
    return out


def bilinear_resize(x, out, scale=1.0):
    lib = _lib.load()
    _lib.check(lib.lssvc_bilinear_resize(byref(x.c()), float(scale), byref(out.c()), _stream()), "bilinear_resize")
    return out


def avgpool2(x, out):
    lib = _lib.load()
    _lib.check(lib.lssvc_avgpool2(byref(x.c()), byref(out.c()), _stream()), "avgpool2")
    return out


def maxpool2(x, out):
    lib = _lib.load()
    _lib.check(lib.lssvc_maxpool2(byref(x.c()), byref(out.c()), _stream()), "maxpool2")
    return out


def spynet_prep(im1, im2, flow_coarse, out8, flow_up):
    lib = _lib.load()
    fc = byref(flow_coarse.c()) if flow_coarse is not None else None
    _lib.check(lib.lssvc_spynet_prep(byref(im1.c()), byref(im2.c()), fc, byref(out8.c()), byref(flow_up.c()), _stream()),
               "spynet_prep")


def offset_diversity(x, off, flow, fusion_w, fusion_b, groups, offset_num, magnitude, out, planar=False):
    """planar: regroup x into a [G][H][W] float4 scratch first so that the per-(group, offset) gathers coalesce along x.
    Measured at 1080p (tools/mem_bench.py, gpurun_out/r2_mem2.log): 1.78 ms against 1.63 ms for the direct gather on smooth
    offsets, 1.86 against 1.92 on i.i.d. ones — coalescing the feature gather is NOT what bounds this kernel, so the direct
    gather (no 566 MB scratch, no regroup pass) stays the default and the planar variant is kept for A/B."""
    lib = _lib.load()
    scratch = torch.empty(groups * x.H * x.W * 4, dtype=torch.float32, device=x.device) if planar else None
    _lib.check(lib.lssvc_offset_diversity(byref(x.c()), byref(off.c()), byref(flow.c()), _ptr(fusion_w), _ptr(fusion_b),
                                          groups, offset_num, float(magnitude), byref(out.c()), _ptr(scratch), _stream()),
               "offset_diversity")
    return out


def _opt(v):
    return byref(v.c()) if v is not None else None


def laplace_quant(y, mean, scale, y_q=None, y_hat=None, bits=None, sym=None, index=None, thresholds=None):
    lib = _lib.load()
    n_thr = 0 if thresholds is None else thresholds.numel()
    _lib.check(lib.lssvc_laplace_quant(byref(y.c()), _opt(mean), byref(scale.c()), _opt(y_q), _opt(y_hat), _ptr(bits),
                                       _ptr(sym), _ptr(index), _ptr(thresholds), n_thr, _stream()), "laplace_quant")


def four_part_step(y, params8, step, y_hat, y_q=None, scales_hat=None, bits=None, sym=None, index=None, thresholds=None):
    lib = _lib.load()
    n_thr = 0 if thresholds is None else thresholds.numel()
    _lib.check(lib.lssvc_four_part_step(byref(y.c()), byref(params8.c()), step, byref(y_hat.c()), _opt(y_q),
                                        _opt(scales_hat), _ptr(bits), _ptr(sym), _ptr(index), _ptr(thresholds), n_thr,
                                        _stream()), "four_part_step")


def four_part_index(params8, step, index, thresholds):
    lib = _lib.load()
    _lib.check(lib.lssvc_four_part_index(byref(params8.c()), step, _ptr(index), _ptr(thresholds), thresholds.numel(),
                                         _stream()), "four_part_index")


def four_part_dec_step(sym, params8, step, y_hat):
    lib = _lib.load()
    _lib.check(lib.lssvc_four_part_dec_step(_ptr(sym), byref(params8.c()), step, byref(y_hat.c()), _stream()),
               "four_part_dec_step")


def scale_index(scale, index, thresholds):
    lib = _lib.load()
    _lib.check(lib.lssvc_scale_index(byref(scale.c()), _ptr(index), _ptr(thresholds), thresholds.numel(), _stream()),
               "scale_index")


def symbols_to_view(sym, add, out):
    lib = _lib.load()
    _lib.check(lib.lssvc_symbols_to_view(_ptr(sym), _opt(add), byref(out.c()), _stream()), "symbols_to_view")
    return out


def gaussian_quant(y, mean, scale, y_hat=None, bits=None, sym=None, index=None, thresholds=None):
    lib = _lib.load()
    n_thr = 0 if thresholds is None else thresholds.numel()
    _lib.check(lib.lssvc_gaussian_quant(byref(y.c()), byref(mean.c()), byref(scale.c()), _opt(y_hat), _ptr(bits),
                                        _ptr(sym), _ptr(index), _ptr(thresholds), n_thr, _stream()), "gaussian_quant")


def bitparm_quant(z, coef, z_hat=None, bits=None, sym=None):
    lib = _lib.load()
    _lib.check(lib.lssvc_bitparm_quant(byref(z.c()), _ptr(coef), _opt(z_hat), _ptr(bits), _ptr(sym), _stream()),
               "bitparm_quant")


def eb_quant(z, coef, z_hat=None, bits=None, sym=None):
    lib = _lib.load()
    _lib.check(lib.lssvc_eb_quant(byref(z.c()), _ptr(coef), _opt(z_hat), _ptr(bits), _ptr(sym), _stream()), "eb_quant")


def range_flag_fetch(dst):
    """dst: 1-element float64 device tensor <- 1.0 if a split-fp16 operand left the fp16 range since the last fetch."""
    lib = _lib.load()
    _lib.check(lib.lssvc_range_flag_fetch(_ptr(dst), _stream()), "range_flag_fetch")


def sse(a, b, out):
    lib = _lib.load()
    _lib.check(lib.lssvc_sse(byref(a.c()), byref(b.c()), _ptr(out), _stream()), "sse")
