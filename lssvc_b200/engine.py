"""Shared execution helpers of the host-side models: weight (re)packing cache and the fused building blocks
(conv + epilogue, ResBlock, DepthConvBlock, GDN, SpyNet, 3-scale extractor / fusion) expressed on the kernels.

A model is a nets.ParamBag (reference-layout tensors) + this mixin.  Activations are `ops.View`s (NHWC fp32)."""
import os

import torch
import torch.nn.functional as F

from . import _lib, nets, ops
from .ops import View


class Engine(nets.ParamBag):
    def __init__(self, spec, model_tag, seed=None):
        super().__init__(spec, seed=seed, gains=nets.model_gains(model_tag))
        self._packs = {}
        self._graphs = {}
        self._graph_pools = {}
        self._tables = None
        # whole-frame CUDA graphs (models.py), opt in with LSSVC_CUDA_GRAPH=1: measured equal to eager launches on one
        # B200 (the GPU never waits for the host: tools/graph_ab.py), useful when the host is the bottleneck
        self.use_graphs = os.environ.get("LSSVC_CUDA_GRAPH", "0") == "1"
        self.fuse_ffn = os.environ.get("LSSVC_FUSE_FFN", "1") != "0"
        self.fuse_pw = os.environ.get("LSSVC_FUSE_PW", "1") != "0"
        self.lazy_act = os.environ.get("LSSVC_LAZY_ACT", "1") != "0"
        self.shape_hr = (256, 256)
        self.scale_factor = 2.0
        self.pad_size = (0, 0, 0, 0)
        self.eval()

    # ---- reference API shared by IntraSS / LSSVC (IntraSS.py:229-232, LSSVC_net.py:266-269) -----------------
    def set_scale_information(self, scale, shape_hr, pad_size):
        new = (scale, tuple(shape_hr), tuple(pad_size))
        if new != (self.scale_factor, self.shape_hr, self.pad_size):     # test.py:212-213 calls this before every frame
            self._graphs = {}
        self.scale_factor, self.shape_hr, self.pad_size = new

    def depad(self, v, p=1):
        """get_depadded_feature (LSSVC_net.py:271-282, IntraSS.py:124-135): F.pad of a base-layer tensor by pad_size / p (left, right,
        top, bottom; negative = crop, the case the inter-layer padding produces) before it is resampled for the enhancement
        layer.  test.py:212-213 always passes zeros, for which this is the identity; otherwise one strided copy (torch plumbing)."""
        if v is None:
            return None
        pads = [int(a / p) for a in self.pad_size]
        if not any(pads):
            return v
        t = F.pad(v.exact().as_tensor().permute(2, 0, 1), pads, mode="constant", value=0).permute(1, 2, 0)
        out = self.new(t.shape[0], t.shape[1], v.real)
        out.exact().as_tensor().copy_(t)
        return out

    # ---- weight cache ---------------------------------------------------------------------------------------
    def _invalidate(self):
        self._packs = {}
        self._graphs = {}
        self._tables = None     # CDF tables follow the bit-estimator / bottleneck parameters: update() rebuilds them

    def _apply(self, fn, *a, **k):
        self._invalidate()
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, state_dict, strict=True):
        self._invalidate()
        return super().load_state_dict(state_dict, strict=strict)

    @property
    def device(self):
        return next(self.parameters()).device

    def _require_cuda(self):
        if self.device.type != "cuda" and not _lib.DRY_RUN:
            raise _lib.LssvcError("lssvc_b200 models run on a CUDA (sm_100a) device only; call .to('cuda') first — "
                                  "there is no CPU fallback")

    def pack(self, name, srcs, stride, ps, transposed, pad, exact_in=False, pair_tile=0):
        key = (name, tuple((v.real, v.C) for v in srcs), stride, ps, transposed, pad, exact_in, pair_tile)
        pc = self._packs.get(key)
        if pc is None:
            pc = ops.PackedConv(self.tensor(name + ".weight"), self.tensor(name + ".bias"), stride=stride, pad=pad,
                                src_channels=[(v.real, v.C) for v in srcs], pixel_shuffle=ps, transposed=transposed,
                                device=self.device, exact_in=exact_in, pair_tile=pair_tile)
            self._packs[key] = pc
        return pc

    def cached(self, key, builder):
        t = self._packs.get(key)
        if t is None:
            t = builder()
            self._packs[key] = t
        return t

    # ---- primitives -----------------------------------------------------------------------------------------
    def new(self, H, W, C):
        return View.alloc_padded(H, W, C, self.device)

    def conv(self, name, srcs, stride=1, act=None, ps=False, transposed=False, pad=None, res1=None, res2=None,
             out=None, act_copy=None, act_copy_out=None, out_scale=1.0, engine=None, in_lrelu=None, exact_in=False,
             entropy=None):
        """conv (+PixelShuffle) with fused epilogue.  Returns the output view, or (out, lrelu(out, act_copy)).
        exact_in: the source holds quantised symbols (ops.PackedConv).  entropy: entropy epilogue (ops.conv)."""
        if isinstance(srcs, View):
            srcs = [srcs]
        # un-materialised activations (ops.LazyAct): one common slope -> LeakyReLU in the conv's operand path, else write them out
        lazy = [s.slope if isinstance(s, ops.LazyAct) else None for s in srcs]
        if any(l is not None for l in lazy):
            if len(set(lazy)) == 1 and in_lrelu is None:
                in_lrelu, srcs = lazy[0], [s.base for s in srcs]
            else:
                srcs = [self.lrelu(s.base, s.slope) if isinstance(s, ops.LazyAct) else s for s in srcs]
        pair_tile = 0
        if (entropy is not None and entropy["mode"] in ("laplace", "fourpart") and out_scale == 1.0 and not ps and res2 is None
                and act_copy is None and entropy.get("y_q") is None and entropy.get("s_hat") is None):
            # 2C parameter channels beyond one channel tile: interleave (scale, mean) per tile at pack time (ops.PackedConv)
            pair_tile = ops.laplace_pair_tile(2 * entropy["y"].C, engine)
        pc = self.pack(name, srcs, stride, ps, transposed, pad, exact_in, pair_tile)
        Hi, Wi = srcs[0].H, srcs[0].W
        Ho = (Hi + 2 * pc.pad - pc.kh) // stride + 1
        Wo = (Wi + 2 * pc.pad - pc.kw) // stride + 1
        f = 2 if ps else 1
        C = pc.cout // 4 if ps else pc.cout
        if out is None:
            out = self.new(Ho * f, Wo * f, C)
        out2 = None
        lazy_copy = act_copy is not None and act_copy_out is None and self.lazy_act and ops.default_engine() == "h2"
        if act_copy is not None and not lazy_copy:
            out2 = act_copy_out if act_copy_out is not None else self.new(Ho * f, Wo * f, C)
        ex = lambda v: None if v is None else v.exact()
        ops.TRACE_NAME = name
        kw = {} if in_lrelu is None else {"in_transform": _lib.IN_LRELU, "in_slope": float(in_lrelu)}
        if entropy is not None:
            kw["entropy"] = entropy
        ops.conv(pc, srcs, out.exact(), act=act, res1=ex(res1), res2=ex(res2), out2=ex(out2),
                 slope2=0.0 if act_copy is None else act_copy, out_scale=out_scale, engine=engine, **kw)
        if lazy_copy:
            return out, ops.LazyAct(out, act_copy)      # the consumer conv applies the activation on the fly
        return (out, out2) if act_copy is not None else out

    def lrelu(self, x, slope, out=None):
        out = out if out is not None else self.new(x.H, x.W, x.real)
        ops.lrelu_copy(x.exact(), slope, out.exact())
        return out

    def copy(self, x, out):
        return self.lrelu(x, 1.0, out)

    def gdn(self, name, x, inverse=False, res1=None, intra=False, out=None):
        """GDN / IGDN: x * (beta + gamma . x^2)^(-/+ 1/2) (+ res1).
        intra: gdn.py:29-44 + others.py:43-67; inter: video_net_component.py:83-105."""
        def build():
            beta, gamma = self.tensor(name + ".beta").detach().float().cpu(), self.tensor(name + ".gamma").detach().float().cpu()
            if intra:
                ped = 2.0 ** -36
                beta = torch.max(beta, torch.tensor((1e-6 + ped) ** 0.5)) ** 2 - ped
                gamma = torch.max(gamma, torch.tensor(ped ** 0.5)) ** 2 - ped
            else:
                ped = (2.0 ** -18) ** 2
                beta = torch.max(beta, torch.ones_like(beta) * ((1e-6 + ped) ** 0.5)) ** 2 - ped
                gamma = torch.max(gamma, torch.ones_like(gamma) * (2.0 ** -18)) ** 2 - ped
            C = beta.numel()
            return ops.PackedConv(gamma.view(C, C, 1, 1), beta, pad=0, src_channels=[(C, x.C)], device=self.device,
                                  coherent=os.environ.get("LSSVC_GDN_COHERENT", "1") != "0")
        pc = self.cached(("gdn", name, x.C), build)
        out = out if out is not None else self.new(x.H, x.W, x.real)
        # split-fp16 tensor-core engine: x^2 is formed and split in the kernel's operand path; the other engines keep
        # the norm pool on the fp32 CUDA cores (ops.conv falls back for input transforms / GDN epilogues)
        ops.TRACE_NAME = name
        ops.conv(pc, [x], out.exact(), in_transform=_lib.IN_SQUARE, epi=_lib.EPI_IGDN if inverse else _lib.EPI_GDN,
                 gdn_x=x.exact(), res1=None if res1 is None else res1.exact())
        return out

    def dwconv(self, name, x):
        def build():
            w = self.tensor(name + ".weight").detach().float()
            C = w.shape[0]
            return (w.reshape(C, 9).t().contiguous().to(self.device), self.tensor(name + ".bias").detach().float().to(self.device))
        w, b = self.cached(("dw", name), build)
        out = self.new(x.H, x.W, x.real)
        ops.dwconv3x3(x.exact(), w, b, out.exact())
        return out

    def pointwise(self, name, x, act=None, res1=None, res2=None, out=None, dw=None):
        """1x1 conv `name` (+ bias, LeakyReLU, residuals), with the depthwise 3x3 `dw` applied to x first when given.
        Runs on the resident-weight kernel (csrc/conv_pw.cu) when the shapes allow, else dwconv + the general conv."""
        w = self.tensor(name + ".weight")
        cout, cin = w.shape[0], w.shape[1]
        # a plain 1x1 runs faster on the general kernel since its lean epilogue (0.19 vs 0.30 ms for 64->64 at 1152x1920):
        # the resident-weight kernel is kept for the fused depthwise front end
        ok = (self.fuse_pw and dw is not None and ops.default_engine() == "h2" and ops.PackedPw.supported(cin, cout, True) and x.real == cin
              and ops.view_aligned(x) and all(r is None or (r.real == cout and ops.view_aligned(r)) for r in (res1, res2, out)))
        if not ok:
            if dw is not None:
                x = self.dwconv(dw, x)
            return self.conv(name, x, act=act, pad=0, res1=res1, res2=res2, out=out)
        pp = self.cached(("pw", name, dw), lambda: ops.PackedPw(
            w, self.tensor(name + ".bias"), self.device,
            dw_w=None if dw is None else self.tensor(dw + ".weight"), dw_b=None if dw is None else self.tensor(dw + ".bias")))
        out = out if out is not None else self.new(x.H, x.W, cout)
        ex = lambda v: None if v is None else v.exact()
        ops.TRACE_NAME = name
        ops.pw(pp, x.exact(), out.exact(), act=act, res1=ex(res1), res2=ex(res2))
        return out

    def deconv_s2(self, name, x, act=None, exact_in=False):
        """nn.ConvTranspose2d(3, stride=2, padding=1, output_padding=1).

        On the tensor-core engine it runs as its sub-pixel decomposition: out[2y+i, 2x+j] only involves in[y+a, x+b]
        with a, b in {0, 1} (ky = 1 for (i,a)=(0,0), 2 for (1,0), 0 for (1,1); same in x), i.e. a stride-1 conv with
        4*cout output channels (taps r = a+1, s = b+1 of a 3x3 / pad 1 kernel, the others zero) + PixelShuffle(2)."""
        if ops.default_engine() == "h2":
            def build_ps():
                w = self.tensor(name + ".weight").detach().float().cpu()       # [cin, cout, 3, 3]
                b = self.tensor(name + ".bias").detach().float().cpu()
                cin, cout = w.shape[0], w.shape[1]
                w3 = torch.zeros(4 * cout, cin, 3, 3)
                kmap = {(0, 0): 1, (1, 0): 2, (1, 1): 0}
                for (i, a), ky in kmap.items():
                    for (j, bb), kx in kmap.items():
                        w3[2 * i + j::4, :, a + 1, bb + 1] = w[:, :, ky, kx].t()
                return ops.PackedConv(w3, b.repeat_interleave(4), pad=1, src_channels=[(x.real, x.C)], pixel_shuffle=True,
                                      device=self.device, exact_in=exact_in)
            pc = self.cached(("deconv_ps", name, x.real, x.C, exact_in), build_ps)
            out = self.new(x.H * 2, x.W * 2, pc.cout // 4)
            ops.TRACE_NAME = name
            ops.conv(pc, [x], out.exact(), act=act)
            return out

        def build():
            w = self.tensor(name + ".weight").detach().float()          # [cin, cout, 3, 3]
            cin, cout = w.shape[0], w.shape[1]
            return (w.permute(2, 3, 0, 1).reshape(9, cin, cout).contiguous().to(self.device),
                    self.tensor(name + ".bias").detach().float().to(self.device), cout)
        w, b, cout = self.cached(("deconv", name), build)
        out = self.new(x.H * 2, x.W * 2, cout)
        ops.deconv3x3_s2(x.exact(), w, b, out.exact(), act=act)
        return out

    def resize(self, x, H, W, scale=1.0):
        out = self.new(H, W, x.real)
        ops.bilinear_resize(x.exact(), out.exact(), scale=scale)
        return out

    def warp(self, src, flow, out=None):
        out = out if out is not None else self.new(src.H, src.W, src.real)
        ops.flow_warp(src.exact(), flow, out.exact())
        return out

    def image_view(self, t):
        """[1,3,H,W] fp32 tensor -> NHWC view with 8 channels (5 zero) readable by every kernel."""
        if t.device != self.device:
            t = t.to(self.device)
        return View.from_nchw(t.float(), C_view=8)

    def feature_view(self, t):
        if t.device != self.device:
            t = t.to(self.device)
        C = t.shape[1]
        shared = View.from_channels_last(t)       # a feature map this package handed out (View.to_nchw_shared): no pass
        if shared is not None:
            return shared
        return View.from_nchw(t.float(), C_view=ops.round_up(C, 8))

    # ---- blocks -----------------------------------------------------------------------------------------------
    def res_block(self, name, x, x_act=None, slope=0.01, start_from_relu=True, end_with_relu=False, res2=None, out=None,
                  act_copy=None, act_copy_out=None):
        """ResBlock (video_net_component.py:170-188, layers.py:229-255): x + last(conv2(lrelu(conv1(first(x))))).
        x_act: lrelu(x, slope) if the producer already wrote it (saves a pass)."""
        inp, in_lrelu = x, None
        if start_from_relu:
            if x_act is not None:
                inp = x_act
            elif ops.default_engine() == "h2":
                in_lrelu = slope        # LeakyReLU of the block input applied in the conv's operand path
            else:
                inp = self.lrelu(x, slope)
        t = self.conv(name + ".conv1", inp, act=slope, in_lrelu=in_lrelu)
        return self.conv(name + ".conv2", t, act=slope if end_with_relu else None, res1=x, res2=res2, out=out,
                         act_copy=act_copy, act_copy_out=act_copy_out)

    def depth_conv_block(self, name, x, res2=None, out=None, entropy=None):
        """DepthConvBlock (lssvc_modules.py:15-72): DepthConv (1x1, lrelu .01, dw3x3, 1x1, + identity/adaptor)
        then ConvFFN (x + lrelu(1x1(lrelu(1x1 x, .1)), .1)).
        entropy: the block emits entropy parameters — its last 1x1 codes the latent in its epilogue (ops.conv `entropy`)."""
        dc, ffn = name + ".block.0", name + ".block.1"
        has_adaptor = (dc + ".adaptor.weight") in self._spec
        t = self.pointwise(dc + ".conv1.0", x, act=0.01)
        identity = self.pointwise(dc + ".adaptor", x) if has_adaptor else x
        o = self.pointwise(dc + ".conv2", t, res1=identity, dw=dc + ".depth_conv")
        w1 = self.tensor(ffn + ".conv.0.weight")
        hidden, C = w1.shape[0], w1.shape[1]
        if (self.fuse_ffn and ops.default_engine() == "h2" and ops.PackedFfn.supported(C, hidden) and o.C == C
                and o.pitch % 4 == 0 and o.coff % 4 == 0):
            # fused ConvFFN: the 4x-wide intermediate stays in tensor memory (csrc/conv_ffn.cu)
            pf = self.cached(("ffn", ffn), lambda: ops.PackedFfn(w1, self.tensor(ffn + ".conv.0.bias"),
                                                                 self.tensor(ffn + ".conv.2.weight"),
                                                                 self.tensor(ffn + ".conv.2.bias"), self.device))
            out = out if out is not None else self.new(o.H, o.W, C)
            ox = out.exact()
            if ox.pitch % 4 == 0 and ox.coff % 4 == 0 and (res2 is None or (res2.pitch % 4 == 0 and res2.coff % 4 == 0)):
                ops.ffn(pf, o.exact(), ox, res2=None if res2 is None else res2.exact())
                if entropy is not None:
                    ops.entropy_standalone(entropy, ox)
                return out
        f = self.conv(ffn + ".conv.0", o, act=0.1, pad=0)
        return self.conv(ffn + ".conv.2", f, act=0.1, pad=0, res1=o, res2=res2, out=out, entropy=entropy)

    def extractor3(self, name, x):
        """conv1/res_block1, conv2 s2/res_block2, conv3 s2/res_block3
        (layers.py:288-308, lssvc_modules.py:157-200, dmc_net.py:11-31)."""
        outs = []
        cur = x
        for i, stride in ((1, 1), (2, 2), (3, 2)):
            t, t_act = self.conv(f"{name}.conv{i}", cur, stride=stride, act_copy=0.01)
            cur = self.res_block(f"{name}.res_block{i}", t, x_act=t_act)
            outs.append(cur)
        return outs

    def fusion3(self, name, c1, c2, c3, out2_slopes=None, outs=None):
        """MultiScaleContextFusion / MultiScaleTextureFusion (lssvc_modules.py:203-232, dmc_net.py:34-62,
        layers.py:311-339).  Returns (c1 + c1_out, c2 + c2_out, c3 + c3_out); outs: optional destination views."""
        outs = outs or (None, None, None)
        t, ta = self.conv(name + ".conv3_up.0", c3, ps=True, act_copy=0.01)
        c3_up = self.res_block(name + ".res_block3_up", t, x_act=ta)
        t, ta = self.conv(name + ".conv3_out", c3, act_copy=0.01)
        o3 = self.res_block(name + ".res_block3_out", t, x_act=ta, res2=c3, out=outs[2])
        t, ta = self.conv(name + ".conv2_up.0", [c3_up, c2], ps=True, act_copy=0.01)
        c2_up = self.res_block(name + ".res_block2_up", t, x_act=ta)
        t, ta = self.conv(name + ".conv2_out", [c3_up, c2], act_copy=0.01)
        o2 = self.res_block(name + ".res_block2_out", t, x_act=ta, res2=c2, out=outs[1])
        t, ta = self.conv(name + ".conv1_out", [c2_up, c1], act_copy=0.01)
        o1 = self.res_block(name + ".res_block1_out", t, x_act=ta, res2=c1, out=outs[0])
        return o1, o2, o3

    def spynet(self, name, im1, im2):
        """ME_Spynet / ME_Spynet_DCVC (video_net_component.py:222-248, 300-326): 4-level coarse-to-fine flow."""
        p1, p2 = [im1], [im2]
        for _ in range(3):
            for p in (p1, p2):
                src = p[-1]
                dst = View.alloc(src.H // 2, src.W // 2, src.C, self.device, pitch=src.pitch)
                dst.real = src.real
                ops.avgpool2(src, dst)
                p.append(dst)
        flow = None
        for level in range(4):
            a, b = p1[3 - level], p2[3 - level]
            x8 = View.alloc(a.H, a.W, 8, self.device, pitch=8)
            flow_up = self.new(a.H, a.W, 2)
            ops.spynet_prep(a, b, flow, x8, flow_up)
            m = f"{name}.moduleBasic.{level}"
            t = x8
            for k in range(1, 5):
                t = self.conv(f"{m}.conv{k}", t, act=0.0, pad=3)
            flow = self.conv(f"{m}.conv5", t, pad=3, res1=flow_up)
        return flow

    def seq2(self, name, x, out=None, **kw):
        """conv, LeakyReLU, conv."""
        return self.conv(name + ".2", self.conv(name + ".0", x, act=0.01), out=out, **kw)
