"""Parameter layout of the reference models (state_dict keys and shapes), declared independently of any
nn.Conv2d tree: the forward pass is hand-written CUDA, so a model here is just a named bag of tensors.

`ParamBag` is an nn.Module whose state_dict()/load_state_dict() use exactly the reference key layout
(IntraSS 334 tensors, LSSVC_extend 926 tensors), so checkpoints are interchangeable.

Default initialisation is a deterministic, well-conditioned synthetic recipe: the
reference's own default init makes the P-frame chain diverge to NaN within 5 frames, which makes parity
checks vacuous.  Gains below keep the recurrent feature path contractive while the latents stay
non-degenerate (non-zero symbols, scales spread over many CDF rows)."""
import json
import math
import os
from collections import OrderedDict

import torch
import torch.nn as nn

G1, G2, G4, G8, G16 = 48, 64, 96, 96, 128   # lssvc_modules.py:8-12


class Spec(OrderedDict):
    """name -> dict(shape=..., kind=..., buffer=bool, **init hints)"""

    def add(self, name, shape, kind, buffer=False, **kw):
        assert name not in self, name
        self[name] = dict(shape=tuple(shape), kind=kind, buffer=buffer, **kw)

    # ---- primitive layers -------------------------------------------------------------------------
    def conv(self, name, cin, cout, k, gain=1.0, bias=0.0, groups=1, in_boost=None):
        """in_boost=(n, f): the first n input channels (the image of a cat([x, context])) are weighted f times
        stronger, so that the current frame, not the recurrent context, drives the latents."""
        self.add(name + ".weight", (cout, cin // groups, k, k), "conv_w", gain=gain, in_boost=in_boost)
        self.add(name + ".bias", (cout,), "bias", value=bias)

    def deconv(self, name, cin, cout, k, gain=1.0):
        self.add(name + ".weight", (cin, cout, k, k), "deconv_w", gain=gain)
        self.add(name + ".bias", (cout,), "bias", value=0.0)

    def subpel(self, name, cin, cout, k, gain=1.0):
        self.conv(name + ".0", cin, cout * 4, k, gain=gain)

    def gdn_inter(self, name, ch):   # video_net_component.py:52-81
        self.add(name + ".beta", (ch,), "gdn_beta_inter")
        self.add(name + ".gamma", (ch, ch), "gdn_gamma_inter")

    def gdn_intra(self, name, ch):   # gdn.py:8-27 + others.py:43-61
        ped = 2.0 ** -36
        self.add(name + ".beta", (ch,), "gdn_beta_intra")
        self.add(name + ".gamma", (ch, ch), "gdn_gamma_intra")
        self.add(name + ".beta_reparam.pedestal", (1,), "const", buffer=True, value=ped)
        self.add(name + ".beta_reparam.lower_bound.bound", (1,), "const", buffer=True, value=(1e-6 + ped) ** 0.5)
        self.add(name + ".gamma_reparam.pedestal", (1,), "const", buffer=True, value=ped)
        self.add(name + ".gamma_reparam.lower_bound.bound", (1,), "const", buffer=True, value=ped ** 0.5)

    def res_block(self, name, ch, bottleneck=False, gain=0.3):
        mid = ch // 2 if bottleneck else ch
        self.conv(name + ".conv1", ch, mid, 3, gain=1.4)
        self.conv(name + ".conv2", mid, ch, 3, gain=gain)

    def depth_conv_block(self, name, cin, cout, gain=0.3):   # lssvc_modules.py:15-72
        dc, ffn = name + ".block.0", name + ".block.1"
        self.conv(dc + ".conv1.0", cin, cin, 1, gain=1.4)
        self.conv(dc + ".depth_conv", cin, cin, 3, groups=cin, gain=1.0)
        self.conv(dc + ".conv2", cin, cout, 1, gain=gain if cin == cout else 0.7)
        if cin != cout:
            self.conv(dc + ".adaptor", cin, cout, 1, gain=0.7)
        internal = max(min(cout * 4, 1024), cout * 2)
        self.conv(ffn + ".conv.0", cout, internal, 1, gain=1.4)
        self.conv(ffn + ".conv.2", internal, cout, 1, gain=gain)

    def bit_estimator(self, name, ch):   # video_entropy_models.py:110-166
        for i in (1, 2, 3, 4):
            self.add(f"{name}.f{i}.h", (1, ch, 1, 1), "bitparm_h")
            self.add(f"{name}.f{i}.b", (1, ch, 1, 1), "bitparm_b")
            if i < 4:
                self.add(f"{name}.f{i}.a", (1, ch, 1, 1), "bitparm_a")

    def entropy_bottleneck(self, name, ch):   # img_entropy_models.py:385-430
        filters = (1, 3, 3, 3, 3, 1)
        scale = 10 ** (1 / 5)
        self.add(name + ".quantiles", (ch, 1, 3), "eb_quantiles")
        for buf in ("_offset", "_quantized_cdf", "_cdf_length"):
            self.add(f"{name}.{buf}", (0,), "empty_int", buffer=True)
        self.add(name + ".target", (3,), "eb_target", buffer=True)
        self.add(name + ".likelihood_lower_bound.bound", (1,), "const", buffer=True, value=1e-9)
        for i in range(5):
            self.add(f"{name}._biases.{i}", (ch, filters[i + 1], 1), "eb_bias")
        for i in range(4):
            self.add(f"{name}._factors.{i}", (ch, filters[i + 1], 1), "zeros")
        for i in range(5):
            self.add(f"{name}._matrices.{i}", (ch, filters[i + 1], filters[i]), "const",
                     value=math.log(math.expm1(1 / scale / filters[i + 1])))

    def gaussian_conditional(self, name):   # img_entropy_models.py:586-607
        for buf in ("_offset", "_quantized_cdf", "_cdf_length"):
            self.add(f"{name}.{buf}", (0,), "empty_int", buffer=True)
        self.add(name + ".scale_bound", (1,), "const", buffer=True, value=0.11)
        self.add(name + ".likelihood_lower_bound.bound", (1,), "const", buffer=True, value=1e-9)
        self.add(name + ".lower_bound_scale.bound", (1,), "const", buffer=True, value=0.11)

    # ---- composite blocks -------------------------------------------------------------------------
    def extractor3(self, name, c0, c1, c2, c3):
        self.conv(name + ".conv1", c0, c1, 3)
        self.res_block(name + ".res_block1", c1)
        self.conv(name + ".conv2", c1, c2, 3)
        self.res_block(name + ".res_block2", c2)
        self.conv(name + ".conv3", c2, c3, 3)
        self.res_block(name + ".res_block3", c3)

    def fusion3(self, name, c1, c2, c3):
        self.subpel(name + ".conv3_up", c3, c2, 3)
        self.res_block(name + ".res_block3_up", c2)
        self.conv(name + ".conv3_out", c3, c3, 3, gain=0.5)
        self.res_block(name + ".res_block3_out", c3)
        self.subpel(name + ".conv2_up", c2 * 2, c1, 3)
        self.res_block(name + ".res_block2_up", c1)
        self.conv(name + ".conv2_out", c2 * 2, c2, 3, gain=0.5)
        self.res_block(name + ".res_block2_out", c2)
        self.conv(name + ".conv1_out", c1 * 2, c1, 3, gain=0.5)
        self.res_block(name + ".res_block1_out", c1)

    def spynet(self, name):   # video_net_component.py:191-210
        for level in range(4):
            m = f"{name}.moduleBasic.{level}"
            self.conv(m + ".conv1", 8, 32, 7, gain=1.4)
            self.conv(m + ".conv2", 32, 64, 7, gain=1.4)
            self.conv(m + ".conv3", 64, 32, 7, gain=1.4)
            self.conv(m + ".conv4", 32, 16, 7, gain=1.4)
            self.conv(m + ".conv5", 16, 2, 7, gain=0.5)

    def prior_encoder(self, name, cin, ch, head_gain):
        self.conv(name + ".0", cin, ch, 3, gain=1.4)
        self.conv(name + ".2", ch, ch, 3, gain=1.4)
        self.conv(name + ".4", ch, ch, 3, gain=head_gain)


# -----------------------------------------------------------------------------------------------------------
# model specs
# -----------------------------------------------------------------------------------------------------------
Y_GAIN = 2.5       # heads that produce latents: spread |y - mu| over a few quantisation bins
Z_GAIN = 2.0
X_BOOST = 15.0    # weight of the image channels relative to the context channels in cat([x, ctx]) convs
SCALE_BIAS = 0.7   # bias of the scale half of parameter heads (scale rows around the middle of the tables)


def _params_head(s, name, cin, cout, k=3):
    """Last conv of a (scales | means) head: scale half biased positive, mean half small."""
    s.conv(name, cin, cout, k, gain=0.5)
    s[name + ".bias"]["kind"] = "scale_mean_bias"


def intra_noar_spec(s, p, N):
    """IntraNoAR   priors.py:112-162."""
    s.entropy_bottleneck(p + "entropy_bottleneck", N)
    cin = 3
    for i in range(0, 6, 2):
        b = f"{p}g_a.{i}"
        s.conv(b + ".conv1", cin, N, 3, gain=1.4)
        s.conv(b + ".conv2", N, N, 3)
        s.gdn_intra(b + ".gdn", N)
        s.conv(b + ".downsample", cin, N, 1, gain=0.7)
        b = f"{p}g_a.{i + 1}"
        s.conv(b + ".conv1", N, N, 3, gain=1.4)
        s.conv(b + ".conv2", N, N, 3, gain=0.4)
        cin = N
    s.conv(p + "g_a.6", N, N, 3, gain=Y_GAIN)
    for i, (ci, co) in zip((0, 2, 4, 6, 8), ((N, N),) * 5):
        s.conv(f"{p}h_a.{i}", ci, co, 3, gain=Z_GAIN if i == 8 else 1.4)
    s.conv(p + "h_s.0", N, N, 3, gain=1.4)
    s.subpel(p + "h_s.2", N, N, 3, gain=1.4)
    s.conv(p + "h_s.4", N, N * 3 // 2, 3, gain=1.4)
    s.subpel(p + "h_s.6", N * 3 // 2, N * 3 // 2, 3, gain=1.4)
    _params_head(s, p + "h_s.8", N * 3 // 2, N * 2)
    for i in range(0, 7, 2):
        b = f"{p}g_s.{i}"
        s.conv(b + ".conv1", N, N, 3, gain=1.4)
        s.conv(b + ".conv2", N, N, 3, gain=0.4)
        if i < 6:
            b = f"{p}g_s.{i + 1}"
            s.subpel(b + ".subpel_conv", N, N, 3, gain=1.4)
            s.conv(b + ".conv", N, N, 3)
            s.gdn_intra(b + ".igdn", N)
            s.subpel(b + ".upsample", N, N, 3, gain=0.7)
    s.subpel(p + "g_s.7", N, 3, 3, gain=0.15)
    s[p + "g_s.7.0.bias"]["value"] = 0.5
    s.gaussian_conditional(p + "gaussian_conditional")


def res_encoder_gdn_spec(s, p, gdn, N=64, M=96):
    """ResEncoder with GDN: layers.py:342-355, dmc_net.py:65-78."""
    s.conv(p + ".conv1", N + 3, N, 3, in_boost=(3, X_BOOST))
    gdn(p + ".gdn1", N)
    s.res_block(p + ".res1", N * 2, bottleneck=True)
    s.conv(p + ".conv2", N * 2, N, 3)
    gdn(p + ".gdn2", N)
    s.res_block(p + ".res2", N * 2, bottleneck=True)
    s.conv(p + ".conv3", N * 2, N, 3)
    gdn(p + ".gdn3", N)
    s.conv(p + ".conv4", N, M, 3, gain=Y_GAIN)


def res_decoder_gdn_spec(s, p, gdn, N=64, M=96):
    """ResDecoder with IGDN: layers.py:370-383, dmc_net.py:93-106."""
    s.subpel(p + ".up1", M, N, 3)
    gdn(p + ".gdn1", N)
    s.subpel(p + ".up2", N, N, 3)
    gdn(p + ".gdn2", N)
    s.res_block(p + ".res1", N * 2, bottleneck=True)
    s.subpel(p + ".up3", N * 2, N, 3)
    gdn(p + ".gdn3", N)
    s.res_block(p + ".res2", N * 2, bottleneck=True)
    s.subpel(p + ".up4", N * 2, 32, 3)


def recon_generation_bl_spec(s, p, ctx=64, res=32, ch=64):
    """ReconGeneration: layers.py:398-406, dmc_net.py:143-151."""
    s.conv(p + ".feature_conv.0", ctx + res, ch, 3, gain=0.7)
    s.res_block(p + ".feature_conv.1", ch)
    s.res_block(p + ".feature_conv.2", ch)
    s.conv(p + ".recon_conv", ch, 3, 3, gain=0.15, bias=0.5)


def intra_ss_spec(channel_BL=192, N=64, M=96):
    """IntraSS   IntraSS.py:74-113 (module registration order of the reference)."""
    s = Spec()
    s.entropy_bottleneck("entropy_bottleneck", N)
    intra_noar_spec(s, "base_layer_model.", channel_BL)
    s.conv("texture_resampler.conv_adaptor.0", 3, 64, 3, gain=1.4)
    s.conv("texture_resampler.conv_adaptor.2", 64, 64, 3)
    s.conv("layer_prior_resampler.conv_adaptor.0", channel_BL, M, 3, gain=1.4)
    s.conv("layer_prior_resampler.conv_adaptor.2", M, M, 3)
    s.extractor3("texture_extractor", 64, 64, 64, 64)
    s.fusion3("context_fusion_net", 64, 64, 64)
    res_encoder_gdn_spec(s, "g_a", s.gdn_intra)
    s.prior_encoder("h_a", M, N, Z_GAIN)
    s.subpel("h_s.0", N, M, 3, gain=1.4)
    s.subpel("h_s.2", M, M * 3 // 2, 3, gain=1.4)
    s.conv("h_s.4", M * 3 // 2, M * 2, 3)
    res_decoder_gdn_spec(s, "g_s", s.gdn_intra)
    recon_generation_bl_spec(s, "recon_net")
    s.conv("prior_fusion_net.context_parameters.0", N, M * 3 // 2, 3, gain=1.4)
    s.conv("prior_fusion_net.context_parameters.2", M * 3 // 2, M * 2, 3)
    s.conv("prior_fusion_net.params_net.0", M * 15 // 3, M * 12 // 3, 3, gain=1.4)
    s.conv("prior_fusion_net.params_net.2", M * 12 // 3, M * 9 // 3, 3, gain=1.4)
    _params_head(s, "prior_fusion_net.params_net.4", M * 9 // 3, M * 6 // 3)
    s.gaussian_conditional("gaussian_conditional")
    return s


def dmc_spec(s, p):
    """DMC (DMCExtend)   dmc_net.py:159-265."""
    mv, N, M = 128, 64, 96
    s.spynet(p + "optic_flow")
    cin = 2
    for base in (0, 4, 8):
        s.conv(f"{p}mv_encoder.{base}", cin, mv, 3)
        s.gdn_inter(f"{p}mv_encoder.{base + 1}", mv)
        s.res_block(f"{p}mv_encoder.{base + 2}", mv)
        cin = mv
    s.conv(p + "mv_encoder.12", mv, mv, 3, gain=Y_GAIN)
    s.prior_encoder(p + "mv_prior_encoder", mv, N, Z_GAIN)
    s.deconv(p + "mv_prior_decoder.0", N, mv, 3, gain=1.4)
    s.deconv(p + "mv_prior_decoder.2", mv, mv * 3 // 2, 3, gain=1.4)
    s.deconv(p + "mv_prior_decoder.4", mv * 3 // 2, mv * 2, 3, gain=0.5)
    s[p + "mv_prior_decoder.4.bias"]["kind"] = "scale_mean_bias"
    s.deconv(p + "mv_decoder.0", mv, mv, 3, gain=1.4)
    s.res_block(p + "mv_decoder.2", mv)
    s.gdn_inter(p + "mv_decoder.3", mv)
    s.deconv(p + "mv_decoder.4", mv, mv, 3)
    s.gdn_inter(p + "mv_decoder.5", mv)
    s.deconv(p + "mv_decoder.6", mv, mv, 3)
    s.gdn_inter(p + "mv_decoder.7", mv)
    s.deconv(p + "mv_decoder.8", mv, 2, 3, gain=0.5)
    s.conv(p + "feature_adaptor_I", 3, N, 3)
    s.conv(p + "feature_adaptor_P", N, N, 1, gain=0.6)
    s.extractor3(p + "feature_extractor", N, N, N, N)
    s.fusion3(p + "context_fusion_net", N, N, N)
    res_encoder_gdn_spec(s, p + "res_encoder", s.gdn_inter)
    s.prior_encoder(p + "res_prior_encoder", M, N, Z_GAIN)
    s.deconv(p + "res_prior_decoder.0", N, M, 3, gain=1.4)
    s.deconv(p + "res_prior_decoder.2", M, M * 3 // 2, 3, gain=1.4)
    s.deconv(p + "res_prior_decoder.4", M * 3 // 2, M * 2, 3)
    t = p + "temporal_prior_encoder"
    s.conv(t + ".conv1", N, N, 3)
    s.gdn_inter(t + ".gdn1", N)
    s.conv(t + ".conv2", N * 2, M, 3)
    s.gdn_inter(t + ".gdn2", M)
    s.conv(t + ".conv3", M + N, M * 3 // 2, 3)
    s.gdn_inter(t + ".gdn3", M * 3 // 2)
    s.conv(t + ".conv4", M * 3 // 2, M * 2, 3)
    s.conv(p + "res_entropy_parameter.0", M * 12 // 3, M * 10 // 3, 3, gain=1.4)
    s.conv(p + "res_entropy_parameter.2", M * 10 // 3, M * 8 // 3, 3, gain=1.4)
    _params_head(s, p + "res_entropy_parameter.4", M * 8 // 3, M * 6 // 3)
    res_decoder_gdn_spec(s, p + "res_decoder", s.gdn_inter)
    recon_generation_bl_spec(s, p + "recon_generation_net")
    s.bit_estimator(p + "bit_estimator_z", N)
    s.bit_estimator(p + "bit_estimator_z_mv", N)


def seq2(s, name, cin, mid, cout, gain2=1.0):
    s.conv(name + ".0", cin, mid, 3, gain=1.4)
    s.conv(name + ".2", mid, cout, 3, gain=gain2)


def unet_spec(s, p, cin, cout):
    """UNet   lssvc_modules.py:295-315."""
    s.depth_conv_block(p + ".conv1", cin, 32)
    s.depth_conv_block(p + ".conv2", 32, 64)
    s.depth_conv_block(p + ".conv3", 64, 128)
    for i in range(4):
        s.depth_conv_block(f"{p}.context_refine.{i}", 128, 128)
    s.subpel(p + ".up3", 128, 64, 1)
    s.depth_conv_block(p + ".up_conv3", 128, 64)
    s.subpel(p + ".up2", 64, 32, 1)
    s.depth_conv_block(p + ".up_conv2", 64, cout)


def lssvc_spec():
    """LSSVC / LSSVC_extend   LSSVC_net.py:12-139 (module registration order of the reference)."""
    s = Spec()
    mv = 64
    dmc_spec(s, "base_layer_model.")
    s.conv("feature_adaptor_EL_I", 3, G1, 3)
    s.conv("feature_adaptor_EL_first_P", 64, G1, 3, gain=0.6)
    s.conv("feature_adaptor_EL", G1, G1, 3, gain=0.6)
    # MvResampler
    seq2(s, "mv_resampler.conv1", 2, 64, 64)
    seq2(s, "mv_resampler.conv2", 64, 64, 64)
    s.depth_conv_block("mv_resampler.feature_refine.0", 64, 64)
    s.depth_conv_block("mv_resampler.feature_refine.1", 64, 64)
    s.conv("mv_resampler.recon_conv", 64, 2, 3, gain=0.5)
    # TextureResampler
    s.conv("texture_resampler.conv_adaptor.base_layer_adaptor", 64, 64, 3)
    s.conv("texture_resampler.conv_adaptor.enhance_layer_adaptor", G1, 64, 3)
    seq2(s, "texture_resampler.conv1", 64, 64, 64)
    seq2(s, "texture_resampler.conv2", 64, 64, 64)
    s.depth_conv_block("texture_resampler.feature_refine.0", 64, 64)
    s.depth_conv_block("texture_resampler.feature_refine.1", 64, 64)
    # LayerPriorResampler
    s.conv("layer_prior_resampler.conv_adaptor.base_layer_adaptor", 96, 96, 3)
    s.conv("layer_prior_resampler.conv_adaptor.enhance_layer_adaptor", G16, 96, 3)
    seq2(s, "layer_prior_resampler.conv1", 96, 96, 96)
    seq2(s, "layer_prior_resampler.conv2", 96, 96, G16)
    s.depth_conv_block("layer_prior_resampler.feature_refine.0", G16, G16)
    s.depth_conv_block("layer_prior_resampler.feature_refine.1", G16, G16)
    s.extractor3("feature_extractor", G1, G1, G2, G4)
    s.extractor3("texture_extractor", 64, G1, G2, G4)
    s.fusion3("context_fusion_net", G1, G2, G4)
    for i, c in ((1, G1), (2, G2), (3, G4)):
        g = f"weight_map_generator.generator{i}"
        s.conv(g + ".0", c * 2, 64, 3, gain=1.4)
        s.res_block(g + ".1", 64)
        s.conv(g + ".2", 64, 2, 3, gain=1.0)
    s.depth_conv_block("prior_fusion_net.prior_fusion_conv.0", G16 * 3, G16 * 3)
    s.depth_conv_block("prior_fusion_net.prior_fusion_conv.1", G16 * 3, G16 * 2)
    for i in (1, 2, 3):
        s.conv(f"y_spatial_prior_adaptor_{i}", G16 * 3, G16 * 3, 1)
    s.depth_conv_block("y_spatial_prior.0", G16 * 3, G16 * 3)
    s.depth_conv_block("y_spatial_prior.1", G16 * 3, G16 * 3)
    s.depth_conv_block("y_spatial_prior.2", G16 * 3, G16 * 2)
    # ResEncoder (no GDN)
    s.conv("res_encoder.conv1", G1 + 3, G2, 3, in_boost=(3, X_BOOST))
    s.res_block("res_encoder.res1", G2 * 2, bottleneck=True)
    s.conv("res_encoder.conv2", G2 * 2, G4, 3)
    s.res_block("res_encoder.res2", G4 * 2, bottleneck=True)
    s.conv("res_encoder.conv3", G4 * 2, G8, 3)
    s.conv("res_encoder.conv4", G8, G16, 3, gain=Y_GAIN)
    s.prior_encoder("res_prior_encoder", G16, G16, Z_GAIN)
    s.conv("res_prior_decoder.0", G16, G16, 3, gain=1.4)
    s.subpel("res_prior_decoder.2", G16, G16, 1, gain=1.4)
    s.conv("res_prior_decoder.4", G16, G16, 3, gain=1.4)
    s.subpel("res_prior_decoder.6", G16, G16, 1, gain=1.4)
    s.conv("res_prior_decoder.8", G16, G16, 3)
    s.conv("temporal_prior_encoder.0", G4, G8, 3, gain=1.4)
    s.conv("temporal_prior_encoder.2", G8, G16, 3)
    # ResDecoder (no GDN)
    s.subpel("res_decoder.up1", G16, G8, 3)
    s.subpel("res_decoder.up2", G8, G4, 3)
    s.res_block("res_decoder.res1", G4 * 2, bottleneck=True)
    s.subpel("res_decoder.up3", G4 * 2, G2, 3)
    s.res_block("res_decoder.res2", G2 * 2, bottleneck=True)
    s.subpel("res_decoder.up4", G2 * 2, 32, 3)
    # ReconGeneration
    s.conv("recon_generation_net.first_conv", G1 + 32, G1, 3, gain=0.7)
    unet_spec(s, "recon_generation_net.unet_1", G1, G1)
    unet_spec(s, "recon_generation_net.unet_2", G1, G1)
    s.conv("recon_generation_net.recon_conv", G1, 3, 3, gain=0.15, bias=0.5)
    s.spynet("optic_flow")
    # OffsetDiversity
    s.conv("align.conv_offset.0", G1 + 3 + 2, G2, 3, gain=1.4)
    s.conv("align.conv_offset.2", G2, G2, 3, gain=1.4)
    s.conv("align.conv_offset.4", G2, 3 * 16 * 2, 3, gain=0.1)
    s.conv("align.fusion", G1 * 2, G1, 1, groups=16)
    s.conv("mv_ctx_transform.transform.0", 2, mv, 3)
    s.res_block("mv_ctx_transform.transform.1", mv)
    # MVResEncoder
    s.conv("mv_encoder.encoder1.0", 2, mv, 3)
    s.gdn_inter("mv_encoder.encoder1.1", mv)
    s.res_block("mv_encoder.encoder1.2", mv)
    s.conv("mv_encoder.encoder2.0", mv * 2, mv, 3)
    s.gdn_inter("mv_encoder.encoder2.1", mv)
    s.res_block("mv_encoder.encoder2.2", mv)
    s.conv("mv_encoder.encoder2.4", mv, mv, 3)
    s.gdn_inter("mv_encoder.encoder2.5", mv)
    s.res_block("mv_encoder.encoder2.6", mv)
    s.conv("mv_encoder.encoder2.8", mv, mv, 3, gain=Y_GAIN)
    s.prior_encoder("mv_prior_encoder", mv, mv, Z_GAIN)
    s.subpel("mv_prior_decoder.0", mv, mv, 3, gain=1.4)
    s.subpel("mv_prior_decoder.2", mv, mv * 3 // 2, 3, gain=1.4)
    s.conv("mv_prior_decoder.4", mv * 3 // 2, mv * 2, 3)
    # MVResDecoder
    s.subpel("mv_decoder.decoder1.0", mv, mv, 3, gain=1.4)
    s.res_block("mv_decoder.decoder1.2", mv)
    s.gdn_inter("mv_decoder.decoder1.3", mv)
    s.subpel("mv_decoder.decoder1.4", mv, mv, 3)
    s.gdn_inter("mv_decoder.decoder1.5", mv)
    s.subpel("mv_decoder.decoder1.6", mv, mv, 3)
    s.gdn_inter("mv_decoder.decoder1.7", mv)
    s.conv("mv_decoder.decoder2.0", mv * 2, mv, 3, gain=1.4)
    s.subpel("mv_decoder.decoder2.2", mv, 2, 3, gain=0.5)
    s.conv("mv_ctx_prior_encoder.0", 2, mv, 3)
    s.gdn_inter("mv_ctx_prior_encoder.1", mv)
    s.conv("mv_ctx_prior_encoder.2", mv, mv, 3)
    s.gdn_inter("mv_ctx_prior_encoder.3", mv)
    s.conv("mv_ctx_prior_encoder.4", mv, mv, 3)
    s.gdn_inter("mv_ctx_prior_encoder.5", mv)
    s.conv("mv_ctx_prior_encoder.6", mv, mv, 3)
    s.conv("mv_prior_fusion.0", mv * 9 // 3, mv * 8 // 3, 3, gain=1.4)
    s.conv("mv_prior_fusion.2", mv * 8 // 3, mv * 7 // 3, 3, gain=1.4)
    _params_head(s, "mv_prior_fusion.4", mv * 7 // 3, mv * 6 // 3)
    s.bit_estimator("bit_estimator_z", G16)
    s.bit_estimator("bit_estimator_z_mv", mv)
    return s


# -----------------------------------------------------------------------------------------------------------
# initialisation + module wrapper
# -----------------------------------------------------------------------------------------------------------
_GAINS = None


def synth_gains():
    """Per-conv weight multipliers measured by tools/calibrate_synth.py (LSUV-style pass through the oracle
    on a synthetic sequence, so that every conv output has the target std declared in the spec)."""
    global _GAINS
    if _GAINS is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "synth_gains.json")
        _GAINS = {}
        if os.path.exists(path):
            with open(path) as f:
                _GAINS = json.load(f)
    return _GAINS


def model_gains(model):
    """model: 'I' (IntraSS) or 'P' (LSSVC)."""
    return synth_gains().get(model, {})


def init_tensor(name, e, gen, gains=None):
    shape, kind = e["shape"], e["kind"]
    if kind in ("conv_w", "deconv_w"):
        rf = shape[2] * shape[3]
        fan_in = (shape[1] if kind == "conv_w" else shape[0]) * rf
        g = (gains or {}).get(name[: -len(".weight")], min(e["gain"], 1.0))
        w = torch.randn(shape, generator=gen) * (g / math.sqrt(fan_in))
        if e.get("in_boost"):
            n, f = e["in_boost"]
            w[:, :n] *= f
        return w
    if kind == "bias":
        return torch.full(shape, float(e["value"]))
    if kind == "scale_mean_bias":
        half = shape[0] // 2
        b = torch.zeros(shape)
        b[:half] = SCALE_BIAS
        return b
    if kind == "const":
        return torch.full(shape, float(e["value"]))
    if kind == "zeros":
        return torch.zeros(shape)
    if kind == "empty_int":
        return torch.zeros(shape, dtype=torch.int32)
    if kind == "gdn_beta_inter":
        return torch.sqrt(torch.ones(shape) + (2.0 ** -18) ** 2)
    if kind == "gdn_gamma_inter":
        return torch.sqrt(0.1 * torch.eye(shape[0]) + (2.0 ** -18) ** 2)
    if kind == "gdn_beta_intra":
        ped = torch.tensor([2.0 ** -36])
        return torch.sqrt(torch.max(torch.ones(shape) + ped, ped))
    if kind == "gdn_gamma_intra":
        ped = torch.tensor([2.0 ** -36])
        return torch.sqrt(torch.max(0.1 * torch.eye(shape[0]) + ped, ped))
    if kind == "bitparm_h":
        return torch.randn(shape, generator=gen) * 0.01 - 0.5
    if kind in ("bitparm_b", "bitparm_a"):
        return torch.randn(shape, generator=gen) * 0.01
    if kind == "eb_quantiles":
        return torch.tensor([-10.0, 0.0, 10.0]).repeat(shape[0], 1, 1)
    if kind == "eb_target":
        t = math.log(2 / 1e-9 - 1)
        return torch.tensor([-t, 0.0, t])
    if kind == "eb_bias":
        return torch.rand(shape, generator=gen) - 0.5
    raise KeyError(kind)


class ParamBag(nn.Module):
    """nn.Module carrying the tensors of a Spec under the reference's dotted names."""

    def __init__(self, spec, seed=None, gains=None):
        super().__init__()
        self._spec = spec
        gen = torch.Generator()
        gen.manual_seed(int(torch.initial_seed() if seed is None else seed) & 0x7FFFFFFF)
        for name, e in spec.items():
            t = init_tensor(name, e, gen, gains)
            mod = self
            parts = name.split(".")
            for part in parts[:-1]:
                if part not in mod._modules:
                    mod.add_module(part, nn.Module())
                mod = mod._modules[part]
            if e["buffer"]:
                mod.register_buffer(parts[-1], t)
            else:
                mod.register_parameter(parts[-1], nn.Parameter(t, requires_grad=False))
        self._version_counter = 0

    def tensor(self, name):
        """Fetch a parameter/buffer by its dotted reference name."""
        mod = self
        parts = name.split(".")
        for part in parts[:-1]:
            mod = mod._modules[part]
        t = mod._parameters.get(parts[-1])
        return t if t is not None else mod._buffers[parts[-1]]
