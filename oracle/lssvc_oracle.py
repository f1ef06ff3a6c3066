"""ORACLE — test infrastructure only (tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg).

A CPU fp32 restatement of the reference's per-frame two-layer coding forward pass (EsakaK/LSSVC), written
functionally over a reference-layout ``state_dict``: no nn.Module tree, plain torch ops in the reference's
operation order.  Each function cites the reference file:line it follows (paths under /root/reference).

Pinning: the reference ships no tests or golden vectors.  This restatement is pinned against
the reference itself, imported unmodified in the build container: tools/make_golden.py runs both on identical
seeded inputs/weights, requires bit-identical outputs, and commits small fixtures to tests/golden/ which
tests/test_oracle.py re-checks wherever the reference is absent (the GPU box).

Nothing under lssvc_b200/ imports this module.
"""
import math

import torch
import torch.nn.functional as F

LN2 = math.log(2.0)


# ---------------------------------------------------------------------------------------------------------
# parameter access + primitive layers
# ---------------------------------------------------------------------------------------------------------
class Params:
    """state_dict with a key prefix; sub('a').sub('b')['weight'] == sd['a.b.weight']."""

    def __init__(self, sd, prefix=""):
        self.sd, self.prefix = sd, prefix

    def sub(self, name):
        return Params(self.sd, f"{self.prefix}{name}.")

    def __getitem__(self, name):
        return self.sd[self.prefix + name]

    def has(self, name):
        return (self.prefix + name) in self.sd


def conv(p, x, stride=1, padding=None):
    w = p["weight"]
    if padding is None:
        padding = w.shape[-1] // 2
    return F.conv2d(x, w, p["bias"], stride=stride, padding=padding)


def deconv(p, x, stride):
    # nn.ConvTranspose2d(k=3, stride, padding=1, output_padding=stride-1)   dmc_net.py:198-246
    return F.conv_transpose2d(x, p["weight"], p["bias"], stride=stride, padding=1, output_padding=stride - 1)


def subpel(p, x, padding=None):
    # conv -> PixelShuffle(2); Sequential index 0 holds the conv   layers.py:41-52, video_net_component.py:21-32
    return F.pixel_shuffle(conv(p.sub("0"), x, padding=padding), 2)


def lrelu(x, slope=0.01):
    return F.leaky_relu(x, slope)


def gdn_intra(p, x, inverse=False):
    # src/IntraModules/gdn.py:29-44 with NonNegativeParametrizer (others.py:43-67)
    C = x.shape[1]
    ped = 2.0 ** -36
    beta = torch.max(p["beta"], torch.tensor((1e-6 + ped) ** 0.5)) ** 2 - ped
    gamma = torch.max(p["gamma"], torch.tensor((0.0 + ped) ** 0.5)) ** 2 - ped
    norm = F.conv2d(x ** 2, gamma.reshape(C, C, 1, 1), beta)
    norm = torch.sqrt(norm) if inverse else torch.rsqrt(norm)
    return x * norm


def gdn_inter(p, x, inverse=False):
    # src/InterModules/video_net_component.py:83-105
    C = x.shape[1]
    ped = (2.0 ** -18) ** 2
    beta = torch.max(p["beta"], torch.ones_like(p["beta"]) * ((1e-6 + ped) ** 0.5)) ** 2 - ped
    gamma = torch.max(p["gamma"], torch.ones_like(p["gamma"]) * (2.0 ** -18)) ** 2 - ped
    norm = torch.sqrt(F.conv2d(x ** 2, gamma.view(C, C, 1, 1), beta))
    return x * norm if inverse else x / norm


def res_block(p, x, slope=0.01, start_from_relu=True, end_with_relu=False):
    # ResBlock: video_net_component.py:170-188, layers.py:229-255
    out = lrelu(x, slope) if start_from_relu else x
    out = lrelu(conv(p.sub("conv1"), out), slope)
    out = conv(p.sub("conv2"), out)
    if end_with_relu:
        out = lrelu(out, slope)
    return x + out


def depth_conv_block(p, x, slope_dc=0.01, slope_ffn=0.1):
    # DepthConvBlock = DepthConv + ConvFFN   lssvc_modules.py:15-72
    dc, ffn = p.sub("block.0"), p.sub("block.1")
    identity = conv(dc.sub("adaptor"), x, padding=0) if dc.has("adaptor.weight") else x
    out = lrelu(conv(dc.sub("conv1.0"), x, padding=0), slope_dc)
    w = dc["depth_conv.weight"]
    out = F.conv2d(out, w, dc["depth_conv.bias"], padding=1, groups=w.shape[0])
    out = conv(dc.sub("conv2"), out, padding=0) + identity
    f = lrelu(conv(ffn.sub("conv.0"), out, padding=0), slope_ffn)
    f = lrelu(conv(ffn.sub("conv.2"), f, padding=0), slope_ffn)
    return out + f


def flow_warp(feature, flow):
    # torch_warp   video_net_component.py:329-352
    N, _, H, W = flow.shape
    hor = torch.linspace(-1.0, 1.0, W, dtype=feature.dtype).view(1, 1, 1, W).expand(N, -1, H, -1)
    ver = torch.linspace(-1.0, 1.0, H, dtype=feature.dtype).view(1, 1, H, 1).expand(N, -1, -1, W)
    base = torch.cat([hor, ver], 1)
    fl = torch.cat([flow[:, 0:1] / ((feature.size(3) - 1.0) / 2.0), flow[:, 1:2] / ((feature.size(2) - 1.0) / 2.0)], 1)
    return F.grid_sample(feature, (base + fl).permute(0, 2, 3, 1), mode="bilinear", padding_mode="border",
                         align_corners=True)


def bilinear(x, size):
    return F.interpolate(x, size=size, mode="bilinear", align_corners=False)


def up2(x):      # bilinearupsacling   video_net_component.py:355-360
    return bilinear(x, (x.shape[2] * 2, x.shape[3] * 2))


def down2(x):    # bilineardownsacling video_net_component.py:363-368
    return bilinear(x, (x.shape[2] // 2, x.shape[3] // 2))


# ---------------------------------------------------------------------------------------------------------
# entropy models
# ---------------------------------------------------------------------------------------------------------
def laplace_bits(q, sigma):
    # get_y_bits_probs   LSSVC_net.py:154-161, dmc_net.py:370-377
    sigma = sigma.clamp(1e-5, 1e10)
    lap = torch.distributions.laplace.Laplace(torch.zeros_like(sigma), sigma)
    probs = lap.cdf(q + 0.5) - lap.cdf(q - 0.5)
    return torch.sum(torch.clamp(-1.0 * torch.log(probs + 1e-5) / LN2, 0, 50))


def bitparm_cdf(p, x):
    # BitEstimator.forward / Bitparm   video_entropy_models.py:110-166
    for name in ("f1", "f2", "f3"):
        f = p.sub(name)
        x = x * F.softplus(f["h"]) + f["b"]
        x = x + torch.tanh(x) * torch.tanh(f["a"])
    f = p.sub("f4")
    return torch.sigmoid(x * F.softplus(f["h"]) + f["b"])


def bitparm_bits(p, z_hat):
    # get_z_bits_probs   LSSVC_net.py:163-167
    prob = bitparm_cdf(p, z_hat + 0.5) - bitparm_cdf(p, z_hat - 0.5)
    return torch.sum(torch.clamp(-1.0 * torch.log(prob + 1e-5) / LN2, 0, 50))


def eb_logits(p, x):
    # EntropyBottleneck._logits_cumulative   img_entropy_models.py:483-502
    logits = x
    for i in range(5):
        logits = torch.matmul(F.softplus(p[f"_matrices.{i}"]), logits)
        logits = logits + p[f"_biases.{i}"]
        if i < 4:
            logits = logits + torch.tanh(p[f"_factors.{i}"]) * torch.tanh(logits)
    return logits


def entropy_bottleneck(p, z):
    # EntropyBottleneck.forward (eval)   img_entropy_models.py:518-554
    xp = z.permute(1, 2, 3, 0).contiguous()
    shape = xp.size()
    values = xp.reshape(xp.size(0), 1, -1)
    med = p["quantiles"][:, :, 1:2]
    out = torch.round(values - med) + med
    lower = eb_logits(p, out - 0.5)
    upper = eb_logits(p, out + 0.5)
    sign = -torch.sign(lower + upper)
    lik = torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower))
    lik = torch.max(lik, torch.tensor(1e-9))
    out = out.reshape(shape).permute(3, 0, 1, 2).contiguous()
    lik = lik.reshape(shape).permute(3, 0, 1, 2).contiguous()
    return out, lik


def gaussian_conditional(y, scales, means):
    # GaussianConditional.forward (eval; `if self.train:` is always true -> d_quant)   img_entropy_models.py:650-685
    out = torch.round(y - means) + means
    values = torch.abs(out - means)
    s = torch.max(scales, torch.tensor(0.11))
    cum = lambda t: 0.5 * torch.erfc(float(-(2 ** -0.5)) * t)
    lik = cum((0.5 - values) / s) - cum((-0.5 - values) / s)
    lik = torch.max(lik, torch.tensor(1e-9))
    r = y - means
    y_hat = r + (torch.round(r) - r) + means      # d_quant   :365-370
    return y_hat, lik


def build_indexes_video(scales):
    # GaussianEncoder.build_indexes   video_entropy_models.py:309-313
    lo, hi = math.log(0.01), math.log(64.0)
    s = torch.maximum(scales, torch.zeros_like(scales) + 1e-5)
    idx = (torch.log(s) - lo) / ((hi - lo) / 255)
    return idx.clamp_(0, 255).int()


def build_indexes_image(scales):
    # GaussianConditional.build_indexes   img_entropy_models.py:687-691
    lo, hi = math.log(0.11), math.log(256.0)
    s = torch.maximum(scales, torch.zeros_like(scales) + 1e-5)
    idx = (torch.log(s) - lo) / ((hi - lo) / 63) + 1
    return idx.clamp_(0, 63).int()


# ---------------------------------------------------------------------------------------------------------
# I-frame: IntraNoAR (base layer) + IntraSS (enhancement layer)
# ---------------------------------------------------------------------------------------------------------
def _rb_stride(p, x):
    # ResidualBlockWithStride   layers.py:60-91
    out = lrelu(conv(p.sub("conv1"), x, stride=2))
    out = gdn_intra(p.sub("gdn"), conv(p.sub("conv2"), out))
    return out + conv(p.sub("downsample"), x, stride=2, padding=0)


def _rb(p, x):
    # ResidualBlock   layers.py:122-145
    out = lrelu(conv(p.sub("conv1"), x))
    out = lrelu(conv(p.sub("conv2"), out))
    return out + x


def _rb_up(p, x):
    # ResidualBlockUpsample   layers.py:94-119
    out = lrelu(subpel(p.sub("subpel_conv"), x))
    out = gdn_intra(p.sub("igdn"), conv(p.sub("conv"), out), inverse=True)
    return out + subpel(p.sub("upsample"), x)


def intra_noar(p, x):
    """IntraNoAR.get_layer_information   priors.py:368-388 (nets :116-159)."""
    g_a = p.sub("g_a")
    y = x
    for i in range(6):
        y = _rb_stride(g_a.sub(str(i)), y) if i % 2 == 0 else _rb(g_a.sub(str(i)), y)
    y = conv(g_a.sub("6"), y, stride=2)
    h_a = p.sub("h_a")
    z = lrelu(conv(h_a.sub("0"), y))
    z = lrelu(conv(h_a.sub("2"), z))
    z = lrelu(conv(h_a.sub("4"), z, stride=2))
    z = lrelu(conv(h_a.sub("6"), z))
    z = conv(h_a.sub("8"), z, stride=2)
    z_hat, z_lik = entropy_bottleneck(p.sub("entropy_bottleneck"), z)
    h_s = p.sub("h_s")
    g = lrelu(conv(h_s.sub("0"), z_hat))
    g = lrelu(subpel(h_s.sub("2"), g))
    g = lrelu(conv(h_s.sub("4"), g))
    g = lrelu(subpel(h_s.sub("6"), g))
    g = conv(h_s.sub("8"), g)
    scales, means = g.chunk(2, 1)
    y_hat, y_lik = gaussian_conditional(y, scales, means)
    g_s = p.sub("g_s")
    x_hat = y_hat
    for i in range(7):
        x_hat = _rb(g_s.sub(str(i)), x_hat) if i % 2 == 0 else _rb_up(g_s.sub(str(i)), x_hat)
    x_hat = subpel(g_s.sub("7"), x_hat)
    bits = (torch.log(y_lik).sum() + torch.log(z_lik).sum()) / (-LN2)
    return {"bits": bits, "x_hat": x_hat, "y_hat": y_hat, "y": y, "z": z, "z_hat": z_hat, "scales": scales,
            "means": means}


def _extractor3(p, x):
    # MultiScaleTextureExtractor layers.py:288-308; FeatureExtractor / TextureExtractor lssvc_modules.py:157-200,
    # dmc_net.py:11-31 — same data flow, channel counts come from the weights
    l1 = res_block(p.sub("res_block1"), conv(p.sub("conv1"), x))
    l2 = res_block(p.sub("res_block2"), conv(p.sub("conv2"), l1, stride=2))
    l3 = res_block(p.sub("res_block3"), conv(p.sub("conv3"), l2, stride=2))
    return l1, l2, l3


def _fusion3(p, c1, c2, c3):
    # MultiScaleTextureFusion layers.py:311-339; MultiScaleContextFusion lssvc_modules.py:203-232, dmc_net.py:34-62
    c3_up = res_block(p.sub("res_block3_up"), subpel(p.sub("conv3_up"), c3))
    c3_out = res_block(p.sub("res_block3_out"), conv(p.sub("conv3_out"), c3))
    cat2 = torch.cat((c3_up, c2), dim=1)
    c2_up = res_block(p.sub("res_block2_up"), subpel(p.sub("conv2_up"), cat2))
    c2_out = res_block(p.sub("res_block2_out"), conv(p.sub("conv2_out"), cat2))
    c1_out = res_block(p.sub("res_block1_out"), conv(p.sub("conv1_out"), torch.cat((c2_up, c1), dim=1)))
    return c1 + c1_out, c2 + c2_out, c3 + c3_out


def _res_encoder_gdn(p, x, c1, c2, c3):
    # ResEncoder with GDN: layers.py:342-367 (intra GDN), dmc_net.py:65-90 (inter GDN)
    gdn = gdn_intra if p.has("gdn1.beta_reparam.pedestal") else gdn_inter
    f = gdn(p.sub("gdn1"), conv(p.sub("conv1"), torch.cat([x, c1], dim=1), stride=2))
    f = res_block(p.sub("res1"), torch.cat([f, c2], dim=1), slope=0.1, start_from_relu=False, end_with_relu=True)
    f = gdn(p.sub("gdn2"), conv(p.sub("conv2"), f, stride=2))
    f = res_block(p.sub("res2"), torch.cat([f, c3], dim=1), slope=0.1, start_from_relu=False, end_with_relu=True)
    f = gdn(p.sub("gdn3"), conv(p.sub("conv3"), f, stride=2))
    return conv(p.sub("conv4"), f, stride=2)


def _res_decoder_gdn(p, x, c2, c3):
    # ResDecoder with IGDN: layers.py:370-395, dmc_net.py:93-118
    gdn = gdn_intra if p.has("gdn1.beta_reparam.pedestal") else gdn_inter
    f = gdn(p.sub("gdn1"), subpel(p.sub("up1"), x), inverse=True)
    f = gdn(p.sub("gdn2"), subpel(p.sub("up2"), f), inverse=True)
    f = res_block(p.sub("res1"), torch.cat([f, c3], dim=1), slope=0.1, start_from_relu=False, end_with_relu=True)
    f = gdn(p.sub("gdn3"), subpel(p.sub("up3"), f), inverse=True)
    f = res_block(p.sub("res2"), torch.cat([f, c2], dim=1), slope=0.1, start_from_relu=False, end_with_relu=True)
    return subpel(p.sub("up4"), f)


def _recon_generation_bl(p, first, second):
    # ReconGeneration(ctx, res) is CALLED as (res_feature, context1): cat order = (first, second)
    # layers.py:398-411, dmc_net.py:143-156; call sites IntraSS.py:161, dmc_net.py:452
    fc = p.sub("feature_conv")
    f = conv(fc.sub("0"), torch.cat((first, second), dim=1))
    f = res_block(fc.sub("1"), f)
    f = res_block(fc.sub("2"), f)
    return f, conv(p.sub("recon_conv"), f)


def depad(feature, pad_size, p=1):
    """get_depadded_feature   LSSVC_net.py:271-282, IntraSS.py:124-135: F.pad by pad_size / p (negative = crop)."""
    if feature is None:
        return None
    return F.pad(feature, tuple(int(a / p) for a in pad_size), mode="constant", value=0)


def intra_ss(sd, x_bl, x_el, shape_hr, pad_size=(0, 0, 0, 0)):
    """IntraSS.forward   IntraSS.py:137-172 (pad_size is (0,0,0,0) on the test.py path, :212-213)."""
    p = Params(sd)
    bl = intra_noar(p.sub("base_layer_model"), x_bl)
    x_hat_bl_full, y_hat_bl_full = bl["x_hat"], bl["y_hat"]
    # de-padded copies feed the enhancement layer; the returned x_hat_bl is the full one   IntraSS.py:145-147
    x_hat_bl, y_hat_bl = depad(x_hat_bl_full, pad_size), depad(y_hat_bl_full, pad_size, 16)
    # multi_scale_context_mining   IntraSS.py:119-122; TextureResampler layers.py:258-270
    tr = p.sub("texture_resampler.conv_adaptor")
    texture = conv(tr.sub("2"), lrelu(conv(tr.sub("0"), x_hat_bl)))
    texture = bilinear(texture, shape_hr)
    t1, t2, t3 = _extractor3(p.sub("texture_extractor"), texture)
    c1, c2, c3 = _fusion3(p.sub("context_fusion_net"), t1, t2, t3)
    y = _res_encoder_gdn(p.sub("g_a"), x_el, c1, c2, c3)
    h_a = p.sub("h_a")
    z = lrelu(conv(h_a.sub("0"), y))
    z = lrelu(conv(h_a.sub("2"), z, stride=2))
    z = conv(h_a.sub("4"), z, stride=2)
    z_hat, z_lik = entropy_bottleneck(p.sub("entropy_bottleneck"), z)
    h_s = p.sub("h_s")
    hyper = lrelu(subpel(h_s.sub("0"), z_hat))
    hyper = lrelu(subpel(h_s.sub("2"), hyper))
    hyper = conv(h_s.sub("4"), hyper)
    # LayerPriorResampler   layers.py:273-285
    lp = p.sub("layer_prior_resampler.conv_adaptor")
    layer_prior = conv(lp.sub("2"), lrelu(conv(lp.sub("0"), y_hat_bl)))
    layer_prior = bilinear(layer_prior, (shape_hr[0] // 16, shape_hr[1] // 16))
    # PriorFusion   layers.py:473-492
    pf = p.sub("prior_fusion_net")
    ctx_p = lrelu(conv(pf.sub("context_parameters.0"), c3, stride=2), 0.1)
    ctx_p = conv(pf.sub("context_parameters.2"), ctx_p, stride=2)
    prm = torch.cat([hyper, layer_prior, ctx_p], dim=1)
    prm = lrelu(conv(pf.sub("params_net.0"), prm))
    prm = lrelu(conv(pf.sub("params_net.2"), prm))
    prm = conv(pf.sub("params_net.4"), prm)
    scales, means = prm.chunk(2, 1)
    y_hat, y_lik = gaussian_conditional(y, scales, means)
    res_hat = _res_decoder_gdn(p.sub("g_s"), y_hat, c2, c3)
    feature, x_hat = _recon_generation_bl(p.sub("recon_net"), res_hat, c1)
    bit_el = (torch.log(y_lik).sum() + torch.log(z_lik).sum()) / (-LN2)
    return {
        "bit_bl": bl["bits"].item(), "bit_el": bit_el.item(),
        "x_hat_bl": x_hat_bl_full, "x_hat_el": x_hat, "feature_el": feature,
        # internals for parity tests
        "bl": bl, "y": y, "z": z, "z_hat": z_hat, "scales": scales, "means": means, "y_hat": y_hat,
        "ctx": (c1, c2, c3),
    }


# ---------------------------------------------------------------------------------------------------------
# P-frame: DMC (base layer) + LSSVC (enhancement layer)
# ---------------------------------------------------------------------------------------------------------
def spynet(p, im1, im2):
    """ME_Spynet / ME_Spynet_DCVC.forward   video_net_component.py:222-248, 300-326."""
    im1s, im2s = [im1], [im2]
    for _ in range(3):
        im1s.append(F.avg_pool2d(im1s[-1], kernel_size=2, stride=2))
        im2s.append(F.avg_pool2d(im2s[-1], kernel_size=2, stride=2))
    fine = im2s[3]
    flow = torch.zeros(im1.shape[0], 2, fine.shape[2] // 2, fine.shape[3] // 2, dtype=torch.float32)
    for level in range(4):
        flow_up = up2(flow) * 2.0
        mb = p.sub(f"moduleBasic.{level}")
        x = torch.cat([im1s[3 - level], flow_warp(im2s[3 - level], flow_up), flow_up], 1)
        for k in range(1, 5):
            x = F.relu(conv(mb.sub(f"conv{k}"), x, padding=3))
        flow = flow_up + conv(mb.sub("conv5"), x, padding=3)
    return flow


def _mv_encoder_bl(p, x):
    # DMC.mv_encoder   dmc_net.py:174-188
    for base in (0, 4, 8):
        x = gdn_inter(p.sub(str(base + 1)), conv(p.sub(str(base)), x, stride=2))
        x = lrelu(res_block(p.sub(str(base + 2)), x, start_from_relu=False), 0.1)
    return conv(p.sub("12"), x, stride=2)


def _mv_decoder_bl(p, x):
    # DMC.mv_decoder   dmc_net.py:208-221
    x = lrelu(deconv(p.sub("0"), x, 2), 0.1)
    x = gdn_inter(p.sub("3"), res_block(p.sub("2"), x, start_from_relu=False), inverse=True)
    x = gdn_inter(p.sub("5"), deconv(p.sub("4"), x, 2), inverse=True)
    x = gdn_inter(p.sub("7"), deconv(p.sub("6"), x, 2), inverse=True)
    return deconv(p.sub("8"), x, 2)


def _prior_encoder(p, x):
    # mv_prior_encoder / res_prior_encoder: conv, lrelu, conv s2, lrelu, conv s2   dmc_net.py:190-196,230-236
    x = lrelu(conv(p.sub("0"), x))
    x = lrelu(conv(p.sub("2"), x, stride=2))
    return conv(p.sub("4"), x, stride=2)


def _prior_decoder_bl(p, x):
    # ConvTranspose s2, lrelu, ConvTranspose s2, lrelu, ConvTranspose s1   dmc_net.py:198-206,238-246
    x = lrelu(deconv(p.sub("0"), x, 2))
    x = lrelu(deconv(p.sub("2"), x, 2))
    return deconv(p.sub("4"), x, 1)


def dmc(p, x, ref_frame, ref_feature):
    """DMC.get_inter_layer_information   dmc_net.py:421-488."""
    est_mv = spynet(p.sub("optic_flow"), x, ref_frame)
    mv_y = _mv_encoder_bl(p.sub("mv_encoder"), est_mv)
    mv_z = _prior_encoder(p.sub("mv_prior_encoder"), mv_y)
    mv_z_hat = torch.round(mv_z)
    mv_params = _prior_decoder_bl(p.sub("mv_prior_decoder"), mv_z_hat)
    mv_scales, mv_means = mv_params.chunk(2, 1)
    mv_y_q = torch.round(mv_y - mv_means)
    mv_y_hat = mv_y_q + mv_means
    mv_hat = _mv_decoder_bl(p.sub("mv_decoder"), mv_y_hat)
    # motion_compensation   dmc_net.py:359-368
    mv2 = down2(mv_hat) / 2
    mv3 = down2(mv2) / 2
    if ref_feature is None:
        feat = conv(p.sub("feature_adaptor_I"), ref_frame)
    else:
        feat = conv(p.sub("feature_adaptor_P"), ref_feature, padding=0)
    f1, f2, f3 = _extractor3(p.sub("feature_extractor"), feat)
    c1, c2, c3 = flow_warp(f1, mv_hat), flow_warp(f2, mv2), flow_warp(f3, mv3)
    c1, c2, c3 = _fusion3(p.sub("context_fusion_net"), c1, c2, c3)
    y = _res_encoder_gdn(p.sub("res_encoder"), x, c1, c2, c3)
    z = _prior_encoder(p.sub("res_prior_encoder"), y)
    z_hat = torch.round(z)
    hier = _prior_decoder_bl(p.sub("res_prior_decoder"), z_hat)
    # TemporalPriorEncoder   dmc_net.py:121-140
    tp = p.sub("temporal_prior_encoder")
    t = gdn_inter(tp.sub("gdn1"), conv(tp.sub("conv1"), c1, stride=2))
    t = gdn_inter(tp.sub("gdn2"), conv(tp.sub("conv2"), torch.cat([t, c2], dim=1), stride=2))
    t = gdn_inter(tp.sub("gdn3"), conv(tp.sub("conv3"), torch.cat([t, c3], dim=1), stride=2))
    temporal = conv(tp.sub("conv4"), t, stride=2)
    ep = p.sub("res_entropy_parameter")
    g = lrelu(conv(ep.sub("0"), torch.cat((temporal, hier), dim=1)))
    g = lrelu(conv(ep.sub("2"), g))
    g = conv(ep.sub("4"), g)
    scales, means = g.chunk(2, 1)
    y_q = torch.round(y - means)
    y_hat = y_q + means
    rec_feat = _res_decoder_gdn(p.sub("res_decoder"), y_hat, c2, c3)
    feature, recon = _recon_generation_bl(p.sub("recon_generation_net"), rec_feat, c1)
    bits = (laplace_bits(y_q, scales) + bitparm_bits(p.sub("bit_estimator_z"), z_hat)
            + laplace_bits(mv_y_q, mv_scales) + bitparm_bits(p.sub("bit_estimator_z_mv"), mv_z_hat))
    return {"bits": bits, "recon_image": recon, "feature": feature, "y_hat": y_hat, "mv_hat": mv_hat,
            "y_q": y_q, "scales": scales, "z_hat": z_hat, "mv_y_q": mv_y_q, "mv_scales": mv_scales,
            "mv_z_hat": mv_z_hat, "est_mv": est_mv, "ctx": (c1, c2, c3)}


def _seq2(p, x):
    # conv, LeakyReLU, conv   lssvc_modules.py:342-351, 375-384, 407-416
    return conv(p.sub("2"), lrelu(conv(p.sub("0"), x)))


def _mv_resampler(p, mv_bl, shape_hr, s):
    # MvResampler   lssvc_modules.py:339-365
    f = _seq2(p.sub("conv1"), mv_bl)
    up = _seq2(p.sub("conv2"), bilinear(f, shape_hr))
    r = depth_conv_block(p.sub("feature_refine.1"), depth_conv_block(p.sub("feature_refine.0"), up))
    return s * conv(p.sub("recon_conv"), r + up)


def _texture_resampler(p, texture_bl, shape_hr):
    # TextureResampler   lssvc_modules.py:368-397
    key = "base_layer_adaptor" if texture_bl.shape[1] == 64 else "enhance_layer_adaptor"
    f = _seq2(p.sub("conv1"), conv(p.sub(f"conv_adaptor.{key}"), texture_bl))
    up = _seq2(p.sub("conv2"), bilinear(f, shape_hr))
    r = depth_conv_block(p.sub("feature_refine.1"), depth_conv_block(p.sub("feature_refine.0"), up))
    return r + up


def _layer_prior_resampler(p, y_hat_bl, shape):
    # LayerPriorResampler   lssvc_modules.py:400-429
    key = "base_layer_adaptor" if y_hat_bl.shape[1] == 96 else "enhance_layer_adaptor"
    f = _seq2(p.sub("conv1"), conv(p.sub(f"conv_adaptor.{key}"), y_hat_bl))
    up = _seq2(p.sub("conv2"), bilinear(f, shape))
    r = depth_conv_block(p.sub("feature_refine.1"), depth_conv_block(p.sub("feature_refine.0"), up))
    return r + up


def _offset_diversity(p, x, aux, flow):
    # OffsetDiversity   lssvc_modules.py:75-112 (G=16 groups, 2 offsets, magnitude 40)
    G, O, mag = 16, 2, 40
    B, C, H, W = x.shape
    co = p.sub("conv_offset")
    out = lrelu(conv(co.sub("0"), aux, stride=2), 0.1)
    out = lrelu(conv(co.sub("2"), out), 0.1)
    out = up2(conv(co.sub("4"), out))
    o1, o2, mask = torch.chunk(out, 3, dim=1)
    mask = torch.sigmoid(mask)
    offset = mag * torch.tanh(torch.cat((o1, o2), dim=1)) + flow.repeat(1, G * O, 1, 1)
    offset = offset.view(B * G * O, 2, H, W)
    mask = mask.view(B * G * O, 1, H, W)
    xx = x.view(B * G, C // G, H, W).repeat(O, 1, 1, 1)
    xx = flow_warp(xx, offset) * mask
    return F.conv2d(xx.view(B, C * O, H, W), p["fusion.weight"], p["fusion.bias"], groups=G)


def _unet(p, x):
    # UNet   lssvc_modules.py:295-336
    x1 = depth_conv_block(p.sub("conv1"), x)
    x2 = depth_conv_block(p.sub("conv2"), F.max_pool2d(x1, 2))
    x3 = depth_conv_block(p.sub("conv3"), F.max_pool2d(x2, 2))
    for i in range(4):
        x3 = depth_conv_block(p.sub(f"context_refine.{i}"), x3)
    d3 = depth_conv_block(p.sub("up_conv3"), torch.cat((x2, subpel(p.sub("up3"), x3, padding=0)), dim=1))
    return depth_conv_block(p.sub("up_conv2"), torch.cat((x1, subpel(p.sub("up2"), d3, padding=0)), dim=1))


MASK_ORDER = ((0, 1, 2, 3), (3, 2, 1, 0), (2, 3, 0, 1), (1, 0, 3, 2))


def _masks(H, W, dtype):
    # get_mask_four_parts   LSSVC_net.py:298-325
    out = []
    for (i, j) in ((0, 0), (0, 1), (1, 0), (1, 1)):
        m = torch.zeros(1, 1, H, W, dtype=dtype)
        m[:, :, i::2, j::2] = 1
        out.append(m)
    return out


def four_part_prior(p, y, common_params):
    """forward_four_part_prior   LSSVC_net.py:338-443 (both the estimate outputs and the write=True outputs)."""
    masks = _masks(y.shape[2], y.shape[3], y.dtype)
    ys = y.chunk(4, 1)
    y_res = [torch.zeros_like(t) for t in ys]
    y_q = [torch.zeros_like(t) for t in ys]
    y_hat = [torch.zeros_like(t) for t in ys]
    s_hat = [torch.zeros_like(t) for t in ys]
    y_q_w, scales_w = [], []
    y_hat_so_far = None
    params8 = common_params
    for step in range(4):
        chunks = params8.chunk(8, 1)
        scales, means = chunks[:4], chunks[4:]
        cur, qw, sw = [], 0, 0
        for k in range(4):
            m = masks[MASK_ORDER[step][k]]
            sh = scales[k] * m
            mh = means[k] * m
            r = (ys[k] - mh) * m
            q = torch.round(r)
            h = q + mh
            y_res[k] = y_res[k] + r
            y_q[k] = y_q[k] + q
            y_hat[k] = y_hat[k] + h
            s_hat[k] = s_hat[k] + sh
            cur.append(h)
            qw = qw + q
            sw = sw + sh
        y_q_w.append(qw)
        scales_w.append(sw)
        cur = torch.cat(cur, dim=1)
        y_hat_so_far = cur if y_hat_so_far is None else y_hat_so_far + cur
        if step < 3:
            t = conv(p.sub(f"y_spatial_prior_adaptor_{step + 1}"), torch.cat((y_hat_so_far, common_params), dim=1),
                     padding=0)
            for i in range(3):
                t = depth_conv_block(p.sub(f"y_spatial_prior.{i}"), t)
            params8 = t
    return {"y_res": torch.cat(y_res, 1), "y_q": torch.cat(y_q, 1), "y_hat": torch.cat(y_hat, 1),
            "scales_hat": torch.cat(s_hat, 1), "y_q_w": y_q_w, "scales_w": scales_w}


def lssvc(sd, x_bl, x_el, dpb, shape_hr, scale_factor, pad_size=(0, 0, 0, 0)):
    """LSSVC.forward_one_frame   LSSVC_net.py:445-528."""
    p = Params(sd)
    bl = dmc(p.sub("base_layer_model"), x_bl, dpb["ref_frame_bl"], dpb["ref_feature_bl"])
    # inter-layer processing on the de-padded base-layer tensors   LSSVC_net.py:453-456
    texture_bl, mv_bl_hat, y_bl_hat = depad(bl["feature"], pad_size), depad(bl["mv_hat"], pad_size), depad(bl["y_hat"], pad_size, 16)
    ref_el, feat_el = dpb["ref_frame_el"], dpb["ref_feature_el"]

    mv_up = _mv_resampler(p.sub("mv_resampler"), mv_bl_hat, shape_hr, scale_factor)
    # mv_ctx_prior_encoder   LSSVC_net.py:108-116
    cp = p.sub("mv_ctx_prior_encoder")
    t = mv_up
    for i in (0, 2, 4):
        t = gdn_inter(cp.sub(str(i + 1)), conv(cp.sub(str(i)), t, stride=2))
    mv_ctx_prior = conv(cp.sub("6"), t, stride=2)
    # MVContextTransformer   lssvc_modules.py:497-508
    tr = p.sub("mv_ctx_transform.transform")
    mv_ctx = res_block(tr.sub("1"), conv(tr.sub("0"), mv_up, stride=2))

    mv = spynet(p.sub("optic_flow"), x_el, ref_el)
    # MVResEncoder   lssvc_modules.py:445-469
    e1, e2 = p.sub("mv_encoder.encoder1"), p.sub("mv_encoder.encoder2")
    f = gdn_inter(e1.sub("1"), conv(e1.sub("0"), mv, stride=2))
    f = lrelu(res_block(e1.sub("2"), f, start_from_relu=False), 0.1)
    f = torch.cat([f, mv_ctx], dim=1)
    for base in (0, 4):
        f = gdn_inter(e2.sub(str(base + 1)), conv(e2.sub(str(base)), f, stride=2))
        f = lrelu(res_block(e2.sub(str(base + 2)), f, start_from_relu=False), 0.1)
    mv_y = conv(e2.sub("8"), f, stride=2)
    mv_z = _prior_encoder(p.sub("mv_prior_encoder"), mv_y)
    mv_z_hat = torch.round(mv_z)
    # mv_prior_decoder   LSSVC_net.py:98-104
    pd = p.sub("mv_prior_decoder")
    hyper = lrelu(subpel(pd.sub("0"), mv_z_hat))
    hyper = lrelu(subpel(pd.sub("2"), hyper))
    hyper = conv(pd.sub("4"), hyper)
    pfu = p.sub("mv_prior_fusion")
    g = lrelu(conv(pfu.sub("0"), torch.cat([hyper, mv_ctx_prior], dim=1)))
    g = lrelu(conv(pfu.sub("2"), g))
    g = conv(pfu.sub("4"), g)
    mv_scales, mv_means = g.chunk(2, 1)
    mv_y_q = torch.round(mv_y - mv_means)
    mv_y_hat = mv_y_q + mv_means
    # MVResDecoder   lssvc_modules.py:472-494
    d1, d2 = p.sub("mv_decoder.decoder1"), p.sub("mv_decoder.decoder2")
    f = lrelu(subpel(d1.sub("0"), mv_y_hat), 0.1)
    f = gdn_inter(d1.sub("3"), res_block(d1.sub("2"), f, start_from_relu=False), inverse=True)
    f = gdn_inter(d1.sub("5"), subpel(d1.sub("4"), f), inverse=True)
    f = gdn_inter(d1.sub("7"), subpel(d1.sub("6"), f), inverse=True)
    f = lrelu(conv(d2.sub("0"), torch.cat([f, mv_ctx], dim=1)), 0.1)
    mv_hat = subpel(d2.sub("2"), f)

    # motion_compensation   LSSVC_net.py:229-244
    warp_frame = flow_warp(ref_el, mv_hat)
    mv2 = down2(mv_hat) / 2
    mv3 = down2(mv2) / 2
    if feat_el is None:
        f0 = conv(p.sub("feature_adaptor_EL_I"), ref_el)
    elif feat_el.shape[1] == 64:
        f0 = conv(p.sub("feature_adaptor_EL_first_P"), feat_el)
    else:
        f0 = conv(p.sub("feature_adaptor_EL"), feat_el)
    rf1, rf2, rf3 = _extractor3(p.sub("feature_extractor"), f0)
    c1_init = flow_warp(rf1, mv_hat)
    c1 = _offset_diversity(p.sub("align"), rf1, torch.cat((c1_init, warp_frame, mv_hat), dim=1), mv_hat)
    c2, c3 = flow_warp(rf2, mv2), flow_warp(rf3, mv3)
    fus = p.sub("context_fusion_net")
    tc1, tc2, tc3 = _fusion3(fus, c1, c2, c3)
    # hybrid_temporal_layer_context_fusion   LSSVC_net.py:246-259
    texture = _texture_resampler(p.sub("texture_resampler"), texture_bl, shape_hr)
    s1, s2, s3 = _extractor3(p.sub("texture_extractor"), texture)
    wg = p.sub("weight_map_generator")
    blended = []
    for i, (t_ctx, s_ctx) in enumerate(((tc1, s1), (tc2, s2), (tc3, s3)), start=1):
        gp = wg.sub(f"generator{i}")
        m = conv(gp.sub("0"), torch.cat([t_ctx, s_ctx], dim=1))
        m = res_block(gp.sub("1"), m, end_with_relu=True)
        w = torch.softmax(conv(gp.sub("2"), m), dim=1)
        w_t, w_s = w.chunk(2, 1)
        blended.append(t_ctx * w_t + s_ctx * w_s)
    c1, c2, c3 = _fusion3(fus, *blended)

    # ResEncoder   lssvc_modules.py:235-254
    re = p.sub("res_encoder")
    f = conv(re.sub("conv1"), torch.cat([x_el, c1], dim=1), stride=2)
    f = res_block(re.sub("res1"), torch.cat([f, c2], dim=1), slope=0.1, end_with_relu=True)
    f = conv(re.sub("conv2"), f, stride=2)
    f = res_block(re.sub("res2"), torch.cat([f, c3], dim=1), slope=0.1, end_with_relu=True)
    y = conv(re.sub("conv4"), conv(re.sub("conv3"), f, stride=2), stride=2)
    z = _prior_encoder(p.sub("res_prior_encoder"), y)
    z_hat = torch.round(z)
    # res_prior_decoder   LSSVC_net.py:63-73
    rd = p.sub("res_prior_decoder")
    h = lrelu(conv(rd.sub("0"), z_hat))
    h = lrelu(subpel(rd.sub("2"), h, padding=0))
    h = lrelu(conv(rd.sub("4"), h))
    h = lrelu(subpel(rd.sub("6"), h, padding=0))
    hier = conv(rd.sub("8"), h)
    te = p.sub("temporal_prior_encoder")
    temporal = conv(te.sub("2"), lrelu(conv(te.sub("0"), c3, stride=2), 0.1), stride=2)
    layer_prior = _layer_prior_resampler(p.sub("layer_prior_resampler"), y_bl_hat, (shape_hr[0] // 16, shape_hr[1] // 16))
    # PriorFusion   lssvc_modules.py:432-442
    pf = p.sub("prior_fusion_net.prior_fusion_conv")
    params = depth_conv_block(pf.sub("1"), depth_conv_block(pf.sub("0"), torch.cat([hier, temporal, layer_prior], dim=1)))
    fp = four_part_prior(p, y, params)
    y_hat = fp["y_hat"]

    # ResDecoder   lssvc_modules.py:257-276
    dd = p.sub("res_decoder")
    f = subpel(dd.sub("up2"), subpel(dd.sub("up1"), y_hat))
    f = res_block(dd.sub("res1"), torch.cat([f, c3], dim=1), slope=0.1, end_with_relu=True)
    f = subpel(dd.sub("up3"), f)
    f = res_block(dd.sub("res2"), torch.cat([f, c2], dim=1), slope=0.1, end_with_relu=True)
    rec_feat = subpel(dd.sub("up4"), f)
    # ReconGeneration (called as (recon_image_feature, context1))   lssvc_modules.py:279-292, LSSVC_net.py:492
    rg = p.sub("recon_generation_net")
    f = conv(rg.sub("first_conv"), torch.cat((rec_feat, c1), dim=1))
    feature = _unet(rg.sub("unet_2"), _unet(rg.sub("unet_1"), f))
    recon = conv(rg.sub("recon_conv"), feature)

    bits_el = (laplace_bits(fp["y_q"], fp["scales_hat"]) + laplace_bits(mv_y_q, mv_scales)
               + bitparm_bits(p.sub("bit_estimator_z"), z_hat) + bitparm_bits(p.sub("bit_estimator_z_mv"), mv_z_hat))
    return {
        "dpb": {"ref_frame_bl": bl["recon_image"], "ref_feature_bl": bl["feature"], "ref_frame_el": recon,
                "ref_feature_el": feature},
        "bit_bl": bl["bits"].item(), "bit_el": bits_el.item(), "mv_hat": mv_hat, "warp_frame": warp_frame,
        # internals for parity tests
        "bl": bl, "mv": mv, "mv_up": mv_up, "mv_y_q": mv_y_q, "mv_scales": mv_scales, "mv_z_hat": mv_z_hat,
        "z_hat": z_hat, "y": y, "params": params, "four_part": fp, "ctx": (c1, c2, c3),
    }
