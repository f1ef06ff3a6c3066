"""Host-logic check without a GPU: the full I-frame and P-frame graphs are walked on CPU-resident buffers through
the real C-ABI argument validation (shapes, channel counts, alignment, weight packing); only the final
"no CUDA device" outcome of each call is tolerated.  Catches wiring errors before any GPU time is spent."""
import pytest
import torch

from lssvc_b200 import _lib


@pytest.fixture()
def dry_run():
    _lib.DRY_RUN = True
    yield
    _lib.DRY_RUN = False


@pytest.mark.parametrize("size", [(128, 128), (256, 384)])
def test_graphs_validate_on_cpu(dry_run, size):
    from lssvc_b200 import IntraSS, LSSVC_extend
    H, W = size
    net_i, net_p = IntraSS(seed=0), LSSVC_extend(seed=1)
    x_bl, x_el = torch.rand(1, 3, H // 2, W // 2), torch.rand(1, 3, H, W)
    for n in (net_i, net_p):
        n.set_scale_information(2.0, (H, W), (0, 0, 0, 0))
    before = _lib.launch_count()
    r = net_i.encode_decode(x_bl, x_el, None, None, H // 2, W // 2, H, W)
    assert r["x_hat_el"].shape == (1, 3, H, W) and r["x_hat_bl"].shape == (1, 3, H // 2, W // 2)
    assert r["feature_el"].shape == (1, 64, H, W)
    dpb = {"ref_frame_bl": r["x_hat_bl"], "ref_frame_el": r["x_hat_el"], "ref_feature_bl": None,
           "ref_feature_el": r["feature_el"]}
    for _ in range(2):   # P after I (64-ch EL feature, no BL feature), then P after P (48-ch / 64-ch features)
        r = net_p.encode_decode(x_bl, x_el, dpb, None, None, W, H, W // 2, H // 2)
        dpb = r["dpb"]
        assert dpb["ref_feature_el"].shape == (1, 48, H, W) and dpb["ref_feature_bl"].shape == (1, 64, H // 2, W // 2)
        assert r["mv_hat"].shape == (1, 2, H, W) and r["warp_frame"].shape == (1, 3, H, W)
    assert _lib.launch_count() == before      # nothing ran: there is no CPU compute path


def test_models_refuse_cpu_outside_dry_run():
    from lssvc_b200 import IntraSS
    with pytest.raises(_lib.LssvcError):
        IntraSS(seed=0).forward(torch.zeros(1, 3, 64, 64), torch.zeros(1, 3, 128, 128))


def test_entropy_passes_are_routed_into_conv_epilogues(dry_run, monkeypatch):
    """Host wiring of the entropy epilogues, checked through the library's own argument validation on CPU buffers:
    a P-frame hands all 11 entropy passes (4 BitEstimator, 3 Laplace — one of them over two channel tiles —, 4 four-part steps) to
    the convolution that produces their parameters, narrow heads go to conv_head only above the size threshold, and with
    LSSVC_NO_ENT_FUSE the same graph runs conv + stand-alone kernel instead."""
    from lssvc_b200 import IntraSS, LSSVC_extend, ops
    H = W = 128
    net_i, net_p = IntraSS(seed=0), LSSVC_extend(seed=1)
    x_bl, x_el = torch.rand(1, 3, H // 2, W // 2), torch.rand(1, 3, H, W)
    for n in (net_i, net_p):
        n.set_scale_information(2.0, (H, W), (0, 0, 0, 0))
    r = net_i.encode_decode(x_bl, x_el, None, None, H // 2, W // 2, H, W)
    dpb = {"ref_frame_bl": r["x_hat_bl"], "ref_frame_el": r["x_hat_el"], "ref_feature_bl": None, "ref_feature_el": r["feature_el"]}

    def p_frame_trace():
        prev, ops.TRACE = ops.TRACE, []
        try:
            net_p.encode_decode(x_bl, x_el, dpb, None, None, W, H, W // 2, H // 2)
            return ops.TRACE
        finally:
            ops.TRACE = prev

    tr = p_frame_trace()
    fused = [t for t in tr if "e" in t.get("extras", "")]
    assert len(fused) == 11, [t["name"] for t in fused]
    names = [t["name"] for t in fused]
    assert sum(n.endswith("prior_encoder.4") for n in names) == 4                                   # z of both layers, mv + residual
    assert "base_layer_model.res_entropy_parameter.4" in names and "mv_prior_fusion.4" in names     # (2 x 96 and 2 x 64 parameters)
    assert sum(n.endswith(".block.1.conv.2") for n in names) == 4                                   # the four-part steps (ConvFFN tails)
    assert not any(t["engine"] == "head" for t in tr)                                               # 128 x 128: below the head threshold
    monkeypatch.setattr(ops, "HEAD_MIN_PIXELS", 0)
    assert sum(t["engine"] == "head" for t in p_frame_trace()) >= 8                                 # flow / reconstruction heads
    monkeypatch.setattr(ops, "ENT_FUSE", False)
    net_p._packs = {}                        # (the two-tile Laplace layer was packed with interleaved channels: pack it again)
    assert not any("e" in t.get("extras", "") for t in p_frame_trace())


def test_conv_head_supported_is_a_pure_host_predicate():
    from ctypes import byref

    from lssvc_b200 import ops
    from lssvc_b200._lib import CConv, load
    lib = load()

    def desc(k, cin, cout, stride=1, n_src=1, epi=0, ps=0):
        d = CConv()
        d.n_src = n_src
        buf = torch.zeros(8 * 8 * cin)
        v = ops.View(buf, 8, 8, cin, cin)
        d.src[0] = v.c()
        w = torch.zeros(k * k * 16 * cin)
        d.weight = w.data_ptr()
        d.kh = d.kw = k
        d.stride, d.pad, d.cout, d.n_pad, d.cin_total, d.epi, d.pixel_shuffle = stride, k // 2, cout, 16, cin, epi, ps
        return d, (buf, w)

    for args, want in (((3, 64, 2), 1), ((3, 48, 3), 1), ((3, 128, 4), 1), ((7, 16, 2), 1), ((7, 16, 3), 0), ((3, 64, 5), 0),
                       ((5, 64, 2), 0), ((3, 64, 2, 2), 0), ((3, 6, 2), 0), ((3, 64, 2, 1, 2), 0), ((3, 64, 2, 1, 1, 1), 0),
                       ((3, 64, 4, 1, 1, 0, 1), 0), ((7, 64, 2), 0)):
        d, keep = desc(*args)
        assert lib.lssvc_conv_head_supported(byref(d)) == want, args
    assert lib.lssvc_conv_head_supported(None) == 0


def test_batch_two_shapes(dry_run):
    """Batch > 1 (estimate mode): results are concatenated along the batch dimension, the DPB is split per item."""
    from lssvc_b200 import IntraSS, LSSVC_extend
    H = W = 128
    net_i, net_p = IntraSS(seed=0), LSSVC_extend(seed=1)
    for n in (net_i, net_p):
        n.set_scale_information(2.0, (H, W), (0, 0, 0, 0))
    x_bl, x_el = torch.rand(2, 3, H // 2, W // 2), torch.rand(2, 3, H, W)
    r = net_i.encode_decode(x_bl, x_el, None, None, H // 2, W // 2, H, W)
    assert r["x_hat_el"].shape == (2, 3, H, W) and r["x_hat_bl"].shape == (2, 3, H // 2, W // 2) and r["feature_el"].shape == (2, 64, H, W)
    dpb = {"ref_frame_bl": r["x_hat_bl"], "ref_frame_el": r["x_hat_el"], "ref_feature_bl": None, "ref_feature_el": r["feature_el"]}
    r = net_p.encode_decode(x_bl, x_el, dpb, None, None, W, H, W // 2, H // 2)
    assert r["dpb"]["ref_feature_el"].shape == (2, 48, H, W) and r["dpb"]["ref_feature_bl"].shape == (2, 64, H // 2, W // 2)
    assert r["mv_hat"].shape == (2, 2, H, W) and "_native" not in r["dpb"]
    assert isinstance(r["bit_bl"], float) and isinstance(r["bit_el"], float)
