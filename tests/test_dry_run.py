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
