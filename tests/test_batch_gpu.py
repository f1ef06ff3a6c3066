"""Batch > 1 in estimate mode (SURVEY 8b "Tensors": the reference's forward passes take [B, 3, H, W]; bits are summed over the
batch): a batch is coded item by item through the same kernels, so it must equal the per-item calls exactly."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_batch_two_equals_two_single_frames(cuda_device):
    from lssvc_b200 import IntraSS, LSSVC_extend, synth
    dev = cuda_device
    H = W = 128
    net_i, net_p = IntraSS(seed=0).to(dev), LSSVC_extend(seed=1).to(dev)
    for n in (net_i, net_p):
        n.set_scale_information(2.0, (H, W), (0, 0, 0, 0))
    seq_a, seq_b = synth.make_sequence(H, W, 2, seed=4), synth.make_sequence(H, W, 2, seed=5)
    xb = torch.cat([seq_a[0][0], seq_b[0][0]]).to(dev)
    xe = torch.cat([seq_a[0][1], seq_b[0][1]]).to(dev)
    single = [net_i.encode_decode(xb[i:i + 1], xe[i:i + 1], None, None, H // 2, W // 2, H, W) for i in range(2)]
    both = net_i.encode_decode(xb, xe, None, None, H // 2, W // 2, H, W)
    assert both["x_hat_el"].shape == (2, 3, H, W) and both["feature_el"].shape[0] == 2
    for k in ("x_hat_bl", "x_hat_el", "feature_el"):
        assert torch.equal(both[k], torch.cat([s[k] for s in single])), k
    for k in ("bit_bl", "bit_el"):
        assert abs(both[k] - (single[0][k] + single[1][k])) <= 1e-10 * both[k]      # (double-precision atomics: order noise)

    dpb = {"ref_frame_bl": both["x_hat_bl"].clamp(0, 1), "ref_frame_el": both["x_hat_el"].clamp(0, 1), "ref_feature_bl": None,
           "ref_feature_el": both["feature_el"].contiguous()}
    xb = torch.cat([seq_a[1][0], seq_b[1][0]]).to(dev)
    xe = torch.cat([seq_a[1][1], seq_b[1][1]]).to(dev)
    item = lambda d, i: {k: (None if v is None else v[i:i + 1]) for k, v in d.items()}
    single = [net_p.encode_decode(xb[i:i + 1], xe[i:i + 1], item(dpb, i), None, None, W, H, W // 2, H // 2) for i in range(2)]
    both = net_p.encode_decode(xb, xe, dpb, None, None, W, H, W // 2, H // 2)
    for k in ("ref_frame_bl", "ref_frame_el", "ref_feature_bl", "ref_feature_el"):
        assert both["dpb"][k].shape[0] == 2
        assert torch.equal(both["dpb"][k], torch.cat([s["dpb"][k] for s in single])), k
    assert both["mv_hat"].shape == (2, 2, H, W)
    for k in ("bit_bl", "bit_el"):
        assert abs(both[k] - (single[0][k] + single[1][k])) <= 1e-10 * both[k]      # (double-precision atomics: order noise)
    with pytest.raises(ValueError):          # bitstream mode codes one frame at a time
        net_p.forward_one_frame(xb, xe, None, None, None, None, _dpb=dpb, _write=object())
