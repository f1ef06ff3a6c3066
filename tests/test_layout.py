"""state_dict layout of the host-side models equals the reference's (golden: tools/make_golden.py)."""
import json
import os

import torch

from lssvc_b200 import nets

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _check(spec, golden):
    bag = nets.ParamBag(spec, seed=3)
    sd = bag.state_dict()
    assert [k for k, _, _ in golden] == list(sd.keys())
    for k, shape, dtype in golden:
        assert list(sd[k].shape) == shape, k
        assert str(sd[k].dtype) == dtype, k
    return bag


def test_intra_ss_layout():
    layout = json.load(open(os.path.join(GOLD, "state_dict_layout.json")))
    bag = _check(nets.intra_ss_spec(), layout["IntraSS"])
    assert len(layout["IntraSS"]) == 334
    assert abs(sum(p.numel() for p in bag.parameters()) / 1e6 - 31.8) < 0.1


def test_lssvc_layout():
    layout = json.load(open(os.path.join(GOLD, "state_dict_layout.json")))
    bag = _check(nets.lssvc_spec(), layout["LSSVC_extend"])
    assert len(layout["LSSVC_extend"]) == 926
    assert abs(sum(p.numel() for p in bag.parameters()) / 1e6 - 29.4) < 0.1


def test_init_is_deterministic_and_seeded():
    a = nets.ParamBag(nets.intra_ss_spec(), seed=5, gains=nets.model_gains("I")).state_dict()
    b = nets.ParamBag(nets.intra_ss_spec(), seed=5, gains=nets.model_gains("I")).state_dict()
    c = nets.ParamBag(nets.intra_ss_spec(), seed=6, gains=nets.model_gains("I")).state_dict()
    assert all(torch.equal(a[k], b[k]) for k in a)
    assert not torch.equal(a["g_a.conv1.weight"], c["g_a.conv1.weight"])
    assert len(nets.model_gains("I")) > 100 and len(nets.model_gains("P")) > 350
