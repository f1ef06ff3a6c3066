"""Run one I-frame + P-frame with every h2 conv executed twice (TMA-store epilogue on / off) and report mismatches."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lssvc_b200 import IntraSS, LSSVC_extend, _lib, ops, synth

dev = torch.device("cuda:0")
orig = ops.conv
bad = []
count = [0]

def conv2(pc, srcs, out, **kw):
    if isinstance(srcs, ops.View):
        srcs = [srcs]
    eng = kw.get("engine") or ops.default_engine()
    if eng != "h2":
        return orig(pc, srcs, out, **kw)
    o2 = kw.get("out2")
    before = out.as_tensor().clone()
    os.environ["LSSVC_H2_NOTMA"] = "1"
    orig(pc, srcs, out, **kw)
    torch.cuda.synchronize()
    ref = out.as_tensor().clone()
    ref2 = o2.as_tensor().clone() if o2 is not None else None
    out.as_tensor().copy_(before)
    del os.environ["LSSVC_H2_NOTMA"]
    orig(pc, srcs, out, **kw)
    torch.cuda.synchronize()
    got = out.as_tensor()
    count[0] += 1
    d = (got - ref).abs().max().item() if torch.isfinite(got).all() else float("nan")
    d2 = 0.0
    if o2 is not None:
        d2 = (o2.as_tensor() - ref2).abs().max().item()
    if not (d == 0.0 and d2 == 0.0):
        info = dict(name=ops.TRACE_NAME, k=pc.kh, stride=pc.stride, src=[(s.H, s.W, s.C, s.pitch, s.coff) for s in srcs],
                    cout=pc.cout, ps=pc.pixel_shuffle, out=(out.H, out.W, out.C, out.pitch, out.coff),
                    res1=kw.get("res1") is not None, res2=kw.get("res2") is not None, out2=o2 is not None, d=d, d2=d2,
                    epi=kw.get("epi"))
        bad.append(info)
        print("MISMATCH", info, flush=True)
    return out

ops.conv = conv2
import lssvc_b200.engine as E
H = W = 256
net_i, net_p = IntraSS(seed=0).to(dev), LSSVC_extend(seed=1).to(dev)
for n in (net_i, net_p):
    n.set_scale_information(2.0, (H, W), (0, 0, 0, 0))
frames = synth.make_sequence(H, W, 2, seed=0)
(b0, e0), (b1, e1) = frames
r = net_i.encode_decode(b0.to(dev), e0.to(dev), None, None, H // 2, W // 2, H, W)
dpb = {"ref_frame_bl": r["x_hat_bl"], "ref_frame_el": r["x_hat_el"], "ref_feature_bl": None, "ref_feature_el": r["feature_el"]}
r = net_p.encode_decode(b1.to(dev), e1.to(dev), dpb, None, None, W, H, W // 2, H // 2)
print(f"{count[0]} h2 convs checked, {len(bad)} mismatches")
