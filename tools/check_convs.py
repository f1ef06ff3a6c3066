"""Debug aid: code I + P frames at a given size and, for every tensor-core conv, re-run it on the fp32 CUDA-core
kernel with the same inputs; print the layers whose outputs differ by more than a threshold."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, nargs=2, default=[256, 256])
    ap.add_argument("--thr", type=float, default=5e-5)
    ap.add_argument("--engine", default=None)
    ap.add_argument("--bias", action="store_true", help="print every layer's signed relative error against the fp32 kernel "
                    "(accumulator-truncation study, ops.acc_comp): < 0 = pulled towards zero")
    args = ap.parse_args()
    import torch
    from lssvc_b200 import IntraSS, LSSVC_extend, _lib, ops, synth
    from lssvc_b200.ops import View
    if args.engine:
        ops.set_engine(args.engine)
    dev = torch.device("cuda:0")
    H, W = args.size
    orig = ops.conv
    stats = {"n": 0, "bad": 0}

    def checked(pc, srcs, out, **kw):
        r = orig(pc, srcs, out, **kw)
        eng = kw.get("engine") or ops.default_engine()
        if eng == "simt" or kw.get("epi", 0) != 0:
            return r
        out2 = kw.get("out2")
        ref = View.alloc(out.H, out.W, out.C, out.device, zero=True)
        kw2 = dict(kw)
        kw2["engine"] = "simt"
        kw2["out2"] = None
        # residuals may alias the output in place: the reference run reads them after the first run wrote -> skip those
        aliased = any(v is not None and v.buf.data_ptr() == out.buf.data_ptr() for v in (kw.get("res1"), kw.get("res2")))
        if aliased:
            return r
        orig(pc, srcs, ref, **kw2)
        a, b = out.as_tensor(), ref.as_tensor()
        err = (a - b).abs().max().item() / (b.abs().max().item() + 1e-20)
        stats["n"] += 1
        if args.bias:
            big = b.abs() > b.abs().mean()
            sh = (((a - b) * b.sign())[big] / b.abs()[big]).mean().item()
            rms = ((a - b).pow(2).mean().sqrt() / (b.pow(2).mean().sqrt() + 1e-30)).item()
            steps = pc.kh * pc.kw * sum((c + 15) // 16 for c in pc.src_c)
            stats.setdefault("rows", []).append((steps, sh * 2 ** 24, rms))
            print(f"  {str(ops.TRACE_NAME):45s} k{pc.kh} s{pc.stride} cin{pc.src_c} cout{pc.cout} {out.H}x{out.W}: steps {steps:4d} "
                  f"signed rel err {sh * 2 ** 24:+7.2f} x 2^-24 ({sh * 2 ** 24 / steps:+.3f}/step)  rms rel {rms:.2e}", flush=True)
        if not (err < args.thr):
            stats["bad"] += 1
            d = (a - b).abs().amax(dim=2)
            ys, xs = torch.nonzero(d > args.thr * b.abs().max(), as_tuple=True)
            print(f"  {ops.TRACE_NAME:55s} k{pc.kh} s{pc.stride} cin{pc.src_c} cout{pc.cout} ps{int(pc.pixel_shuffle)} "
                  f"{out.H}x{out.W}: rel err {err:.2e}; bad pixels {len(ys)} rows {ys.min().item()}..{ys.max().item()} "
                  f"cols {xs.min().item()}..{xs.max().item()}", flush=True)
        return r

    ops.conv = checked
    net_i, net_p = IntraSS(seed=0).to(dev), LSSVC_extend(seed=1).to(dev)
    for n in (net_i, net_p):
        n.set_scale_information(2.0, (H, W), (0, 0, 0, 0))
    frames = synth.make_sequence(H, W, 3, seed=0)
    (b0, e0) = frames[0]
    print("I-frame")
    r = net_i.encode_decode(b0.to(dev), e0.to(dev), None, None, H // 2, W // 2, H, W)
    dpb = {"ref_frame_bl": r["x_hat_bl"].clamp_(0, 1), "ref_frame_el": r["x_hat_el"].clamp_(0, 1), "ref_feature_bl": None,
           "ref_feature_el": r["feature_el"]}
    for i in (1, 2):
        print(f"P-frame {i}")
        b, e = frames[i]
        r = net_p.encode_decode(b.to(dev), e.to(dev), dpb, None, None, W, H, W // 2, H // 2)
        dpb = r["dpb"]
    print(f"checked {stats['n']} convs, {stats['bad']} above {args.thr}")
    if args.bias and stats.get("rows"):
        import numpy as np
        rows = np.array(stats["rows"])
        k = np.polyfit(rows[:, 0], rows[:, 1], 1)
        print(f"signed error / 2^-24 over {len(rows)} layers: mean {rows[:, 1].mean():+.2f}, per step {np.mean(rows[:, 1] / rows[:, 0]):+.4f}, "
              f"least-squares {k[0]:+.4f} * steps {k[1]:+.3f}; mean rms rel {rows[:, 2].mean():.3e}")


if __name__ == "__main__":
    main()
