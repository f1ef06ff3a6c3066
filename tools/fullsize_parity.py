"""1080p I-frame against the CPU oracle (the reference's algorithm in torch fp32; ~10-20 s on the box's host cores): how far
the latents of the fp32 CUDA-core engine and of the default split-fp16 tensor-core engine are from the oracle's before the
quantiser, and how many quantised symbols differ.  Explains the symbol-match figure at full size.
usage: python tools/fullsize_parity.py [H W]"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lssvc_b200 import IntraSS, ops, synth
from oracle import lssvc_oracle as orc

H, W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1152, 1920)
dev = torch.device("cuda:0")
torch.set_num_threads(os.cpu_count() or 8)
net = IntraSS(seed=0)
sd = {k: v.clone() for k, v in net.state_dict().items()}
net.to(dev)
net.set_scale_information(2.0, (H, W), (0, 0, 0, 0))
for seed in (3, 0):
    x_bl, x_el = synth.make_sequence(H, W, 1, seed=seed)[0]
    t0 = time.time()
    with torch.no_grad():
        o = orc.intra_ss(sd, x_bl, x_el, (H, W))
    print(f"== frames seed {seed}: oracle I-frame {H}x{W} in {time.time() - t0:.1f} s on {torch.get_num_threads()} threads", flush=True)
    q_ref = {"bl_z_hat": o["bl"]["z_hat"], "bl_y_q": torch.round(o["bl"]["y"] - o["bl"]["means"]), "z_hat": o["z_hat"],
             "y_q": torch.round(o["y"] - o["means"])}
    for engine in ("simt", ops.default_engine()):
        prev = ops.set_engine(engine)
        try:
            net._debug, net._force, net._force_flips = {}, q_ref, {}
            r = net.encode_decode(x_bl.to(dev), x_el.to(dev), None, None, H // 2, W // 2, H, W)
            dbg, flips = dict(net._debug), dict(net._force_flips)
        finally:
            ops.set_engine(prev)
            net._debug = net._force = None
        total = sum(v.numel() for v in q_ref.values())
        print(f"  {engine:5s} teacher-forced: {sum(flips.values())} of {total} symbols differ ({100 * sum(flips.values()) / total:.4f} %)  {flips}")
        for name, key, ref in (("BL y", "y_bl", o["bl"]["y"]), ("EL y", "y", o["y"])):
            got = dbg[key].to_nchw().cpu()
            d = (got - ref).abs()
            print(f"        {name}: |y| max {ref.abs().max():.2f} rms {ref.pow(2).mean().sqrt():.3f};  |err| max {d.max():.3e} mean {d.mean():.3e} "
                  f"-> expected flips ~ 2 * mean|err| = {200 * d.mean():.4f} %")
        for k in ("x_hat_bl", "x_hat_el"):
            print(f"        {k}: max|d| {(r[k].cpu() - o[k]).abs().max():.3e}")
        print(f"        bits {r['bit_bl']:.1f}/{r['bit_el']:.1f}  oracle {o['bit_bl']:.1f}/{o['bit_el']:.1f}", flush=True)
