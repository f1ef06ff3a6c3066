"""bench.py — two-layer 1080p frames/sec of the LSSVC coding forward pass on N B200s (BASELINE.json metric).

A "step" is the two-layer (BL 960x540 + EL 1920x1080, padded to 576x960 / 1152x1920) coding of ONE frame through the
public model API (`IntraSS.encode_decode` on GOP boundaries, `LSSVC_extend.encode_decode` otherwise, --write_stream 0
semantics: estimated bitrate).  GOPs are independent, so with N > 1 every rank codes its own synthetic sequence
(GOP-sharded, weak scaling) and NCCL is used once, to gather the per-frame rate/distortion rows.

  value : frames/s with the frames already resident in HBM (device tensors in, device tensors out, bits read back)
  e2e   : frames/s through the same API from PINNED HOST buffers: H2D of (x_bl, x_el) and D2H of both reconstructions
          + bits inside the timed region — what test.py's frame loop does (test.py:185-191, 260-263); the copies run on
          two copy streams, overlapped with the coding of the neighbouring frames (software pipelining, all inside the
          timed region)
  --impl reference : the reference's algorithm on the host CPU (oracle port, all host threads), bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# conv FLOPs per padded EL pixel
FLOP_PER_PX_P = 5.2160e6
FLOP_PER_PX_I = 2.2775e6
SIZES = {"1080p": (1080, 1920), "cfg1": (320, 512), "4k": (2160, 3840), "tiny": (128, 128)}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"], "bf16_tflops_sustained": p.get("bf16_tflops_sustained"),
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def frame_is_intra(idx, gop):
    return idx % gop == 0


def make_frames(hw, n, seed):
    from lssvc_b200 import synth
    pad = synth.interlayer_padding(hw[0], hw[1], 2.0)
    H, W = pad["HR_padded_size"]
    return synth.make_sequence(H, W, n, seed=seed), (H, W)


class Coder:
    """The reference's frame loop (test.py:183-250) on the CUDA models."""

    def __init__(self, device, shape_hr):
        from lssvc_b200 import IntraSS, LSSVC_extend
        self.dev = device
        self.H, self.W = shape_hr
        self.net_i = IntraSS(seed=0).to(device)
        self.net_p = LSSVC_extend(seed=1).to(device)
        for n in (self.net_i, self.net_p):
            n.set_scale_information(2.0, shape_hr, (0, 0, 0, 0))
        self.dpb = None

    def step(self, idx, gop, x_bl, x_el):
        H, W = self.H, self.W
        if frame_is_intra(idx, gop):
            r = self.net_i.encode_decode(x_bl, x_el, None, None, H // 2, W // 2, H, W)
            self.dpb = {"ref_frame_bl": r["x_hat_bl"], "ref_frame_el": r["x_hat_el"], "ref_feature_bl": None,
                        "ref_feature_el": r["feature_el"]}
        else:
            r = self.net_p.encode_decode(x_bl, x_el, self.dpb, None, None, W, H, W // 2, H // 2)
            self.dpb = r["dpb"]
        self.dpb["ref_frame_bl"].clamp_(0, 1)
        self.dpb["ref_frame_el"].clamp_(0, 1)
        return r["bit_bl"], r["bit_el"], self.dpb["ref_frame_bl"], self.dpb["ref_frame_el"]


def timed_loop(torch, dist, world, fn, first, count, finish=None):
    """barrier + synchronize, CUDA events around exactly `count` steps on the launching stream, max over ranks.
    finish: called before the closing event (the e2e loop makes the launching stream wait for its copy streams there, so
    that every copy of the timed steps lies between the two events)."""
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rows = [fn(first + i) for i in range(count)]
    if finish is not None:
        finish()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
        dist.barrier()
    return ms, rows


def conv_roofline(torch, device, shape_hr, peaks):
    """The dominant kernel: the 3x3 64->64 conv (28 % of all FLOPs) at half EL resolution, timed
    alone with CUDA events on the launching stream; inputs (2 x 141 MB) exceed nothing but are rotated over 4 buffers
    so that consecutive launches do not re-read an L2-resident tensor."""
    from lssvc_b200 import ops
    H, W = shape_hr[0] // 2, shape_hr[1] // 2
    g = torch.Generator().manual_seed(0)
    w = torch.randn(64, 64, 3, 3, generator=g) / 24.0
    b = torch.zeros(64)
    pc = ops.PackedConv(w, b, device=device)
    srcs = [ops.View(torch.randn(H * W * 64, device=device), H, W, 64, 64) for _ in range(4)]
    out = ops.View.alloc(H, W, 64, device)
    engine = ops.default_engine()
    for i in range(4):
        ops.conv(pc, srcs[i], out, act=0.01, engine=engine)
    torch.cuda.synchronize()
    n = 24
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        ops.conv(pc, srcs[i % 4], out, act=0.01, engine=engine)
    e1.record()
    torch.cuda.synchronize()
    sec = e0.elapsed_time(e1) / 1e3 / n
    flops = 2.0 * H * W * 64 * 64 * 9
    info = ops.engine_info(engine)
    peak = peaks["bf16_tflops"] * info["peak_vs_bf16"]
    achieved = flops / sec / 1e12
    # DRAM bytes per launch of this very layer from the committed ncu --set full capture (profiles/), if there is one
    traffic = None
    try:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "conv_roofline_traffic.json")) as f:
            t = json.load(f)
        if t.get("kernel") == info["kernel"] and t.get("shape") == [H, W, 64, 64, 3]:
            traffic = int(t["dram_bytes_read"] + t["dram_bytes_write"])
    except (OSError, ValueError, KeyError):
        pass
    return {"bound": "tensor", "kernel": info["kernel"], "achieved": round(achieved, 2), "peak": round(peak, 1), "unit": "TFLOP/s",
            "frac": round(achieved / peak, 4), "traffic": traffic,
            "note": f"{info['note']}; peak = {info['peak_vs_bf16']} x bf16 dense GEMM peak ({peaks['source']}); "
                    f"3x3 64->64 conv at {H}x{W}, {flops / 1e9:.1f} GFLOP (algorithmic) per launch, {sec * 1e3:.3f} ms per launch"}


def cpu_reference(steps, warmup, hw_name, as_line, n_gpus=1):
    """The reference's algorithm on the host CPU (oracle port of the PyTorch path, fp32, --write_stream 0), all host
    threads, on a bounded sample: the I/P/P... chain on a 384x512 (padded) crop, scaled by the pixel ratio."""
    import torch
    from lssvc_b200 import nets, synth
    from oracle import lssvc_oracle as orc
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    full = synth.interlayer_padding(*SIZES[hw_name], 2.0)["HR_padded_size"]
    H, W = (384, 512) if full[0] * full[1] > 384 * 512 else full
    n = warmup + steps
    frames = synth.make_sequence(H, W, n, seed=0)
    sd_i = nets.ParamBag(nets.intra_ss_spec(), seed=0, gains=nets.model_gains("I")).state_dict()
    sd_p = nets.ParamBag(nets.lssvc_spec(), seed=1, gains=nets.model_gains("P")).state_dict()
    dpb, times = None, []
    with torch.no_grad():
        for idx in range(n):
            x_bl, x_el = frames[idx]
            t0 = time.perf_counter()
            if idx == 0:
                o = orc.intra_ss(sd_i, x_bl, x_el, (H, W))
                dpb = {"ref_frame_bl": o["x_hat_bl"], "ref_frame_el": o["x_hat_el"], "ref_feature_bl": None,
                       "ref_feature_el": o["feature_el"]}
            else:
                o = orc.lssvc(sd_p, x_bl, x_el, dpb, (H, W), 2.0)
                dpb = o["dpb"]
            dpb["ref_frame_bl"] = dpb["ref_frame_bl"].clamp_(0, 1)
            dpb["ref_frame_el"] = dpb["ref_frame_el"].clamp_(0, 1)
            times.append(time.perf_counter() - t0)
    timed = times[warmup:]
    scale = (full[0] * full[1]) / float(H * W)
    sec_per_frame = sum(timed) / len(timed) * scale
    base = {"value": round(1.0 / sec_per_frame, 5), "unit": "frames/s", "cores": cores, "kind": "port",
            "sample": f"{len(timed)} P-frames (after {warmup} warm-up frames, I first) on a {W}x{H} padded crop, time x {scale:.2f} "
                      f"(pixel ratio to {full[1]}x{full[0]}); oracle port of the reference PyTorch path, fp32, torch threads={cores}"}
    if not as_line:
        return base
    return {"metric": "two-layer 1080p frames/sec (est. bitrate)", "value": base["value"], "unit": "frames/s", "n_gpus": n_gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": round(sec_per_frame * 1e3, 1), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": {"workload": f"LSSVC two-layer {hw_name} (EL {full[1]}x{full[0]} padded, BL half), host CPU",
                       "parallelism": "cpu"},
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", default="1080p", choices=sorted(SIZES))
    ap.add_argument("--gop", type=int, default=12)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    W_ = max(args.warmup, 3)

    if args.impl == "reference":
        if rank == 0:
            print(json.dumps(cpu_reference(min(args.steps, 3), 1, args.size, True, n_gpus=args.gpus)), flush=True)
        return

    # the contract is ONE JSON line on stdout: native libraries (NCCL prints its version banner) write to fd 1 directly,
    # so fd 1 is pointed at stderr for the whole run and the line goes out through the saved descriptor
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    from lssvc_b200 import _lib, ops
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    _lib.check(_lib.load().lssvc_device_check(local), "device_check")
    peaks = load_peaks()

    n_frames = W_ + args.steps
    frames, shape_hr = make_frames(SIZES[args.size], n_frames, seed=rank)      # every rank: its own sequence / GOPs
    H, W = shape_hr
    host = [(b.pin_memory(), e.pin_memory()) for b, e in frames]
    devf = [(b.to(dev), e.to(dev)) for b, e in host]
    coder = Coder(dev, shape_hr)

    stats = []

    def step_resident(idx):
        x_bl, x_el = devf[idx]
        bb, be, rb, re = coder.step(idx, args.gop, x_bl, x_el)
        return (idx, bb, be)

    out_bl = torch.empty(1, 3, H // 2, W // 2).pin_memory()
    out_el = torch.empty(1, 3, H, W).pin_memory()

    # The e2e loop is software-pipelined the way a real encoder front end is: the pinned-host -> device copy of frame t + 1
    # runs on a copy stream while frame t is coded, the device -> pinned-host copy of frame t's reconstructions runs on a
    # second copy stream while frame t + 1 is coded; the bit counts of frame t are read (a device -> host sync) before
    # step t returns.  Every copy of a timed step is issued inside the timed region and completes before its closing event.
    in_stream, out_stream = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    in_bufs = [(torch.empty_like(devf[0][0]), torch.empty_like(devf[0][1])) for _ in range(2)]
    in_free = [torch.cuda.Event(), torch.cuda.Event()]
    in_ready = {}
    out_done = torch.cuda.Event()
    out_keep = []

    def prefetch(idx):
        if idx in in_ready or idx >= len(host):
            return
        slot = idx % 2
        in_stream.wait_event(in_free[slot])          # the frame that last used the slot has been coded
        with torch.cuda.stream(in_stream):
            in_bufs[slot][0].copy_(host[idx][0], non_blocking=True)
            in_bufs[slot][1].copy_(host[idx][1], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(in_stream)
        in_ready[idx] = ev

    def step_host(idx):
        cur = torch.cuda.current_stream()
        prefetch(idx)
        cur.wait_event(in_ready.pop(idx))
        prefetch(idx + 1)
        x_bl, x_el = in_bufs[idx % 2]
        bb, be, rb, re = coder.step(idx, args.gop, x_bl, x_el)      # reads the bit counters: syncs the coding stream
        in_free[idx % 2].record(cur)
        out_stream.wait_stream(cur)
        with torch.cuda.stream(out_stream):
            out_bl.copy_(rb, non_blocking=True)
            out_el.copy_(re, non_blocking=True)
            out_done.record(out_stream)
        out_keep[:] = [rb, re]                       # keep the sources alive until the copies have run
        return (idx, bb, be)

    def e2e_finish():
        torch.cuda.current_stream().wait_event(out_done)
        in_ready.clear()

    # ---- device-resident throughput ("value") ---------------------------------------------------------------
    for i in range(W_):
        step_resident(i)
    l0 = _lib.launch_count()
    with ClockSampler(local) as clocks:
        ms, rows = timed_loop(torch, dist, world, step_resident, W_, args.steps)
    launches = _lib.launch_count() - l0
    # ---- end to end from pinned host memory ("e2e") ------------------------------------------------------------
    coder.dpb = None
    for i in range(W_):
        step_host(i)
    torch.cuda.synchronize()
    in_ready.clear()                                 # the first timed step copies its own input inside the timed region
    ms_e2e, rows_e2e = timed_loop(torch, dist, world, step_host, W_, args.steps, finish=e2e_finish)

    # ---- rate statistics gathered over NCCL (the only collective of the path) ---------------------------------
    t = torch.tensor(rows, dtype=torch.float64, device=dev)
    if world > 1:
        gathered = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(gathered, t)
        t = torch.cat(gathered, 0)
    if rank == 0:
        px = SIZES[args.size][0] * SIZES[args.size][1]
        n_i = sum(1 for i in range(W_, n_frames) if frame_is_intra(i, args.gop))
        flop = (n_i * FLOP_PER_PX_I + (args.steps - n_i) * FLOP_PER_PX_P) * H * W
        fps = world * args.steps / (ms / 1e3)
        fps_e2e = world * args.steps / (ms_e2e / 1e3)
        line = {
            "metric": "two-layer 1080p frames/sec (est. bitrate)", "value": round(fps, 4), "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": W_, "ms_per_step": round(ms / args.steps, 2), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": ops.engine_info(ops.default_engine())["dtype"], "data": "synthetic",
            "config": {"workload": f"LSSVC two-layer {args.size} (BL {W // 2}x{H // 2}, EL {W}x{H} padded), IP{args.gop}, "
                                   f"{args.steps} frames/GPU ({n_i} I + {args.steps - n_i} P), estimated bitrate, random-init synthetic weights",
                       "parallelism": f"gop-sharded x{world}", "conv_engine": ops.default_engine(),
                       "l2": "per-frame working set (GBs of fp32 activations) >> 126 MB L2; inputs differ every step"},
            "e2e": {"value": round(fps_e2e, 4), "unit": "frames/s", "h2d_bytes_per_step": int(host[0][0].numel() + host[0][1].numel()) * 4,
                    "d2h_bytes_per_step": int(out_bl.numel() + out_el.numel()) * 4 + 16},
            "gpu_launches": int(launches),
            "clocks": clocks.summary(),
            "conv_tflops": round(world * flop / (ms / 1e3) / 1e12, 2),
            "mean_bpp": {"bl": round(float(t[:, 1].mean()) / (px / 4), 5), "el": round(float(t[:, 2].mean()) / px, 5)},
        }
        line["roofline"] = conv_roofline(torch, dev, shape_hr, peaks)
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_reference(2, 1, args.size, False)
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
