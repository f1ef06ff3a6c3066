"""bench.py — two-layer 1080p frames/sec of the LSSVC coding forward pass on N B200s (BASELINE.json metric).

The measured thing is the product's job runner (lssvc_b200/runner.py = the reference's frame loop test.py:121-325 + fan-out
:685-748): the job is cut into (sequence, GOP) work units (gop.work_units), a rank takes its round-robin share (gop.shard),
codes it on `--lanes` concurrent coding lanes (independent GOPs in flight on one GPU) through the public model
entry points (`IntraSS.forward` on GOP boundaries, `LSSVC_extend.forward_one_frame` otherwise; --write_stream 0 semantics:
estimated bitrate), and the per-frame rows [seq, frame, is_intra, bits_bl, bits_el, sse_bl, sse_el] are gathered over NCCL
(gop.gather_stats) — the only collective of the path.

  N = 1 : BASELINE config 2 — one 96-frame 1080p sequence per lane, IP12 (BL 960x540 + EL 1920x1080, padded to 576x960 /
          1152x1920)
  N > 1 : BASELINE config 3 — 8 sequences x 96 frames, IP32: 24 work units dealt 12 / 6 / 3 per rank at 2 / 4 / 8 GPUs
  A STEP is one round of the runner: one frame on every lane.  --steps K times exactly K rounds (the job is cut after the
  timed rounds unless --full-job); with a fixed K per rank the per-GPU work is fixed as N grows (weak scaling).

  value : frames/s with the frames already resident in HBM
  e2e   : the same job from PINNED HOST frames: H2D of (x_bl, x_el) and D2H of both reconstructions + the stat rows of every
          frame inside the timed region, on the frame's own lane (copies of one lane overlap the coding of the others)
  roofline : the dominant kernel (conv_hs) — the headline layer timed alone, the in-frame average over every conv_hs launch
          of a P-frame (CUDA events around each launch), both against the measured bf16 peak and against the kernel's own
          3-MMA ceiling
  --impl reference : the reference's algorithm on the host CPU (oracle port, all host threads), full-size frames.
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
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw,power.limit")
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
        def col(i):
            out = []
            for r in self.rows:
                try:
                    out.append(float(r[i]))
                except (IndexError, ValueError):
                    pass
            return sorted(out)
        draw, limit = col(6), col(7)
        # power: the board runs this load AT its power limit (sw_power_cap): frames/s follows energy per frame, not idle gaps
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm), "power_w": round(draw[len(draw) // 2], 1) if draw else None,
                "power_limit_w": round(limit[-1], 1) if limit else None}


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


def _oracle_chain(H, W, n, threads, seed=0):
    """I + (n - 1) P frames of the oracle at (H, W) on `threads` torch threads; returns the per-frame seconds."""
    import torch
    from lssvc_b200 import nets, synth
    from oracle import lssvc_oracle as orc
    torch.set_num_threads(threads)
    frames = synth.make_sequence(H, W, n, seed=seed)
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
            del o
    return times


def _worker_entry(args):
    H, W, n, seed = args
    return _oracle_chain(H, W, n, 1, seed)


def cpu_worker_mode(hw_name, gop):
    """The reference's OWN parallel mode (test.py:642 `torch.set_num_threads(1)`, :686 `ProcessPoolExecutor(max_workers=
    --worker)`): one single-thread process per host core, each coding its own sequence.  Bounded sample: every worker codes
    I + 1 P at config-1 size (384x512 padded); the P-frame time is scaled by the pixel ratio to the named size."""
    import concurrent.futures as cf
    import multiprocessing as mp
    from lssvc_b200 import synth
    cores = os.cpu_count() or 1
    full = synth.interlayer_padding(*SIZES[hw_name], 2.0)["HR_padded_size"]
    H, W = 384, 512
    t0 = time.perf_counter()
    with cf.ProcessPoolExecutor(max_workers=cores, mp_context=mp.get_context("spawn")) as ex:
        res = list(ex.map(_worker_entry, [(H, W, 2, s) for s in range(cores)]))
    wall = time.perf_counter() - t0
    scale = (full[0] * full[1]) / float(H * W)
    t_i = sum(r[0] for r in res) / cores * scale
    t_p = sum(r[1] for r in res) / cores * scale
    per_gop = t_i + (gop - 1) * t_p
    return {"value": round(cores * gop / per_gop, 5), "unit": "frames/s", "workers": cores, "threads_per_worker": 1,
            "sample": f"{cores} single-thread worker processes (test.py:642,686), each I + 1 P at {W}x{H} padded ({wall:.0f} s wall); "
                      f"I {t_i:.0f} s / P {t_p:.0f} s per frame and worker after x{scale:.2f} (pixel ratio to {full[1]}x{full[0]}), IP{gop} mix"}


def cpu_reference(steps, warmup, hw_name, gop, as_line, n_gpus=1, worker_mode=True):
    """The reference's algorithm on the host CPU (oracle port of the PyTorch path, fp32, --write_stream 0) at the FULL padded
    size of the named configuration: mode (i) one process on all host threads codes I + P + ... (a bounded sample: at most 3
    frames), the frame rate of the IP<gop> mix follows from the measured I and P times; mode (ii) the reference's own
    `--worker N` single-thread processes (cpu_worker_mode)."""
    from lssvc_b200 import synth
    cores = os.cpu_count() or 1
    H, W = synth.interlayer_padding(*SIZES[hw_name], 2.0)["HR_padded_size"]
    n = 1 + max(1, min(steps, 2))
    times = _oracle_chain(H, W, n, cores)
    t_i, t_p = times[0], sum(times[1:]) / len(times[1:])
    sec_per_frame = (t_i + (gop - 1) * t_p) / gop
    base = {"value": round(1.0 / sec_per_frame, 5), "unit": "frames/s", "cores": cores, "kind": "port",
            "sample": f"1 I-frame ({t_i:.1f} s) + {n - 1} P-frame(s) ({t_p:.1f} s each) at the full padded size {W}x{H} / {W // 2}x{H // 2}, "
                      f"IP{gop} mix = (I + {gop - 1} P) / {gop}; oracle port of the reference PyTorch path, fp32, one process, torch threads={cores}"}
    if worker_mode:
        base["worker_mode"] = cpu_worker_mode(hw_name, gop)
    if not as_line:
        return base
    return {"metric": "two-layer 1080p frames/sec (est. bitrate)", "value": base["value"], "unit": "frames/s", "n_gpus": n_gpus,
            "steps": n - 1, "warmup": 1, "ms_per_step": round(sec_per_frame * 1e3, 1), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": {"workload": f"LSSVC two-layer {hw_name} (BL {W // 2}x{H // 2}, EL {W}x{H} padded), IP{gop}, estimated bitrate, "
                                   f"random-init synthetic weights, host CPU", "parallelism": "cpu"},
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def in_frame_roofline(torch, coder, frames_dev, peaks, gop):
    """One steady-state P-frame coded eagerly with CUDA events around every operator launch (lssvc_b200/profile.py):
    algorithmic FLOPs of the conv_hs launches / their summed time = the in-frame rate of the dominant kernel."""
    from lssvc_b200 import profile
    coder.dpb = None
    for idx in range(3):                                   # I, P, P: the third frame is a P-after-P frame
        coder.step(idx, gop, *frames_dev[idx])
    with profile.LaunchTimer() as t:
        coder.step(3, gop, *frames_dev[3])
    rows = t.rows()
    fam = profile.conv_summary(rows)
    total_ms = sum(r[2] for r in rows)
    hs = fam.get("hs", {"launches": 0, "ms": 0.0, "flops": 0.0})
    out = {"p_frame_launch_ms": round(total_ms, 2), "p_frame_launches": len(rows)}
    if hs["ms"] > 0:
        tf = hs["flops"] / hs["ms"] / 1e9
        sustained = peaks.get("bf16_tflops_sustained") or peaks["bf16_tflops"]
        out.update({"kernel": "conv_hs_kernel", "launches": hs["launches"], "ms": round(hs["ms"], 2), "achieved": round(tf, 1),
                    "unit": "TFLOP/s", "frac_of_bf16_sustained": round(tf / sustained, 4),
                    "of_kernel_ceiling": round(tf / (sustained / 3.0), 4),
                    "share_of_frame": round(hs["ms"] / total_ms, 4)})
    for k in ("ffn", "pw", "head", "simt"):
        if k in fam and fam[k]["ms"] > 0:
            out[k] = {"launches": fam[k]["launches"], "ms": round(fam[k]["ms"], 2), "tflops": round(fam[k]["flops"] / fam[k]["ms"] / 1e9, 1)}
    other = {}
    for name, info, ms in rows:
        if name not in ("conv", "ffn", "pw"):
            other[name] = round(other.get(name, 0.0) + ms, 3)
    out["other_ms"] = dict(sorted(other.items(), key=lambda kv: -kv[1]))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="timed rounds (one frame per lane each); default: 96 frames / lanes")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", default="1080p", choices=sorted(SIZES))
    ap.add_argument("--gop", type=int, default=None, help="intra period; default 12 at N = 1 (config 2), 32 at N > 1 (config 3)")
    ap.add_argument("--lanes", type=int, default=int(os.environ.get("LSSVC_LANES", "1")),
                    help="concurrent GOPs per GPU (SURVEY H9; measured: 16.8 / 16.4 / 16.4 frames/s at 1 / 2 / 3 lanes, so 1 is the default)")
    ap.add_argument("--frames", type=int, default=96)
    ap.add_argument("--seqs", type=int, default=None, help="sequences of the job; default: lanes at N = 1, 8 at N > 1")
    ap.add_argument("--full-job", action="store_true", help="code this rank's whole share (strong split of the job) instead of K rounds")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graphs", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    W_ = max(args.warmup, 3)
    gop_size = args.gop or (12 if world == 1 else 32)
    lanes = max(1, args.lanes)
    K = args.steps if args.steps is not None else max(1, args.frames // lanes)

    if args.impl == "reference":
        if rank == 0:
            print(json.dumps(cpu_reference(K, 1, args.size, gop_size, True, n_gpus=args.gpus)), flush=True)
        return

    # the contract is ONE JSON line on stdout: native libraries (NCCL prints its version banner) write to fd 1 directly,
    # so fd 1 is pointed at stderr for the whole run and the line goes out through the saved descriptor
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    from lssvc_b200 import _lib, gop, ops
    from lssvc_b200.runner import GopRunner
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    _lib.check(_lib.load().lssvc_device_check(local), "device_check")
    peaks = load_peaks()

    # ---- the job: work units of the named configuration, this rank's share ------------------------------------------
    n_seq = args.seqs or (lanes if world == 1 else 8)
    if world > 1 and args.seqs is None and n_seq * ((args.frames + gop_size - 1) // gop_size) < world * lanes:
        n_seq = world * lanes                           # every lane of every rank needs a unit
    units = gop.work_units(n_seq, args.frames, gop_size)
    mine = gop.shard(units, world, rank)
    # synthetic frames: a pool of whole GOPs resident in HBM (and pinned on the host for the e2e arm); frame f of sequence s
    # is pool[(f + gop * (s + rank)) % len]: a GOP never straddles the wrap, concurrent lanes read different frames
    pool_len = gop_size * max(1, 24 // gop_size)
    pool, shape_hr = make_frames(SIZES[args.size], pool_len, seed=rank)
    H, W = shape_hr
    host = [(b.pin_memory(), e.pin_memory()) for b, e in pool]
    devf = [(b.to(dev), e.to(dev)) for b, e in host]
    where = lambda seq, f: (f + gop_size * (seq + rank)) % pool_len
    coder = Coder(dev, shape_hr)
    runner = GopRunner(coder.net_i, coder.net_p, lanes=lanes, graphs=(lanes > 1 and not args.no_graphs))

    out_host = [(torch.empty(1, 3, H // 2, W // 2).pin_memory(), torch.empty(1, 3, H, W).pin_memory()) for _ in range(lanes)]

    def src_resident(seq, f):
        return devf[where(seq, f)]

    def src_host(seq, f):                              # called under the lane's stream: the H2D copy is part of the lane's work
        hb, he = host[where(seq, f)]
        return hb.to(dev, non_blocking=True), he.to(dev, non_blocking=True)

    def d2h(lane, unit, f, r):                         # both reconstructions back to pinned host memory, on the lane's stream
        d = r["dpb"] if "dpb" in r else None
        rb = d["ref_frame_bl"] if d else r["x_hat_bl"]
        re = d["ref_frame_el"] if d else r["x_hat_el"]
        with torch.cuda.stream(runner.lanes[lane].stream):
            out_host[lane][0].copy_(rb, non_blocking=True)
            out_host[lane][1].copy_(re, non_blocking=True)

    def measure(source, on_frame):
        """cold start (weight packing, graph capture: two short GOPs per lane), W_ warm-up rounds (one more short GOP per
        lane), then K timed rounds of this rank's share STARTING AT A GOP BOUNDARY, so that the timed frames carry the I / P
        mix of the configuration; device time between two events on the current stream, lanes fenced on both sides, max
        over ranks."""
        rows = []
        cold = gop.work_units(lanes, 8, 4)
        for _ in runner.rounds(cold, source, on_frame, rows):
            pass
        runner.finish(rows)
        del rows[:]
        timed_keys = None

        def tap(lane, unit, f, r):
            if timed_keys is not None:
                timed_keys.add((unit.seq, f))
            if on_frame is not None:
                on_frame(lane, unit, f, r)

        for _ in runner.rounds(gop.work_units(lanes, W_, W_), source, on_frame, rows):      # warm-up: W_ rounds
            pass
        runner.finish(rows)
        del rows[:]
        it = runner.rounds(mine, source, tap, rows)
        runner.fence()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = _lib.launch_count()
        e0.record()
        runner.release(e0)
        n_frames = n_rounds = 0
        timed_keys = set()
        while args.full_job or n_rounds < K:
            n = next(it, 0)
            if n == 0:
                break
            n_frames += n
            n_rounds += 1
        runner.fence()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        launches = _lib.launch_count() - l0
        runner.finish(rows)
        it.close()
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
            cnt = torch.tensor([n_frames, n_rounds], device=dev, dtype=torch.float64)
            dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
            n_frames_all = int(cnt[0].item())
            dist.barrier()
        else:
            n_frames_all = n_frames
        timed = [r for r in rows if (r[0], r[1]) in timed_keys]
        assert len(timed) == n_frames, (len(timed), n_frames)
        return {"ms": ms, "frames": n_frames, "frames_all": n_frames_all, "rounds": n_rounds, "launches": launches, "rows": timed}

    # ---- device-resident throughput ("value"), then end to end from pinned host memory ("e2e") -------------------
    d0 = ops.SIMT_DOWNGRADES
    with ClockSampler(local) as clocks:
        res = measure(src_resident, None)
    downgrades = ops.SIMT_DOWNGRADES - d0
    res_e2e = measure(src_host, d2h)

    # ---- rate / distortion rows gathered over NCCL (the only collective of the path) ------------------------------
    table = gop.gather_stats(res["rows"], dist if world > 1 else None, device=dev if world > 1 else "cpu")
    if rank == 0:
        px = SIZES[args.size][0] * SIZES[args.size][1]
        n_i = int(table[:, 2].sum().item())
        n_all = res["frames_all"]
        flop = (n_i * FLOP_PER_PX_I + (n_all - n_i) * FLOP_PER_PX_P) * H * W
        fps = n_all / (res["ms"] / 1e3)
        fps_e2e = res_e2e["frames_all"] / (res_e2e["ms"] / 1e3)
        summ = gop.summarize(table, px, px // 4)
        split = "/".join(str(len(gop.shard(units, world, r))) for r in range(world))
        line = {
            "metric": "two-layer 1080p frames/sec (est. bitrate)", "value": round(fps, 4), "unit": "frames/s", "n_gpus": world,
            "steps": res["rounds"], "warmup": W_, "ms_per_step": round(res["ms"] / max(res["rounds"], 1), 2), "higher_is_better": True,
            "scaling": "strong" if args.full_job else "weak", "vs_baseline": None,
            "dtype": ops.engine_info(ops.default_engine())["dtype"], "data": "synthetic",
            "config": {"workload": f"LSSVC two-layer {args.size} (BL {W // 2}x{H // 2}, EL {W}x{H} padded), IP{gop_size}, "
                                   f"{n_seq} sequence(s) x {args.frames} frames = {len(units)} (sequence, GOP) work units dealt {split} per rank; "
                                   f"a step = one frame on each of {lanes} coding lane(s) per GPU; {res['frames']} frames timed per GPU "
                                   f"({n_i} I + {n_all - n_i} P over all ranks), estimated bitrate, random-init synthetic weights",
                       "parallelism": f"gop-sharded x{world}, {lanes} concurrent GOP lane(s) per GPU" + (", whole-frame CUDA graphs" if runner.graphs else ""),
                       "conv_engine": "hs" if ops.default_engine() == "h2" else ops.default_engine(), "gops_per_gpu": lanes,
                       "frames_per_step": lanes,
                       "l2": "per-frame working set (GBs of fp32 activations) >> 126 MB L2; inputs differ every step"},
            "e2e": {"value": round(fps_e2e, 4), "unit": "frames/s",
                    "h2d_bytes_per_step": int(host[0][0].numel() + host[0][1].numel()) * 4 * lanes,
                    "d2h_bytes_per_step": (int(out_host[0][0].numel() + out_host[0][1].numel()) * 4 + 32) * lanes},
            "gpu_launches": int(res["launches"]),
            "clocks": clocks.summary(),
            "conv_tflops": round(flop / (res["ms"] / 1e3) / 1e12, 2),
            "mean_bpp": {"bl": round(summ.get("bpp_bl", 0.0), 5), "el": round(summ.get("bpp_el", 0.0), 5)},
            "mean_psnr": {"bl": round(summ.get("psnr_bl", 0.0), 3), "el": round(summ.get("psnr_el", 0.0), 3)},
            "hbm_gb": {"allocated_peak": round(torch.cuda.max_memory_allocated(dev) / 2 ** 30, 1),
                       "reserved": round(torch.cuda.memory_reserved(dev) / 2 ** 30, 1)},
        }
        line["roofline"] = conv_roofline(torch, dev, shape_hr, peaks)
        line["roofline"]["in_frame"] = in_frame_roofline(torch, coder, devf, peaks, gop_size)
        sustained = peaks.get("bf16_tflops_sustained") or peaks["bf16_tflops"]
        line["roofline"]["of_kernel_ceiling"] = round(line["roofline"]["achieved"] / (peaks["bf16_tflops"] / 3.0), 4)
        line["roofline"]["whole_step_conv_frac_of_sustained"] = round(line["conv_tflops"] / world / sustained, 4)
        line["roofline"]["simt_downgrades"] = int(downgrades)
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_reference(1, 1, args.size, gop_size, False)
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
