"""Job runner: the reference's per-sequence frame loop (test.py:121-325: I / P decision by `frame_idx % gop_size`, DPB
hand-over, in-place clamp of the reference frames, per-frame bits and PSNR) and its fan-out over workers (test.py:685-748),
re-designed for one process per B200:

  * the job is cut into (sequence, GOP) work units (gop.py) — an I-frame rebuilds the DPB, so units are independent;
  * a rank takes its round-robin share of the units (gop.shard) and codes them on `lanes` concurrent CODING LANES: a lane is a
    CUDA stream + a DPB; lanes share the models (weights and packed weights are read-only) and, for P-frames, replay their
    own whole-frame CUDA graphs (models.LSSVC._forward_graphed, one graph set and one memory pool per lane).  Independent
    GOPs in flight on one GPU is SURVEY.md H9: the 1/16 .. 1/64-resolution launches of a frame cannot fill 148 SMs, and a
    persistent conv kernel leaves SMs idle in its tail; kernels of another lane fill both.  Nothing is batched inside a
    kernel, so every frame is bit-identical to the single-lane result (tests/test_runner_gpu.py);
  * the host never waits for a frame: bits and squared errors of a frame stay on the device, are copied to pinned host
    memory behind the frame on the lane's stream, and are read when the lane comes round again;
  * one row per frame (gop.STAT_COLUMNS: seq, frame, is_intra, bits_bl, bits_el, sse_bl, sse_el), gathered over ranks by
    gop.gather_stats — the only collective of the path.
"""
import torch

from . import _lib, gop


def _ptr(t):
    import ctypes
    return ctypes.c_void_p(t.data_ptr())


def lane_schedule(units, lanes):
    """Issue order of a rank's frames: unit i is coded by lane i % lanes; the lanes advance in lock step, one frame per
    round each (a lane's frames are serial — the P-chain — and lanes are independent).  Returns
    [(lane, unit, frame_idx, is_intra)] in issue order; every frame of every unit appears exactly once."""
    per_lane = [[(u, f, intra) for u in units[k::lanes] for f, intra in gop.frames_of(u)] for k in range(lanes)]
    order = []
    for rnd in range(max((len(q) for q in per_lane), default=0)):
        for k, q in enumerate(per_lane):
            if rnd < len(q):
                order.append((k,) + q[rnd])
    return order


class _Lane:
    def __init__(self, index, device, ring=4):
        self.index = index
        self.stream = torch.cuda.Stream(device=device)
        self.dpb = None
        self.pending = []               # [(slot, meta, kept tensors)]
        self.stats = [torch.zeros(5, dtype=torch.float64).pin_memory() for _ in range(ring)]
        self.events = [torch.cuda.Event() for _ in range(ring)]
        self.slot = 0


class GopRunner:
    """Codes work units on `lanes` concurrent lanes of one GPU.

    net_i / net_p: lssvc_b200.IntraSS / LSSVC(_extend) on the device, `set_scale_information` already called.
    graphs: replay whole-frame CUDA graphs for P-frames (default: when more than one lane runs, so that the single host
    thread keeps every lane fed)."""

    def __init__(self, net_i, net_p, lanes=1, graphs=None):
        assert lanes >= 1
        self.net_i, self.net_p = net_i, net_p
        self.device = net_p.device
        self.lanes = [_Lane(k, self.device) for k in range(lanes)]
        self.graphs = (lanes > 1) if graphs is None else bool(graphs)
        self.H, self.W = net_p.shape_hr
        self._lib = _lib.load()
        self.rows = []

    # ---- one frame on one lane (asynchronous) ------------------------------------------------------------------
    def _issue(self, lane, unit, frame_idx, intra, x_bl, x_el, keep=None):
        H, W = self.H, self.W
        with torch.cuda.stream(lane.stream):
            if intra:
                r = self.net_i.forward(x_bl, x_el, _async=True)
                lane.dpb = {"ref_frame_bl": r["x_hat_bl"], "ref_frame_el": r["x_hat_el"], "ref_feature_bl": None,
                            "ref_feature_el": r["feature_el"]}
            else:
                net = self.net_p
                prev_graphs, prev_lane = net.use_graphs, net.lane
                net.use_graphs, net.lane = self.graphs, lane.index
                try:
                    r = net.forward_one_frame(x_bl, x_el, None, None, None, None, _dpb=lane.dpb, _async=True)
                finally:
                    net.use_graphs, net.lane = prev_graphs, prev_lane
                lane.dpb = r["dpb"]
            # the caller's in-place clamp of the reference frames (test.py:249-250); PSNR is taken on the clamped frames (:253-254)
            rec_bl = lane.dpb["ref_frame_bl"].clamp_(0, 1)
            rec_el = lane.dpb["ref_frame_el"].clamp_(0, 1)
            # bits (2 doubles) | sse_bl | sse_el -> pinned host memory behind the frame
            dev_stats = torch.empty(5, dtype=torch.float64, device=self.device)
            dev_stats[:2].copy_(r["_bits"].t[:2])
            dev_stats[4:5].copy_(r["_bits"].t[2:3])
            for k, (a, b) in enumerate(((x_bl, rec_bl), (x_el, rec_el))):
                _lib.check(self._lib.lssvc_sse_flat(_ptr(a), _ptr(b), a.numel(), _ptr(dev_stats[2 + k:3 + k]),
                                                    torch.cuda.current_stream().cuda_stream), "sse_flat")
            slot = lane.slot
            lane.slot = (slot + 1) % len(lane.stats)
            lane.stats[slot].copy_(dev_stats, non_blocking=True)
            lane.events[slot].record(lane.stream)
            lane.pending.append((slot, (unit.seq, frame_idx, float(intra)), keep))
        return r

    def _collect(self, lane, rows, leave=0):
        while len(lane.pending) > leave:
            slot, meta, _keep = lane.pending.pop(0)
            lane.events[slot].synchronize()
            stats = lane.stats[slot].tolist()
            if stats[4] != 0.0:
                # the lanes run ahead of the host: the frames behind this one already used its DPB, so there is nothing to
                # re-code in place (the synchronous API does that, models._recode_fp32)
                raise _lib.LssvcError(f"sequence {meta[0]} frame {meta[1]}: an activation reached the fp16 limit of the split-fp16 "
                                      "tensor-core engine; run this job with LSSVC_CONV_ENGINE=simt (or through the synchronous API, "
                                      "which re-codes such frames on the fp32 engine)")
            rows.append(meta + tuple(stats[:4]))

    # ---- a rank's share of a job --------------------------------------------------------------------------------
    def rounds(self, units, frames, on_frame=None, rows=None):
        """Generator over the issue schedule: every next() issues ONE ROUND (one frame on every lane that still has work,
        asynchronously) and yields the number of frames issued.  rows: list the finished frames' stat rows are appended to
        (rows of frames still in flight arrive with later rounds / finish())."""
        rows = self.rows if rows is None else rows
        ring = len(self.lanes[0].stats)
        for lane in self.lanes:
            lane.dpb = None
        schedule = lane_schedule(list(units), len(self.lanes))
        i = 0
        while i < len(schedule):
            n = 0
            seen = set()
            while i < len(schedule) and schedule[i][0] not in seen:
                k, unit, frame_idx, intra = schedule[i]
                seen.add(k)
                lane = self.lanes[k]
                self._collect(lane, rows, leave=ring - 1)       # frees the pinned slot this frame will use
                with torch.cuda.stream(lane.stream):
                    x_bl, x_el = frames(unit.seq, frame_idx)
                r = self._issue(lane, unit, frame_idx, intra, x_bl, x_el, keep=(x_bl, x_el))
                if on_frame is not None:
                    on_frame(lane.index, unit, frame_idx, r)
                i += 1
                n += 1
            yield n

    def finish(self, rows=None):
        """Waits for every frame in flight and appends its row."""
        rows = self.rows if rows is None else rows
        for lane in self.lanes:
            self._collect(lane, rows)
        return rows

    def code_units(self, units, frames, on_frame=None):
        """units: this rank's gop.Unit list; frames(seq, frame_idx) -> (x_bl, x_el) fp32 [1, 3, H, W] tensors on the device
        (a source reading pinned host memory issues its own copies: it is called under the lane's stream).
        on_frame(lane_index, unit, frame_idx, result): called right after a frame has been ISSUED (result tensors are valid
        on the lane's stream).  Returns the rows of gop.STAT_COLUMNS in (seq, frame) order."""
        rows = []
        for _ in self.rounds(units, frames, on_frame, rows):
            pass
        self.finish(rows)
        rows.sort(key=lambda r: (r[0], r[1]))
        return rows

    def fence(self, stream=None):
        """Makes `stream` (default: the current stream) wait for everything issued on the lanes so far."""
        stream = stream or torch.cuda.current_stream()
        for lane in self.lanes:
            stream.wait_stream(lane.stream)

    def release(self, event):
        """No lane starts new work before `event` has happened."""
        for lane in self.lanes:
            lane.stream.wait_event(event)

    def synchronize(self):
        for lane in self.lanes:
            lane.stream.synchronize()


def run_job(net_i, net_p, n_seq, n_frames, gop_size, frames, lanes=1, graphs=None, dist=None, rank=0, world=1):
    """The whole job of BASELINE.json's configs (e.g. config 3: 8 sequences x 96 frames, IP32, over `world` ranks): work
    units -> this rank's share -> lanes -> one gathered [n_seq * n_frames, 7] table, identical on every rank."""
    units = gop.shard(gop.work_units(n_seq, n_frames, gop_size), world, rank)
    runner = GopRunner(net_i, net_p, lanes=lanes, graphs=graphs)
    rows = runner.code_units(units, frames)
    runner.synchronize()
    return gop.gather_stats(rows, dist, device=net_p.device if (dist is not None and world > 1) else "cpu")
