"""Per-launch CUDA-event timing of the operator layer (no profiler): every ops.* call made while a LaunchTimer is active is
bracketed by two events on the launching stream.  Events are recorded in stream order and serialise nothing, so the
numbers are warm-cache, in-frame figures.  Used by bench.py (`roofline.in_frame`) and tools/layer_times.py."""
import torch

from . import ops

OPS = ("conv", "ffn", "pw", "dwconv3x3", "deconv3x3_s2", "lrelu_copy", "softmax2_blend", "flow_warp", "bilinear_resize", "avgpool2",
       "maxpool2", "spynet_prep", "offset_diversity", "laplace_quant", "four_part_step", "gaussian_quant", "bitparm_quant",
       "eb_quant", "sse")


class LaunchTimer:
    """with LaunchTimer() as t: <code one frame>; t.rows() -> [(op name, conv trace record or None, milliseconds)]"""

    def __init__(self):
        self.records = []
        self._saved = {}

    def _wrap(self, name, fn):
        def inner(*a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n_before = len(ops.TRACE)
            e0.record()
            r = fn(*a, **k)
            e1.record()
            info = ops.TRACE[-1] if (name in ("conv", "pw", "ffn") and len(ops.TRACE) > n_before) else None
            self.records.append((name, info, e0, e1))
            return r
        return inner

    def __enter__(self):
        self._trace = ops.TRACE
        ops.TRACE = []
        for n in OPS:
            self._saved[n] = getattr(ops, n)
            setattr(ops, n, self._wrap(n, self._saved[n]))
        return self

    def __exit__(self, *exc):
        for n, fn in self._saved.items():
            setattr(ops, n, fn)
        ops.TRACE = self._trace

    def rows(self):
        torch.cuda.synchronize()
        return [(n, i, a.elapsed_time(b)) for n, i, a, b in self.records]


def conv_summary(rows):
    """FLOPs (algorithmic) and milliseconds of the tensor-core convolution launches of `rows`, by kernel family."""
    out = {}
    for name, info, ms in rows:
        if info is None or name not in ("conv", "pw", "ffn"):
            continue
        fam = info["engine"] if name == "conv" else name
        d = out.setdefault(fam, {"launches": 0, "ms": 0.0, "flops": 0.0})
        d["launches"] += 1
        d["ms"] += ms
        d["flops"] += info["flops"]
    return out
