"""Host side of the entropy models: CDF-table construction (once per `update()`), the scale -> CDF-row
threshold tables the index kernels search, coefficient packing for the factorised priors and Python
handles on the C++ rANS coder.  Mirrors src/entropy_models/{video,img}_entropy_models.py of the
reference; tables are built with the same fp32 torch-CPU arithmetic so that bitstreams interoperate."""
import ctypes
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import _lib

PRECISION = 16


# ---- scale -> CDF row -----------------------------------------------------------------------------
def _index_fn(log_min, log_max, levels, plus_one):
    step = (log_max - log_min) / (levels - 1)

    def fn(s):
        s = torch.maximum(s, torch.zeros_like(s) + 1e-5)
        idx = (torch.log(s) - log_min) / step
        if plus_one:
            idx = idx + 1
        return idx.clamp_(0, levels - 1).int()

    return fn


def _thresholds(fn, levels):
    """build_indexes is a non-decreasing step function of the fp32 scale: return, for every row r in
    1..levels-1, the smallest fp32 scale whose row is >= r (bisection over float bit patterns), so that
    row(s) = #{thresholds <= max(s, 1e-5)} reproduces the CPU evaluation exactly."""
    targets = torch.arange(1, levels, dtype=torch.int32)
    lo = torch.full((levels - 1,), np.float32(1e-5).view(np.int32).item(), dtype=torch.int32)
    hi = torch.full((levels - 1,), np.float32(1e6).view(np.int32).item(), dtype=torch.int32)
    reach_hi = fn(hi.view(torch.float32)) >= targets
    for _ in range(34):
        mid = lo + (hi - lo) // 2
        ok = fn(mid.view(torch.float32).clone()) >= targets
        hi = torch.where(ok, mid, hi)
        lo = torch.where(ok, lo, mid)
    thr = hi.view(torch.float32).clone()
    # rows never reached (cannot happen for the two tables in use) get +inf
    thr[~reach_hi] = float("inf")
    # if even the smallest scale already reaches a row, its threshold is that smallest scale
    first = fn(torch.full((1,), 1e-5))[0].item()
    thr[: max(first, 0)] = 0.0
    return thr


_cache = {}


def video_scale_thresholds():
    """GaussianEncoder.build_indexes (video_entropy_models.py:309-313): 256 rows, 0.01 .. 64."""
    if "video" not in _cache:
        _cache["video"] = _thresholds(_index_fn(math.log(0.01), math.log(64.0), 256, False), 256)
    return _cache["video"]


def image_scale_thresholds():
    """GaussianConditional.build_indexes (img_entropy_models.py:687-691): 64 rows, 0.11 .. 256, + 1."""
    if "image" not in _cache:
        _cache["image"] = _thresholds(_index_fn(math.log(0.11), math.log(256.0), 64, True), 64)
    return _cache["image"]


# ---- quantised CDF tables --------------------------------------------------------------------------
def pmf_to_quantized_cdf(pmf, precision=PRECISION):
    """MLCodec_CXX.pmf_to_quantized_cdf (src/cpp/ops/ops.cpp:24-82) through the C-ABI."""
    lib = _lib.load()
    p = np.ascontiguousarray(np.asarray(pmf, dtype=np.float32))
    out = np.empty(p.size + 1, dtype=np.uint32)
    _lib.check(lib.lssvc_pmf_to_quantized_cdf(p.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), p.size, precision,
                                              out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32))), "pmf_to_quantized_cdf")
    return out


def _pmf_to_cdf(pmf, tail_mass, pmf_length, max_length):
    cdf = np.zeros((len(pmf_length), int(max_length) + 2), dtype=np.int32)
    for i in range(len(pmf_length)):
        n = int(pmf_length[i])
        prob = torch.cat((pmf[i, :n], tail_mass[i].reshape(-1)[:1]), dim=0)
        row = pmf_to_quantized_cdf(prob.tolist())
        cdf[i, : row.size] = row.astype(np.int32)
    return cdf


class CdfTable:
    """cdf [rows][stride] int32, sizes [rows], offsets [rows] — what the coder consumes."""

    def __init__(self, cdf, sizes, offsets):
        self.cdf = np.ascontiguousarray(cdf, dtype=np.int32)
        self.sizes = np.ascontiguousarray(np.asarray(sizes).reshape(-1), dtype=np.int32)
        self.offsets = np.ascontiguousarray(np.asarray(offsets).reshape(-1), dtype=np.int32)


def _laplace_cdf(scales, v):
    return 0.5 - 0.5 * torch.sign(v) * torch.expm1(-v.abs() / scales)


def laplace_table():
    """GaussianEncoder.update (video_entropy_models.py:266-307): 256 Laplace rows."""
    table = torch.exp(torch.linspace(math.log(0.01), math.log(64.0), 256))
    center = torch.zeros_like(table) + 50
    for i in range(50, 1, -1):
        probs = _laplace_cdf(table, torch.zeros_like(table) + i)
        center = torch.where(probs > 0.9999, torch.zeros_like(center) + i, center)
    center = center.int()
    length = 2 * center + 1
    max_length = int(length.max())
    samples = (torch.arange(max_length) - center[:, None]).float()
    scales = torch.zeros_like(samples) + table[:, None]
    upper = _laplace_cdf(scales, samples + 0.5)
    lower = _laplace_cdf(scales, samples - 0.5)
    pmf = upper - lower
    tail = 2 * lower[:, :1]
    return CdfTable(_pmf_to_cdf(pmf, tail, length, max_length), (length + 2).numpy(), (-center).numpy())


def bitparm_coef(h, b, a):
    """[C][11] = softplus(h1..4) | b1..4 | tanh(a1..3) from the four Bitparm layers
    (video_entropy_models.py:110-129).  h, b: lists of 4 tensors; a: list of 3."""
    cols = [F.softplus(t.detach().float().cpu().reshape(-1)) for t in h]
    cols += [t.detach().float().cpu().reshape(-1) for t in b]
    cols += [torch.tanh(t.detach().float().cpu().reshape(-1)) for t in a]
    return torch.stack(cols, dim=1).contiguous()


def _bitparm_forward(x, coef):
    # x: [C, n]
    for i in range(3):
        x = x * coef[:, i:i + 1] + coef[:, 4 + i:5 + i]
        x = x + torch.tanh(x) * coef[:, 8 + i:9 + i]
    return torch.sigmoid(x * coef[:, 3:4] + coef[:, 7:8])


def bitparm_table(coef):
    """BitEstimator.update (video_entropy_models.py:168-223)."""
    C = coef.shape[0]
    medians = torch.zeros(C)
    minima = medians + 50
    for i in range(50, 1, -1):
        probs = _bitparm_forward(torch.zeros(C, 1) - i, coef).reshape(-1)
        minima = torch.where(probs < 0.0001, torch.zeros_like(medians) + i, minima)
    maxima = medians + 50
    for i in range(50, 1, -1):
        probs = _bitparm_forward(torch.zeros(C, 1) + i, coef).reshape(-1)
        maxima = torch.where(probs > 0.9999, torch.zeros_like(medians) + i, maxima)
    minima, maxima = minima.int(), maxima.int()
    offset = -minima
    start = medians - minima
    length = maxima + minima + 1
    max_length = int(length.max())
    samples = torch.arange(max_length)[None, :] + start[:, None]
    lower = _bitparm_forward(samples - 0.5, coef)
    upper = _bitparm_forward(samples + 0.5, coef)
    pmf = upper - lower
    tail = lower[:, :1] + (1.0 - upper[:, -1:])
    return CdfTable(_pmf_to_cdf(pmf, tail, length, max_length), (length + 2).numpy(), offset.numpy())


def eb_coef(matrices, biases, factors, quantiles):
    """[C][59] for the EntropyBottleneck kernel (img_entropy_models.py:483-502): softplus(matrices)
    (3, 9, 9, 9, 3) | biases (3, 3, 3, 3, 1) | tanh(factors) (3, 3, 3, 3) | median."""
    C = quantiles.shape[0]
    parts = [F.softplus(m.detach().float().cpu()).reshape(C, -1) for m in matrices]
    parts += [b.detach().float().cpu().reshape(C, -1) for b in biases]
    parts += [torch.tanh(f.detach().float().cpu()).reshape(C, -1) for f in factors]
    parts += [quantiles.detach().float().cpu()[:, 0, 1:2]]
    coef = torch.cat(parts, dim=1).contiguous()
    assert coef.shape[1] == 59, coef.shape
    return coef


def _eb_logits(x, matrices, biases, factors):
    logits = x
    for i in range(len(matrices)):
        logits = torch.matmul(F.softplus(matrices[i]), logits) + biases[i]
        if i < len(factors):
            logits = logits + torch.tanh(factors[i]) * torch.tanh(logits)
    return logits


def eb_table(matrices, biases, factors, quantiles):
    """EntropyBottleneck.update (img_entropy_models.py:436-476)."""
    matrices = [m.detach().float().cpu() for m in matrices]
    biases = [b.detach().float().cpu() for b in biases]
    factors = [f.detach().float().cpu() for f in factors]
    q = quantiles.detach().float().cpu()
    medians = q[:, 0, 1]
    minima = torch.clamp(torch.ceil(medians - q[:, 0, 0]).int(), min=0)
    maxima = torch.clamp(torch.ceil(q[:, 0, 2] - medians).int(), min=0)
    offset = -minima
    start = medians - minima
    length = maxima + minima + 1
    max_length = int(length.max())
    samples = torch.arange(max_length)[None, :] + start[:, None, None]
    lower = _eb_logits(samples - 0.5, matrices, biases, factors)
    upper = _eb_logits(samples + 0.5, matrices, biases, factors)
    sign = -torch.sign(lower + upper)
    pmf = torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower))[:, 0, :]
    tail = torch.sigmoid(lower[:, 0, :1]) + torch.sigmoid(-upper[:, 0, -1:])
    return CdfTable(_pmf_to_cdf(pmf, tail, length, max_length), (length + 2).numpy(), offset.numpy())


def gaussian_table(tail_mass=1e-9):
    """GaussianConditional.update (img_entropy_models.py:623-648): 64 Gaussian rows."""
    import scipy.stats
    table = torch.exp(torch.linspace(math.log(0.11), math.log(256.0), 64))
    multiplier = -scipy.stats.norm.ppf(tail_mass / 2)
    center = torch.ceil(table * multiplier).int()
    length = 2 * center + 1
    max_length = int(length.max())
    samples = torch.abs(torch.arange(max_length).int() - center[:, None]).float()
    scale = table.unsqueeze(1).float()
    cum = lambda t: 0.5 * torch.erfc(float(-(2 ** -0.5)) * t)
    upper = cum((0.5 - samples) / scale)
    lower = cum((-0.5 - samples) / scale)
    pmf = upper - lower
    tail = 2 * lower[:, :1]
    return CdfTable(_pmf_to_cdf(pmf, tail, length, max_length), (length + 2).numpy(), (-center).numpy())


# ---- coder handles ---------------------------------------------------------------------------------
def _i32(a):
    return np.ascontiguousarray(np.asarray(a).reshape(-1), dtype=np.int32)


class RansEncoder:
    """BufferedRansEncoder (src/cpp/rans/rans_interface.cpp:85-172): buffer symbols, flush() -> bytes."""

    def __init__(self):
        self._lib = _lib.load()
        self._h = self._lib.lssvc_rans_encoder_new()

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.lssvc_rans_encoder_free(self._h)
            self._h = None

    def reset(self):
        self._lib.lssvc_rans_encoder_reset(self._h)

    def encode_with_indexes(self, symbols, indexes, table):
        s, i = _i32(symbols), _i32(indexes)
        if s.size != i.size:
            raise ValueError("symbols and indexes must have the same length")
        _lib.check(self._lib.lssvc_rans_encode_with_indexes(
            self._h, s.ctypes.data, i.ctypes.data, s.size, table.cdf.ctypes.data, table.cdf.shape[0], table.cdf.shape[1],
            table.sizes.ctypes.data, table.offsets.ctypes.data), "rans_encode_with_indexes")

    def flush(self):
        data = ctypes.POINTER(ctypes.c_uint8)()
        n = self._lib.lssvc_rans_encoder_flush(self._h, ctypes.byref(data))
        if n < 0:
            _lib.check(int(n), "rans_encoder_flush")
        return ctypes.string_at(data, n)


class RansDecoder:
    """RansDecoder (src/cpp/rans/rans_interface.cpp:176-244): set_stream once, decode_stream in coding order."""

    def __init__(self):
        self._lib = _lib.load()
        self._h = self._lib.lssvc_rans_decoder_new()

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.lssvc_rans_decoder_free(self._h)
            self._h = None

    def set_stream(self, data):
        _lib.check(self._lib.lssvc_rans_decoder_set_stream(self._h, bytes(data), len(data)), "rans_decoder_set_stream")

    def decode_stream(self, indexes, table):
        i = _i32(indexes)
        out = np.empty(i.size, dtype=np.int32)
        _lib.check(self._lib.lssvc_rans_decode_stream(
            self._h, i.ctypes.data, i.size, table.cdf.ctypes.data, table.cdf.shape[0], table.cdf.shape[1], table.sizes.ctypes.data,
            table.offsets.ctypes.data, out.ctypes.data), "rans_decode_stream")
        return out
