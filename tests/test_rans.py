"""Entropy-coder parity: the product coder (C-ABI, lssvc_b200/csrc/rans.cpp) against
  (1) known-answer byte strings produced by the reference's own C++ coder (tests/golden/rans_vectors.npz),
  (2) the plain-C oracle restatement (oracle/rans_oracle.c), built on demand with gcc,
  (3) the reference's pybind11 module in oracle/_ref when it is present,
and the CDF tables against the sha256 of the tables the reference's update() methods build."""
import ctypes
import glob
import hashlib
import importlib.util
import json
import os
import subprocess

import numpy as np
import pytest
import torch

from lssvc_b200 import entropy as E

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def vectors():
    return np.load(os.path.join(GOLD, "rans_vectors.npz"))


@pytest.fixture(scope="module")
def laplace():
    return E.laplace_table()


@pytest.fixture(scope="module")
def oracle_lib():
    path = os.path.join(ROOT, "oracle", "liboracle_rans.so")
    if not os.path.exists(path):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"])
    lib = ctypes.CDLL(path)
    lib.oracle_rans_encode.restype = ctypes.c_long
    lib.oracle_rans_encode.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long, ctypes.c_void_p, ctypes.c_int,
                                       ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_long]
    lib.oracle_rans_decode.argtypes = [ctypes.c_void_p, ctypes.c_long, ctypes.c_void_p, ctypes.c_long, ctypes.c_void_p,
                                       ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    lib.oracle_pmf_to_quantized_cdf.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    return lib


def oracle_encode(lib, sym, idx, t):
    out = np.empty(sym.size * 8 + 64, dtype=np.uint8)
    n = lib.oracle_rans_encode(sym.ctypes.data, idx.ctypes.data, sym.size, t.cdf.ctypes.data, t.cdf.shape[1],
                               t.sizes.ctypes.data, t.offsets.ctypes.data, out.ctypes.data, out.size)
    assert n > 0
    return out[:n].tobytes()


def test_tables_match_reference_hashes(laplace):
    gold = json.load(open(os.path.join(GOLD, "entropy_tables.json")))
    params = torch.load(os.path.join(GOLD, "entropy_params.pt"))
    t = laplace
    assert list(t.cdf.shape) == gold["laplace"]["shape"]
    assert (sha(t.cdf), sha(t.sizes), sha(t.offsets)) == (gold["laplace"]["cdf"], gold["laplace"]["sizes"], gold["laplace"]["offsets"])
    g = E.gaussian_table()
    assert list(g.cdf.shape) == gold["gaussian"]["shape"]
    assert (sha(g.cdf), sha(g.sizes), sha(g.offsets)) == (gold["gaussian"]["cdf"], gold["gaussian"]["sizes"], gold["gaussian"]["offsets"])
    bs = params["bitparm_state"]
    coef = E.bitparm_coef([bs[f"f{i}.h"] for i in (1, 2, 3, 4)], [bs[f"f{i}.b"] for i in (1, 2, 3, 4)],
                          [bs[f"f{i}.a"] for i in (1, 2, 3)])
    b = E.bitparm_table(coef)
    assert list(b.cdf.shape) == gold["bitparm"]["shape"]
    assert (sha(b.cdf), sha(b.sizes), sha(b.offsets)) == (gold["bitparm"]["cdf"], gold["bitparm"]["sizes"], gold["bitparm"]["offsets"])
    es = params["eb_state"]
    e = E.eb_table([es[f"_matrices.{i}"] for i in range(5)], [es[f"_biases.{i}"] for i in range(5)],
                   [es[f"_factors.{i}"] for i in range(4)], es["quantiles"])
    assert list(e.cdf.shape) == gold["eb"]["shape"]
    assert (sha(e.cdf), sha(e.sizes), sha(e.offsets)) == (gold["eb"]["cdf"], gold["eb"]["sizes"], gold["eb"]["offsets"])


@pytest.mark.parametrize("name", ["small", "bypass", "big"])
def test_encoder_reproduces_reference_bytes(vectors, laplace, name):
    sym, idx, ref = vectors[name + "_sym"], vectors[name + "_idx"], vectors[name + "_bytes"].tobytes()
    enc = E.RansEncoder()
    enc.encode_with_indexes(sym, idx, laplace)
    if name == "big":
        enc.encode_with_indexes(sym[:1000], idx[:1000], laplace)
    assert enc.flush() == ref
    dec = E.RansDecoder()
    dec.set_stream(ref)
    assert np.array_equal(dec.decode_stream(idx, laplace), sym)
    if name == "big":
        assert np.array_equal(dec.decode_stream(idx[:1000], laplace), sym[:1000])


def test_pmf_to_quantized_cdf_known_answer(vectors, oracle_lib):
    pmf, ref = vectors["pmf"], vectors["pmf_cdf"]
    assert np.array_equal(E.pmf_to_quantized_cdf(pmf), ref)
    out = np.empty(pmf.size + 1, dtype=np.uint32)
    assert oracle_lib.oracle_pmf_to_quantized_cdf(pmf.ctypes.data, pmf.size, 16, out.ctypes.data) == 0
    assert np.array_equal(out, ref)


def test_product_coder_equals_c_oracle_on_random_streams(oracle_lib, laplace):
    rng = np.random.default_rng(7)
    for n, scale in ((1, 1.0), (2, 0.2), (1000, 0.5), (5000, 8.0), (20000, 60.0)):
        idx = rng.integers(0, 256, size=n).astype(np.int32)
        sym = np.round(rng.laplace(0, scale, size=n)).astype(np.int32)
        ref = oracle_encode(oracle_lib, sym, idx, laplace)
        enc = E.RansEncoder()
        enc.encode_with_indexes(sym, idx, laplace)
        got = enc.flush()
        assert got == ref
        out = np.empty(n, dtype=np.int32)
        oracle_lib.oracle_rans_decode(got, len(got), idx.ctypes.data, n, laplace.cdf.ctypes.data, laplace.cdf.shape[1],
                                      laplace.sizes.ctypes.data, laplace.offsets.ctypes.data, out.ctypes.data)
        assert np.array_equal(out, sym)


def test_encoder_reset_and_reuse(laplace):
    enc = E.RansEncoder()
    enc.encode_with_indexes(np.array([1, 2, 3], np.int32), np.array([5, 6, 7], np.int32), laplace)
    enc.reset()
    enc.encode_with_indexes(np.array([0], np.int32), np.array([9], np.int32), laplace)
    a = enc.flush()
    enc.encode_with_indexes(np.array([0], np.int32), np.array([9], np.int32), laplace)
    assert enc.flush() == a            # flush() empties the buffer like the reference
    assert len(enc.flush()) == 8       # an empty stream is just the 64-bit state


def test_decoder_rejects_truncated_stream(laplace):
    from lssvc_b200._lib import LssvcError
    dec = E.RansDecoder()
    with pytest.raises(LssvcError):
        dec.set_stream(b"\x00\x01\x02")
    enc = E.RansEncoder()
    sym = np.arange(-50, 50, dtype=np.int32)
    idx = np.full(100, 200, np.int32)
    enc.encode_with_indexes(sym, idx, laplace)
    s = enc.flush()
    dec.set_stream(s[:8])
    with pytest.raises(LssvcError):
        dec.decode_stream(idx, laplace)


def test_against_reference_module_when_built(laplace):
    so = glob.glob(os.path.join(ROOT, "oracle", "_ref", "MLCodec_rans*.so"))
    if not so:
        pytest.skip("oracle/_ref not built (needs /root/reference; `make -C oracle ref`)")
    spec = importlib.util.spec_from_file_location("MLCodec_rans", so[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rng = np.random.default_rng(11)
    n = 30000
    idx = rng.integers(0, 256, size=n).astype(np.int32)
    sym = np.round(rng.laplace(0, 4.0, size=n)).astype(np.int32)
    sym[::501] = 99999
    ref_enc = mod.BufferedRansEncoder()
    ref_enc.encode_with_indexes(sym, idx, laplace.cdf, laplace.sizes, laplace.offsets)
    ref = ref_enc.flush()
    enc = E.RansEncoder()
    enc.encode_with_indexes(sym, idx, laplace)
    got = enc.flush()
    assert got == ref
    ref_dec = mod.RansDecoder()
    ref_dec.set_stream(got)
    assert np.array_equal(np.asarray(ref_dec.decode_stream(idx, laplace.cdf, laplace.sizes, laplace.offsets)), sym)


def test_scale_thresholds_reproduce_build_indexes():
    import math
    tv, ti = E.video_scale_thresholds(), E.image_scale_thresholds()
    assert tv.numel() == 255 and ti.numel() == 63
    assert abs(tv[0].item() - 0.01034966) < 1e-7 and abs(ti[-1].item() - 226.35912) < 1e-3
    s = torch.exp(torch.randn(200000, generator=torch.Generator().manual_seed(1)) * 4)
    s[:4] = torch.tensor([-3.0, 0.0, 1e-9, 1e9])
    sv = torch.maximum(s, torch.tensor(1e-5))
    ref_v = ((torch.log(sv) - math.log(0.01)) / ((math.log(64.0) - math.log(0.01)) / 255)).clamp_(0, 255).int()
    ref_i = ((torch.log(sv) - math.log(0.11)) / ((math.log(256.0) - math.log(0.11)) / 63) + 1).clamp_(0, 63).int()
    assert torch.equal(torch.searchsorted(tv, sv, right=True).int(), ref_v)
    assert torch.equal(torch.searchsorted(ti, sv, right=True).int(), ref_i)
