"""Drop-in for the reference's pybind11 module `MLCodec_CXX` (src/cpp/ops/ops.cpp:84-91) over the C-ABI
(`lssvc_pmf_to_quantized_cdf`, csrc/rans.cpp): pmf_to_quantized_cdf(pmf: list[float], precision: int) -> list[int]."""
from . import entropy


def pmf_to_quantized_cdf(pmf, precision):
    return entropy.pmf_to_quantized_cdf(pmf, int(precision)).tolist()
