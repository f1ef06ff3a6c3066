"""Drop-in for the reference's pybind11 module `MLCodec_rans` (src/cpp/rans/rans_interface.cpp:246-261) over the C-ABI
coder of liblssvc_b200.so (csrc/rans.cpp).

Same class and method names, same argument lists (python lists or int32 arrays: symbols, indexes, cdfs[rows][stride],
cdfs_sizes, offsets), same return types, byte-identical streams:

  BufferedRansEncoder().encode_with_indexes(symbols, indexes, cdfs, cdfs_sizes, offsets) -> None   :85-145
                       .flush() -> bytes                                                           :147-172
                       .reset()                                                                    (hpp :64-66)
  RansDecoder().set_stream(bytes)                                                                  :176-182
               .decode_stream(indexes, cdfs, cdfs_sizes, offsets) -> int32 array                   :184-244

and the two calls the image path (`img_entropy_models.py:19-27,309,354`) needs but src/cpp does not define (they exist in
the reference's prebuilt cpython-36 binary; semantics = CompressAI's, whose header rans_interface.cpp:1-20 carries):

  RansEncoder().encode_with_indexes(symbols, indexes, cdfs, cdfs_sizes, offsets) -> bytes   (= buffered encode + flush)
  RansDecoder().decode_with_indexes(bytes, indexes, cdfs, cdfs_sizes, offsets) -> list[int] (= set_stream + decode_stream)

Install in place of the reference's binary with `lssvc_b200.compat.install()`."""
import numpy as np

from . import entropy

_table_cache = {}


def _table(cdfs, cdfs_sizes, offsets):
    """The reference passes the whole CDF table as nested python lists on every call (`_quantized_cdf.tolist()`,
    img_entropy_models.py:312,357): convert once per distinct table object."""
    key = id(cdfs)
    hit = _table_cache.get(key)
    if hit is not None and hit[0] is cdfs and hit[1] is cdfs_sizes and hit[2] is offsets:
        return hit[3]
    cdf = np.asarray(cdfs, dtype=np.int32)
    if cdf.ndim != 2:
        raise ValueError(f"cdfs must be 2-D, got shape {cdf.shape}")
    t = entropy.CdfTable(cdf, cdfs_sizes, offsets)
    if t.sizes.size != cdf.shape[0] or t.offsets.size != cdf.shape[0]:
        raise ValueError("cdfs, cdfs_sizes and offsets disagree on the number of rows")
    if len(_table_cache) > 16:
        _table_cache.clear()
    _table_cache[key] = (cdfs, cdfs_sizes, offsets, t)     # the references keep the ids valid
    return t


class BufferedRansEncoder:
    def __init__(self):
        self._enc = entropy.RansEncoder()

    def encode_with_indexes(self, symbols, indexes, cdfs, cdfs_sizes, offsets):
        self._enc.encode_with_indexes(symbols, indexes, _table(cdfs, cdfs_sizes, offsets))

    def flush(self):
        return self._enc.flush()

    def reset(self):
        self._enc.reset()


class RansEncoder:
    def encode_with_indexes(self, symbols, indexes, cdfs, cdfs_sizes, offsets):
        enc = entropy.RansEncoder()
        enc.encode_with_indexes(symbols, indexes, _table(cdfs, cdfs_sizes, offsets))
        return enc.flush()


class RansDecoder:
    def __init__(self):
        self._dec = entropy.RansDecoder()

    def set_stream(self, encoded):
        self._dec.set_stream(encoded)

    def decode_stream(self, indexes, cdfs, cdfs_sizes, offsets):
        return self._dec.decode_stream(indexes, _table(cdfs, cdfs_sizes, offsets))

    def decode_with_indexes(self, encoded, indexes, cdfs, cdfs_sizes, offsets):
        self._dec.set_stream(encoded)
        return self._dec.decode_stream(indexes, _table(cdfs, cdfs_sizes, offsets)).tolist()
