"""Makes the reference's own entropy-model code run on this package's coder: registers lssvc_b200.MLCodec_rans /
MLCodec_CXX under the module names `src/entropy_models/{video,img}_entropy_models.py` import them by
(`from .MLCodec_rans import ...` :11, img :19; `from .MLCodec_CXX import ...` :16, img :31), replacing the reference's
cpython-36 binaries (src/entropy_models/MLCodec_*.cpython-36m-x86_64-linux-gnu.so)."""
import sys

from . import MLCodec_CXX, MLCodec_rans


def install(package="src.entropy_models"):
    """Call before importing the reference's models.  Returns the two modules."""
    sys.modules[package + ".MLCodec_rans"] = MLCodec_rans
    sys.modules[package + ".MLCodec_CXX"] = MLCodec_CXX
    return MLCodec_rans, MLCodec_CXX
