"""lssvc_b200: B200-native (sm_100a) implementation of LSSVC's per-frame two-layer coding forward pass.

Drop-in for the reference's model API on that path:
    from lssvc_b200 import IntraSS, LSSVC, LSSVC_extend
"""
__version__ = "0.1.0"

from .models import IntraSS, LSSVC, LSSVC_extend  # noqa: E402,F401
