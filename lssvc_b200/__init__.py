"""lssvc_b200: B200-native (sm_100a) implementation of LSSVC's per-frame two-layer coding forward pass."""
__version__ = "0.1.0"
