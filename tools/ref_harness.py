"""Import the UNMODIFIED reference from /root/reference in this container (it does not exist on the GPU box).

Used only by the scripts under tools/ that validate the oracle against the reference and generate the
golden fixtures committed under tests/golden/.  Nothing in the product or in the -m gpu tests imports this.
"""
import os
import sys

REFERENCE = os.environ.get("LSSVC_REFERENCE", "/root/reference")
_STUBS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "stubs")


def available():
    return os.path.isdir(os.path.join(REFERENCE, "src", "models"))


def import_reference():
    """Returns (IntraSS, LSSVC_extend) classes of the reference."""
    if not available():
        raise RuntimeError(f"reference not found at {REFERENCE}")
    for p in (_STUBS, REFERENCE):
        if p not in sys.path:
            sys.path.insert(0, p)
    from src.models.IntraSS import IntraSS
    from src.models.LSSVC_net_extend import LSSVC_extend
    return IntraSS, LSSVC_extend
