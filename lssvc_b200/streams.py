"""--write_stream 1 orchestration (real rANS bitstreams); filled in by the write-stream milestone."""


def intra_encode_decode(*a, **k):
    raise NotImplementedError("write-stream path not built yet")


def inter_encode_decode(*a, **k):
    raise NotImplementedError("write-stream path not built yet")
