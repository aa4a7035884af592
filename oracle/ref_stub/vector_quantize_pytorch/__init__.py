"""Placeholder for the un-vendored pip dependency `vector-quantize-pytorch`
(reference environment.yaml:34, unpinned).  ResidualVQ is out of scope
(parity unpinned)."""


class ResidualVQ:
    def __init__(self, *args, **kwargs):
        raise NotImplementedError("vector_quantize_pytorch is not available offline")
