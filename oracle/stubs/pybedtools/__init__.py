"""Stub of pybedtools (imported at utils.py:28; never reached on the hot path)."""


class BedTool:  # pragma: no cover - placeholder only
    def __init__(self, *a, **k):
        raise RuntimeError("pybedtools stub: not available in the oracle harness")
