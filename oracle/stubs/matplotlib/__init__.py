"""Stub of matplotlib (imported at utils.py:40-43; plotting is out of scope)."""
import contextlib


class _Style:
    @staticmethod
    def context(*a, **k):
        return contextlib.nullcontext()


style = _Style()
rcParams = {}
