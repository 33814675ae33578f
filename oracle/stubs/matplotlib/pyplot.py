"""Stub of matplotlib.pyplot: only switch_backend is called at import time (utils.py:43)."""
import contextlib


def switch_backend(name):
    return None


class style:
    @staticmethod
    def context(*a, **k):
        return contextlib.nullcontext()
