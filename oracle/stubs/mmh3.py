"""Stub of mmh3 (utils.py:1267-1269) on top of scikit-learn's MurmurHash3_x86_32.

Checked against the canonical vector mmh3.hash("foo", 0) == -156908512.
"""
from sklearn.utils import murmurhash3_32


def hash(key, seed=0, signed=True):
    return int(murmurhash3_32(key, seed=seed, positive=not signed))
