"""Scalar restatement of the reference's similarity features.  TEST INFRASTRUCTURE ONLY.

Follows utils.py:892-949 (pac), :1043-1058 (_minhash_similarity, _pac), :1243-1269
(jaccard_similarity, _minhash_kmer) with the reference's exact string semantics (IUPAC-aware
complement, KeyError on lower case).  PINNED: tests/test_host.py checks these functions against the
reference's own outputs (tests/golden/reference_golden.json); the GPU kernel
(tfbs_similarity_sums) is then checked against them.  MurmurHash3_x86_32 (mmh3 in the reference) is
the public-domain algorithm.  Nothing in discover_tfbs_b200/ imports this module.
"""
from collections import Counter
from functools import lru_cache
from math import ceil, log

import numpy as np
from scipy.stats import poisson

_COMP = bytes.maketrans(b"ACGTMRWSYKVHDBXN", b"TGCAKYWSRMBDHVXN")


def murmurhash3_32(key: bytes, seed: int = 0) -> int:
    """MurmurHash3_x86_32 (Austin Appleby, public domain algorithm), unsigned result."""
    c1, c2 = 0xCC9E2D51, 0x1B873593
    h = seed & 0xFFFFFFFF
    n = len(key)
    for i in range(0, n - n % 4, 4):
        k = int.from_bytes(key[i:i + 4], "little")
        k = (k * c1) & 0xFFFFFFFF
        k = ((k << 15) | (k >> 17)) & 0xFFFFFFFF
        k = (k * c2) & 0xFFFFFFFF
        h ^= k
        h = ((h << 13) | (h >> 19)) & 0xFFFFFFFF
        h = (h * 5 + 0xE6546B64) & 0xFFFFFFFF
    tail = key[n - n % 4:]
    k = 0
    if len(tail) >= 3:
        k ^= tail[2] << 16
    if len(tail) >= 2:
        k ^= tail[1] << 8
    if len(tail) >= 1:
        k ^= tail[0]
        k = (k * c1) & 0xFFFFFFFF
        k = ((k << 15) | (k >> 17)) & 0xFFFFFFFF
        k = (k * c2) & 0xFFFFFFFF
        h ^= k
    h ^= n
    h ^= h >> 16
    h = (h * 0x85EBCA6B) & 0xFFFFFFFF
    h ^= h >> 13
    h = (h * 0xC2B2AE35) & 0xFFFFFFFF
    h ^= h >> 16
    return h


def _kmer_len(n: int) -> int:
    # utils.py:906 and :1247 — k = ceil(log4(n * 0.99 / 0.01))
    return ceil(log((n * (1 - 0.01)) / 0.01, 4))


@lru_cache(maxsize=1 << 16)
def _canonical_hash(kmer: str) -> int:
    """utils.py:1259-1269: hash (seed 39, unsigned) of the lexicographically smaller of the k-mer
    and its reverse complement (IUPAC-aware complement; lower case raises KeyError there)."""
    b = kmer.encode("ascii")
    for ch in b:
        if ch not in b"ACGTMRWSYKVHDBXN":
            raise KeyError(chr(ch))
    rc = b.translate(_COMP)[::-1]
    return murmurhash3_32(b if b < rc else rc, 39)


@lru_cache(maxsize=1 << 16)
def _hash_set(seq: str, k: int) -> frozenset:
    return frozenset(_canonical_hash(seq[i:i + k]) for i in range(len(seq) - k + 1))


def jaccard_similarity(a: str, b: str) -> float:
    """utils.py:1243-1255."""
    k = _kmer_len(len(a))
    sa, sb = _hash_set(a, k), _hash_set(b, k)
    return len(sa & sb) / len(sa | sb)


@lru_cache(maxsize=1 << 12)
def _sf(c: int, mu: float) -> float:
    # 1 - poisson.cdf(c - 1, mu), exactly as the reference composes it (utils.py:881-888, 929)
    return 1 - poisson.cdf(c - 1, mu)


@lru_cache(maxsize=1 << 16)
def _word_counts(seq: str, k: int):
    kmers = [seq[i:i + k] for i in range(len(seq) - k + 1)]
    return kmers, Counter(kmers)


def pac(seqA: str, seqB: str):
    """utils.py:892-949 — van Helden's Poisson similarity (additive, product)."""
    k = _kmer_len(len(seqA))
    a_kmers, a_wc = _word_counts(seqA, k)
    b_kmers, b_wc = _word_counts(seqB, k)
    na, nb = len(a_kmers), len(b_kmers)
    prod = None
    sim_sum = None
    for word in a_kmers:
        ca, cb = a_wc.get(word, 0), b_wc.get(word, 0)
        c = min(ca, cb)
        if c > 0:
            mw_a = (ca / na) * na
            mw_b = (cb / nb) * nb
            prob = _sf(c, mw_a) * _sf(c, mw_b)
        else:
            prob = 1
        prod = prob if prod is None else prod * prob
        sim_sum = (1 - prob) if sim_sum is None else sim_sum + (1 - prob)
    return sim_sum / na, 1 - (prod ** (1 / na))


def minhash_similarity(seq: str, true_seqs) -> float:
    """utils.py:1043-1048: mean Jaccard similarity of (true, seq) over the true set."""
    return float(np.mean([jaccard_similarity(t, seq) for t in true_seqs]))


def pac_similarity(seq: str, true_seqs) -> dict:
    """utils.py:1051-1058 (the reference calls pac twice per pair; once suffices)."""
    vals = [pac(t, seq) for t in true_seqs]
    return {"PAS": float(np.mean([v[0] for v in vals])), "PPS": float(np.mean([v[1] for v in vals]))}
