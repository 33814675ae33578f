"""CPU oracle for the PWM path — Biopython PSSM semantics.  TEST INFRASTRUCTURE ONLY.

PARITY STATUS: the arithmetic restated here lives in Biopython (Bio/motifs/matrix.py,
Bio/motifs/_pwm.c), which the reference never calls, which is not installed
here and whose version the reference does not pin — there is no run of Biopython behind this file.
The restatement follows SURVEY.md Appendix A.1; it is cross-checked three ways in
tests/test_oracle.py (C loop, NumPy gather-sum in float64, and the identity
rev(seq) == fwd(revcomp(seq))[::-1]) and PINNED TO THE PUBLISHED KNOWN-ANSWER EXAMPLE of the
Biopython Tutorial (Bio.motifs chapter): counts -> normalize -> log_odds -> calculate (22 float32
scores to the 8 printed decimals) -> search -> max / min / mean / std
(tests/golden/biopython_tutorial.json, which states how the vectors were obtained offline;
tests/test_oracle.py::test_pwm_oracle_reproduces_the_biopython_tutorial).  One published example is not
a run of the library: semantics the example does not exercise (-inf entries, lower case, N handling)
rest on the restatement alone.
"""
import math

import numpy as np

from . import clib

LETTERS = "ACGT"


def normalize(counts, pseudocounts=None) -> np.ndarray:
    """FrequencyPositionMatrix.normalize: counts[m][4] (ACGT) -> probabilities[m][4].

    pseudocounts: None -> 0, number -> same for every letter, dict -> per letter."""
    counts = np.asarray(counts, dtype=np.float64)
    if pseudocounts is None:
        pc = np.zeros(4)
    elif isinstance(pseudocounts, dict):
        pc = np.array([float(pseudocounts[b]) for b in LETTERS])
    else:
        pc = np.full(4, float(pseudocounts))
    p = counts + pc[None, :]
    return p / p.sum(axis=1, keepdims=True)


def log_odds(pwm, background=None) -> np.ndarray:
    """PositionWeightMatrix.log_odds: log2(p / bg); p == 0 -> -inf; bg None -> uniform; a dict
    background is renormalised to sum 1.  (bg == 0: +inf if p > 0 else nan.)"""
    pwm = np.asarray(pwm, dtype=np.float64)
    if background is None:
        bg = np.full(4, 0.25)
    else:
        bg = np.array([float(background[b]) for b in LETTERS])
        bg = bg / bg.sum()
    out = np.empty_like(pwm)
    for j in range(pwm.shape[0]):
        for b in range(4):
            p, q = pwm[j, b], bg[b]
            if q > 0:
                out[j, b] = math.log(p / q, 2) if p > 0 else -math.inf
            else:
                out[j, b] = math.inf if p > 0 else math.nan
    return out


def reverse_complement(pssm) -> np.ndarray:
    """rc[j][b] = M[m-1-j][3-b] (A<->T, C<->G, columns reversed)."""
    pssm = np.asarray(pssm, dtype=np.float64)
    return np.ascontiguousarray(pssm[::-1, ::-1])


def calculate(seq, pssm, threads=None) -> np.ndarray:
    """PositionSpecificScoringMatrix.calculate via the C loop: float32[n-m+1] (always an array;
    Biopython returns the bare scalar when n == m).  threads: None -> the plain loop; an int splits
    the window starts over that many OpenMP threads (0 = all cores), same output."""
    s = clib.as_bytes(seq)
    m = np.ascontiguousarray(pssm, dtype=np.float64)
    n, L = s.size, m.shape[0]
    if n < L:
        raise ValueError("sequence shorter than motif")
    out = np.empty(n - L + 1, dtype=np.float32)
    if threads is None:
        clib.lib().oracle_pwm_calculate(clib.ptr(s), n, clib.ptr(m), L, clib.ptr(out))
    else:
        clib.lib().oracle_pwm_calculate_mt(clib.ptr(s), n, clib.ptr(m), L, clib.ptr(out), int(threads) or host_threads())
    return out


_CODE = np.full(256, -1, dtype=np.int64)
for _i, _b in enumerate(LETTERS):
    _CODE[ord(_b)] = _i
    _CODE[ord(_b.lower())] = _i


def calculate_numpy(seq, pssm) -> np.ndarray:
    """Independent NumPy restatement (gather + sequential float64 sum, then float32)."""
    s = clib.as_bytes(seq)
    m = np.asarray(pssm, dtype=np.float64)
    L = m.shape[0]
    code = _CODE[s]
    nwin = s.size - L + 1
    acc = np.zeros(nwin, dtype=np.float64)
    bad = np.zeros(nwin, dtype=bool)
    for j in range(L):
        c = code[j:j + nwin]
        bad |= c < 0
        acc = acc + m[j, np.where(c < 0, 0, c)]
    out = acc.astype(np.float32)
    out[bad] = np.nan
    return out


def search(seq, pssm, threshold=0.0, both=True):
    """PositionSpecificScoringMatrix.search: list of (position, float32 score); reverse-strand
    hits carry position p - len(seq) (negative).  Ordered by genomic index, fwd before rev."""
    s = clib.as_bytes(seq)
    n = s.size
    thr = np.float32(threshold)
    fwd = calculate(s, pssm)
    rev = calculate(s, reverse_complement(pssm)) if both else None
    hits = []
    for p in range(fwd.size):
        if fwd[p] >= thr:
            hits.append((p, fwd[p]))
        if both and rev[p] >= thr:
            hits.append((p - n, rev[p]))
    return hits


def host_threads() -> int:
    """Host cores this process may run on (not OMP_NUM_THREADS: torchrun exports that as 1)."""
    import os
    try:
        return max(len(os.sched_getaffinity(0)), 1)
    except AttributeError:
        return max(os.cpu_count() or 1, 1)


def search_library(seq, matrices, lens, thr, cap=None, threads=None):
    """Both-strand thresholded scan of one sequence with a motif library via the C loop.
    matrices float64[M][Lmax][4]; returns structured arrays (pos, motif, strand, score)
    in (motif, pos, strand) order.  threads: None -> the single-threaded loop; an int runs the
    OpenMP form with that many threads (0 = all cores), same output."""
    if threads is not None:
        return _search_library_mt(seq, matrices, lens, thr, int(threads) or host_threads())
    s = clib.as_bytes(seq)
    mats = np.ascontiguousarray(matrices, dtype=np.float64)
    lens = np.ascontiguousarray(lens, dtype=np.int32)
    thr = np.ascontiguousarray(thr, dtype=np.float32)
    M, Lmax = mats.shape[0], mats.shape[1]
    L = clib.lib()
    if cap is None:
        cap = L.oracle_pwm_search(clib.ptr(s), s.size, clib.ptr(mats), clib.ptr(lens), M, Lmax,
                                  clib.ptr(thr), None, None, None, None, 0)
    pos = np.empty(cap, dtype=np.int64)
    mot = np.empty(cap, dtype=np.int32)
    strand = np.empty(cap, dtype=np.int32)
    score = np.empty(cap, dtype=np.float32)
    cnt = L.oracle_pwm_search(clib.ptr(s), s.size, clib.ptr(mats), clib.ptr(lens), M, Lmax,
                              clib.ptr(thr), clib.ptr(pos), clib.ptr(mot), clib.ptr(strand),
                              clib.ptr(score), cap)
    k = min(cnt, cap)
    return pos[:k], mot[:k], strand[:k], score[:k], cnt


def _search_library_mt(seq, matrices, lens, thr, threads):
    s = clib.as_bytes(seq)
    mats = np.ascontiguousarray(matrices, dtype=np.float64)
    lens = np.ascontiguousarray(lens, dtype=np.int32)
    thr = np.ascontiguousarray(thr, dtype=np.float32)
    M, Lmax = mats.shape[0], mats.shape[1]
    L = clib.lib()
    cap = 1 << 22  # two passes (count, fill) when the hits fit; one more call with the exact size otherwise
    while True:
        pos = np.empty(cap, dtype=np.int64)
        mot = np.empty(cap, dtype=np.int32)
        strand = np.empty(cap, dtype=np.int32)
        score = np.empty(cap, dtype=np.float32)
        cnt = L.oracle_pwm_search_mt(clib.ptr(s), s.size, clib.ptr(mats), clib.ptr(lens), M, Lmax, clib.ptr(thr),
                                     clib.ptr(pos), clib.ptr(mot), clib.ptr(strand), clib.ptr(score), cap, threads)
        if cnt <= cap:
            break
        cap = int(cnt)
    pos, mot, strand, score = pos[:cnt].copy(), mot[:cnt].copy(), strand[:cnt].copy(), score[:cnt].copy()
    return pos, mot, strand, score, int(cnt)


def scan_count(seq, matrices, lens, thr, threads=0):
    """Timed CPU baseline: multi-threaded both-strand scan, returns (hits, checksum, threads)."""
    import ctypes as C
    s = clib.as_bytes(seq)
    mats = np.ascontiguousarray(matrices, dtype=np.float64)
    lens = np.ascontiguousarray(lens, dtype=np.int32)
    thr = np.ascontiguousarray(thr, dtype=np.float32)
    cs = C.c_double(0.0)
    tu = C.c_int32(0)
    cnt = clib.lib().oracle_pwm_scan_count(clib.ptr(s), s.size, clib.ptr(mats), clib.ptr(lens),
                                           mats.shape[0], mats.shape[1], clib.ptr(thr),
                                           int(threads), C.byref(cs), C.byref(tu))
    return int(cnt), float(cs.value), int(tu.value)


def encode(seq):
    """2-bit packing + invalid mask (see tfbs_oracle.c:oracle_encode)."""
    s = clib.as_bytes(seq)
    n = s.size
    packed = np.zeros((n + 15) // 16, dtype=np.uint32)
    mask = np.zeros((n + 31) // 32, dtype=np.uint32)
    clib.lib().oracle_encode(clib.ptr(s), n, clib.ptr(packed), clib.ptr(mask))
    return packed, mask


def base_counts(seq):
    s = clib.as_bytes(seq)
    out = np.zeros(6, dtype=np.int64)
    clib.lib().oracle_base_counts(clib.ptr(s), s.size, clib.ptr(out))
    return out
