"""Biopython-shaped PSSM API running on the B200 scan kernels.

The north star names Biopython's `Bio.motifs.matrix.PositionSpecificScoringMatrix` as the
scoring contract (the reference itself finds candidates with exact-match BLASTn,
utils.py:340-398).  `PSSM.calculate` / `PSSM.search` keep that call shape; `MotifScanner`
is the batched form (many motifs, one pass over the sequence) the feature builder uses.
"""
import math

import numpy as np

from . import _lib

LETTERS = "ACGT"
_default_ctx = None


def default_context() -> _lib.Context:
    global _default_ctx
    if _default_ctx is None or _default_ctx._h is None:
        _default_ctx = _lib.Context(0)
    return _default_ctx


def normalize(counts, pseudocounts=None) -> np.ndarray:
    """FrequencyPositionMatrix.normalize: counts[L][4] -> probabilities (host, tiny)."""
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
    """PositionWeightMatrix.log_odds: log2(p / bg), p == 0 -> -inf, uniform bg by default."""
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


class PSSM:
    """One position-specific scoring matrix (log-odds[L][4], ACGT order)."""

    def __init__(self, logodds, ctx: _lib.Context = None):
        self.matrix = np.ascontiguousarray(logodds, dtype=np.float64)
        self.length = self.matrix.shape[0]
        self._ctx = ctx
        self._lib = None

    @property
    def ctx(self):
        return self._ctx or default_context()

    def _library(self):
        if self._lib is None or self._lib._h is None:
            self._lib = self.ctx.upload_pwms(self.matrix[None])
        return self._lib

    def reverse_complement(self) -> "PSSM":
        return PSSM(self.matrix[::-1, ::-1], self._ctx)

    @property
    def max(self) -> float:
        return float(self.matrix.max(axis=1).sum())

    @property
    def min(self) -> float:
        return float(self.matrix.min(axis=1).sum())

    def calculate(self, sequence):
        """float32[n-L+1] forward-strand scores (a bare float when n == L, as Biopython)."""
        a = _lib.as_ascii(sequence)
        if a.size < self.length:
            raise ValueError("sequence shorter than the motif")
        enc = self.ctx.encode(a)
        fwd, _ = self.ctx.scan_dense(enc, self._library())
        enc.close()
        out = fwd[0, : a.size - self.length + 1]
        return out[0] if out.size == 1 else out

    def search(self, sequence, threshold=0.0, both=True):
        """Yield (position, score) with score >= threshold; reverse-strand hits carry the
        negative position p - len(sequence) (Biopython convention), ordered by p, fwd first."""
        a = _lib.as_ascii(sequence)
        n = a.size
        enc = self.ctx.encode(a)
        hits = self.ctx.scan_hits(enc, self._library(), [threshold], sort=True)
        enc.close()
        for h in hits:
            p = int(h["pos"])
            if p >= 0:
                yield p, h["score"]
            elif both:
                yield (~p) - n, h["score"]


class MotifScanner:
    """A motif library resident on one GPU; scans encoded buffers in dense or hits mode."""

    def __init__(self, logodds, lens=None, ctx: _lib.Context = None):
        self.ctx = ctx or default_context()
        self.library = self.ctx.upload_pwms(logodds, lens)
        self.lens = np.asarray(self.library.lens)
        self.M = self.library.M

    def dense(self, enc: _lib.EncodedSeq, to_host=True):
        return self.ctx.scan_dense(enc, self.library, to_host=to_host)

    def hits(self, enc: _lib.EncodedSeq, thresholds, **kw):
        return self.ctx.scan_hits(enc, self.library, thresholds, **kw)

    def units(self, n: int) -> int:
        """motif x bp units of one scan of an n-base buffer: sum_m (n - L_m + 1)."""
        return int(sum(max(n - int(L) + 1, 0) for L in self.lens))


# ---- motif input and threshold selection (host helpers; Biopython-shaped) ---------------------------
def read_jaspar(path_or_text, pseudocounts=0.5, background=None):
    """Parse JASPAR-format count matrices (the format `Bio.motifs.parse(handle, "jaspar")` reads):

        >MA0001.1 AGL3
        A  [ 0  3 79 ... ]
        C  [94 75  4 ... ]
        G  [ 1  0  3 ... ]
        T  [ 2 19 11 ... ]

    Returns [(name, logodds float64[L][4]), ...] after normalize(pseudocounts) and log_odds(background)."""
    import os
    import re

    text = open(path_or_text).read() if os.path.exists(str(path_or_text)) else str(path_or_text)
    motifs, name, rows = [], None, {}

    def flush():
        if name is not None and len(rows) == 4:
            counts = np.array([rows[b] for b in LETTERS], dtype=np.float64).T
            motifs.append((name, log_odds(normalize(counts, pseudocounts), background)))

    for line in text.splitlines():
        line = line.strip()
        if not line:
            continue
        if line.startswith(">"):
            flush()
            name, rows = line[1:].strip(), {}
        else:
            m = re.match(r"^([ACGT])\s*\[?\s*([^\]]*)\]?\s*$", line)
            if m:
                rows[m.group(1)] = [float(x) for x in m.group(2).split()]
    flush()
    return motifs


def read_meme(path_or_text, pseudocounts=0.5, background=None):
    """Parse MEME "minimal" motif files (the format `Bio.motifs.parse(handle, "minimal")` reads):

        MOTIF MA0004.1 Arnt
        letter-probability matrix: alength= 4 w= 6 nsites= 20 E= 0
          0.200000  0.800000  0.000000  0.000000
          ...

    Probabilities are turned back into counts with `nsites` (20 when absent), then treated like
    read_jaspar: normalize(pseudocounts), log_odds(background).  Returns [(name, logodds[L][4])]."""
    import os
    import re

    text = open(path_or_text).read() if os.path.exists(str(path_or_text)) else str(path_or_text)
    motifs = []
    lines = text.splitlines()
    i = 0
    name = None
    while i < len(lines):
        line = lines[i].strip()
        if line.startswith("MOTIF"):
            name = line[5:].strip()
        elif line.startswith("letter-probability matrix") and name is not None:
            w = re.search(r"w=\s*(\d+)", line)
            ns = re.search(r"nsites=\s*([0-9.eE+-]+)", line)
            alen = re.search(r"alength=\s*(\d+)", line)
            if alen and int(alen.group(1)) != 4:
                raise ValueError(f"motif {name}: alength {alen.group(1)} is not a DNA alphabet")
            nsites = float(ns.group(1)) if ns else 20.0
            rows = []
            i += 1
            while i < len(lines) and (w is None or len(rows) < int(w.group(1))):
                parts = lines[i].split()
                if len(parts) != 4:
                    if rows or lines[i].strip():
                        break
                    i += 1
                    continue
                rows.append([float(x) for x in parts])
                i += 1
            if not rows or (w is not None and len(rows) != int(w.group(1))):
                raise ValueError(f"motif {name}: expected {w.group(1) if w else 'some'} matrix rows, found {len(rows)}")
            counts = np.asarray(rows, dtype=np.float64) * nsites
            motifs.append((name, log_odds(normalize(counts, pseudocounts), background)))
            name = None
            continue
        i += 1
    return motifs


def score_distribution(logodds, background=None, precision=1000, under="background"):
    """Discretised exact score distribution of a PSSM — the dynamic programme behind Biopython's
    `pssm.distribution(precision=10**3)` (Bio/motifs/thresholds.py): scores are binned into
    `precision` steps between the minimum and maximum attainable score.  under="background": windows
    drawn from the i.i.d. background; under="motif": windows drawn from the motif itself
    (p = background * 2**logodds per column).
    Returns (scores, probability mass), float64[precision + 1 + L] each."""
    m = np.asarray(logodds, dtype=np.float64)
    bg = np.full(4, 0.25) if background is None else np.array([float(background[b]) for b in LETTERS])
    bg = bg / bg.sum()
    if under not in ("background", "motif"):
        raise ValueError("under must be 'background' or 'motif'")
    finite = np.where(np.isfinite(m), m, np.nan)
    lo = float(np.nansum(np.nanmin(finite, axis=1)))
    hi = float(np.nansum(np.nanmax(finite, axis=1)))
    step = (hi - lo) / precision if hi > lo else 1.0
    nbins = precision + 1 + m.shape[0]  # per-column rounding can push the top score a few bins up
    dist = np.zeros(nbins)
    dist[0] = 1.0
    for j in range(m.shape[0]):
        colmin = float(np.nanmin(finite[j]))
        new = np.zeros(nbins)
        for b in range(4):
            if not np.isfinite(m[j, b]):
                continue  # -inf: the window can never reach any finite threshold
            shift = int(round((m[j, b] - colmin) / step))
            weight = bg[b] if under == "background" else bg[b] * 2.0 ** m[j, b]
            if shift == 0:
                new += dist * weight
            else:
                new[shift:] += dist[:-shift] * weight
        dist = new
    return lo + step * np.arange(nbins), dist


def threshold_fpr(logodds, fpr, background=None, precision=1000):
    """Smallest score threshold whose false-positive rate under the background is <= fpr
    (Biopython: `pssm.distribution().threshold_fpr(fpr)`)."""
    scores, mass = score_distribution(logodds, background, precision)
    tail = np.cumsum(mass[::-1])[::-1]  # P(binned score >= scores[i])
    ok = np.flatnonzero(tail <= fpr)
    # a window's binned score is within half a step per column of its real score: err on the safe side
    margin = 0.5 * (scores[1] - scores[0]) * np.asarray(logodds).shape[0]
    return float(scores[ok[0]] + margin) if ok.size else float(scores[-1] + (scores[1] - scores[0]))


def threshold_fnr(logodds, fnr, background=None, precision=1000):
    """Largest score threshold that still keeps the false-NEGATIVE rate <= fnr, i.e. at most a
    fraction fnr of windows drawn from the motif itself score below it
    (Biopython: `pssm.distribution().threshold_fnr(fnr)`)."""
    scores, mass = score_distribution(logodds, background, precision, under="motif")
    mass = mass / mass.sum()
    below = np.cumsum(mass) - mass  # P(binned score < scores[i]) under the motif
    ok = np.flatnonzero(below <= fnr)
    margin = 0.5 * (scores[1] - scores[0]) * np.asarray(logodds).shape[0]  # binning error, safe side
    return float(scores[ok[-1]] - margin)


def threshold_balanced(logodds, rate_proportion=1.0, background=None, precision=1000):
    """Threshold where false-negative rate ~= rate_proportion * false-positive rate
    (Biopython: `pssm.distribution().threshold_balanced(rate_proportion)`)."""
    scores, bg_mass = score_distribution(logodds, background, precision)
    _, mo_mass = score_distribution(logodds, background, precision, under="motif")
    mo_mass = mo_mass / mo_mass.sum()
    fpr = np.cumsum(bg_mass[::-1])[::-1]          # P_bg(score >= s_i)
    fnr = np.cumsum(mo_mass) - mo_mass            # P_motif(score < s_i)
    i = int(np.argmin(np.abs(fnr - rate_proportion * fpr)))
    return float(scores[i])
