"""CPU oracle for the DNA-shape path — DNAshapeR semantics.  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED for the pentamer arithmetic: it lives in Bioconductor DNAshapeR (called at
create_bkg_seqs.py:365-366 and get_DNA_shapes.R:23-25), which is absent here together with its
query table; restated from SURVEY.md Appendix A.2 (C loop in tfbs_oracle.c).  The window slicing
the reference applies to DNAshapeR's output (utils.py:1061-1084, build_features.py:215-241) IS
pinned: tests/golden/shape_golden.json holds the reference's own outputs for token lists
produced here.
"""
import numpy as np

from . import clib


def tracks(seq, tables, kinds) -> np.ndarray:
    """float32[T][n] for ONE record; NaN = NA; per-step tracks use entries [0, n-1)."""
    s = clib.as_bytes(seq)
    tab = np.ascontiguousarray(tables, dtype=np.float32)
    kd = np.ascontiguousarray(kinds, dtype=np.int32)
    out = np.empty((kd.size, s.size), dtype=np.float32)
    clib.lib().oracle_shape_tracks(clib.ptr(s), s.size, clib.ptr(tab), clib.ptr(kd), kd.size, clib.ptr(out))
    return out


def tracks_set(buf, starts, lens, tables, kinds, threads=0) -> np.ndarray:
    """float32[T][len(buf)] for many records laid out in one buffer (multi-threaded)."""
    b = clib.as_bytes(buf)
    st = np.ascontiguousarray(starts, dtype=np.int64)
    ln = np.ascontiguousarray(lens, dtype=np.int64)
    tab = np.ascontiguousarray(tables, dtype=np.float32)
    kd = np.ascontiguousarray(kinds, dtype=np.int32)
    out = np.full((kd.size, b.size), np.nan, dtype=np.float32)
    clib.lib().oracle_shape_tracks_set(clib.ptr(b), clib.ptr(st), clib.ptr(ln), st.size, b.size,
                                       clib.ptr(tab), clib.ptr(kd), kd.size, int(threads), clib.ptr(out))
    return out


def tokens(track: np.ndarray, kind: int, n: int, trailing_empty: bool = False) -> list:
    """DNAshapeR text tokens of one record's track: '%.2f' or 'NA'; n tokens for per-nucleotide
    tracks, n-1 for per-step tracks; optionally the trailing '' that the awk flattening of
    .COMMANDS.txt:290-293 leaves at the end of every one-line record."""
    ntok = n if kind == 0 else n - 1
    out = ["NA" if v != v else "%.2f" % float(v) for v in track[:max(ntok, 0)]]
    if trailing_empty:
        out.append("")
    return out


def get_dna_shape(loc: str, token_list: list, shape: str) -> dict:
    """Restatement of utils.py:1061-1084 (_get_DNA_shape) for one location."""
    chrom, location = loc.split(":")
    strand = location.split("(")[1][0]
    start = int(location.split("-")[0])
    end = int(location.split("(")[0].split("-")[1])
    sl = token_list[start:end + 1] if strand == "+" else token_list[end:start - 1:-1]
    out = {}
    for i, tok in enumerate(sl):
        try:
            out[f"{shape}_{i + 1:02d}"] = float(tok)
        except ValueError:
            out[f"{shape}_{i + 1:02d}"] = 0.0
    return out


def bkg_trim(token_list: list, shape: str) -> dict:
    """Restatement of build_features.py:215-241 for one background record."""
    sl = token_list[2:-2] if shape in ("MGW", "ProT", "EP") else token_list[1:-1]
    out = {}
    for i, tok in enumerate(sl):
        try:
            out["{0}_{1:02d}".format(shape, i + 1)] = float(tok)
        except Exception:
            out["{0}_{1:02d}".format(shape, i + 1)] = 0.0
    return out
