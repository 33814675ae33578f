"""DNAshapeR-shaped host API over the GPU shape kernels.

The reference obtains MGW / ProT / Roll / HelT / EP out of process from the Bioconductor
package DNAshapeR (get_DNA_shapes.R:23-25, create_bkg_seqs.py:365-366) and then parses its
text output (build_features.py:193-198, 215-241).  Here the pentamer lookup runs on the GPU
straight from the 2-bit encoded sequence.  The real 1024-entry query tables are not
redistributable with this repository (and R is absent from the image), so `synthetic_tables`
provides seeded tables in realistic ranges; `load_tables` accepts the real ones.
"""
import numpy as np

from . import _lib

# column-block order of the reference's feature table (utils.py:1134-1170)
SHAPES = ("EP", "HelT", "MGW", "ProT", "Roll")
KIND = {"EP": _lib.PER_NUCLEOTIDE, "MGW": _lib.PER_NUCLEOTIDE, "ProT": _lib.PER_NUCLEOTIDE,
        "HelT": _lib.PER_STEP, "Roll": _lib.PER_STEP}
# approximate value ranges of the published pentamer tables
_RANGE = {"MGW": (2.9, 6.2), "ProT": (-16.5, 0.0), "Roll": (-8.6, 8.6), "HelT": (31.0, 38.0),
          "EP": (-13.0, -4.0)}


def _revcomp_index(idx: np.ndarray) -> np.ndarray:
    """Big-endian pentamer index of the reverse complement."""
    out = np.zeros_like(idx)
    for k in range(5):
        code = (idx >> (2 * (4 - k))) & 3
        out |= (3 - code) << (2 * k)
    return out


def synthetic_tables(shapes=SHAPES, seed: int = 39):
    """Seeded stand-in for the DNAshape query table: float32[T][2][1024] + kinds int32[T].

    Values have two decimals (as the published tables and DNAshapeR's text output do) and obey
    the physical symmetries: a per-nucleotide value is equal for a pentamer and its reverse
    complement; for per-step tracks Y1(P) == Y2(revcomp(P)).  Index = sum code[k] * 4**(4-k)."""
    idx = np.arange(1024)
    rc = _revcomp_index(idx)
    tables = np.zeros((len(shapes), 2, 1024), dtype=np.float32)
    kinds = np.zeros(len(shapes), dtype=np.int32)
    for t, name in enumerate(shapes):
        rng = np.random.default_rng([seed, sum(map(ord, name))])
        lo, hi = _RANGE[name]
        kinds[t] = KIND[name]
        raw = np.round(rng.uniform(lo, hi, size=(2, 1024)), 2)
        if kinds[t] == _lib.PER_NUCLEOTIDE:
            canon = np.minimum(idx, rc)
            tables[t, 0] = raw[0][canon]
            tables[t, 1] = tables[t, 0]
        else:
            y1 = raw[0]
            tables[t, 0] = y1
            tables[t, 1] = y1[rc]
    return tables, kinds


def load_tables(mapping: dict, shapes=SHAPES):
    """Build the table array from {shape: {pentamer str: value or (y1, y2)}} (real tables)."""
    tables = np.full((len(shapes), 2, 1024), np.nan, dtype=np.float32)
    kinds = np.array([KIND[s] for s in shapes], dtype=np.int32)
    code = {"A": 0, "C": 1, "G": 2, "T": 3}
    for t, name in enumerate(shapes):
        for pent, val in mapping[name].items():
            i = 0
            for ch in pent.upper():
                i = i * 4 + code[ch]
            if kinds[t] == _lib.PER_NUCLEOTIDE:
                tables[t, 0, i] = tables[t, 1, i] = float(val)
            else:
                tables[t, 0, i], tables[t, 1, i] = float(val[0]), float(val[1])
    if np.isnan(tables).any():
        raise ValueError("shape table is missing pentamers")
    return tables, kinds


def to_reference_float64(values: np.ndarray) -> np.ndarray:
    """fp32 shape values -> the float64 the reference would hold after DNAshapeR printed them with
    two decimals and Python parsed them back (float("%.2f" % v)).  Exact: a float32 times 100 is
    exactly representable in float64, so rint() performs the same round-half-even on the exact
    binary value that printf does, and c / 100 is the correctly rounded double of the decimal."""
    v = np.asarray(values, dtype=np.float64)
    return np.rint(v * 100.0) / 100.0


def format_tokens(track: np.ndarray, n_tokens: int) -> list:
    """One record's track -> DNAshapeR tokens: '%.2f' or 'NA'."""
    return ["NA" if v != v else "%.2f" % v for v in np.asarray(track[:n_tokens], dtype=np.float64)]


def getDNAShape(fasta_path: str, shape: str, ctx=None, tables=None, kinds=None, shapes=SHAPES):
    """DNAshapeR::getDNAShape(filename, shapeType) (create_bkg_seqs.py:365-366): writes
    `<fasta_path>.<shape>` with one '>name' line per record followed by comma-separated values
    wrapped at 20 per line, NA at the undefined ends."""
    from .utils import parse_fasta

    own = ctx is None
    ctx = _lib.Context(0) if own else ctx
    try:
        if tables is None:
            tables, kinds = synthetic_tables(shapes)
        t = list(shapes).index(shape)
        tab = ctx.upload_shapes(tables[t: t + 1], kinds[t: t + 1])
        names, recs = zip(*parse_fasta(fasta_path))
        from .synth import join_records
        buf, starts, lens = join_records(recs)
        enc = ctx.encode(buf)
        tracks = ctx.shape_tracks(enc, tab, starts, lens)[0]
        out_path = f"{fasta_path}.{shape}"
        with open(out_path, "w") as fh:
            for name, s, ln in zip(names, starts, lens):
                ntok = int(ln) if kinds[t] == _lib.PER_NUCLEOTIDE else int(ln) - 1
                toks = format_tokens(tracks[s: s + ntok], ntok)
                fh.write(f">{name}\n")
                for i in range(0, len(toks), 20):
                    fh.write(",".join(toks[i: i + 20]) + "\n")
        return out_path
    finally:
        if own:
            ctx.close()
