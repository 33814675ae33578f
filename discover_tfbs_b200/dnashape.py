"""DNAshapeR-shaped host API over the GPU shape kernels.

The reference obtains MGW / ProT / Roll / HelT / EP out of process from the Bioconductor
package DNAshapeR (get_DNA_shapes.R:23-25, create_bkg_seqs.py:365-366) and then parses its
text output (build_features.py:193-198, 215-241).  Here the pentamer lookup runs on the GPU
straight from the 2-bit encoded sequence.

PENTAMER TABLES.  The real 1024-entry query tables are not redistributable with this repository (and
R is absent from the build image), so every product entry point REQUIRES tables: pass `tables=` /
`kinds=`, or point the environment variable TFBS_SHAPE_TABLES at a table file; without either the
call raises (`resolve_tables`).  `synthetic_tables` (seeded values in realistic ranges) exists for
tests, bench.py and examples only and must be requested with `synthetic=True`.

Where the real tables come from — DNAshapeR itself, through the text format this module already
reads (no access to DNAshapeR internals needed):
    1. `write_pentamer_fasta("pentamers.fa")` writes the 1024 pentamers as 5-bp records;
    2. in R: `library(DNAshapeR); getShape("pentamers.fa")` (get_DNA_shapes.R:23-25 does exactly
       this per FASTA file) writes pentamers.fa.{MGW,HelT,ProT,Roll,EP};
    3. `tables_from_shape_files("pentamers.fa")` reads them back: a 5-bp record has exactly one
       pentamer, so token 2 of a per-nucleotide track is its table value and tokens 1 / 2 of a per-step
       track are Y1 / Y2;
    4. `save_tables("dnashape_tables.tsv", tables, kinds)`, then `export TFBS_SHAPE_TABLES=...`.

Table file format (`save_tables` / `load_tables`): tab-separated text, one header line
`pentamer<TAB>col...`, one line per pentamer (ACGT, upper case).  A per-nucleotide shape X has one
column `X`; a per-step shape Y has two, `Y_1` (value for the step before the centre) and `Y_2` (step
after it).  Either all 1024 pentamers, or only one of each reverse-complement pair (512): the other
is filled in by symmetry (per-nucleotide: same value; per-step: Y1(P) = Y2(revcomp P)).
"""
import os
import warnings

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


_CODE = {"A": 0, "C": 1, "G": 2, "T": 3}
PENTAMERS = ["".join("ACGT"[(i >> (2 * (4 - k))) & 3] for k in range(5)) for i in range(1024)]


def resolve_tables(tables=None, kinds=None, shapes=SHAPES, synthetic: bool = False):
    """The (tables, kinds) a product entry point uses: the arguments if given; else the file named by
    $TFBS_SHAPE_TABLES; else — only when the caller explicitly passes synthetic=True (tests, bench,
    examples) — the seeded stand-in tables.  Anything else raises: shape values are never fabricated
    silently."""
    if tables is not None:
        if kinds is None:
            kinds = np.array([KIND[s] for s in shapes], dtype=np.int32)
        return tables, kinds
    path = os.environ.get("TFBS_SHAPE_TABLES")
    if path:
        return load_tables(path, shapes)
    if synthetic:
        warnings.warn("using SYNTHETIC pentamer tables: shape values are not DNAshapeR's", stacklevel=3)
        return synthetic_tables(shapes)
    raise ValueError(
        "no DNA-shape pentamer tables: pass tables=/kinds= (dnashape.load_tables(path)) or set "
        "TFBS_SHAPE_TABLES to a table file; see discover_tfbs_b200/dnashape.py for how to export the real "
        "tables from DNAshapeR.  (synthetic=True selects seeded stand-in tables for tests only.)")


def write_pentamer_fasta(path: str) -> str:
    """The 1024 pentamers as 5-bp FASTA records (input for DNAshapeR::getShape, see module docstring)."""
    with open(path, "w") as fh:
        for p in PENTAMERS:
            fh.write(f">{p}\n{p}\n")
    return path


def tables_from_shape_files(fasta_path: str, shapes=SHAPES):
    """Read `<fasta_path>.<shape>` files DNAshapeR wrote for the pentamer FASTA -> (tables, kinds)."""
    from .utils import parse_fasta

    mapping = {}
    for name in shapes:
        rows = {}
        for pent, body in parse_fasta(f"{fasta_path}.{name}"):
            tok = body.strip().split(",")
            rows[pent] = float(tok[2]) if KIND[name] == _lib.PER_NUCLEOTIDE else (float(tok[1]), float(tok[2]))
        mapping[name] = rows
    return load_tables(mapping, shapes)


def save_tables(path: str, tables, kinds, shapes=SHAPES) -> str:
    """Write the table file format described in the module docstring (all 1024 pentamers)."""
    cols = []
    for t, name in enumerate(shapes):
        cols += [name] if kinds[t] == _lib.PER_NUCLEOTIDE else [f"{name}_1", f"{name}_2"]
    with open(path, "w") as fh:
        fh.write("pentamer\t" + "\t".join(cols) + "\n")
        for i, p in enumerate(PENTAMERS):
            vals = []
            for t in range(len(shapes)):
                vals += [tables[t][0][i]] if kinds[t] == _lib.PER_NUCLEOTIDE else [tables[t][0][i], tables[t][1][i]]
            fh.write(p + "\t" + "\t".join(repr(float(np.float32(v))) for v in vals) + "\n")
    return path


def _read_table_file(path: str, shapes):
    with open(path) as fh:
        header = fh.readline().rstrip("\n").split("\t")
        if not header or header[0] != "pentamer":
            raise ValueError(f"{path}: first column of a shape table file must be 'pentamer'")
        col = {name: i for i, name in enumerate(header)}
        mapping = {name: {} for name in shapes}
        for line in fh:
            f = line.rstrip("\n").split("\t")
            if len(f) < 2:
                continue
            for name in shapes:
                if KIND[name] == _lib.PER_NUCLEOTIDE:
                    if name not in col:
                        raise ValueError(f"{path}: no column '{name}'")
                    mapping[name][f[0]] = float(f[col[name]])
                else:
                    if f"{name}_1" not in col or f"{name}_2" not in col:
                        raise ValueError(f"{path}: no columns '{name}_1' / '{name}_2'")
                    mapping[name][f[0]] = (float(f[col[f"{name}_1"]]), float(f[col[f"{name}_2"]]))
    return mapping


def load_tables(source, shapes=SHAPES):
    """Real tables -> (float32[T][2][1024], kinds).  `source` is a table file path (format in the module
    docstring) or a mapping {shape: {pentamer str: value or (y1, y2)}}.  Pentamers missing from the
    input are filled from their reverse complement; anything still missing raises."""
    mapping = _read_table_file(source, shapes) if isinstance(source, (str, os.PathLike)) else source
    tables = np.full((len(shapes), 2, 1024), np.nan, dtype=np.float32)
    kinds = np.array([KIND[s] for s in shapes], dtype=np.int32)
    rc = _revcomp_index(np.arange(1024))
    for t, name in enumerate(shapes):
        for pent, val in mapping[name].items():
            if len(pent) != 5:
                raise ValueError(f"'{pent}' is not a pentamer")
            i = 0
            for ch in pent.upper():
                i = i * 4 + _CODE[ch]
            if kinds[t] == _lib.PER_NUCLEOTIDE:
                tables[t, 0, i] = tables[t, 1, i] = float(val)
            else:
                tables[t, 0, i], tables[t, 1, i] = float(val[0]), float(val[1])
        # reverse-complement symmetry fills the other half of a 512-row table
        miss = np.isnan(tables[t, 0])
        if kinds[t] == _lib.PER_NUCLEOTIDE:
            tables[t, 0, miss] = tables[t, 0, rc[miss]]
            tables[t, 1] = tables[t, 0]
        else:
            tables[t, 0, miss] = tables[t, 1, rc[miss]]
            miss2 = np.isnan(tables[t, 1])
            tables[t, 1, miss2] = tables[t, 0, rc[miss2]]
    if np.isnan(tables).any():
        raise ValueError("shape table is missing pentamers (and their reverse complements)")
    return tables, kinds


def to_reference_float64(values: np.ndarray) -> np.ndarray:
    """fp32 shape values -> the float64 the reference would hold after DNAshapeR printed them with
    two decimals and Python parsed them back (float("%.2f" % v)).  Exact: a float32 times 100 is
    exactly representable in float64, so rint() performs the same round-half-even on the exact
    binary value that printf does, and c / 100 is the correctly rounded double of the decimal."""
    v = np.array(values, dtype=np.float64)   # one fresh float64 copy, then in place (no further temporaries)
    np.multiply(v, 100.0, out=v)
    np.rint(v, out=v)
    np.divide(v, 100.0, out=v)
    return v


def format_tokens(track: np.ndarray, n_tokens: int) -> list:
    """One record's track -> DNAshapeR tokens: '%.2f' or 'NA'."""
    return ["NA" if v != v else "%.2f" % v for v in np.asarray(track[:n_tokens], dtype=np.float64)]


def getDNAShape(fasta_path: str, shape: str, ctx=None, tables=None, kinds=None, shapes=SHAPES, synthetic=False):
    """DNAshapeR::getDNAShape(filename, shapeType) (create_bkg_seqs.py:365-366): writes
    `<fasta_path>.<shape>` with one '>name' line per record followed by comma-separated values
    wrapped at 20 per line, NA at the undefined ends.  Tables: see `resolve_tables` (raises without)."""
    from .utils import parse_fasta

    tables, kinds = resolve_tables(tables, kinds, shapes, synthetic)  # raises before any GPU work
    own = ctx is None
    ctx = _lib.Context(0) if own else ctx
    try:
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
