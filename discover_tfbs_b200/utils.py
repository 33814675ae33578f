"""Drop-in mirror of the reference's `utils.py` names on the feature-extraction hot path.

Same function names, argument meaning and return types as akshayparopkari/discover_tfbs
`utils.py` (reference file:line cited per function), so `build_features.py:17,200,210` and
`predict.py:19,233` can `from discover_tfbs_b200.utils import build_feature_table,
parse_fasta` unchanged.  The arithmetic runs on a B200 through libtfbs_cuda.so (ctypes,
`_lib.py`); there is no CPU fallback — without the library or a GPU these functions raise.

New capabilities are keyword-only with defaults, so existing calls are unaffected:
  * `genome_wide_shape_data` may be a `GenomeShapeSource` (chromosome FASTA + pentamer tables):
    shape features are then computed on the GPU straight from sequence instead of from the 5
    pre-computed DNAshapeR text files (removes the 71.5 M-string parse of
    build_features.py:193-198);
  * `pwms=` / `pwm_names=` add per-site best log-odds scores on both strands.
"""
from os.path import isfile
from sys import exit
from time import strftime

import numpy as np
import pandas as pd

from . import dnashape, similarity
from .pssm import default_context
from .synth import join_records

__version__ = "0.3.7+b200"

# IUPAC complement table of Bio.Seq.ambiguous_dna_complement (upper-case keys only; the
# reference raises KeyError on lower case, utils.py:104)
_COMPLEMENT = {"A": "T", "C": "G", "G": "C", "T": "A", "M": "K", "R": "Y", "W": "W", "S": "S",
               "Y": "R", "K": "M", "V": "B", "H": "D", "D": "H", "B": "V", "X": "X", "N": "N"}


def reverse_complement(seq: str) -> str:
    """utils.py:97-105.  Host string helper (IUPAC-aware); the kernels use 3 - code."""
    return "".join([_COMPLEMENT[nt] for nt in seq[::-1]])


def parse_fasta(fasta_file: str):
    """utils.py:170-194: generator of (header, sequence); multi-line bodies are joined without
    a separator (Biopython SimpleFastaParser semantics); a missing file exits."""
    if not isfile(fasta_file):
        exit("\n{} does not exist. Please provide a valid FASTA file.\n".format(fasta_file))
    with open(fasta_file) as fh:
        title, chunks = None, []
        for line in fh:
            if line.startswith(">"):
                if title is not None:
                    yield title, "".join(chunks).replace(" ", "").replace("\r", "")
                title, chunks = line[1:].rstrip(), []
            elif title is not None:
                chunks.append(line.rstrip())
        if title is not None:
            yield title, "".join(chunks).replace(" ", "").replace("\r", "")


def calculate_gc_content(sequence: str) -> float:
    """utils.py:224-238: (count('G') + count('C')) / len after upper(); GPU popcount."""
    ctx = default_context()
    enc = ctx.encode(sequence)
    gc = enc.gc_counts([0], [len(sequence)])[0]
    enc.close()
    return int(gc) / len(sequence)


def get_kmers(seq, k=6):
    """utils.py:402-425."""
    if isinstance(seq, str):
        return [seq[i: i + k] for i in range(0, len(seq) - (k - 1), 1)]
    return {s: [s[i: i + k] for i in range(0, len(s) - (k - 1), 1)] for s in seq}


# ---- shape sources ----------------------------------------------------------------------------
class GenomeShapeSource:
    """Chromosome sequences + pentamer tables resident on the GPU.

    Pass an instance as `genome_wide_shape_data` to `build_feature_table` / `_get_DNA_shape`
    instead of the reference's {shape: {chrom: [tokens]}} dict."""

    def __init__(self, genome_fasta=None, records=None, tables=None, kinds=None,
                 shapes=dnashape.SHAPES, trailing_empty=True, ctx=None, synthetic=False):
        # pentamer tables are required (argument or $TFBS_SHAPE_TABLES); never fabricated silently
        tables, kinds = dnashape.resolve_tables(tables, kinds, shapes, synthetic)
        self.ctx = ctx or default_context()
        if records is None:
            records = dict(parse_fasta(genome_fasta))
        self.names = list(records)
        buf, starts, lens = join_records([records[k] for k in self.names])
        self.start = dict(zip(self.names, starts.tolist()))
        self.length = dict(zip(self.names, lens.tolist()))
        self.shapes = tuple(shapes)
        self.kinds = np.asarray(kinds, dtype=np.int32)
        self._buf, self._tables_host = buf, np.asarray(tables, dtype=np.float32)
        self._place(self.ctx)
        self._replicas = {}
        # True: token lists come from the awk-flattened one-line files (.COMMANDS.txt:290-293),
        # whose trailing comma adds one '' token per chromosome
        self.trailing_empty = bool(trailing_empty)

    def _place(self, ctx):
        self.ctx = ctx
        self.enc = ctx.encode(self._buf)
        self.tables = ctx.upload_shapes(self._tables_host, self.kinds)

    def on(self, ctx):
        """The same source resident on another context / device (multi-GPU feature builds)."""
        if ctx is self.ctx:
            return self
        rep = object.__new__(GenomeShapeSource)
        rep.__dict__.update({k: v for k, v in self.__dict__.items() if k not in ("enc", "tables", "ctx", "_replicas")})
        rep._replicas = {}
        rep._place(ctx)
        return rep

    def __bool__(self):
        return True

    def __getitem__(self, shape):  # shapedata["EP"] -> a per-shape view, like the dict
        return _ShapeView(self, shape)

    def sites(self, locs, width=None):
        """float32[S][T][W] for FASTA-header locations 'chrom:start-end(strand)'."""
        chrom, start, end, strand = parse_locations(locs)
        rs = np.array([self.start[c] for c in chrom], dtype=np.int64)
        rl = np.array([self.length[c] for c in chrom], dtype=np.int64)
        if width is None:
            width = int((end - start).max()) + 1 if len(start) else 1
        return self.ctx.shape_sites(self.enc, self.tables, rs, rl, start, end, strand,
                                    width=width, mode=0, trailing_empty=self.trailing_empty)


class _ShapeView:
    def __init__(self, src, shape):
        self.src, self.shape = src, shape


def parse_locations(locs):
    """'chrom:start-end(strand)' (bedtools getfasta -s headers, utils.py:138-149, 1065-1068)
    -> (chrom list, start int64[], end int64[], strand int8[] (+1 / -1))."""
    chrom, start, end, strand = [], [], [], []
    for loc in locs:
        c, location = tuple(loc.split(":"))
        chrom.append(c)
        strand.append(1 if location.split("(")[1][0] == "+" else -1)
        start.append(int(location.split("-")[0]))
        end.append(int(location.split("(")[0].split("-")[1]))
    return chrom, np.array(start, dtype=np.int64), np.array(end, dtype=np.int64), np.array(strand, dtype=np.int8)


def _tokens_to_values(tokens) -> np.ndarray:
    """['NA', '5.12', '', ...] -> float64 with the reference's ValueError -> 0.0 rule."""
    return pd.to_numeric(pd.Series(tokens, dtype=object), errors="coerce").fillna(0.0).to_numpy(dtype=np.float64)


def _get_DNA_shape(loc: str, shapedata, shape: str) -> dict:
    """utils.py:1061-1084: {f'{shape}_{i+1:02d}': float} for one location.

    `shapedata` is either the reference's {chrom: [tokens]} dict or a `_ShapeView` of a
    `GenomeShapeSource`; the slicing ('+': [start:end+1], '-': [end:start-1:-1], NA/'' -> 0.0,
    short slices -> fewer keys) happens on the GPU in both cases."""
    row = _shape_rows([loc], shapedata, shape)[0]
    return {f"{shape}_{i + 1:02d}": float(v) for i, v in enumerate(row) if v == v}


def _shape_rows(locs, shapedata, shape, ctx=None) -> np.ndarray:
    """float64[S][W] of reference-format values (NaN = key absent)."""
    chrom, start, end, strand = parse_locations(locs)
    width = int(np.max(end - start)) + 1 if len(locs) else 1
    width = max(width, 1)
    if isinstance(shapedata, _ShapeView):
        src = shapedata.src
        t = src.shapes.index(shape)
        out = src.sites(locs, width)[:, t, :]
        return dnashape.to_reference_float64(out)
    # legacy dict input: parse each needed chromosome's token list once, gather on the GPU
    ctx = ctx or default_context()
    used = sorted(set(chrom))
    vals, off, ntok, pos = [], {}, {}, 0
    for c in used:
        v = _tokens_to_values(shapedata[c])
        vals.append(v)
        off[c], ntok[c] = pos, v.size
        pos += v.size
    values = np.concatenate(vals) if vals else np.zeros(0)
    # float32 holds every 2-decimal shape value to < 1e-6; round back to the parsed decimal
    out = ctx.track_sites(values.astype(np.float32), [off[c] for c in chrom], [ntok[c] for c in chrom],
                          start, end, strand, width)
    return dnashape.to_reference_float64(out)


def get_shape_data(bedfile: str, shapefiles: list) -> dict:
    """utils.py:554-602 (legacy BED-driven API, no caller in the reference): keys
    '{shape}_pos_{i}' for i >= 1 (index 0 of each slice is skipped, utils.py:594)."""
    print(strftime("%x %X | Parsing DNA shape files"))
    shape_data = {f.split(".")[-1]: {h: s.strip().split(",") for h, s in parse_fasta(f)} for f in shapefiles}
    print(strftime("%x %X | Retrieving DNA shape data"))
    locs = []
    with open(bedfile) as inbed:
        for line in inbed:
            chrom, start, end, _name, _score, strand = tuple(line.strip().split("\t"))
            locs.append(chrom + ":" + start + "-" + end + "(" + strand + ")")
    genome_shape = {name: {} for name in locs}
    for shape, data in shape_data.items():
        rows = _shape_rows(locs, data, shape)
        for name, row in zip(locs, rows):
            for i in range(1, row.size):
                if row[i] == row[i]:
                    genome_shape[name]["{0}_pos_{1}".format(shape, i)] = float(row[i])
    return genome_shape


# ---- feature table ----------------------------------------------------------------------------
_SHAPE_LABELS = {"EP": "Electrostatic potential", "HelT": "Helix Twist", "MGW": "Minor groove width",
                 "ProT": "Propeller twist", "Roll": "Roll"}


def _feature_columns(ctx, locs, seqs, true_seqs, shape_data, minhash, pwms, similarity_features, quiet=False):
    """Feature columns of one set (or one record shard) of sequences on one GPU: ordered dict
    name -> float64[len(seqs)] for the scalar columns, plus 'shape' -> {shape: float64[S][W]}."""
    say = (lambda *_: None) if quiet else (lambda msg: print(strftime(f"%x %X | {msg}")))
    cols = {}
    say("Calculating GC content of input sequences")
    buf, starts, lens = join_records(seqs)
    enc = ctx.encode(buf) if len(seqs) else None
    cols["GC_content"] = enc.gc_counts(starts, lens).astype(np.int64) / lens if enc is not None else np.zeros(0)

    if similarity_features:
        if len(seqs):
            if not (similarity.acgt_only(seqs) and similarity.acgt_only(true_seqs)):
                # no host fallback: the all-pairs kernel works on 2-bit k-mers (the reference raises
                # KeyError on lower case as well, utils.py:104)
                raise ValueError("MinHash/PAS/PPS need upper-case ACGT sequences; pass "
                                 "similarity_features=False to skip these columns")
            # all-pairs kernel (tfbs_similarity_sums); k-mer / hash preparation is vectorised host code
            if minhash:
                say("Calculating MinHash similarity score")
            say("Calculating Poisson similarity scores")
            mh, pas, pps = similarity.similarity_columns_gpu(ctx, seqs, true_seqs)
        else:
            mh = pas = pps = np.zeros(0)
        if minhash:
            cols["MinHash"] = mh
        cols["PAS"] = pas
        cols["PPS"] = pps

    if pwms is not None and enc is not None:
        say("Scoring position weight matrices on both strands")
        mats, names = pwms
        lib = ctx.upload_pwms(list(mats))
        # reduced on the device (tfbs_scan_best): S x M x 2 floats cross PCIe, not the dense [M][n] arrays
        best = ctx.scan_best(enc, lib, starts, lens)
        for m, name in enumerate(names):
            cols[f"{name}_fwd_max"] = best[:, m, 0].astype(np.float64)
            cols[f"{name}_rev_max"] = best[:, m, 1].astype(np.float64)
        lib.close()

    shape_rows = {}
    if shape_data:
        for shape in ("EP", "HelT", "MGW", "ProT", "Roll"):
            say(f"Retrieving {_SHAPE_LABELS[shape]} shape values")
            shape_rows[shape] = _shape_rows(locs, shape_data[shape], shape, ctx=ctx)
    if enc is not None:
        enc.close()
    return cols, shape_rows


def build_feature_table(infasta: str, truefasta: str, genome_wide_shape_data=None, minhash=True, *,
                        pwms=None, pwm_names=None, similarity_features=True, ctx=None, devices=None):
    """utils.py:1088-1173.  Returns a DataFrame with columns `location, sequence, GC_content,
    [MinHash,] PAS, PPS, EP_01.., HelT_01.., MGW_01.., ProT_01.., Roll_01..` in the reference's
    order and dtypes (float64).  GC content and every shape column come from the GPU.

    Keyword-only extensions: `pwms` (list of log-odds matrices) adds `<name>_fwd_max`,
    `<name>_rev_max` best-window scores per sequence; `similarity_features=False` skips the
    MinHash/PAC columns; `devices=[0, 1, ...]` shards the records by index over several GPUs of the
    node, one host thread and context per device, rows returned in input order."""
    print(strftime("%x %X | Begin building feature table"))
    records = {header: seq for header, seq in parse_fasta(infasta)}
    locs, seqs = list(records.keys()), list(records.values())
    tf_data = pd.DataFrame({"location": locs, "sequence": seqs})
    true_seqs = {seq for header, seq in parse_fasta(truefasta)}
    pw = None
    if pwms is not None:
        mats = list(pwms)
        pw = (mats, list(pwm_names) if pwm_names else [f"PWM{i + 1}" for i in range(len(mats))])

    if devices is not None and len(devices) > 1 and len(seqs) >= len(devices):
        from . import sharding

        def shard(c, first, last):
            sd = genome_wide_shape_data
            if isinstance(sd, GenomeShapeSource):
                sd = sd.on(c)  # the chromosomes + tables resident on THIS device
            return _feature_columns(c, locs[first:last], seqs[first:last], true_seqs, sd, minhash, pw,
                                    similarity_features, quiet=True)

        print(strftime(f"%x %X | Computing feature columns on {len(devices)} GPUs (records sharded by index)"))
        parts = sharding.map_record_shards(shard, len(seqs), devices)
        cols = {k: np.concatenate([p[0][k] for p in parts]) for k in parts[0][0]}
        width = {sh: max(p[1][sh].shape[1] for p in parts) for sh in parts[0][1]}
        shape_rows = {}
        for sh, w in width.items():  # ragged widths across shards: pad with NaN (= key absent)
            shape_rows[sh] = np.concatenate([np.pad(p[1][sh], ((0, 0), (0, w - p[1][sh].shape[1])), constant_values=np.nan)
                                             for p in parts])
    else:
        cols, shape_rows = _feature_columns(ctx or default_context(), locs, seqs, true_seqs, genome_wide_shape_data,
                                            minhash, pw, similarity_features)

    for name, values in cols.items():
        tf_data[name] = values
    blocks = [tf_data]
    for shape, rows in shape_rows.items():
        # json_normalize in the reference only creates the keys that exist in some row
        keep = np.flatnonzero(~np.isnan(rows).all(axis=0)) if len(locs) else np.zeros(0, dtype=np.int64)
        blocks.append(pd.DataFrame(rows[:, keep], columns=[f"{shape}_{w + 1:02d}" for w in keep], index=tf_data.index))
    if len(blocks) > 1:
        tf_data = pd.concat(blocks, axis=1)   # one concatenation instead of a join per shape
    print(strftime("%x %X | Finished building feature table"))
    return tf_data


def background_shape_table(bkg_fasta: str, tables=None, kinds=None, shapes=dnashape.SHAPES, ctx=None,
                           synthetic=False):
    """GPU equivalent of the driver glue at build_features.py:215-241: DNA-shape columns of the
    2+2-padded background records (MGW/ProT/EP trimmed [2:-2], Roll/HelT trimmed [1:-1], NA ->
    0.0), returned as the `shapes_data_df` the driver merges on 'location' (:243-251)."""
    tables, kinds = dnashape.resolve_tables(tables, kinds, shapes, synthetic)
    ctx = ctx or default_context()
    names, recs = [], []
    for name, seq in parse_fasta(bkg_fasta):
        names.append(name)
        recs.append(seq)
    buf, starts, lens = join_records(recs)
    tab = ctx.upload_shapes(tables, kinds)
    enc = ctx.encode(buf)
    width = int(lens.max()) - 3 if len(recs) else 1
    out = ctx.shape_sites(enc, tab, starts, lens, width=max(width, 1), mode=1)
    out = dnashape.to_reference_float64(out)
    cols = {}
    for t, shape in enumerate(shapes):
        for w in range(out.shape[2]):
            col = out[:, t, w]
            if not np.isnan(col).all():
                cols[f"{shape}_{w + 1:02d}"] = col
    enc.close()
    tab.close()
    df = pd.DataFrame(cols)
    df.insert(0, "location", names)
    return df
