"""Genome-wide candidate sites straight from the GPU scan.

The reference finds candidate sites with exact-match BLASTn of curated motif strings and a chain of
external tools: blastn -> keep 100 %-identity full-length matches (utils.py:340-398) -> BED6 with
0-based start (utils.py:454-485, process_blast.sh:32) -> `bedtools getfasta -s`
(process_blast.sh:56, utils.py:516-550) whose FASTA headers `chrom:start-end(strand)` feed
`build_feature_table` (predict.py:233).  Here the hits-mode scan yields (chrom, start, end, strand)
directly: an exact-match motif is a PWM with 0 / -inf entries scanned at threshold 0, and a real
PWM with a score threshold generalises it.
"""
import numpy as np

from .pssm import default_context
from .synth import join_records

_IUPAC = {"A": "A", "C": "C", "G": "G", "T": "T", "R": "AG", "Y": "CT", "S": "GC", "W": "AT", "K": "GT",
          "M": "AC", "B": "CGT", "D": "AGT", "H": "ACT", "V": "ACG", "N": "ACGT"}
_COMP = bytes.maketrans(b"ACGTacgtNn", b"TGCAtgcaNn")


def exact_match_pwm(motif: str) -> np.ndarray:
    """Motif string (IUPAC codes as in utils.py:109-134) -> log-odds-shaped matrix with 0 for the
    allowed bases and -inf otherwise; scanned with threshold 0 it reproduces BLASTn's 100 %-identity
    full-length matches on both strands."""
    m = np.full((len(motif), 4), -np.inf)
    for j, ch in enumerate(motif.upper()):
        for b in _IUPAC[ch]:
            m[j, "ACGT".index(b)] = 0.0
    return m


def scan_candidates(records: dict, matrices, thresholds, names=None, ctx=None, cap=1 << 22):
    """records: {chrom: sequence}.  Returns a list of dicts (chrom, start, end, strand, motif, name,
    score, sequence) ordered by (chrom order, start, strand); `sequence` is strand-corrected like
    `bedtools getfasta -s`."""
    ctx = ctx or default_context()
    chroms = list(records)
    buf, starts, lens = join_records([records[c] for c in chroms])
    enc = ctx.encode(buf)
    lib = ctx.upload_pwms(list(matrices))
    hits = ctx.scan_hits(enc, lib, thresholds, cap=cap, sort=False)
    mlen = np.asarray(lib.lens)
    enc.close()
    lib.close()
    rev = hits["pos"] < 0
    gpos = np.where(rev, ~hits["pos"], hits["pos"])
    ci = np.searchsorted(starts, gpos, side="right") - 1
    order = np.lexsort((rev, gpos, ci))
    out = []
    for i in order:
        c = chroms[ci[i]]
        s = int(gpos[i] - starts[ci[i]])
        L = int(mlen[hits["motif"][i]])
        site = bytes(buf[gpos[i]: gpos[i] + L])
        if rev[i]:
            site = site.translate(_COMP)[::-1]
        out.append({"chrom": c, "start": s, "end": s + L, "strand": "-" if rev[i] else "+",
                    "motif": int(hits["motif"][i]), "name": (names[hits["motif"][i]] if names else f"motif{hits['motif'][i]}"),
                    "score": float(hits["score"][i]), "sequence": site.decode("ascii")})
    return out


def write_bed6(cands, path):
    """BED6, 0-based half-open, as utils.py:471-480 / process_blast.sh:32 produce."""
    with open(path, "w") as fh:
        for c in cands:
            fh.write(f"{c['chrom']}\t{c['start']}\t{c['end']}\t{c['name']}\t{c['score']:g}\t{c['strand']}\n")


def write_fasta(cands, path):
    """FASTA with `chrom:start-end(strand)` headers — the format `bedtools getfasta -s` emits and
    `build_feature_table` / `_get_DNA_shape` parse (utils.py:138-149, 1065-1068)."""
    with open(path, "w") as fh:
        for c in cands:
            fh.write(f">{c['chrom']}:{c['start']}-{c['end']}({c['strand']})\n{c['sequence']}\n")
