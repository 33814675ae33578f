"""Genome-wide candidate sites from the command line — the GPU stand-in for the reference's
BLASTn + bedtools candidate stage (process_blast.sh:21-56, utils.py:340-398, 454-550):

    python -m discover_tfbs_b200.scan_genome GENOME.fa MOTIFS OUT_PREFIX [--fpr 1e-4]

MOTIFS is a JASPAR (`>id name` + `A [ ... ]` rows) or MEME minimal (`MOTIF` + letter-probability
matrix) file of count / probability matrices, or — with --exact — a text file of motif strings (one
per line, IUPAC codes allowed), which are matched exactly on both strands as the reference's BLASTn
100 %-identity filter does.  Writes OUT_PREFIX.bed (BED6, 0-based, score = log-odds) and
OUT_PREFIX.fasta (`chrom:start-end(strand)` headers, strand-corrected sequences: what
`bedtools getfasta -s` emits and build_feature_table consumes).
"""
import argparse

import numpy as np

from . import candidates, pssm
from .utils import parse_fasta


def load_motifs(path: str, exact: bool, pseudocounts: float, background=None):
    """-> (names, matrices).  Exact-match motif strings become 0 / -inf matrices."""
    if exact:
        words = [ln.split()[0] for ln in open(path) if ln.strip() and not ln.startswith("#")]
        return words, [candidates.exact_match_pwm(w) for w in words]
    text = open(path).read()
    reader = pssm.read_meme if "letter-probability matrix" in text else pssm.read_jaspar
    motifs = reader(text, pseudocounts=pseudocounts, background=background)
    if not motifs:
        raise SystemExit(f"{path}: no JASPAR or MEME motif found")
    return [n for n, _ in motifs], [m for _, m in motifs]


def thresholds_for(matrices, exact: bool, fpr: float, threshold, background=None) -> np.ndarray:
    """One score threshold per motif: 0 for exact matching, a fixed --threshold, or the exact
    false-positive-rate threshold of every matrix (pssm.threshold_fpr)."""
    if exact:
        return np.zeros(len(matrices), dtype=np.float32)
    if threshold is not None:
        return np.full(len(matrices), threshold, dtype=np.float32)
    return np.array([pssm.threshold_fpr(m, fpr, background) for m in matrices], dtype=np.float32)


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("genome_fasta")
    ap.add_argument("motifs")
    ap.add_argument("out_prefix")
    ap.add_argument("--exact", action="store_true", help="MOTIFS holds motif strings to match exactly (BLASTn stage)")
    ap.add_argument("--fpr", type=float, default=1e-4, help="false-positive rate per window and strand (default 1e-4)")
    ap.add_argument("--threshold", type=float, default=None, help="one log-odds threshold for every motif instead of --fpr")
    ap.add_argument("--pseudocounts", type=float, default=0.5)
    ap.add_argument("--gc", type=float, default=None, help="background GC content (default: uniform bases)")
    ap.add_argument("--max-hits", type=int, default=1 << 24, help="initial capacity of the hit buffer")
    args = ap.parse_args(argv)

    background = None
    if args.gc is not None:
        background = {"A": (1 - args.gc) / 2, "C": args.gc / 2, "G": args.gc / 2, "T": (1 - args.gc) / 2}
    names, matrices = load_motifs(args.motifs, args.exact, args.pseudocounts, background)
    thr = thresholds_for(matrices, args.exact, args.fpr, args.threshold, background)
    records = {title.split()[0]: seq for title, seq in parse_fasta(args.genome_fasta)}
    if not records:
        raise SystemExit(f"{args.genome_fasta}: no FASTA record")
    cands = candidates.scan_candidates(records, matrices, thr, names=names, cap=args.max_hits)
    candidates.write_bed6(cands, args.out_prefix + ".bed")
    candidates.write_fasta(cands, args.out_prefix + ".fasta")
    print(f"{len(cands)} candidate sites of {len(matrices)} motifs in {len(records)} records "
          f"({sum(len(s) for s in records.values())} bp) -> {args.out_prefix}.bed / .fasta")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
