#!/usr/bin/env python3
"""Genome -> candidate sites -> feature table, entirely through discover_tfbs_b200.

Mirrors what the reference needs external tools and three scripts for:
  * BLASTn of curated motifs + awk + `bedtools getfasta -s`   (.COMMANDS.txt:226-237, process_blast.sh)
    -> candidates.scan_candidates / write_bed6 / write_fasta
  * DNAshapeR on the whole genome + 71.5 M-string parse       (get_DNA_shapes.R, predict.py:223-228)
    -> utils.GenomeShapeSource (shapes computed on the GPU from sequence)
  * utils.build_feature_table(prediction_FASTA, fg_fasta, shape_data)        (predict.py:233)
    -> same call, same columns.

    python examples/predict_features.py genome.fasta fg_sites.fasta MOTIF [MOTIF ...] --out features.feather
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from discover_tfbs_b200 import candidates, dnashape  # noqa: E402
from discover_tfbs_b200 import utils as U  # noqa: E402


def run(genome_fasta, fg_fasta, motifs, out_prefix, tables=None, kinds=None, ctx=None, synthetic=False):
    records = dict(U.parse_fasta(genome_fasta))
    cands = candidates.scan_candidates(records, [candidates.exact_match_pwm(m) for m in motifs], 0.0,
                                       names=list(motifs), ctx=ctx)
    candidates.write_bed6(cands, out_prefix + ".bed")
    candidates.write_fasta(cands, out_prefix + ".fasta")
    # real pentamer tables (argument or $TFBS_SHAPE_TABLES); seeded stand-ins only on explicit request
    tables, kinds = dnashape.resolve_tables(tables, kinds, synthetic=synthetic)
    shapes = U.GenomeShapeSource(records=records, tables=tables, kinds=kinds, ctx=ctx)
    table = U.build_feature_table(out_prefix + ".fasta", fg_fasta, shapes, minhash=True, ctx=ctx)
    return cands, table


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("genome_fasta")
    ap.add_argument("fg_fasta")
    ap.add_argument("motifs", nargs="+")
    ap.add_argument("--out", default="prediction_features.feather")
    ap.add_argument("--shape-tables", default=None, help="pentamer table file (dnashape.save_tables format); "
                                                         "default: $TFBS_SHAPE_TABLES")
    ap.add_argument("--synthetic-tables", action="store_true", help="demo only: seeded stand-in tables, NOT DNAshapeR's")
    args = ap.parse_args()
    prefix = os.path.splitext(args.out)[0]
    tables = kinds = None
    if args.shape_tables:
        tables, kinds = dnashape.load_tables(args.shape_tables)
    cands, table = run(args.genome_fasta, args.fg_fasta, args.motifs, prefix, tables, kinds,
                       synthetic=args.synthetic_tables)
    table.reset_index().to_feather(args.out)
    print(f"{len(cands)} candidate sites -> {args.out} ({table.shape[1]} columns)")


if __name__ == "__main__":
    main()
