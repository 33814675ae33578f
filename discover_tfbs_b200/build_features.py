#!/usr/bin/env python3
"""Drop-in for the reference's `build_features.py` driver (same positional command line).

    python -m discover_tfbs_b200.build_features PROTEIN fg.fasta fg.bed bkg.fasta \\
        bkg.fasta.{EP,HelT,MGW,ProT,Roll} genome_oneline.fasta.{EP,HelT,MGW,ProT,Roll} \\
        training.feather pca_plot.pdf  [--genome-fasta genome.fasta [--shape-tables tables.json]]

Reproduces build_features.py:166-267: foreground table with genome-wide shapes, background table
merged with the trimmed background shapes, concat, dropna(axis=1), feather.  With
`--genome-fasta` all DNA-shape values are computed on the GPU from sequence and the ten
pre-computed DNAshapeR files are not read (pass any placeholder for them).  The PCA scatter plot
(build_features.py:283-333) is classifier reporting, out of scope: it is drawn only if matplotlib
is importable.
"""
import argparse
import json
from os.path import abspath, isfile
from sys import exit
from time import strftime

import pandas as pd

from . import dnashape
from .utils import GenomeShapeSource, background_shape_table, build_feature_table, parse_fasta


def handle_program_options(argv=None):
    parser = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    parser.add_argument("protein_name", type=str, help="name of the transcription factor")
    parser.add_argument("fg_fasta_file", type=str, help="true binding-site sequences (FASTA)")
    parser.add_argument("fg_bed_file", type=str, help="true binding events (BED6); kept for CLI compatibility")
    parser.add_argument("bkg_fasta_file", type=str, help="background sequences from create_bkg_seqs.py (FASTA)")
    parser.add_argument("bkg_shape_fasta_file", nargs=5, help="DNAshapeR output of the background FASTA (5 files)")
    parser.add_argument("genome_wide_shape_fasta_file", nargs=5, help="genome-wide one-line DNAshapeR files (5 files)")
    parser.add_argument("save_training_data", type=str, help="output feather file")
    parser.add_argument("save_pca_plot", type=str, help="output PCA plot (only drawn if matplotlib is available)")
    parser.add_argument("--genome-fasta", default=None, help="compute every DNA-shape value on the GPU from this genome")
    parser.add_argument("--shape-tables", default=None,
                        help="JSON {shape: {pentamer: value | [y1, y2]}} with the real pentamer tables "
                             "(default: the seeded synthetic tables)")
    return parser.parse_args(argv)


def _bkg_shapes_from_files(shapefiles):
    """build_features.py:215-241 verbatim in behaviour: text parse + trims ([2:-2] / [1:-1])."""
    bkg_shapes = dict()
    for shapefile in shapefiles:
        whichshape = shapefile.split(".")[-1]
        lo, hi = (2, -2) if whichshape in ["MGW", "ProT", "EP"] else (1, -1)
        for name, shape in parse_fasta(abspath(shapefile)):
            shape = (shape.strip() if lo == 2 else shape).split(",")
            row = bkg_shapes.setdefault(name, dict())
            for i, val in enumerate(shape[lo:hi]):
                try:
                    row["{0}_{1:02d}".format(whichshape, i + 1)] = float(val)
                except Exception:
                    row["{0}_{1:02d}".format(whichshape, i + 1)] = 0.0
    return pd.DataFrame.from_dict(bkg_shapes, orient="index").reset_index().rename(columns={"index": "location"})


def main(argv=None):
    print("#" * 90, strftime("%x %X | BUILD FEATURE TABLE\n"), sep="\n\n")
    args = handle_program_options(argv)
    if not (isfile(args.fg_fasta_file) and isfile(args.bkg_fasta_file)):
        print("Error: Please check supplied FASTA file")
        exit()
    tables = kinds = None
    if args.shape_tables:
        with open(args.shape_tables) as fh:
            tables, kinds = dnashape.load_tables(json.load(fh))

    print(strftime("%x %X | Processing foreground FASTA file"))
    if args.genome_fasta:
        shape_data = GenomeShapeSource(args.genome_fasta, tables=tables, kinds=kinds)
    else:
        shape_data = {f.split(".")[-1]: {h: s.strip().split(",") for h, s in parse_fasta(f)}
                      for f in args.genome_wide_shape_fasta_file}
    print("=" * 90, sep="\n")
    positive_data_df = build_feature_table(args.fg_fasta_file, args.fg_fasta_file, shape_data, minhash=True)
    positive_data_df.insert(1, "seq_type", "True")

    print(strftime("\n%x %X | Processing background FASTA file"))
    print("=" * 90, sep="\n")
    negative_data_df = build_feature_table(args.bkg_fasta_file, args.fg_fasta_file, minhash=True)
    print(strftime("%x %X | Processing DNA shape data"))
    if args.genome_fasta:
        shapes_data_df = background_shape_table(args.bkg_fasta_file, tables, kinds)
    else:
        shapes_data_df = _bkg_shapes_from_files(args.bkg_shape_fasta_file)
    print(strftime("%x %X | Creating background training dataset"))
    negative_data_df = negative_data_df.merge(shapes_data_df, how="left", left_on="location", right_on="location")
    negative_data_df.insert(1, "seq_type", "Not_True")

    training_data = pd.concat([positive_data_df, negative_data_df], sort=False)
    training_data = training_data.dropna(axis=1)  # drop columns with any NaN
    training_data = training_data.reset_index()
    print(strftime(f"%x %X | Saving {args.protein_name} background training dataset to {args.save_training_data}"))
    training_data.to_feather(args.save_training_data)
    try:
        import matplotlib  # noqa: F401
    except ImportError:
        print(strftime("%x %X | matplotlib not available: skipping the PCA plot"))
    return training_data


if __name__ == "__main__":
    main()
