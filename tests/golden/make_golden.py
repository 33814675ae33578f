#!/usr/bin/env python3
"""Generate tests/golden/reference_golden.json by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

It imports `/root/reference/utils.py` through oracle/refshim.py (stub modules for the absent
Biopython / mmh3 / pybedtools / matplotlib), feeds it seeded synthetic inputs and records what
the reference's own functions return.  The driver-glue block build_features.py:214-251
(background shape ingest + merge) is executed from the reference's source text, not copied.
The inputs are stored next to the outputs so the tests need neither the reference nor this
script at run time.
"""
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from discover_tfbs_b200 import dnashape  # noqa: E402  (seeded synthetic pentamer tables)
from oracle import refshim, shape_oracle  # noqa: E402

SHAPES = ("EP", "HelT", "MGW", "ProT", "Roll")


def rand_seq(rng, n, alphabet="ACGT", p=None):
    return "".join(rng.choice(list(alphabet), size=n, p=p))


def write_fasta(path, records):
    with open(path, "w") as fh:
        for name, seq in records:
            fh.write(f">{name}\n{seq}\n")


def main():
    ref = refshim.load_reference_utils()
    assert ref.__version__ == "0.3.7"
    rng = np.random.default_rng(39)
    tables, kinds = dnashape.synthetic_tables(SHAPES, seed=39)
    gold = {"meta": {"reference_version": ref.__version__, "seed": 39, "shapes": list(SHAPES),
                     "table_seed": 39}}

    # ---- primitives (utils.py:97-105, 224-238, 402-425) -------------------------------------
    prim = {"reverse_complement": [], "calculate_gc_content": [], "get_kmers": []}
    for n in (1, 5, 12, 31, 200):
        s = rand_seq(rng, n, "ACGTRYSWKMBDHVN", [.22, .22, .22, .22] + [.12 / 11] * 11)
        prim["reverse_complement"].append([s, ref.reverse_complement(s)])
    for n in (1, 6, 12, 33, 200, 1000):
        s = rand_seq(rng, n, "ACGTacgtNn", [.2, .2, .2, .2, .04, .04, .04, .04, .02, .02])
        prim["calculate_gc_content"].append([s, ref.calculate_gc_content(s)])
    for n, k in ((12, 6), (12, 4), (20, 5)):
        s = rand_seq(rng, n)
        prim["get_kmers"].append([s, k, ref.get_kmers(s, k)])
    gold["primitives"] = prim

    # ---- similarity features ("next" rows; utils.py:892-949, 1243-1269) ----------------------
    sim = {"minhash_kmer": [], "jaccard": [], "pac": []}
    for _ in range(8):
        kmer = rand_seq(rng, 6)
        sim["minhash_kmer"].append([kmer, ref._minhash_kmer(kmer)])
    base = rand_seq(rng, 12)
    for _ in range(8):
        a = rand_seq(rng, 12)
        b = "".join(c if rng.random() > 0.2 else rng.choice(list("ACGT")) for c in base)
        sim["jaccard"].append([a, b, ref.jaccard_similarity(a, b)])
        sim["jaccard"].append([base, b, ref.jaccard_similarity(base, b)])
        sim["pac"].append([base, b, list(ref.pac(base, b))])
        sim["pac"].append([a, b, list(ref.pac(a, b))])
    gold["similarity"] = sim

    # ---- a two-chromosome genome and its DNAshapeR-format token lists --------------------------
    genome = {"chrA": rand_seq(rng, 400), "chrB": rand_seq(rng, 257)}
    g = list(genome["chrA"])
    g[100:104] = "NNNN"          # an N-run: NA pentamers
    g[200] = "r"                 # IUPAC ambiguity code
    g[300:310] = [c.lower() for c in g[300:310]]  # soft-masked bases stay valid
    genome["chrA"] = "".join(g)
    shape_tokens = {}
    for t, shape in enumerate(SHAPES):
        shape_tokens[shape] = {}
        for chrom, seq in genome.items():
            tr = shape_oracle.tracks(seq, tables[t:t + 1], kinds[t:t + 1])[0]
            shape_tokens[shape][chrom] = shape_oracle.tokens(tr, int(kinds[t]), len(seq), trailing_empty=True)
    gold["genome"] = genome
    gold["shape_tokens"] = shape_tokens

    # ---- _get_DNA_shape (utils.py:1061-1084) incl. the quirks of SURVEY §8(a4) -----------------
    locs = []
    for chrom, seq in genome.items():
        n = len(seq)
        for _ in range(12):
            L = int(rng.integers(6, 21))
            s = int(rng.integers(0, n - L))
            locs.append(f"{chrom}:{s}-{s + L}({'+' if rng.random() < 0.5 else '-'})")
        locs += [f"{chrom}:0-12(+)", f"{chrom}:0-12(-)", f"{chrom}:1-13(-)",
                 f"{chrom}:{n - 12}-{n}(+)", f"{chrom}:{n - 12}-{n}(-)",
                 f"{chrom}:{n - 6}-{n + 6}(+)", f"{chrom}:{n - 6}-{n + 6}(-)",
                 f"{chrom}:{n + 5}-{n + 17}(+)", f"{chrom}:{n + 5}-{n + 17}(-)"]
    locs += ["chrA:95-107(+)", "chrA:95-107(-)", "chrA:195-207(+)", "chrA:298-312(-)"]
    gds = []
    for loc in locs:
        chrom = loc.split(":")[0]
        for shape in SHAPES:
            out = ref._get_DNA_shape(loc, shape_tokens[shape], shape)
            mine = shape_oracle.get_dna_shape(loc, shape_tokens[shape][chrom], shape)
            assert out == mine, (loc, shape)
            gds.append({"loc": loc, "shape": shape, "out": out})
    gold["get_dna_shape"] = gds

    # ---- build_feature_table (utils.py:1088-1173) ---------------------------------------------
    with tempfile.TemporaryDirectory() as tmp:
        fg = []
        for i in range(10):
            chrom = "chrA" if i % 2 == 0 else "chrB"
            s = int(rng.integers(5, len(genome[chrom]) - 20))
            strand = "+" if i % 3 else "-"
            site = genome[chrom][s:s + 12].upper().replace("N", "A").replace("R", "A")
            if strand == "-":
                site = ref.reverse_complement(site)
            fg.append((f"{chrom}:{s}-{s + 12}({strand})", site))
        cand = list(fg[:4])
        for i in range(8):
            chrom = "chrB" if i % 2 == 0 else "chrA"
            s = int(rng.integers(5, len(genome[chrom]) - 20))
            strand = "+" if i % 2 else "-"
            site = genome[chrom][s:s + 12].upper().replace("N", "A").replace("R", "A")
            if strand == "-":
                site = ref.reverse_complement(site)
            cand.append((f"{chrom}:{s}-{s + 12}({strand})", site))
        fg_path, cand_path = os.path.join(tmp, "fg.fasta"), os.path.join(tmp, "cand.fasta")
        write_fasta(fg_path, fg)
        write_fasta(cand_path, cand)
        df_fg = ref.build_feature_table(fg_path, fg_path, shape_tokens, minhash=True)
        df_cand = ref.build_feature_table(cand_path, fg_path, shape_tokens, minhash=True)
        df_noshape = ref.build_feature_table(cand_path, fg_path, minhash=False)

        def pack(df):
            return {"columns": list(df.columns), "dtypes": [str(t) for t in df.dtypes],
                    "rows": df.astype(object).where(df.notna(), None).values.tolist()}  # repr floats

        gold["feature_table"] = {"fg": fg, "cand": cand, "fg_table": pack(df_fg),
                                 "cand_table": pack(df_cand), "cand_noshape_table": pack(df_noshape)}

        # ---- background shape ingest: build_features.py:214-251, run from the reference source ---
        bkg = []
        for i, (loc, site) in enumerate(fg):
            core = "".join(rng.permutation(list(site)))
            pad = rand_seq(rng, 4)
            bkg.append((f"dinuc_shuffled_bkg_seq_for_{loc}", pad[:2] + core + pad[2:]))
        bkg_path = os.path.join(tmp, "bkg.fasta")
        write_fasta(bkg_path, bkg)
        shape_files, shape_file_text = [], {}
        for t, shape in enumerate(SHAPES):
            path = f"{bkg_path}.{shape}"
            lines = []
            for name, seq in bkg:
                tr = shape_oracle.tracks(seq, tables[t:t + 1], kinds[t:t + 1])[0]
                toks = shape_oracle.tokens(tr, int(kinds[t]), len(seq))
                lines.append(f">{name}")
                lines.append(",".join(toks))  # 16-bp records: under DNAshapeR's 20-per-line wrap
            text = "\n".join(lines) + "\n"
            with open(path, "w") as fh:
                fh.write(text)
            shape_files.append(path)
            shape_file_text[shape] = text
        src = open(os.path.join(refshim.REFERENCE_DIR, "build_features.py")).read().split("\n")
        block = src[213:251]  # lines 214..251
        assert block[0].strip() == "# collate all DNA shape values", block[0]
        assert block[-1].strip() == ")", block[-1]
        import textwrap
        from os.path import abspath
        from time import strftime

        import pandas as pd

        class Args:
            bkg_shape_fasta_file = shape_files

        negative_data_df = ref.build_feature_table(bkg_path, fg_path, minhash=True)
        ns = {"args": Args, "parse_fasta": ref.parse_fasta, "abspath": abspath, "pd": pd,
              "strftime": strftime, "negative_data_df": negative_data_df, "print": lambda *a, **k: None}
        exec(textwrap.dedent("\n".join(block)), ns)
        gold["background"] = {"records": bkg, "shape_file_text": shape_file_text,
                              "shapes_data_df": pack(ns["shapes_data_df"]),
                              "negative_data_df": pack(ns["negative_data_df"])}

    # ---- legacy BED-driven API get_shape_data (utils.py:554-602; no caller in the reference) ------
    with tempfile.TemporaryDirectory() as tmp:
        shapefiles = []
        for shape in SHAPES:
            path = os.path.join(tmp, f"genome_oneline.fasta.{shape}")
            with open(path, "w") as fh:
                for chrom in genome:
                    fh.write(f">{chrom}\n" + ",".join(shape_tokens[shape][chrom]) + "\n")  # tokens end with ''
            shapefiles.append(path)
        bed_rows = []
        for chrom, seq in genome.items():
            n = len(seq)
            for _ in range(6):
                L = int(rng.integers(6, 21))
                s0 = int(rng.integers(1, n - L - 1))
                bed_rows.append((chrom, s0, s0 + L, "site", "1.0", "+" if rng.random() < 0.5 else "-"))
            bed_rows.append((chrom, n - 10, n, "edge", "1.0", "+"))
            bed_rows.append((chrom, 1, 13, "edge", "1.0", "-"))
        bed_path = os.path.join(tmp, "sites.bed")
        with open(bed_path, "w") as fh:
            for r in bed_rows:
                fh.write("\t".join(str(x) for x in r) + "\n")
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()):
            gsd = ref.get_shape_data(bed_path, shapefiles)
        gold["get_shape_data"] = {"bed": [list(r) for r in bed_rows], "out": gsd}

    out = os.path.join(HERE, "reference_golden.json")
    with open(out, "w") as fh:
        json.dump(gold, fh, indent=0, separators=(",", ":"))
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
