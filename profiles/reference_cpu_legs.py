#!/usr/bin/env python3
"""CPU baseline of the shape-window and feature-table legs with the REAL reference (BASELINE.md §3
item 1, VERDICT r01 #2d): the unmodified /root/reference/utils.py, imported through oracle/refshim.py
(stubs for the absent Biopython / mmh3 / pybedtools / matplotlib), single-threaded exactly as the
reference runs it, on config C1 inputs (200-bp peaks, 12-bp sites).  The reference cannot travel to
the GPU box, so this runs in the BUILD CONTAINER and writes profiles/r02/reference_cpu_legs.json,
which bench.py quotes (with this provenance) beside the GPU numbers.

    python profiles/reference_cpu_legs.py
"""
import contextlib
import io
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from discover_tfbs_b200 import dnashape, synth  # noqa: E402
from oracle import refshim, shape_oracle  # noqa: E402


def main():
    ref = refshim.load_reference_utils()
    rng = np.random.default_rng(39)
    genome_len = 2_000_000
    genome = synth.genome(39, 0, genome_len, gc=0.335, n_runs=False, lowercase=False)
    tables, kinds = dnashape.synthetic_tables()
    tracks = shape_oracle.tracks(genome, tables, kinds)
    # the {shape: {chrom: [tokens]}} dict the reference's drivers build (build_features.py:193-198)
    t0 = time.perf_counter()
    shapedata = {sh: {"chrS": shape_oracle.tokens(tracks[t], int(kinds[t]), genome_len, trailing_empty=True)}
                 for t, sh in enumerate(dnashape.SHAPES)}
    t_tokens = time.perf_counter() - t0
    out = {"where": "build container (no GPU)", "host": f"{os.cpu_count()} vCPU, 1 thread used (the reference is "
           "single-threaded)", "python": sys.version.split()[0], "reference": "unmodified /root/reference/utils.py via "
           "oracle/refshim.py", "script": "profiles/reference_cpu_legs.py"}

    # ---- _get_DNA_shape (utils.py:1061-1084): 5 000 C1 peaks (200 bp -> 201 values) x 5 shapes -------------
    starts = np.sort(rng.choice(genome_len - 400, size=5000, replace=False))
    locs = [f"chrS:{s}-{s + 200}({'+' if i % 2 else '-'})" for i, s in enumerate(starts)]
    t0 = time.perf_counter()
    n_values = 0
    for sh in dnashape.SHAPES:
        for loc in locs:
            n_values += len(ref._get_DNA_shape(loc, shapedata[sh], sh))
    dt = time.perf_counter() - t0
    out["get_DNA_shape_c1"] = {"sites": len(locs), "shapes": 5, "values": n_values, "seconds": dt,
                               "us_per_site_per_shape": dt / (5 * len(locs)) * 1e6, "ns_per_value": dt / n_values * 1e9}
    # 12-bp sites (predict.py at genome scale: 13 values per site and shape)
    locs12 = [f"chrS:{s}-{s + 12}({'+' if i % 2 else '-'})" for i, s in enumerate(starts)]
    t0 = time.perf_counter()
    n_values = sum(len(ref._get_DNA_shape(loc, shapedata["MGW"], "MGW")) for loc in locs12)
    dt = time.perf_counter() - t0
    out["get_DNA_shape_12bp"] = {"sites": len(locs12), "values": n_values, "seconds": dt,
                                 "us_per_site": dt / len(locs12) * 1e6, "ns_per_value": dt / n_values * 1e9}

    # ---- build_feature_table (utils.py:1088-1173) on a bounded C1 sample ---------------------------------
    # all-pairs MinHash / PAC dominate: 60 candidates x 300 true 200-bp peaks, then
    # extrapolated linearly in the number of (candidate, true) pairs to the full C1 (55 000 x 5 000)
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    recs = []
    for i, s in enumerate(starts[:300]):
        site = genome[s:s + 200].tobytes()
        recs.append((locs[i], (site.translate(comp)[::-1] if locs[i].endswith("(-)") else site).decode()))
    tmp = tempfile.mkdtemp()
    true_path, cand_path = os.path.join(tmp, "true.fasta"), os.path.join(tmp, "cand.fasta")
    with open(true_path, "w") as fh:
        fh.write("".join(f">{n}\n{s}\n" for n, s in recs))
    with open(cand_path, "w") as fh:
        fh.write("".join(f">{n}\n{s}\n" for n, s in recs[:60]))
    with contextlib.redirect_stdout(io.StringIO()):
        t0 = time.perf_counter()
        table = ref.build_feature_table(cand_path, true_path, shapedata, minhash=True)
        dt = time.perf_counter() - t0
        t0 = time.perf_counter()
        ref.build_feature_table(cand_path, true_path, shapedata, minhash=False)
        dt_nomh = time.perf_counter() - t0
    pairs = 60 * 300
    full_pairs = 55_000 * 5_000
    out["build_feature_table_c1_sample"] = {
        "candidates": 60, "true": 300, "pairs": pairs, "columns": int(table.shape[1]), "seconds": dt,
        "us_per_pair": dt / pairs * 1e6, "seconds_without_minhash": dt_nomh,
        "extrapolated_full_c1_hours": dt / pairs * full_pairs / 3600,
        "extrapolation": "linear in (candidate, true) pairs: 55 000 x 5 000 for config C1 (fg + bkg against the fg set)"}
    out["shape_file_token_build_seconds"] = {"seconds": t_tokens, "values": int(5 * genome_len),
                                             "note": "oracle tracks -> DNAshapeR tokens for the reference's dict input"}
    dst = os.path.join(ROOT, "profiles", "r02", "reference_cpu_legs.json")
    with open(dst, "w") as fh:
        json.dump(out, fh, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
