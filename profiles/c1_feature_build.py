#!/usr/bin/env python3
"""Config C1 of BASELINE.json at full size through the drop-in API:
one 12-bp TF PWM, 5 000 foreground 200-bp peaks + 50 000 shuffled 2+2-padded backgrounds,
GC + MinHash/PAS/PPS + PWM + DNA-shape features, assembled exactly like build_features.py:193-258.
Prints per-stage wall-clock and, beside it, the host-CPU cost of the same stages measured on a
sample with the scalar restatement of the reference's loops (oracle/similarity_oracle.py)."""
import os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import io, contextlib
import numpy as np, pandas as pd
from discover_tfbs_b200 import _lib, dnashape, synth
from oracle import similarity_oracle as sm
from discover_tfbs_b200 import utils as U

def main():
    rng = np.random.default_rng(39)
    ctx = _lib.Context(0)
    genome_len = 14_300_000
    genome = synth.genome(39, 0, genome_len, gc=0.335, n_runs=False, lowercase=False)
    pwm, motif = synth.consensus_pwm(39, 12)
    starts = np.sort(rng.choice(genome_len - 400, size=5000, replace=False))
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    fg = []
    for s in starts:
        strand = "+" if rng.random() < 0.5 else "-"
        site = genome[s:s + 200].tobytes()
        if strand == "-":
            site = site.translate(comp)[::-1]
        fg.append((f"chrS:{s}-{s + 200}({strand})", site.decode()))
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    bkg = []
    for loc, site in fg:
        arr = np.frombuffer(site.encode(), dtype=np.uint8)
        for k in range(10):
            pad = acgt[rng.integers(0, 4, 4)]
            core = rng.permutation(arr)
            bkg.append((f"dinuc_shuffled_bkg_seq{k}_for_{loc}", bytes(pad[:2]) .decode() + core.tobytes().decode() + bytes(pad[2:]).decode()))
    tmp = tempfile.mkdtemp()
    fg_path, bkg_path = os.path.join(tmp, "fg.fasta"), os.path.join(tmp, "bkg.fasta")
    for path, recs in ((fg_path, fg), (bkg_path, bkg)):
        with open(path, "w") as fh:
            fh.write("".join(f">{n}\n{s}\n" for n, s in recs))
    tables, kinds = dnashape.synthetic_tables()
    t = {}
    quiet = io.StringIO()
    t0 = time.perf_counter()
    src = U.GenomeShapeSource(records={"chrS": genome.tobytes().decode()}, tables=tables, kinds=kinds, ctx=ctx)
    t["genome encode + tables upload"] = time.perf_counter() - t0
    with contextlib.redirect_stdout(quiet):
        t0 = time.perf_counter()
        pos = U.build_feature_table(fg_path, fg_path, src, minhash=True, pwms=[pwm], pwm_names=["TF"], ctx=ctx)
        pos.insert(1, "seq_type", "True")
        t["fg feature table (5 000 x 5 000 pairs, 5 shapes x 201)"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        neg = U.build_feature_table(bkg_path, fg_path, minhash=True, pwms=[pwm], pwm_names=["TF"], ctx=ctx)
        t["bkg feature table (50 000 x 5 000 pairs)"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        shapes_df = U.background_shape_table(bkg_path, tables, kinds, ctx=ctx)
        neg = neg.merge(shapes_df, how="left", left_on="location", right_on="location")
        neg.insert(1, "seq_type", "Not_True")
        t["bkg shape table + merge"] = time.perf_counter() - t0
    t0 = time.perf_counter()
    training = pd.concat([pos, neg], sort=False).dropna(axis=1).reset_index()
    t["concat + dropna(axis=1)"] = time.perf_counter() - t0
    total = sum(t.values())
    print(f"training table: {training.shape[0]} rows x {training.shape[1]} columns")
    for k, v in t.items():
        print(f"  {k:58s} {v:8.2f} s")
    print(f"  {'TOTAL (GPU path, wall clock incl. FASTA parse + pandas)':58s} {total:8.2f} s")
    # host-CPU cost of the similarity stage, scalar restatement of the reference's loops, on a sample
    true = {s for _, s in fg}
    sample = [s for _, s in bkg[:2]]
    sub = list(true)[:200]
    t0 = time.perf_counter()
    for sq in sample:
        sm.minhash_similarity(sq, sub); sm.pac_similarity(sq, sub)
    per_pair = (time.perf_counter() - t0) / (len(sample) * len(sub))
    pairs = 55_000 * len(true)
    print(f"  host scalar similarity: {per_pair * 1e6:.0f} us/pair -> {pairs:.3g} pairs = {per_pair * pairs / 3600:.1f} h on one core "
          f"(the reference itself: 160 us/pair for 12-mers, BASELINE.md)")
    cols = list(training.columns)
    print("  columns:", cols[:8], "...", cols[-3:])

if __name__ == "__main__":
    main()
