"""Candidate inflation of the three block forms of the hits filter, measured on the CPU with the
kernel's own tables and field arithmetic (tfbs_debug_hits_filter; no GPU needed).

Bench library (800 PWMs, L 6..30), thresholds at a 1e-4 hit rate, 60 kb of the synthetic genome.
For every motif length: true hits, and candidates the wide (16-bit fields), narrow (8-bit fields) and
narrow+sticky (8-bit fields with sticky flags from 5 groups on) filters pass.

    python profiles/filter_forms_cpu.py [rate] > profiles/r01/filter_forms_cpu.txt
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from discover_tfbs_b200 import _lib, synth  # noqa: E402

rate = float(sys.argv[1]) if len(sys.argv) > 1 else 1e-4
n = 60_000
logodds, lens = synth.pwm_library(39, 800, 6, 30)
thr = synth.thresholds_for_rate(logodds, lens, rate=rate, gc=0.41, seed=39, sample=100_000)
seq = synth.genome(39, 0, n, gc=0.41)
codes = np.zeros(256, np.uint8)
codes[list(b"ACGT")] = [0, 1, 2, 3]
codes = codes[seq]

# exact scores (float64 sum -> float32, as the oracle compares them) for the hit counts
hits = np.zeros((n, 800, 2), dtype=bool)
for m in range(800):
    L = int(lens[m])
    nw = n - L + 1
    f = np.zeros(nw)
    r = np.zeros(nw)
    for j in range(L):
        c = codes[j:j + nw]
        f += logodds[m, j, c]
        r += logodds[m, L - 1 - j, 3 - c]
    hits[:nw, m, 0] = f.astype(np.float32) >= thr[m]
    hits[:nw, m, 1] = r.astype(np.float32) >= thr[m]

inside = (np.arange(n)[:, None] + lens[None, :] <= n)[:, :, None]
names = {0: "wide", 1: "narrow", 2: "narrow+sticky"}
cand = {}
for form in names:
    c = _lib.hits_filter_host(logodds, lens, thr, form, codes).astype(bool) & inside
    assert not (hits & ~c).any(), f"form {form} lost a hit"
    cand[form] = c
plans = {form: _lib.hits_block_plan(lens, 0 if form == 0 else 9) for form in names}
print(f"rate {rate:g}, {n} bp x 800 PWMs x 2 strands; hits {int(hits.sum())}")
print("blocks (wide plan):  ", [a + b for a, b, _, _ in plans[0]])
print("blocks (narrow plan):", [a + b for a, b, _, _ in plans[1]], "(sticky where >= 5)")
print(f"{'L':>3} {'groups':>6} {'hits':>8} " + " ".join(f"{names[f]:>14}" for f in names) + "   (candidates / hits)")
for L in sorted(set(int(x) for x in lens)):
    sel = lens == L
    h = int(hits[:, sel].sum())
    g = sum(_lib.hits_group_plan(L))
    row = " ".join(f"{int(cand[f][:, sel].sum()) / max(h, 1):14.2f}" for f in names)
    print(f"{L:3d} {g:6d} {h:8d} {row}")
tot = int(hits.sum())
print("all".rjust(3), " " * 6, f"{tot:8d} " + " ".join(f"{int(cand[f].sum()) / tot:14.2f}" for f in names))
