"""CPU fuzz campaign for the hits scan's host-side planning (no GPU): random libraries (soft, sharp,
with -inf entries, lengths 1..32), random threshold vectors (rate quantiles, thresholds set exactly on a
window's score, +-inf, NaN) and sequences with homopolymer runs.  For every case the oracle's hits must be

  * EXACTLY the direct path's keys for the motifs of at most 12 columns it took (chosen per motif),
  * a subset of the direct path's 12-column windows for the 13 / 14-column motifs it took,
  * a subset of the table filter's candidates for every motif the plan keeps on the tables (form 4) and
    for every motif when all stay on the tables (form 3).

Usage: python tests/fuzz_hits_host.py [--cases 200] [--seed 1]   (oracle = test infrastructure)."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

from discover_tfbs_b200 import _lib, synth  # noqa: E402
from oracle import pssm_oracle as po  # noqa: E402


def one_case(rng, case):
    M = int(rng.choice([33, 40, 64, 65, 96, 130]))
    kind = case % 4
    if kind == 0:      # short library: everything may go direct
        lmin, lmax = int(rng.integers(4, 9)), int(rng.integers(9, 15))
    elif kind == 1:    # mixed
        lmin, lmax = int(rng.integers(5, 10)), int(rng.integers(15, 33))
    elif kind == 2:    # around the 12 / 13 / 14 boundary
        lmin, lmax = int(rng.integers(10, 13)), int(rng.integers(13, 17))
    else:
        lmin, lmax = 6, 30
    alpha = float(rng.choice([0.05, 0.1, 0.3, 1.0]))
    nsites = int(rng.choice([20, 200, 50000]))
    pc = float(rng.choice([0.5, 0.01, 0.001]))
    logodds, lens = synth.pwm_library(int(rng.integers(1 << 30)), M, lmin=lmin, lmax=lmax, alpha=alpha,
                                      pseudocount=pc, nsites=nsites)
    if case % 3 == 0:
        holes = rng.random(logodds.shape) < 0.02
        logodds = np.where(holes, -np.inf, logodds)
    n = 20000
    codes = rng.integers(0, 4, size=n).astype(np.uint8)
    a = int(rng.integers(0, n - 400))
    codes[a:a + 300] = int(rng.integers(0, 4))           # homopolymer run
    # implant consensus-like windows so sharp motifs have hits
    for m in range(0, M, 3):
        L = int(lens[m])
        fin = np.where(np.isinf(logodds[m, :L]), -1e9, logodds[m, :L])
        w = fin.argmax(axis=1)
        k = int(rng.integers(0, 3))
        idx = rng.choice(L, size=min(k, L), replace=False)
        w = w.copy()
        w[idx] = rng.integers(0, 4, size=idx.size)
        p = int(rng.integers(0, n - L))
        codes[p:p + L] = w
    seq = np.frombuffer(b"ACGT", dtype=np.uint8)[codes]
    rate = float(rng.choice([1e-5, 1e-4, 1e-4, 1e-3, 1e-2]))
    fin = np.where(np.isinf(logodds), -20.0, logodds)
    thr = synth.thresholds_for_rate(fin, lens, rate=rate, gc=0.5, seed=int(rng.integers(1 << 30)), sample=20000)
    # thresholds exactly on a window's score (a hit by equality), one ulp above it (not a hit)
    for m in range(1, M, 4) if case % 3 else ():         # (every third case: sparse thresholds only, whole-set plan)
        L = int(lens[m])
        p = int(rng.integers(0, n - L))
        with np.errstate(invalid="ignore"):
            s = po.calculate(seq[p:p + L], logodds[m, :L])[0]
        if np.isfinite(s):
            thr[m] = s if (m // 4) % 2 else np.nextafter(np.float32(s), np.float32(np.inf))
    if case % 3:
        thr[7::31] = -np.inf
    thr[11::31] = np.inf
    thr[13::31] = np.nan
    with np.errstate(invalid="ignore"):
        pos, mot, strand, score, cnt = po.search_library(seq, logodds, lens, thr, threads=0)
    want = np.zeros((n, M, 2), dtype=np.uint8)
    want[pos, mot, strand] = 1
    inside = np.arange(n)[:, None] + lens[None, :] <= n

    cand, info = _lib.direct_lookup_host(logodds, lens, thr, codes)
    direct = _lib.direct_motifs_host(logodds, lens, thr)
    assert int(direct.sum()) == info["motifs"] and (info["motifs"] == 0 or info["motifs"] >= 32)
    if info["motifs"]:
        assert int(lens[direct].max()) == info["max_len"] <= 14
        exact = direct & (lens <= 12)
        if not np.array_equal(cand[:, exact][inside[:, exact]], want[:, exact][inside[:, exact]]):
            bad = np.argwhere(cand[:, exact] != want[:, exact])
            raise AssertionError(f"case {case}: direct path differs from the oracle at {bad[:5].tolist()}")
        longer = direct & (lens > 12)
        if longer.any() and not np.all(cand[:, longer][inside[:, longer]] >= want[:, longer][inside[:, longer]]):
            raise AssertionError(f"case {case}: direct path lost a hit of a 13/14-column motif")
    assert cand[:, ~direct].sum() == 0
    # the table side of the real plan (form 4) must cover every hit of the motifs it keeps; form 3 (every
    # motif on the tables) must cover all of them
    info_plan, real_plan = _lib.hits_plan_checksum_host(logodds, lens, thr)   # the scan's own (parallel) planner
    assert info_plan["direct_motifs"] == info["motifs"] and info_plan["direct_keys"] == info["keys"]
    for form in (3, 4):
        filt, plan, _ = _lib.hits_filter_host(logodds, lens, thr, form, codes, with_plan=True)
        if form == 4:
            assert plan == real_plan, f"case {case}: parallel and sequential planners disagree"

        keep = np.ones(cnt, dtype=bool) if form == 3 else ~direct[mot]
        lost = int((filt[pos[keep], mot[keep], strand[keep]] == 0).sum())
        if lost:
            raise AssertionError(f"case {case}: the table filter (form {form}) lost {lost} hits")
        if form == 4:
            assert filt[:, direct].sum() == 0
    return cnt, info["motifs"], int(direct.sum()) < int((lens <= 14).sum()), int(filt.sum())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=200)
    ap.add_argument("--seed", type=int, default=1)
    args = ap.parse_args()
    for v in ("TFBS_HITS_DIRECT", "TFBS_HITS_NARROW", "TFBS_HITS_STICKY"):
        os.environ.pop(v, None)
    rng = np.random.default_rng(args.seed)
    t0 = time.time()
    tot_hits = tot_direct = tot_partial = 0
    for case in range(args.cases):
        cnt, nd, partial, fc = one_case(rng, case)
        tot_hits += cnt
        tot_direct += 1 if nd else 0
        tot_partial += 1 if nd and partial else 0
        if case % 20 == 19:
            print(f"{case + 1} cases, {tot_hits} oracle hits, direct path taken in {tot_direct} "
                  f"(leaving eligible motifs on the tables in {tot_partial}), {time.time() - t0:.0f} s", flush=True)
    print(f"OK: {args.cases} cases, {tot_hits} oracle hits, direct path taken in {tot_direct} cases ({tot_partial} partially)")


if __name__ == "__main__":
    main()
