"""CPU simulation: candidate inflation of the hits filter if its packed deficit fields were 8 bits
wide (64 motifs per block, 4 fields per LUT word) instead of 16 (32 motifs, 2 fields).

Halving the field width would halve shared-memory bytes per (window, motif) — the resource the hits
kernel is bound by (profiles/r01/SUMMARY.md) — but the no-carry constraint leaves only
Dq <= (128 - G) / (G - 1) quantisation steps for the deficit budget D, and every group rounds DOWN,
so windows up to G steps over budget still pass.  This script measures how many more candidates the
exact verifier would see, on the bench library (800 PWMs, L 6..30, thresholds at 1e-4).

Pure NumPy, no GPU, no oracle: python profiles/field8_simulation.py
"""
import sys

import numpy as np

sys.path.insert(0, ".")
from discover_tfbs_b200 import synth  # noqa: E402

SEED, N = 39, 400_000


def plan(L):
    best = None
    for a in range(0, 7):
        for b in range(0, 12):
            if a + b < 1 or 4 * a + 3 * b < L or a * 32768 + b * 8192 > 192 * 1024:
                continue
            if best is None or a + b < best[0] + best[1] or (a + b == best[0] + best[1] and a > best[0]):
                best = (a, b)
    return best


def main():
    logodds, lens = synth.pwm_library(SEED, 800, 6, 30)
    thr = synth.thresholds_for_rate(logodds, lens, rate=1e-4, gc=0.41, seed=SEED, sample=100_000)
    rng = np.random.default_rng(1)
    pb = np.array([0.295, 0.205, 0.205, 0.295])
    seq = rng.choice(4, size=N + 32, p=pb)
    order = np.argsort(-lens, kind="stable")
    rows = {}
    for width, per_block in ((16, 32), (8, 64)):
        for bi in range(0, 800, per_block):
            blk = order[bi:bi + per_block]
            Lmax = int(lens[blk].max())
            A, B = plan(Lmax)
            G = A + B
            if width == 16:
                dq_max = 0x7FFF // G - 3
            else:
                dq_max = max((128 - G) // max(G - 1, 1) - 1, 1)
            gcol = [4 * g if g < A else 4 * A + 3 * (g - A) for g in range(G)]
            gk = [4 if g < A else 3 for g in range(G)]
            for m in blk[:: max(per_block // 8, 1)]:  # 8 motifs per block are enough
                L = int(lens[m])
                mat = logodds[m, :L]
                score = np.zeros(N)
                for j in range(L):
                    score += mat[j, seq[j:j + N]]
                best = 0.0
                qsum = np.zeros(N, dtype=np.int64)
                D = None
                parts = []
                for g in range(G):
                    if gcol[g] >= L:
                        break
                    cols = range(gcol[g], min(gcol[g] + gk[g], L))
                    part = np.zeros(N)
                    mx = 0.0
                    for j in cols:
                        part += mat[j, seq[j:j + N]]
                        mx += mat[j].max()
                    best += mx
                    parts.append(mx - part)
                D = best - float(thr[m])
                if D < 0:
                    continue
                scale = dq_max / max(D, 1e-300)
                Dq = min(np.floor(D * scale) + 1, dq_max + 1)
                for d in parts:
                    qsum += np.minimum(np.floor(d * scale), dq_max + 2).astype(np.int64)
                cand = int((qsum <= Dq).sum())
                hits = int((score >= thr[m]).sum())
                key = (width, G)
                r = rows.setdefault(key, [0, 0, 0])
                r[0] += cand
                r[1] += hits
                r[2] += N
    print("field_bits groups candidates hits windows cand_rate hit_rate inflation")
    for (width, G), (c, h, n) in sorted(rows.items()):
        print(f"{width:10d} {G:6d} {c:10d} {h:6d} {n:9d} {c / n:9.2e} {h / n:8.2e} {c / max(h, 1):9.2f}")


if __name__ == "__main__":
    main()
