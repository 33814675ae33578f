#!/usr/bin/env python3
"""Sensitivity of the hits kernel: hit-rate sweep (thresholds) and genome-size sweep, 800 PWMs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from discover_tfbs_b200 import _lib, synth
ctx = _lib.Context(0)
logodds, lens = synth.pwm_library(39, 800, 6, 30)
lib = ctx.upload_pwms(logodds, lens)
def run(n, thr, cap):
    dev = ctx.synth_ascii(39, 0, n, gc=0.41)
    enc = ctx.encode_device(dev.ptr, n)
    ms = []
    for _ in range(3):
        ctx.flush_l2()
        h = ctx.scan_hits(enc, lib, thr, cap=cap, to_host=False)
        ms.append(ctx.last_kernel_ms())
    units = sum(n - int(L) + 1 for L in lens)
    enc.close(); dev.close()
    return min(ms), units / min(ms) / 1e9, h, ctx.last_scan_candidates()
print("| rate target | ms (100 Mb) | T units/s | hits | candidates |\n|---|---|---|---|---|")
for rate in (1e-2, 1e-3, 1e-4, 1e-5, 1e-6):
    thr = synth.thresholds_for_rate(logodds, lens, rate=rate, gc=0.41, seed=39, sample=400_000)
    ms, tu, h, c = run(100_000_000, thr, cap=int(max(4e8 * rate * 800 / 0.8, 1 << 20)) if rate <= 1e-3 else 1 << 20)
    print(f"| {rate:g} | {ms:.2f} | {tu:.3f} | {h} | {c} |", flush=True)
thr = synth.thresholds_for_rate(logodds, lens, rate=1e-4, gc=0.41, seed=39, sample=100_000)
print("\n| genome bp | ms | T units/s |\n|---|---|---|")
for n in (1_000_000, 10_000_000, 100_000_000, 387_500_000, 1_000_000_000):
    ms, tu, h, c = run(n, thr, cap=int(2.4 * 800 * n * 1e-4) + (1 << 20))
    print(f"| {n} | {ms:.2f} | {tu:.3f} |", flush=True)
