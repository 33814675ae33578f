#!/usr/bin/env python3
"""End-to-end (host buffers) time of the C4 workload vs the number of pipeline chunks."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from discover_tfbs_b200 import _lib, synth
ctx = _lib.Context(0)
n = 100_000_000
logodds, lens = synth.pwm_library(39, 800, 6, 30)
thr = synth.thresholds_for_rate(logodds, lens, rate=1e-4, gc=0.41, seed=39, sample=100_000)
lib = ctx.upload_pwms(logodds, lens)
dev = ctx.synth_ascii(39, 0, n, gc=0.41)
pin = _lib.PinnedBuffer(n); pin.array[:] = dev.download()
enc = ctx.alloc_seq(n)
cap = 1 << 26
out = _lib.PinnedBuffer(cap * 16); hits = out.view(_lib.HIT_DTYPE)
for chunks in (1, 2, 4, 6, 8, 12, 16, 24):
    ms = []
    for i in range(6):
        ctx.flush_l2()
        t0 = time.perf_counter()
        got = ctx.scan_hits_host_shard(pin.array, enc, lib, thr, hits[:cap], owned=n, shard_start=0, chunks=chunks)
        ms.append((time.perf_counter() - t0) * 1e3)
    print(f"chunks {chunks:2d}: {min(ms[2:]):.2f} ms (all {[round(m, 2) for m in ms]}), hits {got.size}", flush=True)
