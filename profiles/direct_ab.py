#!/usr/bin/env python3
"""A/B of the direct path's reach (bitmap density cap) on the C4 workload."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from discover_tfbs_b200 import _lib, synth
ctx = _lib.Context(0)
n = 100_000_000
logodds, lens = synth.pwm_library(39, 800, 6, 30)
thr = synth.thresholds_for_rate(logodds, lens, rate=1e-4, gc=0.41, seed=39, sample=100_000)
dev = ctx.synth_ascii(39, 0, n, gc=0.41)
enc = ctx.encode_device(dev.ptr, n)
for dens in ("0.35", "0.5", "0.7"):
    os.environ["TFBS_DIRECT_DENSITY"] = dens
    lib = ctx.upload_pwms(logodds, lens)
    ms = []
    for i in range(4):
        ctx.flush_l2()
        hits = ctx.scan_hits(enc, lib, thr, cap=1 << 26, to_host=False)
        ms.append((ctx.last_kernel_ms(), ctx.last_direct_ms(), *ctx.last_hits_ms()))
    best = min(ms)
    print(f"density cap {dens}: total {best[0]:.2f} ms (direct {best[1]:.2f}, scan {best[2]:.2f}, verify {best[3]:.2f}), hits {hits}, "
          f"candidates {ctx.last_scan_candidates()}, direct {lib.hits_direct()}, loads/base {sum(b[0] + b[1] for b in lib.hits_blocks())}", flush=True)
    lib.close()
