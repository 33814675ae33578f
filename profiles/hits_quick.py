#!/usr/bin/env python3
"""Quick timing + parity spot-check of the hits kernel on the C4 workload (100 Mb x 800 PWMs)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from discover_tfbs_b200 import _lib, synth
from oracle import pssm_oracle as po
ctx = _lib.Context(0)
logodds, lens = synth.pwm_library(39, 800, 6, 30)
thr = synth.thresholds_for_rate(logodds, lens, rate=1e-4, gc=0.41, seed=39, sample=100_000)
lib = ctx.upload_pwms(logodds, lens)
# parity spot check on 300 kb
small = synth.genome(39, 0, 300_000)
e = ctx.encode(small)
h = ctx.scan_hits(e, lib, thr, sort=True)
pos, mot, strand, score, cnt = po.search_library(small, logodds, lens, thr)
gp = np.where(h["pos"] < 0, ~h["pos"], h["pos"])
assert cnt == h.size and np.array_equal(gp, pos) and np.array_equal(h["motif"], mot) and np.array_equal(h["score"], score)
print("parity ok on 300 kb:", cnt, "hits")
n = 100_000_000
dev = ctx.synth_ascii(39, 0, n, gc=0.41)
enc = ctx.encode_device(dev.ptr, n)
units = sum(n - int(L) + 1 for L in lens)
MODES = {"auto": {}, "narrow": {"TFBS_HITS_NARROW": "9"}, "sticky>=5": {"TFBS_HITS_NARROW": "4", "TFBS_HITS_STICKY": "5"},
         "wide": {"TFBS_HITS_NARROW": "0"}}  # cost model (default) and the three forced block forms
for narrow, env in MODES.items():
    os.environ.pop("TFBS_HITS_NARROW", None)
    os.environ.pop("TFBS_HITS_STICKY", None)
    os.environ.update(env)
    lib2 = ctx.upload_pwms(logodds, lens)
    ms = []
    for i in range(4):
        ctx.flush_l2()
        hits = ctx.scan_hits(enc, lib2, thr, cap=1 << 25, to_host=False)
        ms.append(ctx.last_kernel_ms())
    print(f"scan_hits {narrow}: {min(ms):.2f} ms (all {[round(m,2) for m in ms]}), {units / min(ms) / 1e6:.1f} G units/s, "
          f"hits {hits}, candidates {ctx.last_scan_candidates()}")
    lib2.close()
