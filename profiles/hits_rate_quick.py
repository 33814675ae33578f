"""Hits kernel vs hit density (100 Mb x 800 PWMs); thresholds precomputed in hits_rate_thresholds.npz."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from discover_tfbs_b200 import _lib, synth
ctx = _lib.Context(0)
logodds, lens = synth.pwm_library(39, 800, 6, 30)
thrs = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "hits_rate_thresholds.npz"))
n = 100_000_000
dev = ctx.synth_ascii(39, 0, n, gc=0.41)
enc = ctx.encode_device(dev.ptr, n)
for narrow in ("", "9", "0"):  # cost model (default); narrow forced for every block; wide blocks only
    os.environ.pop("TFBS_HITS_NARROW", None)
    if narrow:
        os.environ["TFBS_HITS_NARROW"] = narrow
    lib = ctx.upload_pwms(logodds, lens)
    for rate in ("0.01", "0.001", "0.0001", "1e-05", "1e-06"):
        thr = thrs[rate]
        ms = []
        for _ in range(3):
            h = ctx.scan_hits(enc, lib, thr, cap=1 << 28, to_host=False)
            ms.append(ctx.last_kernel_ms())
        plan = lib.hits_blocks()
        print(f"narrow<={narrow or "auto"} rate {rate}: {min(ms):.2f} ms, hits {h}, candidates {ctx.last_scan_candidates()}, "
              f"blocks {len(plan)} ({sum(b[2] for b in plan)} narrow: groups {sorted(set(b[0] + b[1] for b in plan if b[2]))})", flush=True)
    lib.close()
