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
MODES = {"auto": {}, "narrow": {"TFBS_HITS_NARROW": "9"}, "sticky>=5": {"TFBS_HITS_NARROW": "4", "TFBS_HITS_STICKY": "5"},
         "wide": {"TFBS_HITS_NARROW": "0"}}  # cost model (default) and the three forced block forms
for narrow, env in MODES.items():
    if len(sys.argv) > 1 and narrow not in sys.argv[1:]:
        continue
    os.environ.pop("TFBS_HITS_NARROW", None)
    os.environ.pop("TFBS_HITS_STICKY", None)
    os.environ.update(env)
    lib = ctx.upload_pwms(logodds, lens)
    for rate in ("0.01", "0.001", "0.0001", "1e-05", "1e-06"):
        thr = thrs[rate]
        ms = []
        for _ in range(3):
            h = ctx.scan_hits(enc, lib, thr, cap=1 << 28, to_host=False)
            ms.append(ctx.last_kernel_ms())
        plan = lib.hits_blocks()
        print(f"{narrow} rate {rate}: {min(ms):.2f} ms, hits {h}, candidates {ctx.last_scan_candidates()}, "
              f"direct {lib.hits_direct()}, blocks {len(plan)} (narrow groups {sorted(b[0] + b[1] for b in plan if b[2] == 1)}, sticky groups {sorted(b[0] + b[1] for b in plan if b[2] == 2)})", flush=True)
    lib.close()
