"""Historical (r01): timed the hits kernel with the last column group of long-motif blocks dropped from the
filter (TFBS_HITS_DROP_MIN_G).  The switch was a net loss and has been removed from the kernel, so this script
no longer changes anything; results in profiles/r01/drop_group_experiment.txt."""
import os, sys
sys.path.insert(0, '/root/repo')
from discover_tfbs_b200 import _lib, synth
ctx = _lib.Context(0)
n = 100_000_000
logodds, lens = synth.pwm_library(39, 800, 6, 30)
thr = synth.thresholds_for_rate(logodds, lens, rate=1e-4, gc=0.41, seed=39, sample=100_000)
dev = ctx.synth_ascii(39, 0, n, gc=0.41)
enc = ctx.encode_device(dev.ptr, n)
for dm in (0, 9, 7, 5, 4, 3, 2):
    os.environ["TFBS_HITS_DROP_MIN_G"] = str(dm)
    lib = ctx.upload_pwms(logodds, lens)
    ms = []
    for i in range(3):
        h = ctx.scan_hits(enc, lib, thr, cap=1 << 25, to_host=False)
        ms.append(ctx.last_kernel_ms())
    print(f"drop_min_g={dm}: {min(ms):.2f} ms, hits {h}, candidates {ctx.last_scan_candidates()}", flush=True)
    lib.close()
