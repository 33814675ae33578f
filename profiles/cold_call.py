"""Host wall clock of the public end-to-end call (tfbs_scan_hits_host_shard) on ONE GPU with a threshold
vector the table cache has not seen ("cold": direct-path enumeration + block plan + filter tables on the
host, uploads, then the scan) next to the warm call, on the bench workload C4 (800 PWMs x 100 Mb).

Usage (GPU box): python profiles/cold_call.py [--genome-mb 100] [--rates 1e-4,1e-3]"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from discover_tfbs_b200 import _lib, sharding, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genome-mb", type=float, default=100.0)
    ap.add_argument("--motifs", type=int, default=800)
    ap.add_argument("--rates", default="1e-4")
    ap.add_argument("--repeats", type=int, default=4)
    args = ap.parse_args()
    ctx = _lib.Context(0)
    logodds, lens = synth.pwm_library(bench.SEED, args.motifs, 6, 30)
    lib = ctx.upload_pwms(logodds, lens)
    n = int(args.genome_mb * 1e6)
    shard = sharding.plan_shards(n, 1, int(lens.max()) - 1)[0]
    out = {"host_cores": len(os.sched_getaffinity(0)), "genome_bp": n, "motifs": args.motifs, "rates": {}}
    for rate in (float(r) for r in args.rates.split(",")):
        thr = synth.thresholds_for_rate(logodds, lens, rate=rate, gc=0.41, seed=bench.SEED, sample=100_000)
        job = bench.HitsJob(ctx, lib, thr, lens, shard, n, rate, 6)
        job.step_e2e()
        warm, cold = [], []
        for k in range(args.repeats):
            t0 = time.perf_counter()
            job.step_e2e()
            warm.append((time.perf_counter() - t0) * 1e3)
            job.thr = (thr + np.float32(1e-3 * (k + 1))).astype(np.float32)   # a vector the cache has not seen
            t0 = time.perf_counter()
            job.step_e2e()
            cold.append((time.perf_counter() - t0) * 1e3)
            job.thr = thr
            job.step_e2e()
        out["rates"][f"{rate:g}"] = {"warm_ms": round(min(warm), 2), "cold_ms_min": round(min(cold), 2),
                                     "cold_ms_all": [round(c, 2) for c in cold], "hits": int(job.got.size),
                                     "direct_motifs": lib.hits_direct()[0], "blocks": len(lib.hits_blocks())}
        job.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
