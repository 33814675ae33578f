#!/usr/bin/env python3
"""Launch every hot-path kernel a couple of times on a profiling-sized input (for ncu).

    python profiles/profile_kernels.py [--genome-mb 10] [--motifs 800]

Used as:  <cmd> && ncu --set full ... -k regex:'scan_hits_kernel|verify_candidates|scan_dense_kernel|encode_kernel|shape_tracks' <cmd>
"""
import argparse
import os
import sys


sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from discover_tfbs_b200 import _lib, dnashape, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--genome-mb", type=float, default=10.0)
ap.add_argument("--motifs", type=int, default=800)
ap.add_argument("--reps", type=int, default=2)
args = ap.parse_args()

ctx = _lib.Context(0)
n = int(args.genome_mb * 1e6)
logodds, lens = synth.pwm_library(39, args.motifs, 6, 30)
thr = synth.thresholds_for_rate(logodds, lens, rate=1e-4, gc=0.41, seed=39, sample=100_000)  # = bench.py workload()
lib = ctx.upload_pwms(logodds, lens)
dev = ctx.synth_ascii(39, 0, n, gc=0.41)
enc = ctx.encode_device(dev.ptr, n)
tables, kinds = dnashape.synthetic_tables(("MGW", "ProT", "Roll", "HelT"))
tab = ctx.upload_shapes(tables, kinds)
mats = [synth.consensus_pwm(39 + i, L)[0] for i, L in enumerate((8, 10, 12, 14, 17, 20))]
lib6 = ctx.upload_pwms(mats)
for rep in range(args.reps):
    enc.reencode_device(dev.ptr)
    t_enc = ctx.last_kernel_ms()
    hits = ctx.scan_hits(enc, lib, thr, cap=1 << 26, to_host=False)
    t_hits = ctx.last_kernel_ms()
    t_scan, t_verify = ctx.last_hits_ms()
    ctx.scan_dense(enc, lib6, to_host=False)
    t_dense = ctx.last_kernel_ms()
    ctx.shape_tracks(enc, tab, to_host=False)
    t_shape = ctx.last_kernel_ms()
    print(f"rep {rep}: encode {t_enc:.3f} ms, scan_hits {t_hits:.3f} ms = filter {t_scan:.3f} + verify {t_verify:.3f} "
          f"({hits} hits, {ctx.last_scan_candidates()} candidates), "
          f"scan_dense(6 PWMs) {t_dense:.3f} ms, shape_tracks {t_shape:.3f} ms")
ctx.close()
