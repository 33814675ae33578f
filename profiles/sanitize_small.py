#!/usr/bin/env python3
"""Smallest end-to-end pass of every kernel, for compute-sanitizer (one tool per gpurun call)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from discover_tfbs_b200 import _lib, dnashape, synth, similarity as sm
ctx = _lib.Context(0)
seq = synth.genome(39, 0, 40_000)
enc = ctx.encode(seq)
enc.base_counts(); enc.gc_counts([0, 100], [50, 33])
logodds, lens = synth.pwm_library(39, 70, 1, 32)
thr = synth.thresholds_for_rate(logodds, lens, rate=2e-3, sample=5000)
lib = ctx.upload_pwms(logodds, lens)
print("hits", ctx.scan_hits(enc, lib, thr).size)
print("flood", ctx.scan_hits(enc, lib, np.full(70, -1e30, np.float32), cap=1 << 23, to_host=False))
f, r = ctx.scan_dense(enc, lib)
tables, kinds = dnashape.synthetic_tables()
tab = ctx.upload_shapes(tables, kinds)
ctx.shape_tracks(enc, tab, [0, 20_000], [19_999, 20_000])
ctx.shape_sites(enc, tab, [0, 0], [40_000, 40_000], [5, 100], [17, 112], [1, -1], width=13)
ctx.track_sites(np.arange(100, dtype=np.float32), [0, 0], [100, 100], [5, 90], [17, 102], [1, -1], 13)
c = ["".join(np.random.default_rng(i).choice(list("ACGT"), 12)) for i in range(9)]
print(ctx.similarity_sums(sm.prepare(c, 6), sm.prepare(c[:4], 6)).shape)
print("best", ctx.scan_best(enc, lib, [0, 100, 39_990], [50, 3000, 10], with_pos=True)[0].shape)
out = np.empty(1 << 16, dtype=_lib.HIT_DTYPE)
e2 = ctx.alloc_seq(seq.size)
print("host shard", ctx.scan_hits_host_shard(seq, e2, lib, thr, out, owned=30_000, shard_start=7, chunks=3, sort=True).size)
print("narrow cap", ctx.scan_hits(enc, lib, thr, cap=16, grow=True).size)
from discover_tfbs_b200.classify import GpuSVC
g = GpuSVC(np.ones(5), np.zeros(5), np.zeros(5), np.ones(5), "rbf", np.random.default_rng(0).normal(size=(7, 5)),
           np.ones(7), 0.1, 0.0, ["a", "b"], ctx=ctx)
print("svc", g.decision_function(np.random.default_rng(1).normal(size=(33, 5))).shape)
d = ctx.synth_ascii(39, 5, 1000); d.download(); d.close()
print("sanitize pass done")
