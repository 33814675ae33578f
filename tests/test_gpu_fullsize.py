"""BASELINE.json full-size configurations checked through size-independent properties (the CPU
oracle would need hours there): cross-kernel consistency, shard invariance, idempotence."""
import numpy as np
import pytest

from discover_tfbs_b200 import dnashape, sharding, synth

pytestmark = pytest.mark.gpu


def _c3_genome():
    """Config C3: 8 chromosomes totalling 14.3 Mb, GC 33.5 %, N-runs, 5 % lower case."""
    props = np.array([3.19, 2.23, 1.80, 1.60, 1.19, 1.03, 0.95, 2.29])
    sizes = np.round(props / props.sum() * 14.3e6).astype(np.int64)
    chroms, off = [], 0
    for s in sizes:
        chroms.append(synth.genome(39, off, int(s), gc=0.335))
        off += int(s)
    return synth.join_records([c.tobytes().decode("ascii") for c in chroms])


def test_c3_whole_genome_hits_agree_with_dense_scores(ctx):
    buf, starts, lens_rec = _c3_genome()
    mats = [synth.consensus_pwm(39 + i, L)[0] for i, L in enumerate((8, 10, 12, 14, 17, 20))]
    lib = ctx.upload_pwms(mats)
    lens = np.asarray(lib.lens)
    thr = np.array([0.8 * m.max(axis=1).sum() for m in mats], dtype=np.float32)  # 0.8 x max score
    enc = ctx.encode(buf)
    hits = ctx.scan_hits(enc, lib, thr, cap=1 << 22, sort=True)
    again = ctx.scan_hits(enc, lib, thr, cap=1 << 22, sort=True)
    assert np.array_equal(hits, again)  # idempotent / deterministic after sorting
    fwd, rev = ctx.scan_dense(enc, lib)
    gpos = np.where(hits["pos"] < 0, ~hits["pos"], hits["pos"])
    isrev = hits["pos"] < 0
    dense_at_hit = np.where(isrev, rev[hits["motif"], gpos], fwd[hits["motif"], gpos])
    assert np.max(np.abs(dense_at_hit - hits["score"])) <= 1e-4      # fp32 path vs exact re-score
    assert np.all(hits["score"] >= thr[hits["motif"]])
    for m in range(lib.M):
        for arr, flag in ((fwd[m], False), (rev[m], True)):
            with np.errstate(invalid="ignore"):
                sure = int((arr >= thr[m] + 1e-4).sum())
                maybe = int((arr >= thr[m] - 1e-4).sum())
            got = int(((hits["motif"] == m) & (isrev == flag)).sum())
            assert sure <= got <= maybe, (m, flag, sure, got, maybe)
    # no hit window may cross a chromosome boundary or touch an N
    ci = np.searchsorted(starts, gpos, side="right") - 1
    assert np.all(gpos + lens[hits["motif"]] <= starts[ci] + lens_rec[ci])
    assert hits.size > 100
    enc.close()
    lib.close()


def test_c4_scale_scan_is_shard_invariant(ctx):
    """100 Mb x 800 PWMs: one scan == three chunk+halo shards (count, checksums of pos/motif/score)."""
    n = 100_000_000
    logodds, lens = synth.pwm_library(39, 800, 6, 30)
    thr = synth.thresholds_for_rate(logodds, lens, rate=1e-4, gc=0.41, seed=39, sample=50_000)
    lib = ctx.upload_pwms(logodds, lens)

    def digest(h):
        g = np.where(h["pos"] < 0, ~h["pos"], h["pos"]).astype(np.uint64)
        key = g * np.uint64(1601) + h["motif"].astype(np.uint64) * np.uint64(2) + (h["pos"] < 0)
        return h.size, int(np.bitwise_xor.reduce(key)), int(key.sum(dtype=np.uint64)), float(h["score"].astype(np.float64).sum())

    dev = ctx.synth_ascii(39, 0, n, gc=0.41)
    enc = ctx.encode_device(dev.ptr, n)
    whole = digest(ctx.scan_hits(enc, lib, thr, cap=1 << 25, sort=False))
    enc.close()
    dev.close()
    parts = []
    for sh in sharding.plan_shards(n, 3, int(lens.max()) - 1):
        d = ctx.synth_ascii(39, sh.start, sh.scanned, gc=0.41)  # every shard regenerates its own bases
        e = ctx.encode_device(d.ptr, sh.scanned)
        parts.append(sharding.filter_owned(ctx.scan_hits(e, lib, thr, cap=1 << 24, sort=False), sh))
        e.close()
        d.close()
    merged = digest(np.concatenate(parts))
    assert merged[:3] == whole[:3] and abs(merged[3] - whole[3]) < 1e-6 * abs(whole[3])
    assert 1.0e7 < whole[0] < 3.0e7
    lib.close()


def test_c5_shard_shape_tracks_properties(ctx):
    """387.5 Mb shard (config C5 / 8 GPUs) x 4 tracks: sampled windows equal the oracle."""
    from oracle import shape_oracle as so

    n = 387_500_000
    shapes = ("MGW", "ProT", "Roll", "HelT")
    tables, kinds = dnashape.synthetic_tables(shapes)
    tab = ctx.upload_shapes(tables, kinds)
    dev = ctx.synth_ascii(39, 0, n, gc=0.41)
    enc = ctx.encode_device(dev.ptr, n)
    tracks = ctx.shape_tracks(enc, tab)
    rng = np.random.default_rng(5)
    for s in list(rng.integers(0, n - 5000, size=12)) + [0, n - 5000]:
        s = int(s)
        seg = synth.genome(39, s, 5000, gc=0.41)
        ref = so.tracks(seg, tables, kinds)
        # interior of the window is independent of the window ends
        assert np.array_equal(tracks[:, s + 3: s + 4996], ref[:, 3:4996], equal_nan=True)
    assert np.isnan(tracks[0, :2]).all() and np.isnan(tracks[0, -2:]).all()
    enc.close()
    dev.close()
    tab.close()


def test_c5_scale_positions_beyond_2_31(ctx):
    """One flat 3.1 Gb sequence (config C5 as a single record): hit positions are int64 and the
    part of the scan beyond 2^31 equals an independent scan of the same bases regenerated from the
    position-addressable generator."""
    n = 3_100_000_000
    logodds, lens = synth.pwm_library(5, 4, lmin=10, lmax=20)
    thr = synth.thresholds_for_rate(logodds, lens, rate=2e-6, gc=0.41, seed=5, sample=2_000_000)
    lib = ctx.upload_pwms(logodds, lens)
    dev = ctx.synth_ascii(39, 0, n, gc=0.41)
    enc = ctx.encode_device(dev.ptr, n)
    dev.close()
    hits = ctx.scan_hits(enc, lib, thr, cap=1 << 22, sort=True)
    enc.close()
    gpos = np.where(hits["pos"] < 0, ~hits["pos"], hits["pos"])
    assert hits.size > 1000 and gpos.max() > 2**31 and gpos.min() >= 0
    assert np.all(gpos[hits["motif"] == 0][1:] >= gpos[hits["motif"] == 0][:-1])  # sorted per motif
    # independent check of the last 50 Mb (all positions > 2^31)
    start = n - 50_000_000
    tail = ctx.synth_ascii(39, start, 50_000_000, gc=0.41)
    e2 = ctx.encode_device(tail.ptr, 50_000_000)
    tail.close()
    part = ctx.scan_hits(e2, lib, thr, cap=1 << 20, sort=True)
    e2.close()
    lib.close()
    lp = np.where(part["pos"] < 0, ~part["pos"], part["pos"]) + start
    sel = gpos >= start
    # windows starting in the first Lmax-1 bases of the tail see the same bases in both scans
    assert np.array_equal(gpos[sel], lp) and np.array_equal(hits["motif"][sel], part["motif"])
    assert np.array_equal(hits["score"][sel], part["score"])
    assert np.array_equal(hits["pos"][sel] < 0, part["pos"] < 0)
