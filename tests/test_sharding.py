"""Multi-GPU sharding logic: CPU tests (incl. a world_size-2 gloo run) with the oracle standing in
for the per-rank scan, and a GPU test that runs the shards one after another on one device."""
import os
import socket

import numpy as np
import pytest

from discover_tfbs_b200 import _lib, sharding, synth
from oracle import pssm_oracle as po


def _oracle_hits(seq, logodds, lens, thr):
    pos, mot, strand, score, cnt = po.search_library(seq, logodds, lens, thr)
    h = np.zeros(cnt, dtype=_lib.HIT_DTYPE)
    h["pos"] = np.where(strand == 1, ~pos, pos)
    h["motif"] = mot
    h["score"] = score
    return h


def _workload(n=60_000, M=9):
    logodds, lens = synth.pwm_library(3, M, lmin=6, lmax=30)
    thr = synth.thresholds_for_rate(logodds, lens, rate=2e-3, sample=20_000)
    genome = synth.genome(3, 0, n)
    return genome, logodds, lens, thr


@pytest.mark.parametrize("total,n,halo", [(1000, 1, 29), (1000, 3, 29), (100_000_000, 8, 29), (130, 4, 5), (5, 8, 2)])
def test_plan_shards_partition(total, n, halo):
    shards = sharding.plan_shards(total, n, halo)
    assert len(shards) == n
    assert shards[0].start == 0 and shards[-1].owned_end == total
    for a, b in zip(shards, shards[1:]):
        assert a.owned_end == b.start
    for s in shards:
        assert s.start % 128 == 0 or s.start == total
        assert s.scan_end == min(s.owned_end + halo, total)
    lens = [6, 12, 30]
    assert sum(sharding.units_of_shard(s, lens, total) for s in shards) == sum(max(total - L + 1, 0) for L in lens)


@pytest.mark.parametrize("n_shards", [1, 2, 3, 8])
def test_sharded_oracle_equals_unsharded(n_shards):
    genome, logodds, lens, thr = _workload()
    want = sharding.merge_shard_hits([_oracle_hits(genome, logodds, lens, thr)])
    parts = []
    for sh in sharding.plan_shards(genome.size, n_shards, int(lens.max()) - 1):
        local = _oracle_hits(genome[sh.start: sh.scan_end], logodds, lens, thr)
        parts.append(sharding.filter_owned(local, sh))
    got = sharding.merge_shard_hits(parts)
    assert got.size == want.size and np.array_equal(got, want)


def test_ordered_concatenation_of_sorted_shards_equals_global_sort():
    """Shards are position-ordered, so per-shard (motif, pos, strand)-sorted arrays with global
    positions join by ordered concatenation: same result as one global lexsort."""
    genome, logodds, lens, thr = _workload()
    for n_shards in (1, 2, 5):
        parts = []
        for sh in sharding.plan_shards(genome.size, n_shards, int(lens.max()) - 1):
            local = _oracle_hits(genome[sh.start: sh.scan_end], logodds, lens, thr)  # oracle order is (motif, pos, strand)
            parts.append(sharding.filter_owned(local, sh))
        got = sharding.concat_sorted_shards(parts, len(lens))
        assert np.array_equal(got, sharding.merge_shard_hits(parts))
    empty = np.zeros(0, dtype=_lib.HIT_DTYPE)
    assert sharding.concat_sorted_shards([empty, empty], len(lens)).size == 0
    # threaded copies into a preallocated output give the same array
    rng = np.random.default_rng(1)
    big = []
    for k in range(3):
        h = np.zeros(700_000, dtype=_lib.HIT_DTYPE)
        h["motif"] = np.sort(rng.integers(0, 50, size=h.size))
        h["pos"] = np.arange(h.size) + k * 10**7
        h["score"] = rng.random(h.size)
        big.append(h)
    want = sharding.merge_shard_hits(big)
    buf = np.empty(want.size + 5, dtype=_lib.HIT_DTYPE)
    got = sharding.concat_sorted_shards(big, 50, out=buf, threads=4)
    assert got.base is not None and np.array_equal(got, want)
    # motif boundaries by vectorised binary search on the strided field (no copy of the column): every
    # motif present, some absent, all absent, one row
    for col in (big[0], big[0][big[0]["motif"] % 3 == 0], big[0][:1], big[0][:0]):
        if col.size:
            assert np.array_equal(sharding._motif_bounds(col["motif"], 50), np.searchsorted(col["motif"], np.arange(51)))
    assert np.array_equal(sharding.concat_sorted_shards([big[0][:0], big[1]], 50), big[1])
    assert np.array_equal(sharding.concat_sorted_shards(big, 50, threads=1), want)
    with pytest.raises(ValueError):
        sharding.concat_sorted_shards(big, 50, out=buf[:10])


@pytest.mark.parametrize("n,k", [(0, 3), (1, 4), (55_000, 8), (7, 7), (10, 3)])
def test_record_shards_cover_every_record_once(n, k):
    r = sharding.plan_record_shards(n, k)
    assert len(r) == k and r[0][0] == 0 and r[-1][1] == n
    assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
    sizes = [b - a for a, b in r]
    assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gloo_rank(rank, world, port, out_path):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    genome, logodds, lens, thr = _workload()
    # every rank regenerates its own chunk from the position-addressable generator
    sh = sharding.plan_shards(genome.size, world, int(lens.max()) - 1)[rank]
    chunk = synth.genome(3, sh.start, sh.scanned)
    assert np.array_equal(chunk, genome[sh.start: sh.scan_end])
    mine = sharding.filter_owned(_oracle_hits(chunk, logodds, lens, thr), sh)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)  # host-side gather only; no data-path collective
    if rank == 0:
        np.save(out_path, sharding.merge_shard_hits(gathered))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_gather_matches_single_process(tmp_path):
    import torch.multiprocessing as mp

    out = str(tmp_path / "hits.npy")
    mp.spawn(_gloo_rank, args=(2, _free_port(), out), nprocs=2, join=True)
    genome, logodds, lens, thr = _workload()
    want = sharding.merge_shard_hits([_oracle_hits(genome, logodds, lens, thr)])
    got = np.load(out)
    assert np.array_equal(got, want)


@pytest.mark.gpu
def test_shards_run_sequentially_on_one_gpu_match_single_scan(ctx):
    genome, logodds, lens, thr = _workload(n=300_000, M=40)
    lib = ctx.upload_pwms(logodds, lens)
    enc = ctx.encode(genome)
    want = sharding.merge_shard_hits([ctx.scan_hits(enc, lib, thr, sort=False)])
    enc.close()
    parts = []
    for sh in sharding.plan_shards(genome.size, 4, int(lens.max()) - 1):
        e = ctx.encode(genome[sh.start: sh.scan_end])
        parts.append(sharding.filter_owned(ctx.scan_hits(e, lib, thr, sort=False), sh))
        e.close()
    got = sharding.merge_shard_hits(parts)
    assert np.array_equal(got, want)
    assert np.array_equal(got, sharding.merge_shard_hits([_oracle_hits(genome, logodds, lens, thr)]))
    lib.close()
    multi = sharding.scan_genome_multi_gpu(genome, logodds, lens, thr, devices=[0])
    assert np.array_equal(multi, want)


@pytest.mark.gpu
@pytest.mark.parametrize("sort", [True, False])
def test_shard_entry_point_drops_halo_and_shifts_positions_on_the_device(ctx, sort):
    """tfbs_scan_hits_host_shard, shard after shard on one GPU: outputs carry global positions, no halo
    duplicates, and (sorted) join by ordered concatenation into exactly the oracle's whole-genome list."""
    genome, logodds, lens, thr = _workload(n=300_000, M=70)
    lib = ctx.upload_pwms(logodds, lens)
    want = sharding.merge_shard_hits([_oracle_hits(genome, logodds, lens, thr)])
    for n_shards in (1, 3, 7):
        parts = []
        for sh in sharding.plan_shards(genome.size, n_shards, int(lens.max()) - 1):
            enc = ctx.alloc_seq(sh.scanned)
            out = np.empty(1 << 18, dtype=_lib.HIT_DTYPE)
            part = ctx.scan_hits_host_shard(genome[sh.start: sh.scan_end], enc, lib, thr, out, owned=sh.owned,
                                            shard_start=sh.start, chunks=3, sort=sort)
            g = np.where(part["pos"] < 0, ~part["pos"], part["pos"])
            assert part.size == 0 or (g.min() >= sh.start and g.max() < sh.owned_end)
            parts.append(part.copy())
            enc.close()
        got = sharding.concat_sorted_shards(parts, len(lens)) if sort else sharding.merge_shard_hits(parts)
        assert np.array_equal(got, want), n_shards
    lib.close()


@pytest.mark.gpu
def test_scanner_is_reusable_across_genomes_and_thresholds(ctx):
    genome, logodds, lens, thr = _workload(n=200_000, M=40)
    ndev = _lib.device_count()
    with sharding.MultiGpuScanner(genome.size, logodds, lens, devices=list(range(ndev)), cap_per_device=1 << 18) as sc:
        pinned = _lib.PinnedBuffer(genome.size)
        for seed, scale in ((3, 1.0), (4, 1.0), (3, 1.05)):
            g = synth.genome(seed, 0, genome.size)
            pinned.array[:] = g
            t = (thr * scale).astype(np.float32)
            got = sc.scan(pinned.array, t)
            assert np.array_equal(got, sharding.merge_shard_hits([_oracle_hits(g, logodds, lens, t)]))
            assert sc.last_ms["scan_wall_ms"] > 0
        pinned.close()


@pytest.mark.gpu
def test_record_shards_reproduce_the_single_gpu_feature_columns(ctx):
    """Shard-by-record-index (configs C1 / C2): per-record best windows computed shard by shard on
    every visible GPU equal the single-context result."""
    fg, bkg = synth.peak_set(5, 300, 200, 2)
    buf, starts, lens_rec = synth.join_records(fg)
    mats = [synth.consensus_pwm(5 + i, L)[0] for i, L in enumerate((8, 12, 20))]
    enc = ctx.encode(buf)
    lib = ctx.upload_pwms(mats)
    want = ctx.scan_best(enc, lib, starts, lens_rec)

    def per_shard(c, first, last):
        b, st, ln = synth.join_records(fg[first:last])
        e = c.encode(b)
        l2 = c.upload_pwms(mats)
        return c.scan_best(e, l2, st, ln)

    got = np.concatenate(sharding.map_record_shards(per_shard, fg.shape[0]))
    assert np.array_equal(got, want)
    enc.close()
    lib.close()


@pytest.mark.gpu
def test_threads_drive_every_visible_gpu(ctx):
    """One host thread per device (ctypes releases the GIL): with 1 GPU this degenerates to the
    single-device case, with N it exercises N contexts concurrently."""
    ndev = _lib.device_count()
    genome, logodds, lens, thr = _workload(n=400_000, M=33)
    got = sharding.scan_genome_multi_gpu(genome, logodds, lens, thr, devices=list(range(ndev)))
    want = sharding.merge_shard_hits([_oracle_hits(genome, logodds, lens, thr)])
    assert np.array_equal(got, want)


@pytest.mark.gpu
def test_scanner_with_more_devices_than_the_genome_has_128_base_blocks(ctx):
    """A 200-base genome planned for three workers (contexts on the same GPU when the box has one): the
    third shard is empty and gets no worker; the hits equal the oracle's."""
    genome, logodds, lens, thr = _workload(n=200, M=9)
    thr = np.full(lens.size, -5.0, dtype=np.float32)
    got = sharding.scan_genome_multi_gpu(genome, logodds, lens, thr, devices=[0, 0, 0], cap=1 << 14)
    want = sharding.merge_shard_hits([_oracle_hits(genome, logodds, lens, thr)])
    assert want.size > 20 and np.array_equal(got, want)
