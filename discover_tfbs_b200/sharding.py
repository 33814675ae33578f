"""Chunk + halo sharding of a genome scan across GPUs (one process or thread per GPU).

The path has no cross-shard reduction: a window belongs to the shard that
contains its start, so every shard scans `[start, owned_end + halo)` independently and keeps
only hits that start before `owned_end`; results are concatenated on the host in shard order.
No collective is used on the data path (torch.distributed only provides the barrier / timing
reduction in bench.py).
"""
from dataclasses import dataclass

import numpy as np

from . import _lib


@dataclass(frozen=True)
class Shard:
    index: int
    start: int        # first base scanned (global coordinate)
    owned_end: int    # windows starting in [start, owned_end) belong to this shard
    scan_end: int     # last base read + 1 = min(owned_end + halo, total)

    @property
    def owned(self) -> int:
        return self.owned_end - self.start

    @property
    def scanned(self) -> int:
        return self.scan_end - self.start


def plan_shards(total: int, n_shards: int, halo: int, align: int = 128):
    """Split [0, total) into n_shards contiguous ranges whose boundaries are multiples of `align`
    bases (16-byte aligned in packed space), each extended by a right halo of `halo` bases
    (Lmax - 1 for PWM windows; the pentamer/step lookups need 2 left + 3 right, covered by
    scanning from `start - 2` when shapes are requested)."""
    if n_shards < 1:
        raise ValueError("n_shards must be >= 1")
    per = -(-total // n_shards)
    per = -(-per // align) * align
    shards = []
    for i in range(n_shards):
        start = min(i * per, total)
        owned_end = min((i + 1) * per, total)
        shards.append(Shard(i, start, owned_end, min(owned_end + halo, total)))
    return shards


def units_of_shard(shard: Shard, lens, total: int) -> int:
    """motif x bp units owned by a shard: windows starting in [start, owned_end) that fit in the
    genome, summed over motifs."""
    u = 0
    for L in lens:
        last = total - int(L)  # last valid start
        u += max(min(shard.owned_end - 1, last) - shard.start + 1, 0)
    return u


def filter_owned(hits: np.ndarray, shard: Shard) -> np.ndarray:
    """Drop halo duplicates (windows that start at or after owned_end) and shift shard-local
    positions to global coordinates.  Reverse-strand hits keep the ~pos encoding."""
    local = np.where(hits["pos"] < 0, ~hits["pos"], hits["pos"])
    keep = local < shard.owned
    out = hits[keep].copy()
    g = local[keep] + shard.start
    out["pos"] = np.where(out["pos"] < 0, ~g, g)
    return out


def merge_shard_hits(per_shard: list) -> np.ndarray:
    """Concatenate per-shard hit arrays (already global) and order them (motif, pos, strand)."""
    if not per_shard:
        return np.zeros(0, dtype=_lib.HIT_DTYPE)
    allh = np.concatenate(per_shard)
    pos = np.where(allh["pos"] < 0, ~allh["pos"], allh["pos"])
    order = np.lexsort(((allh["pos"] < 0), pos, allh["motif"]))
    return allh[order]


def scan_genome_multi_gpu(genome: np.ndarray, logodds, lens, thresholds, devices=None, cap=1 << 22):
    """Hits-mode scan of one host-resident genome across several GPUs of one node, one host thread
    per device (ctypes releases the GIL).  Returns globally ordered hits."""
    import threading

    ndev = _lib.device_count()
    devices = list(range(ndev)) if devices is None else list(devices)
    if not devices:
        raise RuntimeError("no CUDA device; there is no CPU fallback")
    lens = np.asarray(lens)
    shards = plan_shards(genome.size, len(devices), int(lens.max()) - 1)
    results = [None] * len(shards)
    errors = []

    def work(i, dev):
        try:
            sh = shards[i]
            with _lib.Context(dev) as ctx:
                lib = ctx.upload_pwms(logodds, lens)
                enc = ctx.encode(genome[sh.start: sh.scan_end])
                hits = ctx.scan_hits(enc, lib, thresholds, cap=cap, sort=False)
                results[i] = filter_owned(hits, sh)
                enc.close()
                lib.close()
        except Exception as exc:  # surfaced below
            errors.append(exc)

    threads = [threading.Thread(target=work, args=(i, d)) for i, d in enumerate(devices)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return merge_shard_hits(results)
