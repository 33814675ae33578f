"""Chunk + halo sharding of a genome scan across GPUs (one process or thread per GPU), and
record-index sharding of record sets.

The path has no cross-shard reduction: a window belongs to the shard that
contains its start, so every shard scans `[start, owned_end + halo)` independently and keeps
only hits that start before `owned_end` (dropped on the device by `tfbs_scan_hits_host_shard`);
results are concatenated on the host in shard order.  No collective is used on the data path
(torch.distributed only provides the barrier / timing reduction in bench.py).
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
    """Concatenate per-shard hit arrays (already global) and order them (motif, pos, strand) with one
    global sort.  Generic form (any per-shard order); `concat_sorted_shards` is the fast path."""
    if not per_shard:
        return np.zeros(0, dtype=_lib.HIT_DTYPE)
    allh = np.concatenate(per_shard)
    pos = np.where(allh["pos"] < 0, ~allh["pos"], allh["pos"])
    order = np.lexsort(((allh["pos"] < 0), pos, allh["motif"]))
    return allh[order]


def _motif_bounds(motif_col: np.ndarray, n_motifs: int) -> np.ndarray:
    """Row of the first hit of every motif 0..n_motifs (lower bounds) in a motif-sorted hit array.  `motif_col`
    is the strided 'motif' field of the structured array: np.searchsorted would first copy all of it (74 MB per
    18.6 M hits), a vectorised binary search reads 25 x (n_motifs + 1) elements."""
    n = int(motif_col.size)
    q = np.arange(n_motifs + 1, dtype=np.int64)
    lo = np.zeros(n_motifs + 1, dtype=np.int64)
    hi = np.full(n_motifs + 1, n, dtype=np.int64)
    while True:
        active = lo < hi
        if not active.any():
            return lo
        mid = (lo + hi) >> 1
        right = active & (motif_col[np.minimum(mid, n - 1)] < q)
        lo = np.where(right, mid + 1, lo)
        hi = np.where(active & ~right, mid, hi)


def concat_sorted_shards(per_shard: list, n_motifs: int, out: np.ndarray = None, threads: int = None) -> np.ndarray:
    """Whole-genome (motif, pos, strand) order from per-shard arrays that are each in that order and
    hold global positions: shards are position-ordered, so for every motif the shard pieces are simply
    laid end to end — an ordered concatenation, no comparison sort.  The block copies of disjoint motif
    ranges run on `threads` host threads (default: the usable cores, at most 16; NumPy releases the GIL
    while it copies); `out` may be a
    preallocated HIT_DTYPE array of at least the total size (saves the page faults of a fresh one)."""
    if not per_shard:
        return np.zeros(0, dtype=_lib.HIT_DTYPE)
    if len(per_shard) == 1 and out is None:
        return per_shard[0]
    bounds = np.stack([_motif_bounds(h["motif"], n_motifs) if h.size else np.zeros(n_motifs + 1, dtype=np.int64)
                       for h in per_shard])                           # [shard][motif] -> row
    counts = np.diff(bounds, axis=1)                                  # [shard][motif]
    total = int(counts.sum())
    if out is None:
        out = np.empty(total, dtype=_lib.HIT_DTYPE)
    elif out.size < total or out.dtype != _lib.HIT_DTYPE:
        raise ValueError(f"out must be a HIT_DTYPE array of at least {total} entries")
    out = out[:total]
    # destination of the first hit of motif m: everything of the motifs before it
    motif_base = np.concatenate([[0], np.cumsum(counts.sum(axis=0))])
    # (copied as plain 2 x uint64 rows: slice assignment of a structured dtype goes field by field)
    raw_out = out.view(np.uint64).reshape(-1, 2)
    raw_in = [np.ascontiguousarray(h).view(np.uint64).reshape(-1, 2) for h in per_shard]

    def copy_motifs(m0, m1):
        for m in range(m0, m1):
            at = int(motif_base[m])
            for k in range(len(per_shard)):
                c = int(counts[k][m])
                if c:
                    lo = int(bounds[k][m])
                    raw_out[at: at + c] = raw_in[k][lo: lo + c]
                    at += c

    if threads is None:
        import os

        threads = min(16, len(os.sched_getaffinity(0)))
    threads = max(1, min(int(threads), n_motifs))
    if threads == 1 or total < (1 << 20):
        copy_motifs(0, n_motifs)
    else:
        import threading

        # motif ranges of about equal numbers of hits
        cuts = np.searchsorted(motif_base, np.linspace(0, total, threads + 1)[1:-1]).tolist()
        edges = [0] + [min(max(int(c), 0), n_motifs) for c in cuts] + [n_motifs]
        pool = [threading.Thread(target=copy_motifs, args=(edges[i], edges[i + 1])) for i in range(threads)
                if edges[i + 1] > edges[i]]
        for t in pool:
            t.start()
        for t in pool:
            t.join()
    return out


class _DeviceWorker:
    """One GPU of the node: context, motif library, shard-sized device buffers, pinned output."""

    def __init__(self, device, logodds, lens, shard: Shard, cap: int):
        _lib.bind_thread_near(device)   # this thread allocates the GPU's pinned output: keep it in the GPU's socket
        self.ctx = _lib.Context(device)
        self.lib = self.ctx.upload_pwms(logodds, lens)
        self.enc = self.ctx.alloc_seq(shard.scanned)
        self.out = _lib.PinnedBuffer(cap * _lib.HIT_DTYPE.itemsize)
        self.hits = self.out.view(_lib.HIT_DTYPE)
        self.shard = shard

    def close(self):
        self.enc.close()
        self.lib.close()
        self.out.close()
        self.ctx.close()


class MultiGpuScanner:
    """Hits-mode scan of host-resident genomes across the GPUs of one node (chunk +
    (Lmax - 1) halo shards, no collective).  One host thread per device (ctypes releases the GIL);
    every thread runs the chunk-pipelined host path `tfbs_scan_hits_host_shard` on its shard — H2D,
    encode + scan and D2H overlap on three streams per GPU; halo duplicates are dropped and positions
    made global ON THE DEVICE; per-shard outputs (device-sorted) are joined by ordered concatenation.
    Contexts, libraries, device buffers and pinned outputs persist across `scan()` calls."""

    def __init__(self, n_bases: int, logodds, lens, devices=None, cap_per_device: int = 1 << 22):
        ndev = _lib.device_count()
        self.devices = list(range(ndev)) if devices is None else list(devices)
        if not self.devices:
            raise RuntimeError("no CUDA device; there is no CPU fallback")
        self.lens = np.asarray(lens, dtype=np.int32)
        self.n = int(n_bases)
        self.shards = plan_shards(self.n, len(self.devices), int(self.lens.max()) - 1)
        self.workers = [None] * len(self.devices)
        self._run(lambda i: self._make(i, logodds, cap_per_device))
        self.last_ms = None

    def _make(self, i, logodds, cap):
        # (a genome shorter than n_devices x 128 bases leaves the last shards empty: no worker for them)
        if self.shards[i].scanned > 0:
            self.workers[i] = _DeviceWorker(self.devices[i], logodds, self.lens, self.shards[i], cap)

    def _run(self, fn):
        import threading

        errors = []

        def guarded(i):
            try:
                fn(i)
            except Exception as exc:  # surfaced below
                errors.append(exc)

        threads = [threading.Thread(target=guarded, args=(i,)) for i in range(len(self.devices))]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errors:
            raise errors[0]

    def scan(self, genome: np.ndarray, thresholds, sort: bool = True, chunks: int = 6, merge: bool = True):
        """genome: uint8[n_bases] in host memory (page-locked for full PCIe rate: `_lib.PinnedBuffer` or
        `_lib.RegisteredHost`).  Returns the whole-genome hit array (a fresh array owned by the caller) —
        (motif, pos, strand) order when `sort` — or, with merge=False, the list of per-shard views into
        the pinned outputs (zero-copy; valid until the next scan() or close())."""
        import time

        if genome.size != self.n:
            raise ValueError(f"genome has {genome.size} bases, scanner was planned for {self.n}")
        parts = [None] * len(self.workers)
        t0 = time.perf_counter()

        def work(i):
            w = self.workers[i]
            if w is None:   # empty shard
                parts[i] = np.zeros(0, dtype=_lib.HIT_DTYPE)
                return
            sh = w.shard
            _lib.bind_thread_near(self.devices[i])
            parts[i] = w.ctx.scan_hits_host_shard(genome[sh.start: sh.scan_end], w.enc, w.lib, thresholds, w.hits,
                                                  owned=sh.owned, shard_start=sh.start, chunks=chunks, sort=sort)

        self._run(work)
        t1 = time.perf_counter()
        if not merge:
            out = parts   # views into the workers' pinned buffers: valid until the next scan() / close()
        elif len(parts) == 1:
            out = parts[0].copy()
        elif sort:
            out = concat_sorted_shards(parts, int(self.lens.size))
        else:
            out = np.concatenate(parts)
        self.last_ms = {"scan_wall_ms": (t1 - t0) * 1e3, "merge_wall_ms": (time.perf_counter() - t1) * 1e3}
        return out

    def close(self):
        for w in self.workers:
            if w is not None:
                w.close()
        self.workers = []

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def scan_genome_multi_gpu(genome: np.ndarray, logodds, lens, thresholds, devices=None, cap=1 << 22, sort=True,
                          chunks: int = 6):
    """One-shot form of `MultiGpuScanner`: page-locks `genome` in place for the duration of the call,
    scans it on every visible GPU (or `devices`) and returns globally ordered hits."""
    genome = np.ascontiguousarray(genome, dtype=np.uint8)
    with MultiGpuScanner(genome.size, logodds, lens, devices, cap) as sc:
        with _lib.RegisteredHost(genome):
            return sc.scan(genome, thresholds, sort=sort, chunks=chunks)


# ---- record sets (configs C1 / C2): shard by record index ---------------------------------------------
def plan_record_shards(n_records: int, n_shards: int):
    """[(first, last_exclusive), ...]: contiguous, near-equal ranges of record indices (SURVEY.md §8e:
    "small-sequence workloads shard by sequence index instead").  Records are independent, so there is
    no halo."""
    if n_shards < 1:
        raise ValueError("n_shards must be >= 1")
    edges = [(n_records * i) // n_shards for i in range(n_shards + 1)]
    return list(zip(edges[:-1], edges[1:]))


def map_record_shards(fn, n_records: int, devices=None):
    """Run fn(ctx, first, last) for every record shard on its own GPU (one thread + Context per device)
    and return the results in shard order.  fn must only use the Context it is given."""
    import threading

    ndev = _lib.device_count()
    devices = list(range(ndev)) if devices is None else list(devices)
    if not devices:
        raise RuntimeError("no CUDA device; there is no CPU fallback")
    ranges = plan_record_shards(n_records, len(devices))
    results, errors = [None] * len(devices), []

    def work(i):
        try:
            first, last = ranges[i]
            if last > first:
                with _lib.Context(devices[i]) as ctx:
                    results[i] = fn(ctx, first, last)
        except Exception as exc:
            errors.append(exc)

    threads = [threading.Thread(target=work, args=(i,)) for i in range(len(devices))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return [r for r in results if r is not None]
