"""Seeded synthetic inputs for tests and bench.py (there is no network for real genomes).

Everything is a pure function of (seed, position) so that any shard on any GPU — and the CPU
oracle — can regenerate exactly the same bases.  The base generator is mirrored bit for bit by
`synth_kernel` in csrc/encode.cu (tfbs_synth_ascii).  Seed 39 echoes the reference's
`random_state=39` (create_bkg_seqs.py:346, cross_validate_model.py).
"""
import numpy as np

_U64 = np.uint64
_MASK53 = (1 << 53)


def splitmix64(z: np.ndarray) -> np.ndarray:
    z = np.asarray(z, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = z + _U64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> _U64(30))) * _U64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> _U64(27))) * _U64(0x94D049BB133111EB)
        return z ^ (z >> _U64(31))


def _mix(seed: int, tag: int, i: np.ndarray) -> np.ndarray:
    s = splitmix64(np.array([(seed + tag) & 0xFFFFFFFFFFFFFFFF], dtype=np.uint64))[0]
    return splitmix64(np.asarray(i, dtype=np.uint64) ^ s)


def genome(seed: int, start: int, n: int, gc: float = 0.41, n_runs: bool = True,
           lowercase: bool = True) -> np.ndarray:
    """ASCII bases for positions [start, start+n) of the synthetic genome `seed` (uint8 array).

    ACGT with the requested GC fraction; optionally ~0.1 % of bases inside N-runs of 100-10 000
    (one candidate run per 16 384-base block) and ~5 % lower-case (64-base blocks)."""
    out = np.empty(n, dtype=np.uint8)
    two53 = float(_MASK53)
    t_c = np.uint64(int(gc / 2.0 * two53))
    t_g = np.uint64(int(gc * two53))
    t_a = np.uint64(int((gc + (1.0 - gc) / 2.0) * two53))
    step = 1 << 22
    for lo in range(0, n, step):
        hi = min(n, lo + step)
        i = np.arange(start + lo, start + hi, dtype=np.uint64)
        r = _mix(seed, 1, i) >> _U64(11)
        ch = np.where(r < t_c, ord("C"), np.where(r < t_g, ord("G"), np.where(r < t_a, ord("A"), ord("T")))).astype(np.uint8)
        if lowercase:
            rl = _mix(seed, 3, i >> _U64(6))
            ch = np.where((rl & _U64(0xFFFF)) < _U64(3277), ch | np.uint8(0x20), ch)
        if n_runs:
            rn = _mix(seed, 2, i >> _U64(14))
            has = (rn & _U64(0xFFFF)) < _U64(213)
            off = (i & _U64(16383)).astype(np.int64)
            s = ((rn >> _U64(16)) & _U64(16383)).astype(np.int64)
            ln = 100 + ((rn >> _U64(32)) % _U64(9901)).astype(np.int64)
            ch = np.where(has & (off >= s) & (off < s + ln), np.uint8(ord("N")), ch)
        out[lo:hi] = ch
    return out


# ---- motif libraries -------------------------------------------------------------------------
def pwm_library(seed: int, M: int, lmin: int = 6, lmax: int = 30, alpha: float = 0.3,
                pseudocount: float = 0.5, nsites: int = 200):
    """M count matrices with Dirichlet(alpha) columns -> log-odds (uniform background).

    Returns (logodds float64[M][Lmax][4], lens int32[M]).  The normalise/log2 arithmetic is the
    Biopython definition: p = (count + pc) / colsum, log2(p / 0.25)."""
    rng = np.random.default_rng(seed)
    lens = rng.integers(lmin, lmax + 1, size=M).astype(np.int32)
    L = int(lens.max())
    out = np.zeros((M, L, 4), dtype=np.float64)
    for m in range(M):
        probs = rng.dirichlet(np.full(4, alpha), size=int(lens[m]))
        counts = np.stack([rng.multinomial(nsites, p) for p in probs]).astype(np.float64)
        p = counts + pseudocount
        p = p / p.sum(axis=1, keepdims=True)
        out[m, : lens[m]] = np.log2(p / 0.25)
    return out, lens


def consensus_pwm(seed: int, L: int = 12, nsites: int = 200, mutation: float = 0.15,
                  pseudocount: float = 0.5):
    """Config C1: count matrix from `nsites` sampled L-mers of a random consensus with
    per-position mutation rate `mutation`.  Returns (logodds[L][4], consensus bytes)."""
    rng = np.random.default_rng(seed)
    cons = rng.integers(0, 4, size=L)
    sites = np.where(rng.random((nsites, L)) < mutation, rng.integers(0, 4, size=(nsites, L)), cons[None, :])
    counts = np.stack([(sites == b).sum(axis=0) for b in range(4)], axis=1).astype(np.float64)
    p = counts + pseudocount
    p = p / p.sum(axis=1, keepdims=True)
    return np.log2(p / 0.25), np.frombuffer(b"ACGT", dtype=np.uint8)[cons]


def thresholds_for_rate(logodds, lens, rate: float = 1e-4, gc: float = 0.41, seed: int = 39,
                        sample: int = 200_000) -> np.ndarray:
    """Per-motif score thresholds with an expected hit rate of about `rate` per position per
    strand on a background of the given GC content (Monte-Carlo quantile, seeded)."""
    rng = np.random.default_rng(seed)
    M = logodds.shape[0]
    pb = np.array([(1 - gc) / 2, gc / 2, gc / 2, (1 - gc) / 2])
    thr = np.empty(M, dtype=np.float32)
    for m in range(M):
        L = int(lens[m])
        codes = rng.choice(4, size=(sample, L), p=pb)
        sc = logodds[m, np.arange(L)[None, :], codes].sum(axis=1)
        thr[m] = np.float32(np.quantile(sc, 1.0 - rate))
    return thr


# ---- sequence sets (configs C1 / C2) ------------------------------------------------------------
def peak_set(seed: int, n_fg: int = 5000, length: int = 200, bkg_per_fg: int = 10,
             motif: np.ndarray = None):
    """Foreground peaks (uniform ACGT, optional motif implanted once at a random offset/strand)
    and `bkg_per_fg` mononucleotide-shuffled, 2+2-padded backgrounds per peak (the reference pads
    background sequences with 2 random nt per side: create_bkg_seqs.py:245,264).
    Returns (fg uint8[n_fg][length], bkg uint8[n_fg*bkg_per_fg][length+4])."""
    rng = np.random.default_rng(seed)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    fg = acgt[rng.integers(0, 4, size=(n_fg, length))]
    if motif is not None:
        L = motif.size
        comp = {65: 84, 67: 71, 71: 67, 84: 65}
        rc = np.array([comp[int(c)] for c in motif[::-1]], dtype=np.uint8)
        offs = rng.integers(0, length - L + 1, size=n_fg)
        strands = rng.integers(0, 2, size=n_fg)
        for i in range(n_fg):
            fg[i, offs[i]: offs[i] + L] = motif if strands[i] == 0 else rc
    core = np.repeat(fg, bkg_per_fg, axis=0)
    core = rng.permuted(core, axis=1)
    pad = acgt[rng.integers(0, 4, size=(core.shape[0], 4))]
    bkg = np.concatenate([pad[:, :2], core, pad[:, 2:]], axis=1)
    return np.ascontiguousarray(fg), np.ascontiguousarray(bkg)


def join_records(records, sep: int = ord("\n")):
    """Concatenate equal- or ragged-length records with one invalid separator byte between them.
    Returns (buffer uint8, starts int64, lens int64)."""
    if isinstance(records, np.ndarray) and records.ndim == 2:
        S, L = records.shape
        buf = np.full((S, L + 1), sep, dtype=np.uint8)
        buf[:, :L] = records
        starts = np.arange(S, dtype=np.int64) * (L + 1)
        lens = np.full(S, L, dtype=np.int64)
        return buf.reshape(-1)[:-1].copy() if S else np.zeros(0, np.uint8), starts, lens
    recs = [np.frombuffer(r.encode("ascii") if isinstance(r, str) else bytes(r), dtype=np.uint8) for r in records]
    lens = np.array([r.size for r in recs], dtype=np.int64)
    starts = np.zeros(len(recs), dtype=np.int64)
    if len(recs):
        starts[1:] = np.cumsum(lens[:-1] + 1)
    total = int(lens.sum() + max(len(recs) - 1, 0))
    buf = np.full(total, sep, dtype=np.uint8)
    for r, s in zip(recs, starts):
        buf[s: s + r.size] = r
    return buf, starts, lens
