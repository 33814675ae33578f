"""MinHash-Jaccard and Poisson (PAC) similarity columns of the feature table.

Produced by the reference at utils.py:892-949 (pac), :1043-1058 (_minhash_similarity, _pac) and
:1243-1269 (jaccard_similarity, _minhash_kmer) — 98 % of its own feature-build time, but not part
of the path BASELINE.json north_star names, hence SURVEY.md §8(f) "next" rank 1.

`similarity_columns_gpu` evaluates all (candidate, true) pairs on the GPU (tfbs_similarity_sums,
csrc/similarity.cu); the k-mer multisets and canonical MurmurHash3_x86_32 (seed 39) sets it needs
are prepared here with vectorised NumPy.  There is no scalar host path: sequences with characters
other than upper-case ACGT raise (the reference itself raises KeyError on lower case,
utils.py:104).  The scalar restatement that serves as the test oracle is test infrastructure, not product.
"""
from math import ceil, log

import numpy as np

def _kmer_len(n: int) -> int:
    # utils.py:906 and :1247 — k = ceil(log4(n * 0.99 / 0.01))
    return ceil(log((n * (1 - 0.01)) / 0.01, 4))


# ---- GPU path: host preparation (vectorised) + tfbs_similarity_sums ---------------------------------
_CODE = np.full(256, 255, dtype=np.uint8)
for _i, _b in enumerate(b"ACGT"):
    _CODE[_b] = _i
_ASCII = np.frombuffer(b"ACGT", dtype=np.uint8)


def _murmur3_rows(keys: np.ndarray, seed: int = 39) -> np.ndarray:
    """MurmurHash3_x86_32 of every row of a uint8 matrix [N, k] (vectorised)."""
    n, k = keys.shape
    c1, c2 = np.uint32(0xCC9E2D51), np.uint32(0x1B873593)
    h = np.full(n, seed, dtype=np.uint32)
    with np.errstate(over="ignore"):
        def rotl(x, r):
            return (x << np.uint32(r)) | (x >> np.uint32(32 - r))
        for i in range(0, k - k % 4, 4):
            blk = keys[:, i:i + 4].astype(np.uint32)
            kk = blk[:, 0] | (blk[:, 1] << np.uint32(8)) | (blk[:, 2] << np.uint32(16)) | (blk[:, 3] << np.uint32(24))
            kk = rotl(kk * c1, 15) * c2
            h = rotl(h ^ kk, 13) * np.uint32(5) + np.uint32(0xE6546B64)
        t = k % 4
        if t:
            tail = keys[:, k - t:].astype(np.uint32)
            kk = np.zeros(n, dtype=np.uint32)
            if t >= 3:
                kk ^= tail[:, 2] << np.uint32(16)
            if t >= 2:
                kk ^= tail[:, 1] << np.uint32(8)
            kk ^= tail[:, 0]
            h ^= rotl(kk * c1, 15) * c2
        h ^= np.uint32(k)
        h ^= h >> np.uint32(16)
        h *= np.uint32(0x85EBCA6B)
        h ^= h >> np.uint32(13)
        h *= np.uint32(0xC2B2AE35)
        h ^= h >> np.uint32(16)
    return h


def _csr_unique(rows: np.ndarray):
    """Per-row sorted unique values + counts of a 2-D integer array -> (values, counts, offsets)."""
    srt = np.sort(rows, axis=1)
    first = np.ones(srt.shape, dtype=bool)
    first[:, 1:] = srt[:, 1:] != srt[:, :-1]
    flat_first = first.reshape(-1)
    values = srt.reshape(-1)[flat_first]
    starts = np.flatnonzero(flat_first)
    counts = np.diff(np.append(starts, flat_first.size))
    per_row = first.sum(axis=1)
    off = np.zeros(rows.shape[0] + 1, dtype=np.int64)
    np.cumsum(per_row, out=off[1:])
    return values, counts, off


def acgt_only(seqs) -> bool:
    return all(not s.strip("ACGT") for s in seqs)


def prepare(seqs, k: int) -> dict:
    """k-mer multisets and canonical-hash sets of equal- or mixed-length ACGT sequences, in the
    CSR layout tfbs_similarity_sums expects (sequence order preserved)."""
    S = len(seqs)
    lens = np.array([len(s) for s in seqs], dtype=np.int64)
    out_words, out_counts, out_hash = [None] * S, [None] * S, [None] * S
    nk = np.maximum(lens - k + 1, 0).astype(np.int32)
    for n in np.unique(lens):
        idx = np.flatnonzero(lens == n)
        m = int(n) - k + 1
        if m <= 0:
            raise ValueError(f"sequence of length {n} is shorter than k = {k}")
        mat = np.frombuffer("".join(seqs[i] for i in idx).encode("ascii"), dtype=np.uint8).reshape(idx.size, int(n))
        codes = _CODE[mat].astype(np.uint32)
        kmer = np.zeros((idx.size, m), dtype=np.uint32)
        rc = np.zeros((idx.size, m), dtype=np.uint32)
        for j in range(k):
            kmer |= codes[:, j:j + m] << np.uint32(2 * (k - 1 - j))
            rc |= (np.uint32(3) - codes[:, k - 1 - j: k - 1 - j + m]) << np.uint32(2 * (k - 1 - j))
        w, c, woff = _csr_unique(kmer)
        canon = np.minimum(kmer, rc).reshape(-1)
        keys = np.empty((canon.size, k), dtype=np.uint8)
        for j in range(k):
            keys[:, j] = _ASCII[(canon >> np.uint32(2 * (k - 1 - j))) & np.uint32(3)]
        hv, _, hoff = _csr_unique(_murmur3_rows(keys).reshape(idx.size, m))
        for r, i in enumerate(idx):
            out_words[i] = w[woff[r]:woff[r + 1]]
            out_counts[i] = c[woff[r]:woff[r + 1]]
            out_hash[i] = hv[hoff[r]:hoff[r + 1]]
    woff = np.zeros(S + 1, dtype=np.int32)
    hoff = np.zeros(S + 1, dtype=np.int32)
    np.cumsum([a.size for a in out_words], out=woff[1:])
    np.cumsum([a.size for a in out_hash], out=hoff[1:])
    return {"words": np.concatenate(out_words).astype(np.uint32) if S else np.zeros(0, np.uint32),
            "counts": np.concatenate(out_counts).astype(np.uint16) if S else np.zeros(0, np.uint16),
            "woff": woff, "nk": nk,
            "hash": np.concatenate(out_hash).astype(np.uint32) if S else np.zeros(0, np.uint32), "hoff": hoff}


def similarity_columns_gpu(ctx, seqs, true_seqs):
    """MinHash / PAS / PPS columns for `seqs` against the set `true_seqs` on the GPU.
    k follows the TRUE sequence of every pair (utils.py:906,1247), so the true set is grouped by k."""
    true_list = list(true_seqs)
    U = len(true_list)
    total = np.zeros((len(seqs), 3), dtype=np.float64)
    by_k = {}
    for t in true_list:
        by_k.setdefault(_kmer_len(len(t)), []).append(t)
    for k, group in by_k.items():
        total += ctx.similarity_sums(prepare(list(seqs), k), prepare(group, k))
    total /= U
    return total[:, 0], total[:, 1], total[:, 2]
