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


# ASCII bytes of 4 consecutive bases as one little-endian uint32 (the 4-byte block MurmurHash3 consumes),
# indexed by their 2-bit codes with the FIRST base in the two most significant bits (k-mer code order)
_BLOCK4 = np.zeros(256, dtype=np.uint32)
for _x in range(256):
    _BLOCK4[_x] = sum(int(_ASCII[(_x >> (2 * (3 - _j))) & 3]) << (8 * _j) for _j in range(4))


def _murmur3_kmers(codes: np.ndarray, k: int, seed: int = 39) -> np.ndarray:
    """MurmurHash3_x86_32 of the ASCII spelling of every k-mer given as a 2-bit code (first base most
    significant), vectorised; equals mmh3.hash(kmer_string, seed) as an unsigned 32-bit value
    (reference: utils.py:1259-1269).  The 4-byte blocks come from a 256-entry table, the k-mer strings are
    never materialised."""
    codes = np.ascontiguousarray(codes, dtype=np.uint32).reshape(-1)
    c1, c2 = np.uint32(0xCC9E2D51), np.uint32(0x1B873593)
    h = np.full(codes.size, seed, dtype=np.uint32)

    def rotl(x, r):
        return (x << np.uint32(r)) | (x >> np.uint32(32 - r))

    with np.errstate(over="ignore"):
        for i in range(0, k - k % 4, 4):                       # bases i .. i+3 of the k-mer
            kk = _BLOCK4[(codes >> np.uint32(2 * (k - 4 - i))) & np.uint32(255)]
            kk = rotl(kk * c1, 15) * c2
            h = rotl(h ^ kk, 13) * np.uint32(5) + np.uint32(0xE6546B64)
        t = k % 4
        if t:                                                   # tail: bases k-t .. k-1, low bytes first
            kk = np.zeros(codes.size, dtype=np.uint32)
            for j in range(t):
                kk |= _ASCII[(codes >> np.uint32(2 * (t - 1 - j))) & np.uint32(3)].astype(np.uint32) << np.uint32(8 * j)
            h ^= rotl(kk * c1, 15) * c2
        h ^= np.uint32(k)
        h ^= h >> np.uint32(16)
        h *= np.uint32(0x85EBCA6B)
        h ^= h >> np.uint32(13)
        h *= np.uint32(0xC2B2AE35)
        h ^= h >> np.uint32(16)
    return h


def _csr_unique(rows: np.ndarray):
    """Per-row sorted unique values + counts of a 2-D integer array -> (values, counts, per-row sizes)."""
    srt = np.sort(rows, axis=1)
    first = np.ones(srt.shape, dtype=bool)
    first[:, 1:] = srt[:, 1:] != srt[:, :-1]
    flat_first = first.reshape(-1)
    values = srt.reshape(-1)[flat_first]
    starts = np.flatnonzero(flat_first)
    counts = np.diff(np.append(starts, flat_first.size))
    return values, counts, first.sum(axis=1)


def acgt_only(seqs) -> bool:
    return all(not s.strip("ACGT") for s in seqs)


def _scatter_rows(dst_off: np.ndarray, idx: np.ndarray, sizes: np.ndarray, *pairs):
    """Copy the CSR rows of one length group (row r of the group = sequence idx[r], `sizes[r]` entries, rows
    laid end to end in every `src`) to their places in the whole-set arrays: pairs = (dst, src), ..."""
    if idx.size and int(idx[-1]) - int(idx[0]) + 1 == idx.size:       # a run of consecutive sequences: one block
        lo = int(dst_off[idx[0]])
        for dst, src in pairs:
            dst[lo: lo + src.size] = src
        return
    start = np.zeros(sizes.size, dtype=np.int64)
    np.cumsum(sizes[:-1], out=start[1:])
    where = np.repeat(dst_off[idx] - start, sizes) + np.arange(int(sizes.sum()), dtype=np.int64)
    for dst, src in pairs:
        dst[where] = src


def prepare(seqs, k: int) -> dict:
    """k-mer multisets and canonical-hash sets of equal- or mixed-length ACGT sequences, in the
    CSR layout tfbs_similarity_sums expects (sequence order preserved)."""
    S = len(seqs)
    lens = np.array([len(s) for s in seqs], dtype=np.int64)
    nk = np.maximum(lens - k + 1, 0).astype(np.int32)
    groups = []
    wsize = np.zeros(S, dtype=np.int64)
    hsize = np.zeros(S, dtype=np.int64)
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
        w, c, wn = _csr_unique(kmer)
        hv, _, hn = _csr_unique(_murmur3_kmers(np.minimum(kmer, rc), k).reshape(idx.size, m))
        wsize[idx], hsize[idx] = wn, hn
        groups.append((idx, w, c, wn, hv, hn))
    woff = np.zeros(S + 1, dtype=np.int64)
    hoff = np.zeros(S + 1, dtype=np.int64)
    np.cumsum(wsize, out=woff[1:])
    np.cumsum(hsize, out=hoff[1:])
    words = np.zeros(int(woff[-1]), dtype=np.uint32)
    counts = np.zeros(int(woff[-1]), dtype=np.uint16)
    hashes = np.zeros(int(hoff[-1]), dtype=np.uint32)
    for idx, w, c, wn, hv, hn in groups:
        _scatter_rows(woff, idx, wn, (words, w), (counts, c))
        _scatter_rows(hoff, idx, hn, (hashes, hv))
    return {"words": words, "counts": counts, "woff": woff.astype(np.int32), "nk": nk,
            "hash": hashes, "hoff": hoff.astype(np.int32)}


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
