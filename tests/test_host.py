"""CPU tests of the host side: C-ABI surface, bit tricks, generators, host-only feature code."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from discover_tfbs_b200 import _lib, dnashape, similarity, synth
from discover_tfbs_b200 import utils as U

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    text = open(os.path.join(ROOT, "include", "tfbs_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tfbs_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(_lib.LIB_PATH)
    names = _header_functions()
    assert len(names) >= 30
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/tfbs_cuda.h but not exported"
    # and the ctypes binding covers exactly the declared surface
    assert sorted(_lib.exported_symbols()) == names


def test_hit_struct_is_16_bytes():
    assert _lib.HIT_DTYPE.itemsize == 16


def test_no_cpu_fallback_without_gpu():
    """On a box without a GPU every compute entry point must fail loudly."""
    if _lib.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError, match="libtfbs_cuda error"):
        _lib.Context(0)
    with pytest.raises(RuntimeError):
        U.calculate_gc_content("ACGT")


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "discover_tfbs_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "liboracle" not in text and "oracle/" not in text.replace("the oracle/", ""), f


def test_swar_encode_bit_tricks_exhaustive():
    lib = _lib.load()
    codes, valid = C.c_uint32(0), C.c_uint32(0)
    expect = {ord(c): i for i, c in enumerate("ACGT")}
    expect.update({ord(c): i for i, c in enumerate("acgt")})
    rng = np.random.default_rng(0)
    for pos in range(4):
        for b in range(256):
            others = rng.integers(0, 256, size=4)
            others[pos] = b
            w = int(others[0]) | int(others[1]) << 8 | int(others[2]) << 16 | int(others[3]) << 24
            assert lib.tfbs_debug_swar4(w, C.byref(codes), C.byref(valid)) == 0
            for k in range(4):
                byte = int(others[k])
                ok = byte in expect
                assert ((valid.value >> k) & 1) == int(ok), (hex(w), k)
                assert ((codes.value >> (2 * k)) & 3) == (expect[byte] if ok else 0), (hex(w), k)
    assert valid.value < 16 and codes.value < 256


def test_synth_genome_is_position_addressable():
    a = synth.genome(39, 0, 100_000)
    b = synth.genome(39, 40_000, 20_000)
    assert np.array_equal(a[40_000:60_000], b)
    assert set(np.unique(a)) <= set(b"ACGTacgtN")
    frac_n = (a == ord("N")).mean()
    up = np.char.upper(a.view("S1"))
    gc = np.isin(up, [b"G", b"C"]).sum() / np.isin(up, [b"A", b"C", b"G", b"T"]).sum()
    assert 0.38 < gc < 0.44 and frac_n < 0.05
    big = synth.genome(7, 0, 2_000_000)
    assert 0.03 < np.isin(big, list(b"acgt")).mean() < 0.07
    # N-runs: ~0.3 % of the 16 384-base blocks carry one run of 100..10 000 Ns
    blocks = np.arange(200_000, dtype=np.uint64)
    has = (synth._mix(7, 2, blocks) & np.uint64(0xFFFF)) < np.uint64(213)
    assert 0.002 < has.mean() < 0.005
    b = int(np.flatnonzero(has)[0])
    blk = synth.genome(7, b * 16384, 16384)
    n_count = int((blk == ord("N")).sum())
    assert 1 <= n_count <= 10_000
    runs = np.flatnonzero(np.diff(np.concatenate([[0], (blk == ord("N")).astype(int), [0]])))
    assert runs.size == 2  # one contiguous run


def test_join_records_layout():
    buf, starts, lens = synth.join_records(["ACGT", "GG", "TTTAA"])
    assert bytes(buf) == b"ACGT\nGG\nTTTAA" and starts.tolist() == [0, 5, 8] and lens.tolist() == [4, 2, 5]
    arr = np.frombuffer(b"ACGTTGCA", dtype=np.uint8).reshape(2, 4)
    buf2, s2, l2 = synth.join_records(arr)
    assert bytes(buf2) == b"ACGT\nTGCA" and s2.tolist() == [0, 5] and l2.tolist() == [4, 4]


def test_pwm_library_and_thresholds():
    lo, lens = synth.pwm_library(39, 20)
    assert lo.shape[0] == 20 and lens.min() >= 6 and lens.max() <= 30
    assert np.isfinite(lo[0, : lens[0]]).all()
    thr = synth.thresholds_for_rate(lo, lens, rate=1e-3, sample=20_000)
    assert thr.dtype == np.float32 and np.isfinite(thr).all()


def test_synthetic_shape_tables_symmetry():
    tables, kinds = dnashape.synthetic_tables()
    idx = np.arange(1024)
    rc = dnashape._revcomp_index(idx)
    assert np.array_equal(dnashape._revcomp_index(rc), idx)
    for t, k in enumerate(kinds):
        if k == 0:
            assert np.array_equal(tables[t, 0], tables[t, 0][rc])
        else:
            assert np.array_equal(tables[t, 1], tables[t, 0][rc])
    assert np.array_equal(np.round(tables.astype(np.float64), 2).astype(np.float32), tables)


def test_to_reference_float64_equals_printf_roundtrip():
    rng = np.random.default_rng(1)
    a = np.round(rng.uniform(-20, 40, size=5000), 2).astype(np.float32)
    b = np.round(rng.uniform(-20, 40, size=5000), 2).astype(np.float32)
    v = np.concatenate([a, (a + b) * np.float32(0.5)])
    got = dnashape.to_reference_float64(v)
    want = np.array([float("%.2f" % float(x)) for x in v])
    assert np.array_equal(got, want)


def test_primitives_vs_reference_golden(golden):
    for s, out in golden["primitives"]["reverse_complement"]:
        assert U.reverse_complement(s) == out
    with pytest.raises(KeyError):
        U.reverse_complement("acgt")  # the reference's table has upper-case keys only
    for s, k, out in golden["primitives"]["get_kmers"]:
        assert U.get_kmers(s, k) == out


def test_similarity_vs_reference_golden(golden):
    """The scalar restatement (oracle/similarity_oracle.py) against the reference's own outputs; the
    GPU kernel is then checked against it (tests/test_gpu_parity.py)."""
    from oracle import similarity_oracle as similarity

    for kmer, h in golden["similarity"]["minhash_kmer"]:
        assert similarity._canonical_hash(kmer) == h
    assert similarity.murmurhash3_32(b"foo", 0) == (-156908512) & 0xFFFFFFFF
    for a, b, v in golden["similarity"]["jaccard"]:
        assert similarity.jaccard_similarity(a, b) == v
    for a, b, (pas, pps) in golden["similarity"]["pac"]:
        got = similarity.pac(a, b)
        assert abs(got[0] - pas) < 1e-12 and abs(got[1] - pps) < 1e-12


def test_similarity_host_preparation_equals_the_scalar_restatement():
    """The vectorised host preparation the all-pairs kernel is fed with (k-mer multisets and canonical
    MurmurHash3 sets in CSR layout, discover_tfbs_b200/similarity.py:prepare) against the scalar restatement
    pinned to the reference's outputs above: mixed lengths (rows are scattered back into sequence order),
    every k mod 4 (block and tail paths of the hash), duplicated k-mers."""
    from collections import Counter

    from discover_tfbs_b200 import similarity as product
    from oracle import similarity_oracle as oracle

    rng = np.random.default_rng(8)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    for k, lo, hi, n in ((3, 5, 9, 30), (4, 20, 20, 16), (5, 18, 40, 40), (6, 12, 13, 25), (7, 30, 31, 20), (8, 200, 204, 12),
                         (9, 9, 33, 20), (13, 13, 14, 6)):
        seqs = [bytes(acgt[rng.integers(0, 4 if i % 3 else 2, size=int(rng.integers(lo, hi + 1)))]).decode() for i in range(n)]
        p = product.prepare(seqs, k)
        assert p["woff"].dtype == np.int32 and p["hoff"].dtype == np.int32 and p["counts"].dtype == np.uint16
        assert np.array_equal(p["nk"], [len(s) - k + 1 for s in seqs])
        for i, sq in enumerate(seqs):
            want_hash = np.array(sorted({oracle._canonical_hash(sq[j:j + k]) for j in range(len(sq) - k + 1)}), dtype=np.uint32)
            assert np.array_equal(p["hash"][p["hoff"][i]: p["hoff"][i + 1]], want_hash), (k, i)
            wc = Counter(sq[j:j + k] for j in range(len(sq) - k + 1))
            code = {w: sum("ACGT".index(ch) << (2 * (k - 1 - j)) for j, ch in enumerate(w)) for w in wc}
            order = sorted(wc, key=code.get)
            assert np.array_equal(p["words"][p["woff"][i]: p["woff"][i + 1]], [code[w] for w in order]), (k, i)
            assert np.array_equal(p["counts"][p["woff"][i]: p["woff"][i + 1]], [wc[w] for w in order]), (k, i)
    with pytest.raises(ValueError):
        product.prepare(["ACGT"], 6)


def test_pssm_host_functions_reproduce_the_biopython_tutorial(biopython_tutorial):
    """The product's normalize / log_odds / max / min (host side of pssm.PSSM) against the published
    Bio.motifs example (tests/golden/biopython_tutorial.json); calculate / search are the GPU test
    tests/test_gpu_parity.py::test_pssm_calculate_and_search_reproduce_the_biopython_tutorial."""
    from discover_tfbs_b200 import pssm

    t = biopython_tutorial
    counts = np.array(t["counts_ACGT_by_column"], dtype=np.float64)
    u, b = t["example_uniform"], t["example_background"]
    pwm = pssm.normalize(counts, u["pseudocounts"])
    assert np.array_equal(np.round(pwm.T, 2), np.array(u["pwm_2dec_rows_ACGT"]))
    assert np.array_equal(np.round(pssm.log_odds(pwm).T, 2), np.array(u["pssm_2dec_rows_ACGT"]))
    pwm = pssm.normalize(counts, b["pseudocounts"])
    assert np.array_equal(np.round(pwm.T, 2), np.array(b["pwm_2dec_rows_ACGT"]))
    lo = pssm.log_odds(pwm, b["background"])
    assert np.array_equal(np.round(lo.T, 2), np.array(b["pssm_2dec_rows_ACGT"]))
    m = pssm.PSSM(lo)
    assert "%4.2f" % m.max == b["max_2dec"] and "%4.2f" % m.min == b["min_2dec"]
    rc = m.reverse_complement().matrix
    assert np.array_equal(rc, lo[::-1, ::-1])


def test_parse_fasta_and_locations(tmp_path):
    p = tmp_path / "x.fasta"
    p.write_text(">chr1:5-17(-) extra\nACGT\nTT GG\n>b\n\nAC\n")
    assert list(U.parse_fasta(str(p))) == [("chr1:5-17(-) extra", "ACGTTTGG"), ("b", "AC")]
    chrom, s, e, sd = U.parse_locations(["Ca22chr1A:293970-293981(-)", "c:0-12(+)"])
    assert chrom == ["Ca22chr1A", "c"] and s.tolist() == [293970, 0] and e.tolist() == [293981, 12]
    assert sd.tolist() == [-1, 1]
    with pytest.raises(SystemExit):
        next(U.parse_fasta(str(tmp_path / "missing.fasta")))


def test_tokens_to_values_rule():
    v = U._tokens_to_values(["NA", "5.12", "", "-3.40", "x"])
    assert v.tolist() == [0.0, 5.12, 0.0, -3.4, 0.0]


def test_hits_group_plan_is_covered_by_kernel_instantiations():
    """Every (A, B) the planner can return for L = 1..32 must have a scan_segment<A, B> case."""
    src = open(os.path.join(ROOT, "discover_tfbs_b200", "csrc", "scan_hits.cu")).read()
    cases = {(int(a), int(b), int(n)) for a, b, n in re.findall(r"TFBS_CASE\((\d+), (\d+), (\d+)\)", src)}
    narrow_max = int(re.search(r"#define TFBS_NARROW_MAX_GROUPS (\d+)",
                               open(os.path.join(ROOT, "discover_tfbs_b200", "csrc", "common.cuh")).read()).group(1))
    sticky_min = int(re.search(r"#define TFBS_STICKY_MIN_GROUPS (\d+)",
                               open(os.path.join(ROOT, "discover_tfbs_b200", "csrc", "common.cuh")).read()).group(1))
    for L in range(1, 33):
        a, b = _lib.hits_group_plan(L)
        assert 4 * a + 3 * b >= L and 32 * a + 8 * b <= 192 and (a, b, 0) in cases, (L, a, b)
        if a + b <= narrow_max:  # such blocks may be scanned in narrow form (64 motifs, 8-bit fields)
            assert (a, b, 1) in cases, (L, a, b)
        if a + b >= sticky_min:  # ... and these in the sticky narrow form
            assert (a, b, 2) in cases, (L, a, b)
        # optimal: no admissible grouping with fewer loads
        assert all(not (4 * x + 3 * y >= L and 32 * x + 8 * y <= 192 and 1 <= x + y < a + b)
                   for x in range(7) for y in range(12))


def test_read_jaspar_and_threshold_fpr():
    from discover_tfbs_b200 import pssm

    text = """>MA0004.1 Arnt
A  [ 4 19  0  0  0  0 ]
C  [16  0 20  0  0  0 ]
G  [ 0  1  0 20  0 20 ]
T  [ 0  0  0  0 20  0 ]
>MA0006.1 Ahr::Arnt
A  [ 3  0  0  0  0  0 ]
C  [ 8  0 23  0  0  0 ]
G  [ 2 23  0 23  0 24 ]
T  [11  1  1  1 24  0 ]
"""
    motifs = pssm.read_jaspar(text, pseudocounts=0.5)
    assert [n for n, _ in motifs] == ["MA0004.1 Arnt", "MA0006.1 Ahr::Arnt"]
    lo = motifs[0][1]
    assert lo.shape == (6, 4) and np.isclose(lo[0, 1], np.log2((16.5 / 22) / 0.25))
    # the DP distribution agrees with brute-force enumeration of all 4^6 windows
    scores, mass = pssm.score_distribution(lo, precision=2000)
    assert abs(mass.sum() - 1.0) < 1e-12
    idx = np.indices((4,) * 6).reshape(6, -1)
    brute = lo[np.arange(6)[:, None], idx].sum(axis=0)
    for fpr in (0.2, 0.05, 0.01, 0.001):
        thr = pssm.threshold_fpr(lo, fpr, precision=2000)
        achieved = (brute >= thr).mean()
        assert achieved <= fpr * 1.05 + 1e-3 and achieved >= fpr * 0.3 - 1e-3, (fpr, achieved)


def test_hits_block_plan_covers_every_motif_once():
    """Block cutting (host logic of the hits-table builder): narrow blocks only up to the given group
    limit and only while more than 32 motifs are left; every motif lands in exactly one block."""
    rng = np.random.default_rng(5)
    for M in (1, 31, 32, 33, 64, 65, 97, 800):
        lens = rng.integers(1, 33, size=M)
        for limit in (0, 3, 5, 6, 9):
            blocks = _lib.hits_block_plan(lens, limit)
            assert sum(b[3] for b in blocks) == M
            srt = np.sort(lens)[::-1]
            i = 0
            for a, b, narrow, count in blocks:
                assert (a, b) == _lib.hits_group_plan(int(srt[i]))  # sized for the longest motif of the block
                assert count == min(M - i, 64 if narrow else 32)
                assert bool(narrow) == (a + b <= limit and M - i > 32)
                i += count
    assert _lib.hits_block_plan([8] * 100, 5) == [(2, 0, 1, 64), (2, 0, 1, 36)]
    assert _lib.hits_block_plan([30] * 40, 6) == [(5, 4, 0, 32), (5, 4, 0, 8)]
    assert _lib.hits_block_plan([30] * 40, 9) == [(5, 4, 1, 40)]
    assert _lib.hits_block_plan([8] * 100, 0) == [(2, 0, 0, 32)] * 3 + [(2, 0, 0, 4)]
    assert _lib.hits_block_plan([24] * 100, 99) == [(6, 0, 1, 64), (6, 0, 1, 36)]
    assert _lib.hits_block_plan([25] * 100, 6)[0] == (5, 2, 0, 32)


@pytest.mark.parametrize("seed,lmin,lmax,rate,minus_inf", [(1, 3, 32, 3e-3, False), (2, 17, 32, 1e-2, False),
                                                          (3, 1, 12, 1e-2, True), (4, 20, 30, 3e-4, True)])
def test_hits_filter_never_loses_a_hit_in_any_block_form(seed, lmin, lmax, rate, minus_inf):
    """The hits kernel's candidate filter (its own tables and 32-bit field arithmetic, run on the
    host by tfbs_debug_hits_filter) must pass every window the oracle scores >= threshold, in wide,
    narrow and sticky blocks: fields must not carry into their neighbours (wide, narrow) and the
    carry slack of sticky fields must hold."""
    from oracle import pssm_oracle as po
    from discover_tfbs_b200 import synth

    rng = np.random.default_rng(seed)
    M, n = 100, 6000
    logodds, lens = synth.pwm_library(700 + seed, M, lmin=lmin, lmax=lmax, alpha=float(rng.choice([0.1, 0.3, 1.0])))
    if minus_inf:
        logodds = np.where(rng.random(logodds.shape) < 0.03, -np.inf, logodds)
    thr = synth.thresholds_for_rate(np.where(np.isinf(logodds), -20.0, logodds), lens, rate=rate, sample=5000, seed=seed)
    thr[::17] = -np.inf   # everything passes
    thr[5::17] = np.inf   # nothing can
    thr[9::17] = np.nan
    codes = rng.integers(0, 4, size=n).astype(np.uint8)
    codes[1000:1300] = 0  # a poly-A stretch: extreme deficits in every group at once
    seq = np.frombuffer(b"ACGT", dtype=np.uint8)[codes]
    with np.errstate(invalid="ignore"):
        pos, mot, strand, score, cnt = po.search_library(seq, logodds, lens, thr)
    assert cnt > 1000
    passed = {}
    for form in (0, 1, 2):
        cand = _lib.hits_filter_host(logodds, lens, thr, form, codes)
        assert cand[pos, mot, strand].all(), (form, int((cand[pos, mot, strand] == 0).sum()))
        inside = np.arange(n)[:, None] + lens[None, :] <= n  # windows that end inside the sequence
        passed[form] = int(cand[inside].sum())
        assert not cand[:, 5::17].any() and not cand[:, 9::17].any()  # +inf / NaN thresholds: poisoned fields
    # the wide filter is sharp; the 8-bit forms pass more, the sticky one less than the plain one
    finite = np.isfinite(thr)
    n_hits_finite = int(np.isin(mot, np.flatnonzero(finite)).sum())
    always = int(sum(2 * (n - int(L) + 1) for L, t in zip(lens, thr) if t == -np.inf))
    assert passed[0] - always <= 1.25 * n_hits_finite + 50
    assert passed[1] >= passed[2] >= passed[0] >= cnt


def test_hits_cost_model_plan_and_candidate_prediction(monkeypatch):
    """The table builder predicts the candidates a narrow / sticky block passes (exact convolution of
    the quantised deficits under uniform bases) and picks the block form from it.  Run on the host:
    the prediction must match what the filter really passes on a uniform random sequence, flood
    thresholds must keep every block wide, sparse thresholds must use the two-motifs-per-lane forms."""
    from discover_tfbs_b200 import synth

    monkeypatch.delenv("TFBS_HITS_NARROW", raising=False)
    monkeypatch.delenv("TFBS_HITS_STICKY", raising=False)
    rng = np.random.default_rng(11)
    M, n = 200, 30_000
    logodds, lens = synth.pwm_library(711, M, lmin=6, lmax=30)
    codes = rng.integers(0, 4, size=n).astype(np.uint8)
    order = np.argsort(-lens, kind="stable")

    sparse = synth.thresholds_for_rate(logodds, lens, rate=2e-4, gc=0.5, sample=40_000)
    cand, plan, expected = _lib.hits_filter_host(logodds, lens, sparse, 3, codes, with_plan=True)
    assert sum(b[3] for b in plan) == M and any(b[2] for b in plan)
    first = 0
    checked = 0
    for (a, b, form, count), exp in zip(plan, expected):
        motifs = order[first:first + count]
        first += count
        if form == 0:
            assert exp == 0.0
            continue
        observed = cand[:, motifs].sum() / n
        assert abs(observed - exp) <= 0.25 * exp + 0.005, (a, b, form, observed, exp)
        checked += 1
    assert checked >= 1

    flood = np.full(M, -1e30, dtype=np.float32)
    _, plan_flood, _ = _lib.hits_filter_host(logodds, lens, flood, 3, codes[:64], with_plan=True)
    assert all(form == 0 for _, _, form, _ in plan_flood) and len(plan_flood) == (M + 31) // 32

    strict = synth.thresholds_for_rate(logodds, lens, rate=2e-5, gc=0.5, sample=50_000)
    _, plan_strict, _ = _lib.hits_filter_host(logodds, lens, strict, 3, codes[:64], with_plan=True)
    assert sum(1 for b in plan_strict if b[2]) >= sum(1 for b in plan if b[2])

    monkeypatch.setenv("TFBS_HITS_NARROW", "0")  # the environment override reaches the same code path
    _, plan_off, _ = _lib.hits_filter_host(logodds, lens, sparse, 3, codes[:64], with_plan=True)
    assert all(form == 0 for _, _, form, _ in plan_off)


def test_read_meme_and_motif_side_thresholds():
    from discover_tfbs_b200 import pssm

    text = """MEME version 4

ALPHABET= ACGT

MOTIF MA0004.1 Arnt
letter-probability matrix: alength= 4 w= 6 nsites= 20 E= 0
 0.200000  0.800000  0.000000  0.000000
 0.950000  0.000000  0.050000  0.000000
 0.000000  1.000000  0.000000  0.000000
 0.000000  0.000000  1.000000  0.000000
 0.000000  0.000000  0.000000  1.000000
 0.000000  0.000000  1.000000  0.000000
URL http://jaspar.genereg.net/matrix/MA0004.1

MOTIF second
letter-probability matrix: alength= 4 w= 2
 0.25 0.25 0.25 0.25
 0.10 0.20 0.30 0.40
"""
    motifs = pssm.read_meme(text, pseudocounts=0.5)
    assert [n for n, _ in motifs] == ["MA0004.1 Arnt", "second"]
    jaspar = pssm.read_jaspar(">x\nA [ 4 19 0 0 0 0 ]\nC [16 0 20 0 0 0 ]\nG [ 0 1 0 20 0 20 ]\nT [ 0 0 0 0 20 0 ]\n")
    assert np.allclose(motifs[0][1], jaspar[0][1])  # same counts, same log-odds
    assert motifs[1][1].shape == (2, 4)
    lo = motifs[0][1]
    # motif-side distribution against brute force over all 4^6 windows weighted by the motif itself
    idx = np.indices((4,) * 6).reshape(6, -1)
    brute = lo[np.arange(6)[:, None], idx].sum(axis=0)
    p_motif = np.prod(0.25 * 2.0 ** lo[np.arange(6)[:, None], idx], axis=0)
    assert abs(p_motif.sum() - 1.0) < 1e-9
    for fnr in (0.01, 0.1, 0.3):
        thr = pssm.threshold_fnr(lo, fnr, precision=4000)
        missed = p_motif[brute < thr - 1e-9].sum()
        assert missed <= fnr + 2e-3, (fnr, missed)
        assert p_motif[brute < thr + 0.05].sum() >= fnr * 0.5 - 2e-3  # ... and not needlessly low
    bal = pssm.threshold_balanced(lo, 1.0, precision=4000)
    def gap_at(t):
        return abs(p_motif[brute < t].sum() - (brute >= t).mean())

    # scores are discrete and the DP bins them (6 columns x half a step of 0.01): as balanced as an
    # attainable threshold can be, within the binning error
    best = min(gap_at(t) for t in np.unique(brute))
    assert min(gap_at(bal + d) for d in (-0.05, 0.0, 0.05)) <= best + 0.01
    with pytest.raises(ValueError):
        pssm.read_meme("MOTIF p\nletter-probability matrix: alength= 20 w= 1\n" + " ".join(["0.05"] * 20) + "\n")


def test_scan_genome_cli_wiring(tmp_path, monkeypatch):
    """scan_genome.py = motif file -> matrices + thresholds -> candidates.scan_candidates -> BED6 +
    FASTA.  The scan itself is the GPU path (tested in test_gpu_parity); here it is replaced by a
    recorder so that file parsing, threshold selection and the writers are checked without a GPU."""
    from discover_tfbs_b200 import candidates, pssm, scan_genome

    fa = tmp_path / "g.fa"
    fa.write_text(">chr1 test\nACGTACGTAC\nGTCACGTGAA\n>chr2\nTTTTCACGTG\n")
    jaspar = tmp_path / "m.jaspar"
    jaspar.write_text(">MA0004.1 Arnt\nA [ 4 19 0 0 0 0 ]\nC [16 0 20 0 0 0 ]\nG [ 0 1 0 20 0 20 ]\nT [ 0 0 0 0 20 0 ]\n")
    words = tmp_path / "w.txt"
    words.write_text("# curated\nCACGTG\nCANNTG extra\n")
    seen = {}

    def fake_scan(records, matrices, thresholds, names=None, ctx=None, cap=0):
        seen.update(records=records, matrices=matrices, thresholds=np.asarray(thresholds), names=names)
        return [{"chrom": "chr1", "start": 13, "end": 19, "strand": "+", "motif": 0, "name": names[0],
                 "score": 9.5, "sequence": "CACGTG"},
                {"chrom": "chr2", "start": 4, "end": 10, "strand": "-", "motif": 0, "name": names[0],
                 "score": 9.5, "sequence": "CACGTG"}]

    monkeypatch.setattr(candidates, "scan_candidates", fake_scan)
    out = str(tmp_path / "out")
    assert scan_genome.main([str(fa), str(jaspar), out, "--fpr", "0.01"]) == 0
    assert seen["records"] == {"chr1": "ACGTACGTACGTCACGTGAA", "chr2": "TTTTCACGTG"}
    assert seen["names"] == ["MA0004.1 Arnt"] and seen["matrices"][0].shape == (6, 4)
    assert np.isclose(seen["thresholds"][0], pssm.threshold_fpr(seen["matrices"][0], 0.01), atol=1e-6)
    assert open(out + ".bed").read().splitlines() == ["chr1\t13\t19\tMA0004.1 Arnt\t9.5\t+", "chr2\t4\t10\tMA0004.1 Arnt\t9.5\t-"]
    assert open(out + ".fasta").read() == ">chr1:13-19(+)\nCACGTG\n>chr2:4-10(-)\nCACGTG\n"

    assert scan_genome.main([str(fa), str(words), out, "--exact"]) == 0
    assert seen["names"] == ["CACGTG", "CANNTG"] and list(seen["thresholds"]) == [0.0, 0.0]
    m = seen["matrices"][1]
    assert np.array_equal(np.isfinite(m).sum(axis=1), [1, 1, 4, 4, 1, 1])

    assert scan_genome.main([str(fa), str(jaspar), out, "--threshold", "7.5", "--gc", "0.4"]) == 0
    assert list(seen["thresholds"]) == [7.5]
    assert np.isclose(seen["matrices"][0][0, 1], np.log2((16.5 / 22) / 0.2))


# ---- dense scan: fixed-point arithmetic replayed on the host --------------------------------------
def adversarial_dense_cases():
    """Sharp, long motifs: consensus-only columns with large counts and small
    pseudocounts give entries near -20 .. -25 and window scores down to -800, where summing
    fp32-rounded group entries in fp32 drifts past 1e-4.  -> list of (name, log-odds [L][4])."""
    from oracle import pssm_oracle as po

    cases = []
    for L in (30, 32):
        for total, pc in ((5e4, 0.01), (1e5, 0.001), (1e5, 0.01), (5e4, 0.001)):
            rng = np.random.default_rng([L, int(total), int(pc * 1e4)])
            counts = np.zeros((L, 4))
            counts[np.arange(L), rng.integers(0, 4, size=L)] = total
            cases.append((f"L{L}_n{int(total)}_pc{pc}", po.log_odds(po.normalize(counts, pc))))
    return cases


def adversarial_dense_sequence(pssm, n_random=3_000_000, seed=1):
    """n_random uniform bases followed by the all-worst-base window and the consensus window."""
    rng = np.random.default_rng(seed)
    worst = np.argmin(pssm, axis=1)
    best = np.argmax(pssm, axis=1)
    codes = np.concatenate([rng.integers(0, 4, size=n_random), worst, best, worst[::-1], 3 - worst[::-1]])
    return np.frombuffer(b"ACGT", dtype=np.uint8)[codes]


def test_dense_fixed_point_replay_on_adversarial_motifs():
    """The dense kernel's arithmetic — int32 sums of the per-motif scaled k-mer entries, one
    int->float conversion, one exact multiply by 2^-e — replayed in NumPy with the kernel's own tables
    (tfbs_debug_dense_tables) against the double-accumulating oracle: <= 1e-4 on every window."""
    from discover_tfbs_b200 import _lib
    from oracle import pssm_oracle as po

    for name, pssm in adversarial_dense_cases():
        L = pssm.shape[0]
        K, G, e, neg, ent = _lib.dense_tables(pssm)
        assert K == 3 and G == (L + 2) // 3 and neg == -2**31
        # every window sum must fit in int32
        assert np.abs(ent.astype(np.int64)).max(axis=(1, 2)).sum() < 2**31
        seq = adversarial_dense_sequence(pssm, 300_000)
        code = np.searchsorted(np.frombuffer(b"ACGT", dtype=np.uint8), seq)
        nwin = seq.size - L + 1
        for strand, mat in ((0, pssm), (1, po.reverse_complement(pssm))):
            acc = np.zeros(nwin, dtype=np.int64)
            for g in range(G):
                x = np.zeros(nwin, dtype=np.int64)
                for j in range(K):
                    if K * g + j < L:
                        x |= code[K * g + j: K * g + j + nwin].astype(np.int64) << (2 * j)
                acc += ent[g, x, strand]
            assert np.abs(acc).max() < 2**31
            got = acc.astype(np.int32).astype(np.float32) * np.float32(2.0 ** -e)
            ref = po.calculate(seq, mat)
            err = float(np.max(np.abs(got.astype(np.float64) - ref.astype(np.float64))))
            assert err <= 1e-4, (name, strand, err)
            assert float(ref.min()) < -550.0  # the regime the fp32 path failed in


def test_dense_fixed_point_minus_inf_sentinel():
    """A motif with -inf entries: any window with a -inf group sums below neg_thr, every finite window
    stays at or above it, and no sum leaves int32."""
    from discover_tfbs_b200 import _lib
    from oracle import pssm_oracle as po

    rng = np.random.default_rng(8)
    for L in (5, 12, 24, 25, 32):
        counts = rng.integers(0, 1000, size=(L, 4)).astype(float)
        counts[rng.random((L, 4)) < 0.3] = 0.0
        counts[counts.sum(axis=1) == 0, 1] = 7.0
        pssm = po.log_odds(po.normalize(counts))
        assert np.isneginf(pssm).any()
        K, G, e, neg, ent = _lib.dense_tables(pssm)
        e64 = ent.astype(np.int64)
        sentinel = e64.min()
        finite = e64 != sentinel
        rq = sum(np.abs(np.where(finite[g], e64[g], 0)).max() for g in range(G))
        assert neg == -rq and sentinel == -(2 * rq + 1)
        assert G * (2 * rq + 1) + rq < 2**31
        assert e >= 16  # still far finer than 1e-4 / G


# ---- pentamer tables are never fabricated silently ------------------------------------------------
def test_shape_entry_points_require_tables(tmp_path, monkeypatch):
    """VERDICT r01 / ADVICE: GenomeShapeSource, background_shape_table and getDNAShape without tables
    must raise (before touching the GPU) instead of substituting the synthetic stand-in tables."""
    monkeypatch.delenv("TFBS_SHAPE_TABLES", raising=False)
    fa = tmp_path / "g.fasta"
    fa.write_text(">chr1\nACGTACGTACGTACGT\n")
    with pytest.raises(ValueError, match="pentamer tables"):
        U.GenomeShapeSource(str(fa))                      # the exact call of INTEGRATION.md §2 without tables
    with pytest.raises(ValueError, match="pentamer tables"):
        U.background_shape_table(str(fa))
    with pytest.raises(ValueError, match="pentamer tables"):
        dnashape.getDNAShape(str(fa), "MGW")
    with pytest.raises(ValueError, match="pentamer tables"):
        dnashape.resolve_tables()
    with pytest.warns(UserWarning, match="SYNTHETIC"):
        t, k = dnashape.resolve_tables(synthetic=True)
    assert t.shape == (5, 2, 1024) and list(k) == [0, 1, 0, 0, 1]


def test_shape_table_file_roundtrip_half_tables_and_env(tmp_path, monkeypatch):
    tables, kinds = dnashape.synthetic_tables()
    path = dnashape.save_tables(str(tmp_path / "tables.tsv"), tables, kinds)
    t2, k2 = dnashape.load_tables(path)
    assert np.array_equal(t2, tables) and np.array_equal(k2, kinds)
    # a 512-row file (one pentamer of every reverse-complement pair) is completed by symmetry
    lines = open(path).read().split("\n")
    rc = dnashape._revcomp_index(np.arange(1024))
    half = [lines[0]] + [lines[1 + i] for i in range(1024) if i < rc[i]]
    assert len(half) == 513
    (tmp_path / "half.tsv").write_text("\n".join(half) + "\n")
    t3, _ = dnashape.load_tables(str(tmp_path / "half.tsv"))
    assert np.array_equal(t3, tables)
    # a subset of shapes, and the environment variable route
    t4, k4 = dnashape.load_tables(path, ("MGW", "Roll"))
    assert np.array_equal(t4[0], tables[2]) and np.array_equal(t4[1], tables[4]) and list(k4) == [0, 1]
    monkeypatch.setenv("TFBS_SHAPE_TABLES", path)
    t5, _ = dnashape.resolve_tables()
    assert np.array_equal(t5, tables)
    # a file that lacks a pentamer pair raises
    (tmp_path / "bad.tsv").write_text("\n".join(half[:-1]) + "\n")
    with pytest.raises(ValueError, match="missing pentamers"):
        dnashape.load_tables(str(tmp_path / "bad.tsv"))


def test_tables_from_dnashaper_output_of_the_pentamer_fasta(tmp_path):
    """The documented export route: DNAshapeR run on the 1024 five-base records gives back the tables.
    DNAshapeR is played by the oracle (SURVEY.md A.2 semantics, text format of create_bkg_seqs.py:365)."""
    from oracle import shape_oracle as so

    tables, kinds = dnashape.synthetic_tables()
    fa = dnashape.write_pentamer_fasta(str(tmp_path / "pentamers.fa"))
    recs = list(U.parse_fasta(fa))
    assert len(recs) == 1024 and recs[27] == ("AACGT", "AACGT")
    for t, shape in enumerate(dnashape.SHAPES):
        with open(f"{fa}.{shape}", "w") as fh:
            for name, seq in recs:
                tr = so.tracks(seq, tables[t: t + 1], kinds[t: t + 1])[0]
                fh.write(f">{name}\n" + ",".join(so.tokens(tr, int(kinds[t]), 5)) + "\n")
    got, k = dnashape.tables_from_shape_files(fa)
    assert np.array_equal(k, kinds)
    for t in range(5):
        assert np.array_equal(got[t, 0], tables[t, 0])
        if kinds[t]:
            assert np.array_equal(got[t, 1], tables[t, 1])


def test_shim_runs_a_reference_driver_with_the_hot_path_swapped(tmp_path, capsys):
    """discover_tfbs_b200.shim executes the reference's own driver unchanged: `from utils import ...`
    resolves to the reference's utils.py with only the hot-path names replaced."""
    from discover_tfbs_b200 import shim

    (tmp_path / "utils.py").write_text(
        "def build_feature_table(*a, **k):\n    return 'reference'\n"
        "def parse_fasta(f):\n    return 'reference'\n"
        "def random_dna(n):\n    return 'N' * n\n")
    (tmp_path / "driver.py").write_text(
        "import sys\nfrom utils import build_feature_table, parse_fasta, random_dna\n"
        "print(build_feature_table.__module__, parse_fasta.__module__, random_dna(3), sys.argv[1:])\n")
    shim.main([str(tmp_path / "driver.py"), "tec1", "fg.fasta"])
    out = capsys.readouterr().out.strip()
    assert out == "discover_tfbs_b200.utils discover_tfbs_b200.utils NNN ['tec1', 'fg.fasta']"
    import sys
    sys.modules.pop("utils", None)


@pytest.mark.parametrize("form", [0, 1, 2, 3])
def test_sticky_carry_slack_adversarial(form):
    """Sharp 30-column motifs at thresholds set exactly on a window's score: the window is a hit of
    motif 0 (forward), while the reverse-strand field of the same motif and the fields of the motif that
    shares its lane collect the largest deficits the tables can hold.  In the sticky form those fields
    wrap and carry into the hit's field; round 1 under-counted the carries (one hit in 2.6 M lost at a
    1e-6 hit rate on B200, 206 of 400 of these cases on the host).  No form may lose the hit."""
    from oracle import pssm_oracle as po

    rng = np.random.default_rng(1)
    L, M = 30, 64

    def sharp(cons):
        counts = np.full((L, 4), 1.0)
        counts[np.arange(L), cons] += 200
        return po.log_odds(po.normalize(counts, 0.5))

    cons = rng.integers(0, 4, size=L)
    mats = [sharp(rng.integers(0, 4, size=L)) for _ in range(M)]
    mats[0] = sharp(cons)
    anti = np.full((L, 4), 30.0)
    anti[np.arange(L), (cons + 2) % 4] += 200
    anti[np.arange(L), cons] = 0.0
    mats[32] = po.log_odds(po.normalize(anti, 0.01))      # shares lane 0 with motif 0 in a narrow block
    lo, lens = np.stack(mats), np.full(M, L, dtype=np.int32)
    for trial in range(120):
        w = cons.copy()
        k = int(rng.integers(1, 6))
        idx = rng.choice(L, size=k, replace=False)
        w[idx] = rng.integers(0, 4, size=k)
        seq = np.frombuffer(b"ACGT", dtype=np.uint8)[w]
        thr = np.full(M, 1e9, dtype=np.float32)
        thr[0] = po.calculate(seq, mats[0])[0]             # the window scores exactly the threshold: a hit
        thr[32] = po.calculate(seq, mats[32])[0] if trial % 2 else -1e9
        cand = _lib.hits_filter_host(lo, lens, thr, form, w.astype(np.uint8))
        assert cand[0, 0, 0] == 1, (form, trial)
        if trial % 2:
            assert cand[0, 32, 0] == 1, (form, trial)


def test_pair_sum_rounding_guard_is_sufficient():
    """The verify pass of the hits scan sums column PAIRS (half the loads) and keeps the result only when
    it lies farther than guard = 2^-46 * sum_j max_b |entry| from both ends of its float's rounding
    interval (csrc/scan_hits.cu: verify_candidates_kernel; tables built in csrc/scan.cu).  Replayed here
    in NumPy on soft and sharp matrices: whenever the guard accepts, the pair-order float equals the
    float of the oracle's left-to-right double sum; and it accepts almost always."""
    from oracle import pssm_oracle as po

    rng = np.random.default_rng(17)
    accepted = total = 0
    for trial in range(24):
        L = int(rng.integers(5, 33))
        counts = rng.integers(0, 4, size=(L, 4)).astype(float) + 1e-3
        if trial % 2:
            counts[np.arange(L), rng.integers(0, 4, size=L)] += rng.integers(10, 100000)
        mat = po.log_odds(po.normalize(counts, 0.001 if trial % 3 else 0.5))
        n = 200_000
        codes = rng.integers(0, 4, size=n + L)
        # oracle order: left-to-right double sum
        ref = np.zeros(n)
        for j in range(L):
            ref = ref + mat[j, codes[j: j + n]]
        # pair order
        fast = np.zeros(n)
        for j in range(0, L, 2):
            pair = mat[j, codes[j: j + n]] + (mat[j + 1, codes[j + 1: j + 1 + n]] if j + 1 < L else 0.0)
            fast = fast + pair
        f = fast.astype(np.float32)
        bits = f.view(np.uint32)
        ex = bits & np.uint32(0x7F800000)
        ok_form = (ex != 0) & (ex != 0x7F800000) & ((bits & np.uint32(0x007FFFFF)) != 0)
        half_ulp = ex.view(np.float32).astype(np.float64) * 2.0 ** -24
        guard = 2.0 ** -46 * np.abs(mat).max(axis=1).sum()
        accept = ok_form & (half_ulp - np.abs(fast - f.astype(np.float64)) > guard)
        assert np.array_equal(f[accept].view(np.uint32), ref.astype(np.float32)[accept].view(np.uint32)), trial
        accepted += int(accept.sum())
        total += n
    assert accepted > 0.9999 * total


# ---- hits scan, direct path: enumerated 12-mer hash for the shortest motifs ---------------------------
@pytest.mark.parametrize("rate", [1e-4, 1e-3])
def test_direct_path_lookup_equals_the_oracle_hits(rate, monkeypatch):
    """At sparse thresholds the motifs of at most 12 columns bypass the table filter: every L-mer whose
    oracle score reaches the threshold is enumerated into a hash of 12-mers (scan_direct_kernel looks
    positions up in it).  The host replay of plan + lookup must flag EXACTLY the oracle's hits of those
    motifs (windows fully inside the sequence), on both strands, and leave the longer motifs alone."""
    from oracle import pssm_oracle as po

    for v in ("TFBS_HITS_DIRECT", "TFBS_HITS_NARROW", "TFBS_HITS_STICKY"):
        monkeypatch.delenv(v, raising=False)
    logodds, lens = synth.pwm_library(39, 200, 6, 30)
    thr = synth.thresholds_for_rate(logodds, lens, rate=rate, gc=0.41, seed=39, sample=50_000)
    seq = synth.genome(39, 0, 400_000, gc=0.41, n_runs=False, lowercase=False)
    codes = np.searchsorted(np.frombuffer(b"ACGT", dtype=np.uint8), seq).astype(np.uint8)
    cand, info = _lib.direct_lookup_host(logodds, lens, thr, codes)
    dlen = info["max_len"]
    if rate == 1e-4:
        assert dlen in (12, 14) and info["motifs"] >= 32 and info["bitmap_bits"] <= 0.5 * 2**20
    assert info["motifs"] == (int((lens <= dlen).sum()) if info["motifs"] else 0)
    pos, mot, strand, score, cnt = po.search_library(seq, logodds, lens, thr, threads=0)
    want = np.zeros_like(cand)
    want[pos, mot, strand] = 1
    inside = np.arange(seq.size)[:, None] + lens[None, :] <= seq.size       # windows that fit
    exact = lens <= min(dlen, 12)                                           # enumerated L-mers: exactly the hits
    assert np.array_equal(cand[:, exact][inside[:, exact]], want[:, exact][inside[:, exact]])
    longer = (lens > 12) & (lens <= dlen)                                   # 12-column windows: a superset
    if longer.any():
        sub_c, sub_w = cand[:, longer][inside[:, longer]], want[:, longer][inside[:, longer]]
        assert np.all(sub_c >= sub_w) and sub_c.sum() < 20 * max(int(sub_w.sum()), 50)
    assert cand[:, lens > dlen, :].sum() == 0
    if rate == 1e-4:
        assert want[:, exact].sum() > 1000


def test_direct_path_is_not_taken_when_it_cannot_pay(monkeypatch):
    """Flood thresholds, very short motifs (4^(12-L) keys per hit L-mer) and the A/B switches keep every
    motif on the table filter."""
    for v in ("TFBS_HITS_DIRECT", "TFBS_HITS_NARROW", "TFBS_HITS_STICKY"):
        monkeypatch.delenv(v, raising=False)
    logodds, lens = synth.pwm_library(5, 64, 6, 12)
    codes = np.zeros(64, dtype=np.uint8)
    thr = synth.thresholds_for_rate(logodds, lens, rate=1e-4, sample=20_000)
    assert _lib.direct_lookup_host(logodds, lens, thr, codes)[1]["motifs"] == 64
    assert _lib.direct_lookup_host(logodds, lens, np.full(64, -1e30, np.float32), codes)[1]["motifs"] == 0
    assert _lib.direct_lookup_host(logodds, lens, np.full(64, np.nan, np.float32), codes)[1]["keys"] == 0
    tiny, tl = synth.pwm_library(6, 64, 1, 3)
    assert _lib.direct_lookup_host(tiny, tl, np.zeros(64, np.float32), codes)[1]["motifs"] == 0
    monkeypatch.setenv("TFBS_HITS_DIRECT", "0")
    assert _lib.direct_lookup_host(logodds, lens, thr, codes)[1]["motifs"] == 0


def test_direct_path_is_chosen_per_motif(monkeypatch):
    """One loose threshold must not keep the rest of the library off the direct path: a motif whose own keys
    are too many (dense rate, -inf threshold) stays on the table filter, the others still go direct.  The
    union of the two sides covers every oracle hit; the table side (form 4) passes nothing for direct motifs."""
    from oracle import pssm_oracle as po

    for v in ("TFBS_HITS_DIRECT", "TFBS_HITS_NARROW", "TFBS_HITS_STICKY"):
        monkeypatch.delenv(v, raising=False)
    logodds, lens = synth.pwm_library(39, 200, 6, 30)
    thr = synth.thresholds_for_rate(logodds, lens, rate=1e-4, gc=0.41, seed=39, sample=50_000)
    homogeneous = _lib.direct_motifs_host(logodds, lens, thr)
    assert np.array_equal(homogeneous, lens <= 14)            # sparse everywhere: every motif of <= 14 columns
    loose = synth.thresholds_for_rate(logodds, lens, rate=2e-2, gc=0.41, seed=40, sample=20_000)
    short = np.flatnonzero(lens <= 14)
    thr[short[::4]] = loose[short[::4]]
    thr[short[1]] = -np.inf
    thr[short[2]] = np.nan                                     # no keys at all: costs nothing, goes direct
    direct = _lib.direct_motifs_host(logodds, lens, thr)
    expect = homogeneous.copy()
    expect[short[::4]] = False
    expect[short[1]] = False
    assert np.array_equal(direct, expect) and direct.sum() >= 32

    seq = synth.genome(39, 0, 60_000, gc=0.41, n_runs=False, lowercase=False)
    codes = np.searchsorted(np.frombuffer(b"ACGT", dtype=np.uint8), seq).astype(np.uint8)
    with np.errstate(invalid="ignore"):
        pos, mot, strand, score, cnt = po.search_library(seq, logodds, lens, thr, threads=0)
    cand, info = _lib.direct_lookup_host(logodds, lens, thr, codes)
    assert info["motifs"] == int(direct.sum()) and info["max_len"] == int(lens[direct].max())
    assert cand[:, ~direct].sum() == 0
    on_direct = direct[mot]
    assert cand[pos[on_direct], mot[on_direct], strand[on_direct]].all()
    filt, plan, _ = _lib.hits_filter_host(logodds, lens, thr, 4, codes, with_plan=True)
    assert sum(b[3] for b in plan) == 200 - int(direct.sum())
    assert filt[:, direct].sum() == 0
    assert filt[pos[~on_direct], mot[~on_direct], strand[~on_direct]].all()
    # the short motifs left on the tables close the length-sorted table order: the last block is sized for them
    left_short = int((~direct & (lens <= 14)).sum())
    assert left_short >= 10 and plan[-1][0] + plan[-1][1] <= 4


def test_hits_plan_is_the_same_on_any_number_of_host_threads(monkeypatch):
    """The complete plan of a threshold vector (tfbs_debug_hits_plan_checksum runs the scan's own
    plan_hits_tables: direct-path choice, speculative block planning split by lane ranges over the host's
    cores, the cut, every filter table) must not depend on how many threads built it, and its blocks must be
    the ones the sequential route (choose_form, block after block: tfbs_debug_hits_filter form 4) cuts."""
    for v in ("TFBS_HITS_DIRECT", "TFBS_HITS_NARROW", "TFBS_HITS_STICKY", "TFBS_PLAN_THREADS"):
        monkeypatch.delenv(v, raising=False)
    rng = np.random.default_rng(77)
    codes = np.zeros(40, dtype=np.uint8)
    seen_forms = set()
    for case in range(6):
        M = (33, 64, 200, 300, 130, 97)[case]
        logodds, lens = synth.pwm_library(100 + case, M, lmin=(6, 1, 6, 10, 25, 4)[case], lmax=(30, 14, 30, 32, 32, 20)[case],
                                          alpha=(0.3, 0.3, 0.1, 1.0, 0.3, 0.1)[case])
        if case == 3:
            logodds = np.where(rng.random(logodds.shape) < 0.02, -np.inf, logodds)
        fin = np.where(np.isinf(logodds), -20.0, logodds)
        thr = synth.thresholds_for_rate(fin, lens, rate=(1e-4, 1e-4, 1e-3, 1e-4, 3e-3, 1e-2)[case], gc=0.5, seed=case, sample=10_000)
        if case in (2, 5):
            thr[3::11], thr[5::11], thr[7::11] = -np.inf, np.inf, np.nan
        monkeypatch.delenv("TFBS_PLAN_THREADS", raising=False)
        info, plan = _lib.hits_plan_checksum_host(logodds, lens, thr)
        for threads in ("1", "3"):
            monkeypatch.setenv("TFBS_PLAN_THREADS", threads)
            assert _lib.hits_plan_checksum_host(logodds, lens, thr) == (info, plan), (case, threads)
        _, seq_plan, _ = _lib.hits_filter_host(logodds, lens, thr, 4, codes, with_plan=True)
        assert plan == seq_plan, case
        assert sum(b[3] for b in plan) + info["direct_motifs"] == M
        assert info["direct_motifs"] == int(_lib.direct_motifs_host(logodds, lens, thr).sum())
        seen_forms |= {b[2] for b in plan}
    assert seen_forms == {0, 1, 2}       # wide, narrow and sticky blocks all occurred


@pytest.mark.parametrize("rate", [1e-4, 1e-3])
def test_bench_library_plan_never_loses_a_hit_on_the_host(rate, monkeypatch):
    """The bench workload's library (800 PWMs, L 6-30, seed 39) at the thresholds of the GPU rate sweeps: the plan a scan makes
    (direct path for the short motifs at 1e-4, 8-bit narrow blocks for most of the rest) replayed on the host
    over 30 kb — direct keys == oracle hits for the motifs of <= 12 columns, table candidates (form 4) cover
    every oracle hit of the motifs left on the tables, and the parallel planner's blocks are the ones the
    sequential route cuts."""
    from oracle import pssm_oracle as po

    for v in ("TFBS_HITS_DIRECT", "TFBS_HITS_NARROW", "TFBS_HITS_STICKY", "TFBS_PLAN_THREADS"):
        monkeypatch.delenv(v, raising=False)
    logodds, lens = synth.pwm_library(39, 800, 6, 30)
    # thresholds of that library as the rate sweeps on the GPU used them (profiles/hits_rate_quick.py); computing
    # them here (synth.thresholds_for_rate, 100 000 samples x 800 motifs) would take half a minute
    thr = np.load(os.path.join(ROOT, "profiles", "hits_rate_thresholds.npz"))[f"{rate:g}"]
    assert thr.shape == (800,) and thr.dtype == np.float32
    seq = synth.genome(39, 0, 30_000, gc=0.41, n_runs=False, lowercase=False)
    codes = np.searchsorted(np.frombuffer(b"ACGT", dtype=np.uint8), seq).astype(np.uint8)
    pos, mot, strand, score, cnt = po.search_library(seq, logodds, lens, thr, threads=0)
    info, plan = _lib.hits_plan_checksum_host(logodds, lens, thr)
    direct = _lib.direct_motifs_host(logodds, lens, thr)
    filt, seq_plan, _ = _lib.hits_filter_host(logodds, lens, thr, 4, codes, with_plan=True)
    assert plan == seq_plan and sum(b[3] for b in plan) + int(direct.sum()) == 800
    assert any(b[2] == 1 for b in plan)                               # narrow blocks in the plan
    if rate == 1e-4:
        assert direct.sum() == int((lens <= 14).sum()) == info["direct_motifs"] and info["direct_keys"] > 10**6
    on_direct = direct[mot]
    assert filt[pos[~on_direct], mot[~on_direct], strand[~on_direct]].all() and filt[:, direct].sum() == 0
    cand, _ = _lib.direct_lookup_host(logodds, lens, thr, codes)
    want = np.zeros_like(cand)
    want[pos, mot, strand] = 1
    inside = np.arange(seq.size)[:, None] + lens[None, :] <= seq.size
    exact = direct & (lens <= 12)
    assert np.array_equal(cand[:, exact][inside[:, exact]], want[:, exact][inside[:, exact]])
    longer = direct & (lens > 12)
    assert np.all(cand[:, longer][inside[:, longer]] >= want[:, longer][inside[:, longer]])
    assert cnt > 2000


def test_host_fuzz_campaign_cases(monkeypatch):
    """A few cases of tests/fuzz_hits_host.py (random libraries with -inf entries, thresholds set exactly on
    window scores, +-inf / NaN / loose thresholds, homopolymer runs): the plan a scan makes never loses an
    oracle hit.  The full campaign (hundreds of cases) is run by hand: python tests/fuzz_hits_host.py."""
    import importlib.util

    for v in ("TFBS_HITS_DIRECT", "TFBS_HITS_NARROW", "TFBS_HITS_STICKY", "TFBS_PLAN_THREADS"):
        monkeypatch.delenv(v, raising=False)
    spec = importlib.util.spec_from_file_location("fuzz_hits_host", os.path.join(ROOT, "tests", "fuzz_hits_host.py"))
    fuzz = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(fuzz)
    rng = np.random.default_rng(2026)
    hits = sum(fuzz.one_case(rng, case)[0] for case in range(6))
    assert hits > 10_000


def test_hits_table_capacity_covers_any_per_threshold_plan():
    """The device tables are sized at upload for 32-motif blocks of the whole length-sorted library; a
    per-threshold plan may start block i at a SHORTER motif, whose table can be the larger one (24 columns:
    6 x 32 KB; 25 columns: 5 x 32 + 2 x 8 KB).  The sizing takes the running maximum over all shorter lengths."""
    words = {}
    for L in range(1, 33):
        a, b = _lib.hits_group_plan(L)
        words[L] = a * 8192 + b * 2048
    assert words[24] > words[25]                               # the non-monotone step the envelope exists for
    assert all(words[L] <= words[L + 1] for L in range(1, 16)) # monotone over the lengths the direct path can remove


# ---- bench.py contract (CPU side) ---------------------------------------------------------------------
def test_bench_reference_arm_prints_one_json_line_with_all_host_cores():
    """`bench.py --impl reference`: ONE JSON line on stdout with the contract's keys; the CPU arm takes its
    thread count from sched_getaffinity even when OMP_NUM_THREADS=1 (what torch.distributed.run exports)."""
    import json
    import subprocess
    import sys

    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--motifs", "24"], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
    assert "10000000 bp" in d["cpu_baseline"]["sample"]                 # >= 10 Mb per step
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_bench_gpu_arm_fails_loudly_without_a_gpu():
    import subprocess
    import sys

    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1", "--motifs", "8",
                          "--genome-mb", "1", "--no-extras"], capture_output=True, text=True, timeout=600)
    assert out.returncode != 0 and out.stdout.strip() == ""
    assert "no CUDA device" in out.stderr or "libtfbs_cuda" in out.stderr
