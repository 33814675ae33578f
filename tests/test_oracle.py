"""CPU tests of the oracle itself: internal cross-checks and the reference's golden vectors."""
import numpy as np
import pytest

from discover_tfbs_b200 import dnashape
from oracle import pssm_oracle as po
from oracle import shape_oracle as so


def _rand_seq(rng, n, p_bad=0.05):
    alphabet = np.frombuffer(b"ACGTacgtNRY-", dtype=np.uint8)
    p = np.array([0.2, 0.2, 0.2, 0.2, 0.04, 0.04, 0.04, 0.04] + [p_bad / 4] * 4)
    return alphabet[rng.choice(alphabet.size, size=n, p=p / p.sum())]


@pytest.mark.parametrize("L", [1, 6, 12, 20, 30])
def test_pwm_three_way_crosscheck(L):
    rng = np.random.default_rng(L)
    pssm = po.log_odds(po.normalize(rng.integers(0, 60, size=(L, 4)), 0.5))
    seq = _rand_seq(rng, 700)
    c = po.calculate(seq, pssm)
    n = po.calculate_numpy(seq, pssm)
    assert c.dtype == np.float32 and c.size == seq.size - L + 1
    assert np.array_equal(c, n, equal_nan=True)
    # identity: reverse-strand score at i == forward score of revcomp(seq) at n-L-i
    comp = np.zeros(256, dtype=np.uint8)
    for a, b in zip(b"ACGTacgtNRY-", b"TGCAtgcaNYR-"):
        comp[a] = b
    rc_seq = comp[seq][::-1]
    rev = po.calculate(seq, po.reverse_complement(pssm))
    fwd_on_rc = po.calculate(rc_seq, pssm)
    # same multiset of addends, summed in the opposite order: equal to float32 rounding
    assert np.allclose(rev, fwd_on_rc[::-1], atol=2e-5, rtol=0, equal_nan=True)
    assert np.array_equal(np.isnan(rev), np.isnan(fwd_on_rc[::-1]))


def test_pwm_minus_inf_and_case():
    counts = np.array([[10, 0, 0, 0], [0, 10, 0, 0], [3, 3, 2, 2]], dtype=float)
    pssm = po.log_odds(po.normalize(counts))  # no pseudocount -> -inf entries
    assert np.isneginf(pssm[0, 1])
    s = po.calculate("ACGacTAcN", pssm)
    assert s[0] == po.calculate("acg", pssm)[0]
    assert np.isneginf(s[1]) and np.isnan(s[-1])


def test_search_conventions():
    rng = np.random.default_rng(5)
    pssm = po.log_odds(po.normalize(rng.integers(0, 40, size=(8, 4)), 1.0))
    seq = _rand_seq(rng, 400)
    hits = po.search(seq, pssm, threshold=2.0)
    fwd = po.calculate(seq, pssm)
    rev = po.calculate(seq, po.reverse_complement(pssm))
    exp = int((fwd >= np.float32(2.0)).sum() + (rev >= np.float32(2.0)).sum())
    assert len(hits) == exp
    assert all(p < 0 or fwd[p] == s for p, s in hits)
    assert all(p >= 0 or rev[p + seq.size] == s for p, s in hits)
    pos, mot, strand, score, cnt = po.search_library(seq, pssm[None], [8], [2.0])
    assert cnt == exp and np.all(mot == 0)


def test_normalize_log_odds_definitions():
    counts = np.array([[1, 2, 3, 4]], dtype=float)
    p = po.normalize(counts, {"A": 1, "C": 0, "G": 0, "T": 1})
    assert np.allclose(p, np.array([[2, 2, 3, 5]]) / 12)
    lo = po.log_odds(p, {"A": 0.3, "C": 0.2, "G": 0.2, "T": 0.3})
    assert np.isclose(lo[0, 0], np.log2((2 / 12) / 0.3))


def test_encode_oracle_layout():
    packed, mask = po.encode("ACGTNacgt")
    assert packed[0] == (0 | 1 << 2 | 2 << 4 | 3 << 6 | 0 << 8 | 0 << 10 | 1 << 12 | 2 << 14 | 3 << 16)
    assert mask[0] == (0xFFFFFFFF & ~0x1FF) | (1 << 4)
    assert po.base_counts("ACGTNacgtR-").tolist() == [2, 2, 2, 2, 1, 2]


def test_shape_tokens_match_reference_golden(golden):
    """The oracle's tracks for the golden genome reproduce the token lists the reference parsed."""
    tables, kinds = dnashape.synthetic_tables(tuple(golden["meta"]["shapes"]), seed=golden["meta"]["table_seed"])
    for t, shape in enumerate(golden["meta"]["shapes"]):
        for chrom, seq in golden["genome"].items():
            tr = so.tracks(seq, tables[t:t + 1], kinds[t:t + 1])[0]
            assert so.tokens(tr, int(kinds[t]), len(seq), trailing_empty=True) == golden["shape_tokens"][shape][chrom]


def test_get_dna_shape_restatement_vs_reference(golden):
    for case in golden["get_dna_shape"]:
        chrom = case["loc"].split(":")[0]
        mine = so.get_dna_shape(case["loc"], golden["shape_tokens"][case["shape"]][chrom], case["shape"])
        assert mine == case["out"], case["loc"]


def test_bkg_trim_restatement_vs_reference(golden):
    bg = golden["background"]
    cols = bg["shapes_data_df"]["columns"]
    rows = {r[0]: dict(zip(cols[1:], r[1:])) for r in bg["shapes_data_df"]["rows"]}
    for shape, text in bg["shape_file_text"].items():
        lines = text.strip().split("\n")
        for name, toks in zip(lines[0::2], lines[1::2]):
            mine = so.bkg_trim(toks.split(","), shape)
            for k, v in mine.items():
                assert rows[name[1:]][k] == v


def test_shape_edge_rules():
    tables, kinds = dnashape.synthetic_tables()
    seq = "ACGTACGTAC"
    tr = so.tracks(seq, tables, kinds)
    n = len(seq)
    for t, k in enumerate(kinds):
        if k == 0:
            assert np.isnan(tr[t, [0, 1, n - 2, n - 1]]).all() and not np.isnan(tr[t, 2:n - 2]).any()
        else:
            assert np.isnan(tr[t, [0, n - 2, n - 1]]).all() and not np.isnan(tr[t, 1:n - 2]).any()
    # an N poisons every pentamer that contains it, and every step those pentamers cover
    tr2 = so.tracks("ACGTANGTACGT", tables, kinds)
    assert np.isnan(tr2[2, 3:8]).all() and not np.isnan(tr2[2, 8:10]).any()
    assert np.isnan(tr2[1, 2:8]).all() and not np.isnan(tr2[1, 8])


def test_openmp_forms_equal_the_plain_loops():
    """oracle_pwm_search_mt / oracle_pwm_calculate_mt (used by the config-size GPU tests) give exactly
    the output of the single-threaded loops, in the same order, with more hits than fit one pass."""
    from discover_tfbs_b200 import synth

    seq = synth.genome(7, 0, 150_000)
    lo, lens = synth.pwm_library(7, 12, 6, 30)
    thr = synth.thresholds_for_rate(lo, lens, rate=2e-3, sample=20_000)
    a = po.search_library(seq, lo, lens, thr)
    b = po.search_library(seq, lo, lens, thr, threads=3)
    assert a[4] == b[4] > 1000
    for x, y in zip(a[:4], b[:4]):
        assert np.array_equal(x, y)
    for m in (0, 5):
        mat = lo[m, : lens[m]]
        assert np.array_equal(po.calculate(seq, mat), po.calculate(seq, mat, threads=0), equal_nan=True)


def _tutorial_counts(t):
    counts = np.zeros((5, 4))
    for inst in t["instances"]:
        for j, ch in enumerate(inst):
            counts[j, "ACGT".index(ch)] += 1
    assert counts.tolist() == t["counts_ACGT_by_column"]
    return counts


def test_pwm_oracle_reproduces_the_biopython_tutorial(biopython_tutorial):
    """Known-answer test against the PUBLISHED outputs of the dependency the PWM path replaces (Biopython
    Tutorial, Bio.motifs chapter; provenance in tests/golden/biopython_tutorial.json): counts -> normalize
    (scalar and per-letter pseudocounts) -> log_odds (uniform and explicit background) -> calculate (all 22
    float32 scores to the 8 printed decimals) -> search (both strands, negative reverse positions, order)
    -> max / min / mean / std."""
    t = biopython_tutorial
    counts = _tutorial_counts(t)
    u = t["example_uniform"]
    pwm = po.normalize(counts, u["pseudocounts"])
    assert np.array_equal(np.round(pwm.T, 2), np.array(u["pwm_2dec_rows_ACGT"]))
    assert np.array_equal(np.round(po.log_odds(pwm).T, 2), np.array(u["pssm_2dec_rows_ACGT"]))
    b = t["example_background"]
    pwm = po.normalize(counts, b["pseudocounts"])
    assert np.array_equal(np.round(pwm.T, 2), np.array(b["pwm_2dec_rows_ACGT"]))
    pssm = po.log_odds(pwm, b["background"])
    assert np.array_equal(np.round(pssm.T, 2), np.array(b["pssm_2dec_rows_ACGT"]))
    seq = np.frombuffer(b["test_seq"].encode(), dtype=np.uint8)
    scores = po.calculate(seq, pssm)
    assert scores.dtype == np.float32 and ["%.8f" % v for v in scores] == b["calculate_float32_8dec"]
    assert np.array_equal(po.calculate_numpy(seq, pssm), scores)
    hits = [[int(p), "%5.3f" % s] for p, s in po.search(seq, pssm, b["search_threshold"])]
    assert hits == b["search_position_score_3dec"]
    assert "%4.2f" % pssm.max(axis=1).sum() == b["max_2dec"] and "%4.2f" % pssm.min(axis=1).sum() == b["min_2dec"]
    # PositionSpecificScoringMatrix.mean / .std: p = background * 2^logodds per letter and column
    bg = np.array([b["background"][ch] for ch in "ACGT"])
    p = bg[None, :] * np.exp2(pssm)
    sx, sxx = (p * pssm).sum(axis=1), (p * pssm * pssm).sum(axis=1)
    assert "%0.2f" % sx.sum() == b["mean_2dec"] and "%0.2f" % np.sqrt((sxx - sx * sx).sum()) == b["std_2dec"]
