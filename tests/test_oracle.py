"""The oracle (oracle/) against the golden vectors produced by the reference itself
(tests/golden/gen_golden.py) and against the brute-force enumerator.  CPU only."""
import numpy as np
import pytest

import oracle
from oracle import bruteforce as bf

DMV_CASES = ["dmv_tiny_ragged", "dmv_cfg1", "dmv_cfg1_ragged", "dmv_ties_q025", "dmv_ties_q1", "dmv_ties_zero",
             "dmv_len40"]
# tolerances from BASELINE.json north_star
Z_RTOL, MARG_ATOL = 1e-4, 1e-5


@pytest.mark.parametrize("name", DMV_CASES)
def test_merge_matches_reference(golden, name):
    g = golden(name)
    md, ma = oracle.merge(g["dec"], g["attach"], g["root"])
    assert md.dtype == np.float32 and ma.dtype == np.float32
    np.testing.assert_array_equal(md, g["merged_dec"])
    np.testing.assert_array_equal(ma, g["merged_attach"])


@pytest.mark.parametrize("trim", [False, True])
@pytest.mark.parametrize("name", DMV_CASES)
def test_log_semiring_matches_reference(golden, name, trim):
    g = golden(name)
    Z, gdec, gatt = oracle.dmv_log(g["merged_dec"], g["merged_attach"], g["lengths"], trim=trim)
    np.testing.assert_allclose(Z, g["partition"][:, 0], rtol=Z_RTOL, atol=0)
    np.testing.assert_allclose(gatt, g["grad_attach"], rtol=0, atol=MARG_ATOL)
    np.testing.assert_allclose(gdec, g["grad_dec"], rtol=0, atol=MARG_ATOL * 4)
    # invariants
    L = g["lengths"]
    for b in range(len(L)):
        m = gatt[b].sum(-1)
        np.testing.assert_allclose(m[:, 1:L[b] + 1].sum(0), 1.0, atol=1e-4)
        assert m[:, L[b] + 1:].sum() == 0 and m[L[b] + 1:].sum() == 0 and m[:, 0].sum() == 0


BIG_CASES = ["dmv_cfg2_full", "dmv_n64", "dmv_n128"]


@pytest.mark.parametrize("name", BIG_CASES)
def test_big_reference_fixtures(golden, name):
    """The full cfg2 batch (B = 128, len <= 40) and the cfg3 upper end (n = 64, 128) recorded from the reference.
    Max semiring: bit-exact.  Log semiring: at these lengths chart values are ~ -170 .. -670, one fp32 ulp there is
    1.5e-5 .. 6e-5, and two fp32 sweeps that do not share every rounding differ by more than the north star's 1e-5:
    measured |fp32 oracle - reference| = 1.41e-5 (cfg2), |reference - fp64 oracle| = 1.00e-5 / 1.42e-5 / 2.03e-5.
    What is pinned here: the reference agrees with the oracle's fp64 sweep up to that fp32 noise, and the fp32
    restatement is no farther from the reference than the reference is from the exact result (x 1.5)."""
    g = golden(name)
    md, ma = oracle.merge(g["dec"], g["attach"], g["root"])
    L = g["lengths"]
    Z, gdec, gatt = oracle.dmv_log(md, ma, L, trim=True)
    Z64, gdec64, gatt64 = oracle.dmv_log(md, ma, L, trim=True, f64=True)
    np.testing.assert_allclose(Z, g["partition"][:, 0], rtol=Z_RTOL, atol=0)
    np.testing.assert_allclose(Z64, g["partition"][:, 0], rtol=Z_RTOL, atol=0)
    ref_err = np.abs(g["grad_attach"] - gatt64).max()
    assert ref_err <= 2.5e-5, ref_err
    assert np.abs(gatt - g["grad_attach"]).max() <= 1.5 * max(ref_err, MARG_ATOL)
    assert np.abs(g["grad_dec"] - gdec64).max() <= 5e-5
    best, heads, arcs, vgdec = oracle.dmv_viterbi(md, ma, L, trim=True)
    np.testing.assert_array_equal(best, g["max"][:, 0])
    np.testing.assert_array_equal(heads, g["heads"])
    np.testing.assert_array_equal(vgdec, g["vgrad_dec"])
    B, N = heads.shape
    val = np.full((B, N), -1, dtype=np.int8)
    for b, h, c, v in np.argwhere(arcs > 0):
        val[b, c] = v
    np.testing.assert_array_equal(val, g["arc_valence"])


@pytest.mark.parametrize("trim", [False, True])
@pytest.mark.parametrize("name", DMV_CASES)
def test_viterbi_bit_exact(golden, name, trim):
    g = golden(name)
    best, heads, arcs, gdec = oracle.dmv_viterbi(g["merged_dec"], g["merged_attach"], g["lengths"], trim=trim)
    np.testing.assert_array_equal(best, g["max"][:, 0])  # same fp32 association -> identical bits
    np.testing.assert_array_equal(heads, g["heads"])
    B, N = heads.shape
    val = np.full((B, N), -1, dtype=np.int8)
    for b, h, c, v in np.argwhere(arcs > 0):
        val[b, c] = v
    np.testing.assert_array_equal(val, g["arc_valence"])
    np.testing.assert_array_equal(gdec, g["vgrad_dec"])


def test_zero_scores_kat(golden):
    g = golden("dmv_ties_zero")
    _, heads, arcs, _ = oracle.dmv_viterbi(g["merged_dec"], g["merged_attach"], g["lengths"])
    for b, L in enumerate(g["lengths"]):
        assert heads[b, 1:L + 1].tolist() == list(range(0, L))  # right-branching chain
        assert arcs[b, :, :, oracle.HASCHILD].sum() == 0


def test_f64_truth_brackets_reference(golden):
    """The reference's own fp32 noise vs an fp64 evaluation: sizes the tolerances we use."""
    g = golden("dmv_len40")
    Z64, gd64, ga64 = oracle.dmv_log(g["merged_dec"], g["merged_attach"], g["lengths"], f64=True)
    assert np.abs(g["grad_attach"] - ga64).max() < MARG_ATOL
    np.testing.assert_allclose(g["partition"][:, 0], Z64, rtol=Z_RTOL)


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5])
def test_bruteforce_enumeration(n):
    rng = np.random.default_rng(n)
    dec = np.log(rng.dirichlet([1, 1], (n, 2, 2))).astype(np.float32)
    attach = rng.normal(size=(n, n, 2)).astype(np.float32)
    root = rng.normal(size=(n,)).astype(np.float32)
    md, ma = oracle.merge(dec[None], attach[None], root[None])
    Z, _, gatt = oracle.dmv_log(md, ma, [n])
    best, heads, _, _ = oracle.dmv_viterbi(md, ma, [n])
    logZ, bmax, bheads, marg, ntrees = bf.brute(md[0], ma[0], n)
    assert ntrees == [1, 2, 7, 30, 143][n - 1]
    assert abs(Z[0] - logZ) < 1e-4 * max(1, abs(logZ))
    assert abs(best[0] - bmax) < 1e-4
    assert heads[0, 1:n + 1].tolist() == bheads
    np.testing.assert_allclose(gatt[0].sum(-1), marg, atol=1e-5)


def test_bruteforce_inside_padded_batch():
    """A short sentence inside a longer padded batch (the reference sweeps the padding too)."""
    rng = np.random.default_rng(7)
    n, L = 6, 4
    dec = np.log(rng.dirichlet([1, 1], (n, 2, 2))).astype(np.float32)
    attach = rng.normal(size=(n, n, 2)).astype(np.float32)
    root = rng.normal(size=(n,)).astype(np.float32)
    md, ma = oracle.merge(dec[None], attach[None], root[None])
    Z, _, gatt = oracle.dmv_log(md, ma, [L])
    logZ, _, _, marg, _ = bf.brute(md[0], ma[0], L)
    assert abs(Z[0] - logZ) < 1e-4
    np.testing.assert_allclose(gatt[0].sum(-1), marg, atol=1e-5)


@pytest.mark.parametrize("name", ["deptree_rand", "deptree_mbr", "deptree_ties"])
def test_deptree_matches_reference(golden, name):
    g = golden(name)
    # runtime fill is the global NEGINF (-1e20 after setup_inf), the root mask the class attr (-1e12)
    Z, marg, _ = oracle.deptree(g["arc"], g["lengths"], semiring="log", fill=-1e20, mask_zero=-1e12)
    np.testing.assert_allclose(Z, g["partition"], rtol=1e-4)
    np.testing.assert_allclose(marg, g["marginals"], atol=2e-5)
    mx, ind, heads = oracle.deptree(g["arc"], g["lengths"], semiring="max", fill=-1e20, mask_zero=-1e12)
    np.testing.assert_array_equal(mx, g["max"])
    np.testing.assert_array_equal(ind.astype(np.int8), g["argmax"])


@pytest.mark.parametrize("name", ["align_small", "align_mid"])
def test_alignment_matches_reference(golden, name):
    g = golden(name)
    att = oracle.gather_logit_simple(g["vis_feat"], g["vis_mask"], g["txt_feat"], g["txt_mask"])
    np.testing.assert_allclose(att, g["attmap"], rtol=1e-5, atol=1e-4)
    assert ((att == -1e20) == (g["attmap"] == -1e20)).all()
    red = oracle.gather_logit_reduced(g["vis_feat"], g["vis_mask"], g["txt_feat"], g["txt_mask"], g["txt_marginal"])
    np.testing.assert_allclose(red, g["reduced"], rtol=1e-4, atol=1e-4)


def test_grounding_loss_restatement(golden):
    """oracle.loss_grounding_factor_ce against the reference's loss lines (joint.py:439-491, re-typed in gen_golden.py),
    without and with the POS prior."""
    g = golden("align_loss")
    att = oracle.gather_logit_simple(g["vis_feat"], g["vis_mask"], g["txt_feat"], g["txt_mask"])
    off = np.concatenate([[0], np.cumsum(g["vis_split"])])
    prior = [(g["pos_obj"], off[0], off[1]), (g["pos_rel"], off[1], off[2]), (g["pos_attr"], off[2], off[3])]
    t2v, v2t = oracle.loss_grounding_factor_ce(att, g["txt_marginal"], g["vis_mask"])
    np.testing.assert_allclose([t2v, v2t], [g["txt2vis_plain"], g["vis2txt_plain"]], rtol=1e-5)
    t2v, v2t = oracle.loss_grounding_factor_ce(att, g["txt_marginal"], g["vis_mask"], prior)
    np.testing.assert_allclose([t2v, v2t], [g["txt2vis_prior"], g["vis2txt_prior"]], rtol=1e-5)
