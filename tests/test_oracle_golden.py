"""The CPU oracle against the golden vectors written by the unmodified reference
(tests/golden/make_golden.py).  This is what pins oracle/gq_oracle.c."""
import numpy as np
import pytest

import oracle
from conftest import EVAL_CASES, TRAIN_CASES, golden_codebook, load_golden


def rows_to_indices(idx_rows, g):
    """rows [b*l*ncb] -> the reference's index tensor layout (gaussian.py:154-158)."""
    if str(g["format"]) == "bchw":
        b, ncb, h, w = g["indices"].shape
        return idx_rows.reshape(b, h * w, ncb).transpose(0, 2, 1).reshape(b, ncb, h, w)
    return idx_rows.reshape(g["indices"].shape)


@pytest.mark.parametrize("name", EVAL_CASES)
def test_quantize_matches_reference_bit_exact(name):
    g = load_golden(name)
    cb, nlp = golden_codebook(g)
    assert np.array_equal(oracle.normal_log_prob(cb), nlp)
    idx, zhat, best, gap = oracle.quantize(g["mu_rows"], g["std_rows"], g["log_std_rows"], cb, nlp, float(g["beta"]))
    assert idx.dtype == np.int64
    assert np.array_equal(rows_to_indices(idx, g), g["indices"])
    # top-2 gap of the oracle equals the reference's (same fp32 scores)
    assert np.array_equal(gap, g["gap"])
    # zhat rows are codebook rows; dequant(indices) == zhat exactly in the reference
    assert np.array_equal(g["dequant"], g["zhat"])
    rows = oracle.gather(idx, cb)
    assert np.array_equal(rows, zhat)


@pytest.mark.parametrize("name", ["eval_bchw_g16_n1024", "eval_bchw_g4_n4096"])
def test_dequant_layout(name):
    g = load_golden(name)
    cb, _ = golden_codebook(g)
    b, ncb, h, w = g["indices"].shape
    grp = int(g["group"])
    idx_rows = g["indices"].reshape(b, ncb, h * w).transpose(0, 2, 1).reshape(-1)
    rows = oracle.gather(idx_rows, cb)  # [b*l*ncb, g]
    z = rows.reshape(b, h * w, ncb, grp).transpose(0, 1, 3, 2).reshape(b, h * w, ncb * grp)
    z = z.transpose(0, 2, 1).reshape(b, ncb * grp, h, w)
    assert np.array_equal(z, g["dequant"])


def test_gather_rejects_out_of_range():
    cb = np.zeros((8, 4), np.float32)
    with pytest.raises(IndexError):
        oracle.gather(np.array([0, 8]), cb)
    with pytest.raises(IndexError):
        oracle.gather(np.array([-1]), cb)


@pytest.mark.parametrize("name", ["legacy_b1", "legacy_b07"])
def test_legacy_scores_match_reference_closed_form(name):
    g = load_golden(name)
    for contract in (True, False):
        out = oracle.scores_legacy(g["mu"], g["std"], g["noise"], float(g["beta"]), contract)
        # test_extension.py:5-6 evaluates the same closed form with torch's op order; the
        # kernel's order differs (running sum, fma), so agreement is to fp32 rounding
        np.testing.assert_allclose(out, g["result"], rtol=2e-5, atol=2e-3)
        assert np.array_equal(out.argmax(1), g["result"].argmax(1))


@pytest.mark.parametrize("name", TRAIN_CASES)
def test_kl_matches_reference(name):
    g = load_golden(name)
    fmt = str(g["format"])
    for s in range(int(g["steps"])):
        z = g[f"z{s}"]
        if fmt == "bchw":
            b, c2, h, w = z.shape
            z = z.transpose(0, 2, 3, 1).reshape(b, h * w, c2)
        lam, lam_min, lam_max = g[f"lam_before{s}"]
        r = oracle.kl(z, int(g["group"]), 16, 0.5, lam, lam_min, lam_max)
        assert r["kl_loss"] == pytest.approx(float(g[f"kl_loss{s}"]), rel=1e-5)
        assert r["bits_mean"] == pytest.approx(float(g[f"bits_mean{s}"]), rel=1e-5)
        assert r["bits_min"] == pytest.approx(float(g[f"bits_min{s}"]), rel=1e-5, abs=1e-5)
        assert r["bits_max"] == pytest.approx(float(g[f"bits_max{s}"]), rel=1e-5)


def test_scores_torch_small_and_ties():
    rng = np.random.RandomState(0)
    cb = rng.randn(64, 4).astype(np.float32)
    cb[10] = cb[3]  # duplicate codeword: exact tie, first index must win
    nlp = oracle.normal_log_prob(cb)
    mu = cb[[3, 20]] + 0.01
    std = np.full((2, 4), 0.3, np.float32)
    S = oracle.scores_torch(mu, std, np.log(std), cb, nlp)
    idx, _, best, gap = oracle.quantize(mu, std, np.log(std), cb, nlp)
    assert np.array_equal(idx, S.argmax(1))
    assert idx[0] == 3 and S[0, 3] == S[0, 10] and gap[0] == 0.0
    assert np.array_equal(best, S.max(1))


def test_empty_rows():
    cb = np.random.RandomState(1).randn(16, 8).astype(np.float32)
    nlp = oracle.normal_log_prob(cb)
    e = np.zeros((0, 8), np.float32)
    idx, zhat, _, _ = oracle.quantize(e, e, e, cb, nlp)
    assert idx.shape == (0,) and zhat.shape == (0, 8)


# ---- index stream after the quantizer -------------------------------------------
@pytest.mark.parametrize("name", ["usage_n65536", "usage_n1000"])
def test_usage_histogram_and_entropy_match_reference(name):
    """oracle.index_histogram / usage_entropy against fixtures produced with the reference's own
    torch calls (eval.py:127,137-141,152-154; tests/golden/make_golden.py usage_case)."""
    g = load_golden(name)
    N = int(g["N"])
    hist = None
    for bi in range(int(g["batches"])):
        hist = oracle.index_histogram(g[f"idx{bi}"], N, hist)
    assert np.array_equal(hist, g["hist"].astype(np.int64))
    usage, ent = oracle.usage_entropy(hist)
    assert usage == pytest.approx(float(g["usage"]), rel=1e-6)
    assert ent == pytest.approx(float(g["entropy"]), rel=1e-5)


def test_histogram_drops_out_of_range_and_pack_roundtrip():
    idx = np.array([0, 5, 5, 9, -1, 10, 3], dtype=np.int64)
    assert oracle.index_histogram(idx, 10).tolist() == [1, 0, 0, 1, 0, 2, 0, 0, 0, 1]
    for N in (10, 65536, 65537, 1 << 18):
        v = np.random.RandomState(N % 97).randint(0, N, size=1001).astype(np.int64)
        p = oracle.pack_indices(v, N)
        assert p.dtype == (np.uint16 if N <= 65536 else np.uint32)
        assert np.array_equal(oracle.unpack_indices(p), v)
