"""GPU parity tests of the feature selectors next to ``:vst`` (``find_variable_features(...; method = :dispersion |
:meanvarplot | :saunders)``, variablefeatures.jl:52-103,135-155): the data sweeps run on the device (relative counts, order-exact
moments, UMI totals — each already bit-exact against the oracle), the gene-length arithmetic on the host (pinned on the CPU in
tests/test_oracle_golden.py::test_variable_feature_selectors_host_arithmetic)."""
import numpy as np
import pytest

from conftest import planted_counts

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("method", ["dispersion", "meanvarplot", "saunders"])
def test_selectors_against_oracle(sv, orc, method):
    X = planted_counts(600, 1500, 6, seed=5)
    ref = orc.variable_feature_metric(X, method)
    got = sv.api._variable_feature_metric(X, method, {})
    np.testing.assert_allclose(got, ref, rtol=1e-12, atol=1e-14)
    nsel = 50
    sel = sv.find_variable_features(X, nsel, method=method)
    order = np.argsort(-ref, kind="stable")
    if ref[order[nsel - 1]] - ref[order[nsel]] > 1e-9:                 # no tie at the cut
        assert set(sel.tolist()) == set(order[:nsel].tolist())
    assert np.all(np.diff(got[sel]) <= 0)                               # decreasing metric (partialsortperm(..., rev = true))
    # a precomputed normalised matrix through `norm =` (variablefeatures.jl:136-139)
    Y = sv.normalize_cells(X, method="relativecounts", scale_factor=1.0)
    assert np.array_equal(sv.find_variable_features(X, nsel, method=method, norm=Y), sel)
    # labelled input -> labelled selection
    genes = ["g%d" % j for j in range(X.shape[1])]
    named = sv.NamedArray(X, (["c%d" % i for i in range(X.shape[0])], genes), ("cells", "features"))
    lab = sv.find_variable_features(named, nsel, method=method)
    assert isinstance(lab, sv.NamedArray) and np.array_equal(lab.array, sel) and lab.names[0] == [genes[j] for j in sel]


def test_selector_keywords_and_errors(sv, orc):
    X = planted_counts(300, 500, 4, seed=9)
    np.testing.assert_allclose(sv.api._variable_feature_metric(X, "meanvarplot", {"num_bins": 7}),
                               orc.variable_feature_metric(X, "meanvarplot", num_bins=7), rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(sv.api._variable_feature_metric(X, "saunders", {"alpha_thresh": 0.99}),
                               orc.variable_feature_metric(X, "saunders", alpha_thresh=0.99), rtol=1e-12, atol=1e-14)
    with pytest.raises(ValueError):
        sv.find_variable_features(X, 10, method="nope")


def test_vst_dtype_float32(sv, orc):
    # variablefeatures.jl:34,37: mean_std(dtype, A) — the Float32 Welford of scaling.jl:18-34 on the counts, bit for bit
    from conftest import planted_counts
    X = planted_counts(3000, 500, 5, seed=2)
    mu32, var32 = orc.mean_var(X, dtype=np.float32)
    g_mu, g_sd = sv.mean_std(X.astype(np.float32))
    np.testing.assert_array_equal(g_mu, mu32)
    np.testing.assert_array_equal(g_sd, np.sqrt(var32))
    m64 = sv.variance_stabilizing_transformation(X)
    m32 = sv.variance_stabilizing_transformation(X, dtype=np.float32)
    assert m32.dtype == np.float64 and m32.shape == m64.shape
    np.testing.assert_allclose(m32, m64, rtol=2e-4)
    top64 = set(sv.find_variable_features(X, 100).tolist())
    top32 = set(sv.find_variable_features(X, 100, dtype=np.float32).tolist())
    assert len(top64 & top32) >= 97
    with pytest.raises(TypeError):
        sv.variance_stabilizing_transformation(X, dtype=np.int32)
