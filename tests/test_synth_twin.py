"""CPU: the host twin of the synthetic generator (oracle/csrc/synth_twin.c) is self-consistent — the streaming statistics
pass equals the oracle's own sweeps over the explicit matrix, a shard equals the rows of the whole, and the distribution is
the Poisson one it claims (mean number of nonzeros per cell as requested)."""
import numpy as np
import scipy.sparse as sp


def test_twin_stats_pass_equals_oracle_sweeps(orc):
    m, g = 20_000, 3000
    X = orc.synthetic_counts(m, g, 300.0, programs=12, seed=7)
    assert 0.9 * 300 < X.nnz / m < 1.1 * 300
    t = orc.synth_tables(m, g, 300.0, 12, 6.0, 7)
    libsize, gene_nnz, mean, var, hist = orc.synth_stats(t)
    assert np.array_equal(libsize, np.asarray(X.sum(axis=1)).ravel())
    assert np.array_equal(gene_nnz, np.diff(X.indptr))
    mo, vo = orc.mean_var(X)
    assert np.array_equal(mean, mo) and np.array_equal(var, vo)          # the reference's Welford order, bit for bit
    assert np.array_equal(hist[:, 1:].sum(axis=1), gene_nnz)
    sd = np.sqrt(var)
    ex = sd.copy()
    ex[sd > 0] = np.sqrt(mean[sd > 0] * (1.0 + 0.5 * mean[sd > 0]))
    a = orc.stdvar_clipped_hist(m, hist, gene_nnz, mean, ex)
    b = orc.standardized_var_clipped(X, mean, ex)
    assert np.max(np.abs(a - b) / np.spacing(np.maximum(np.abs(b), 1e-300))) <= 2.0


def test_twin_shard_is_rows_of_the_whole(orc):
    X = orc.synthetic_counts(9000, 800, 120.0, programs=5, seed=3)
    t = orc.synth_tables(9000, 800, 120.0, 5, 6.0, 3, rows=(4000, 7003))
    cp, rv, v = orc.synth_columns(t, np.arange(800))
    Xs = sp.csc_matrix((v.astype(np.int64), rv, cp), shape=(3003, 800))
    assert (Xs != X[4000:7003]).nnz == 0
    sel = np.array([5, 700, 3])
    cp, rv, v = orc.synth_columns(orc.synth_tables(9000, 800, 120.0, 5, 6.0, 3), sel)
    assert (sp.csc_matrix((v.astype(np.int64), rv, cp), shape=(9000, 3)) != X[:, sel]).nnz == 0


def test_twin_fixed_order_exp_is_accurate(orc):
    # the sampler's exp is a fixed-order polynomial (bit-reproducible on the device): P(X = 0) must still be exp(-lam)
    t = orc.synth_tables(200_000, 4, 2.0, 1, 1.0, 11)      # 4 genes, one program: lam_j = L_i * scale * p_j
    cp, rv, v = orc.synth_columns(t, np.arange(4))
    for j in range(4):
        lam = t.lib * t.lamtab[j]
        expect = np.mean(1.0 - np.exp(-lam))
        got = (cp[j + 1] - cp[j]) / 200_000
        assert abs(got - expect) < 5 * np.sqrt(expect * (1 - expect) / 200_000) + 1e-9
