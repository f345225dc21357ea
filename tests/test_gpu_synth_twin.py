"""GPU: the device generator of the benchmark inputs (csrc/synth.cu) and its host twin (oracle/csrc/synth_twin.c) produce the
SAME count matrix, bit for bit — the property that lets bench.py's `--impl reference` arm build the CUDA arm's input on the
host cores without loading the product library."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("m,g,nnz,K,seed,rows", [
    (5003, 700, 60.0, 7, 1, None),              # ragged tail (not a multiple of 4 / 1024)
    (40_000, 2500, 400.0, 50, 20260103, None),  # dense genes: intensities above 1, long inverse-CDF walks
    (40_000, 2500, 400.0, 50, 20260103, (12_000, 29_004)),   # a rank's shard of the same matrix
    (3000, 300, 250.0, 3, 99, None),            # very dense: the intensity clamp of the sampler is reached
])
def test_device_generator_equals_host_twin(sv, orc, m, g, nnz, K, seed, rows):
    d = sv.synthetic_counts(m, g, nnz, programs=K, fold=6.0, seed=seed, rows=rows)
    X = d.to_host()
    T = orc.synthetic_counts(m, g, nnz, programs=K, fold=6.0, seed=seed, rows=rows)
    assert X.shape == T.shape
    assert np.array_equal(X.indptr, T.indptr) and np.array_equal(X.indices, T.indices) and np.array_equal(X.data, T.data)
    d.free()


def test_twin_streaming_stats_equal_device_preprocessing(sv, orc):
    # the one pass the reference arm makes over all genes: library sizes, Welford moments, clipped variance
    m, g = 30_000, 4000
    t = orc.synth_tables(m, g, 300.0, 20, 6.0, 5)
    libsize, gene_nnz, mean, var, hist = orc.synth_stats(t)
    d = sv.synthetic_counts(m, g, 300.0, programs=20, fold=6.0, seed=5)
    s = np.empty(m, dtype=np.int64)
    sv._lib.check(sv.lib().svb_row_sums(d._h, sv._lib.ptr(s)))
    assert np.array_equal(s, libsize)
    mu_g, var_g = sv.mean_var(d)
    np.testing.assert_array_equal(mu_g, mean)
    np.testing.assert_array_equal(var_g, var)
    sd = np.sqrt(var)
    ex = sd.copy()
    ex[sd > 0] = np.sqrt(mean[sd > 0] * (1.0 + 0.5 * mean[sd > 0]))
    a = orc.stdvar_clipped_hist(m, hist, gene_nnz, mean, ex)
    b = sv.standardized_var_clipped(d, mean, ex)
    assert np.max(np.abs(a - b) / np.spacing(np.maximum(np.abs(a), np.abs(b)))) <= 2.0   # two <=1-ulp evaluations of one sum
    d.free()
