"""CPU: the reference arm of bench.py at a small configuration — the input comes from the host twin of the generator, goes through
the oracle's restatement of the reference pre-processing (no product code is imported) and is solved by the oracle's IRLBA;
the result must satisfy the reference's own acceptance criterion (test/test_irlba.jl:30) and the trend stand-in shared by both
bench arms must be a deterministic function of the per-gene moments."""
import numpy as np

import bench
from oracle import severo_oracle as orc


def test_hvg_trend_is_deterministic_and_interpolates_bin_medians():
    rng = np.random.default_rng(0)
    mu = 10.0 ** rng.uniform(-3, 2, 5000)
    sd = np.sqrt(mu) * 10.0 ** rng.normal(0, 0.1, mu.size)
    a = bench.hvg_trend(mu, sd)
    b = bench.hvg_trend(mu.copy(), sd.copy())
    np.testing.assert_array_equal(a, b)                              # pure function of (mu, sd)
    perm = rng.permutation(mu.size)
    np.testing.assert_array_equal(bench.hvg_trend(mu[perm], sd[perm]), a[perm])   # no dependence on the gene order
    assert np.all(a > 0) and np.all(np.isfinite(a))
    # the trend follows sqrt(mu) (the planted relation) to within the scatter of the bin medians
    assert np.median(np.abs(np.log10(a) - 0.5 * np.log10(mu))) < 0.03


def test_reference_arm_small_configuration():
    cfg = dict(m=6000, g=1200, nnz=150.0, n=200, nu=6, programs=6, desc="test")
    C, info = bench.reference_problem(cfg)
    m, n = C.shape
    assert (m, n) == (cfg["m"], cfg["n"]) and 0 < info["z"] <= info["Z"]
    init = np.random.default_rng(bench.SEED).standard_normal(n)
    S = orc.irlba(C, cfg["nu"], init=init, tol=1e-7)
    assert np.all(np.diff(S.S) <= 0) and S.S[-1] > 0
    # ||S'U - V Sigma|| / ||sigma|| < tol-ish (the bench's residual test, with the oracle's products)
    StU = C.mul(np.asfortranarray(S.U), trans=True)
    assert np.linalg.norm(StU - S.V * S.S) / np.linalg.norm(S.S) < 1e-5
    # the same call again gives the same matrix: the generator twin is deterministic
    C2, info2 = bench.reference_problem(cfg)
    assert info2["z"] == info["z"] and info2["Z"] == info["Z"]


def test_reference_arm_never_loads_the_product(tmp_path):
    """`bench.py --impl reference` on the reference's own CPU-runnable configuration (C1): one JSON line with impl = reference,
    and neither the product package nor libsevero_b200.so in the process afterwards."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys, json\n"
        "sys.argv = ['bench.py', '--impl', 'reference', '--config', 'C1', '--steps', '1', '--warmup', '0']\n"
        "import bench\n"
        "lines = []\n"
        "bench.emit = lines.append\n"
        "bench.main()\n"
        "line = lines[-1]\n"
        "maps = open('/proc/self/maps').read()\n"
        "print(json.dumps({'impl': line.get('impl'), 'value': line.get('value'), 'metric': line.get('metric'),\n"
        "                  'kind': line.get('cpu_baseline', {}).get('kind'),\n"
        "                  'product_module': any(k.startswith('severo_jl_b200') or k.startswith('severo.jl_b200') for k in sys.modules),\n"
        "                  'product_so': 'libsevero_b200' in maps}))\n")
    r = subprocess.run([sys.executable, "-c", code], cwd=root, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    # (bench.main() points fd 1 at stderr to keep the real stdout for its one line: the summary arrives on either stream)
    out = json.loads([ln for ln in (r.stdout + r.stderr).splitlines() if ln.startswith('{"impl"')][-1])
    assert out["impl"] == "reference" and out["kind"] == "port" and out["value"] > 0
    assert out["metric"] == bench.METRIC
    assert not out["product_module"] and not out["product_so"]
