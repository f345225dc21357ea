"""CPU: the C-ABI library loads and exports every symbol the header declares; host-side logic."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "severo_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(svb_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    import severo_jl_b200 as sv
    lib = sv.load()
    syms = _header_symbols()
    assert len(syms) >= 40
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/severo_b200.h but not exported"
    assert sorted(sv.SIGNATURES) == syms, "python binding table out of sync with the header"
    counts = _header_param_counts()
    for name, (_, argtypes) in sv.SIGNATURES.items():
        assert len(argtypes) == counts[name], f"{name}: ctypes table has {len(argtypes)} arguments, C declares {counts[name]}"


def _header_param_counts():
    txt = open(os.path.join(ROOT, "include", "severo_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    out = {}
    for name, params in re.findall(r"\b(svb_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;", txt, flags=re.S):
        params = params.strip()
        out[name] = 0 if params in ("", "void") else params.count(",") + 1
    return out


def _split_top_level(text):
    parts, depth, cur = [], 0, ""
    for ch in text:
        if ch in "({[":
            depth += 1
        elif ch in ")}]":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        parts.append(cur.strip())
    return parts


def test_julia_overlay_binds_declared_symbols_with_matching_arity():
    # Julia is not in the image: the overlay cannot run here, but every `ccall` in it must name a symbol the header
    # declares and pass as many argument types as the C declaration has parameters.
    counts = _header_param_counts()
    src = open(os.path.join(ROOT, "julia", "SeveroB200.jl")).read()
    seen = 0
    for m in re.finditer(r"ccall\(\(:(svb_[a-z0-9_]+), libsvb\),", src):
        name = m.group(1)
        assert name in counts, f"julia overlay calls {name}, which include/severo_b200.h does not declare"
        rest = src[m.end():]
        # rest = " RetType, (T1, T2, ...), args...)": the argument-type tuple is the first top-level parenthesis group
        start = rest.index("(")
        depth, end = 0, None
        for i in range(start, len(rest)):
            depth += rest[i] == "("
            depth -= rest[i] == ")"
            if depth == 0:
                end = i
                break
        types = _split_top_level(rest[start + 1:end])
        assert len(types) == counts[name], f"{name}: julia passes {len(types)} argument types, C declares {counts[name]}"
        seen += 1
    assert seen >= 20


def test_no_cpu_fallback_without_device():
    import severo_jl_b200 as sv
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(sv.SeveroB200Error):
        sv.init()
    with pytest.raises(sv.SeveroB200Error):
        sv.normalize_cells(__import__("scipy.sparse").sparse.csc_matrix(np.eye(3, dtype=np.int64)))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "severo.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), f"{f} mentions the oracle"


def test_shard_bounds():
    from severo_jl_b200 import sharding
    b = sharding.shard_bounds(1306127, 8)
    assert b[0][0] == 0 and b[-1][1] == 1306127
    assert all(lo % 4 == 0 for lo, _ in b) and all(b[i][1] == b[i + 1][0] for i in range(7))
    nnz = np.r_[np.full(500, 10), np.full(500, 1)]
    b2 = sharding.shard_bounds(1000, 2, row_nnz=nnz)
    assert b2[0][1] < 500 and b2[0][1] % 4 == 0
    assert sharding.shard_bounds(10, 1) == [(0, 10)]


def test_shard_bounds_properties():
    # any sizes: the ranges tile [0, m) in order, inner boundaries are multiples of 4, nnz-balanced cuts stay within one cell
    # quad + one cell of the ideal split
    from hypothesis import given, settings, strategies as st
    from severo_jl_b200 import sharding

    @settings(max_examples=150, deadline=None)
    @given(st.integers(0, 5000), st.integers(1, 9), st.booleans(), st.integers(0, 2 ** 31))
    def check(m, nranks, weighted, seed):
        nnz = np.random.default_rng(seed).integers(0, 50, m) if weighted else None
        b = sharding.shard_bounds(m, nranks, row_nnz=nnz)
        assert len(b) == nranks and b[0][0] == 0 and b[-1][1] == m
        assert all(lo <= hi for lo, hi in b) and all(b[i][1] == b[i + 1][0] for i in range(nranks - 1))
        assert all(lo % 4 == 0 for lo, _ in b[1:])
        if weighted and m:
            csum = np.concatenate([[0], np.cumsum(nnz)])
            for r in range(1, nranks):
                ideal = int(np.searchsorted(csum, csum[-1] * r / nranks))
                assert ideal - 4 < b[r][0] <= ideal
    check()
    with pytest.raises(ValueError):
        sharding.shard_bounds(10, 0)


def test_tools_and_studies_parse():
    import ast
    for d in ("tools", os.path.join("tests", "studies")):
        for f in sorted(os.listdir(os.path.join(ROOT, d))):
            if f.endswith(".py"):
                ast.parse(open(os.path.join(ROOT, d, f)).read(), filename=f)


def test_loess_recovers_smooth_trend():
    from severo_jl_b200.loess import loess_fit_predict
    rng = np.random.default_rng(0)
    x = np.sort(rng.uniform(-3, 1, 3000))
    y = 0.5 * x + 0.1 * x ** 2 + 0.01 * rng.standard_normal(3000)
    f = loess_fit_predict(x, y, span=0.5)
    assert np.max(np.abs(f - (0.5 * x + 0.1 * x ** 2))) < 0.02


def test_svd_flip_matches_reference_rule():
    import severo_jl_b200 as sv
    U = np.array([[0.1, -0.9], [-0.8, 0.2]])
    Vt = np.array([[1.0, 2.0], [3.0, -4.0]])
    S = sv.svd_flip(sv.SVD(U.copy(), np.ones(2), Vt.copy()))
    assert S.U[1, 0] > 0 and S.U[0, 1] > 0
    np.testing.assert_array_equal(S.Vt, -Vt)


_GLOO = r"""
import os, sys
import numpy as np, scipy.sparse as sp
import torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from oracle import severo_oracle as orc
from severo_jl_b200 import sharding
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
rng = np.random.default_rng(0)
X = sp.random(403, 60, 0.2, random_state=1, format="csc")
mu = rng.standard_normal(60)
bounds = sharding.shard_bounds(403, world)
lo, hi = bounds[rank]
full = orc.CenteredMatrix(X, mu)
local = orc.CenteredMatrix(X[lo:hi], mu)
v = rng.standard_normal(60); w = rng.standard_normal(403)
# S*v is shard-local; S'*w needs the sum over ranks (SURVEY 8e)
yl = local.mul(v)
assert np.allclose(yl, full.mul(v)[lo:hi], rtol=1e-13, atol=1e-13)
t = torch.from_numpy(local.mul(w[lo:hi], trans=True).copy())
dist.all_reduce(t)
assert np.allclose(t.numpy(), full.mul(w, trans=True), rtol=1e-12, atol=1e-12)
U = sharding.gather_rows(np.asfortranarray(yl[:, None]), bounds, rank)
assert np.allclose(U[:, 0], full.mul(v), rtol=1e-13, atol=1e-13)
dist.destroy_process_group()
print("OK", rank)
"""


def test_cell_sharding_two_ranks_gloo(tmp_path):
    script = tmp_path / "gloo_shard.py"
    script.write_text(_GLOO)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29611", str(script), ROOT],
                       capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("OK") == 2
