"""GPU: the C ABI exercised WITHOUT Python in the process — tests/abi/abi_harness.c is compiled with gcc against
include/severo_b200.h and the in-tree library and drives svb_csc_upload(index_base = 1, Int64) -> svb_operator_create ->
svb_irlba the way Julia's `ccall` would (src/irlba.jl:66-71), checking test/test_irlba.jl:30's criterion in plain C."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_c_harness_drives_the_abi(tmp_path):
    libdir = os.path.join(ROOT, "severo.jl_b200")
    exe = str(tmp_path / "abi_harness")
    subprocess.run(["gcc", "-O2", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "abi", "abi_harness.c"), "-o", exe,
                    "-L", libdir, "-lsevero_b200", f"-Wl,-rpath,{libdir}", "-lm"], check=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ABI HARNESS OK" in r.stdout
