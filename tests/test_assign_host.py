"""CPU: the build-time replica assignment of the count-level operator (csrc/assign.cuh compiles for host and device) — validity of
every output code, the pass count it reports, and optimality of the matching against an independent augmenting-path matching."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_replica_assignment_valid_and_optimal(tmp_path):
    exe = tmp_path / "assign_check"
    subprocess.run(["g++", "-O2", "-o", str(exe), os.path.join(ROOT, "tests", "assign", "assign_check.cpp")], check=True)
    r = subprocess.run([str(exe), "40000"], capture_output=True, text=True)
    sys.stdout.write(r.stdout)
    assert r.returncode == 0, r.stdout + r.stderr
    assert " 0 failures" in r.stdout
