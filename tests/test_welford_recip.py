"""CPU: the reciprocal recurrence of the order-exact Welford kernels (csrc/preprocess.cu, WelfordStep<double>::run4) — from
RN(1/c) the reciprocals of c+1 .. c+8 by three fused Newton steps equal the IEEE reciprocal, exhaustively for every count the
test has time for (all c < 2^28 here; tools/studies/recip_check.c with no argument covers c < 2^32)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reciprocal_recurrence_is_correctly_rounded(tmp_path):
    exe = tmp_path / "recip_check"
    try:
        has_fma = " fma " in open("/proc/cpuinfo").read()
    except OSError:
        has_fma = False
    flags = ["-mfma"] if has_fma else []  # (without the instruction glibc's fma() is exact too, only slower)
    subprocess.run(["gcc", "-O2", *flags, "-ffp-contract=off", "-fopenmp", "-o", str(exe),
                    os.path.join(ROOT, "tools", "studies", "recip_check.c"), "-lm"], check=True)
    r = subprocess.run([str(exe), "28" if has_fma else "22"], capture_output=True, text=True)
    sys.stdout.write(r.stdout)
    assert r.returncode == 0, r.stdout + r.stderr
    assert ": 0 mismatches" in r.stdout
