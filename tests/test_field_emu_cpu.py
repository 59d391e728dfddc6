"""polymath_b200/csrc/field.cuh on the host: tests/csrc/field_emu_test.cpp replaces the PTX wrappers by functions with an
explicit carry flag and runs the SAME template code — word-serial Montgomery product, dedicated squaring (sqr_wide), the
Karatsuba product, wide reduction — against an independent 32-bit schoolbook Montgomery product, for Fr and Fq, on random
and boundary operands (0, 1, p - 1, all-ones patterns).  The device runs of the same code are tests/test_kernels_gpu.py."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_field_templates_on_the_host(tmp_path):
    exe = str(tmp_path / "field_emu")
    src = os.path.join(ROOT, "tests", "csrc", "field_emu_test.cpp")
    subprocess.run(["g++", "-O2", "-std=c++17", "-Wno-unknown-pragmas", "-o", exe, src], check=True)
    out = subprocess.run([exe, "40000"], capture_output=True, text=True)
    sys.stdout.write(out.stdout)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "Fr: 40000 iterations, 0 mismatches" in out.stdout and "Fq: 40000 iterations, 0 mismatches" in out.stdout
