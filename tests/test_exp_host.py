"""CPU: the table-driven exp of the GRIN kernels (csrc/pyr_exp.cuh, __host__ __device__)
compiled for the host and compared with the C library's expl over 8e6 samples."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="nvcc not on PATH")
def test_table_exp_is_within_one_ulp(tmp_path):
    # 512-entry table + economised quartic: measured 1.0015 ulp against expl (the table entry's own
    # rounding and the final fma account for ~1 ulp; the polynomial for < 0.01)
    exe = str(tmp_path / "test_exp")
    subprocess.check_call(["nvcc", "-O2", "-Wno-deprecated-gpu-targets", "-o", exe,
                           os.path.join(ROOT, "tools", "micro", "test_exp.cu")])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "max_ulp_err" in out.stdout
    assert float(out.stdout.split("max_ulp_err")[1].split()[0]) <= 1.01
