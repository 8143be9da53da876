"""A C99 program written like INTEGRATION.md section 2 must compile against include/davidson_b200.h with
`gcc -std=c99 -pedantic -Wall -Wextra -Werror` and link against the shared library: the boundary is a plain C ABI
(SURVEY.md section 8b), not a C++ or torch interface.  (Collected last on purpose.)"""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "fortran_davidson_b200")


def _build(tmp_path):
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    if not os.path.exists(os.path.join(LIBDIR, "libdavidson_b200.so")):
        pytest.skip("library not built")
    exe = str(tmp_path / "consumer")
    cmd = [gcc, "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "c_consumer", "consumer.c"), "-o", exe, "-L", LIBDIR, "-ldavidson_b200",
           "-Wl,-rpath," + LIBDIR]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stderr
    return exe


def test_c99_consumer_compiles_links_and_fails_loudly_without_a_gpu(tmp_path):
    import torch
    exe = _build(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("GPU present: the computing variant runs in the gpu-marked test")
    p = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert p.returncode == 0, p.stdout + p.stderr
    out = p.stdout
    assert "version 100" in out and "rows 0 12544 87808 100000" in out
    assert "(DAV_ERR_CUDA)" in out and "no CPU fallback" in out


@pytest.mark.gpu
def test_c99_consumer_solves_on_the_gpu(tmp_path):
    exe = _build(tmp_path)
    p = subprocess.run([exe, "run"], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    line = [ln for ln in p.stdout.splitlines() if ln.startswith("rc ")][0].split()
    assert line[1] == "0" and abs(float(line[3]) - 1.0) < 1e-3
