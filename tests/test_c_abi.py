"""The drop-in boundary from C: tests/c/gemm_literals.c is a plain C99 translation unit compiled with gcc against
include/ccv_nnc_sm100.h and linked with libccv_nnc_sm100.so.  It calls ccv_nnc_cmd_exec with BY-VALUE ccv_nnc_cmd_t /
ccv_nnc_hint_t (the reference's signature, lib/nnc/ccv_nnc.h:315 -- the Python tests reach it through a pointer wrapper)
on the literal known-answer GEMM cases of test/unit/nnc/gemm.tests.c, moving data with CMD_DATA_TRANSFER as
test/int/nnc/cublas.tests.c does.  Without a GPU only its host-side half runs (struct sizes, registration record,
NO_KERNEL for host tensors)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    exe = str(tmp_path / "gemm_literals")
    lib_dir = os.path.join(ROOT, "ccv_b200")
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c", "gemm_literals.c"), "-o", exe,
           "-L", lib_dir, "-lccv_nnc_sm100", "-lm", "-Wl,-rpath," + lib_dir]
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert out.returncode == 0, out.stdout
    return exe


def test_header_compiles_as_c_and_host_side_of_the_boundary(tmp_path):
    exe = _build(tmp_path)
    out = subprocess.run([exe, "--no-gpu"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120)
    assert out.returncode == 0, out.stdout


@pytest.mark.gpu
def test_by_value_cmd_exec_from_c_on_the_reference_gemm_literals(gpu, tmp_path):
    exe = _build(tmp_path)
    out = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert out.returncode == 0, out.stdout
    assert "0 failure(s)" in out.stdout
