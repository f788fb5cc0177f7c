"""The drop-in boundary from plain C: include/jpetra_b200.h compiles as C99 and the library is usable through the
header's own prototypes (examples/c_abi_smoke.c).  Without a GPU the program must report the loud JP_CUDA_ERROR of
jp_init (no CPU fallback); on a B200 (-m gpu) it runs the reference's 2x2 apply."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    import jpetra_b200 as jp
    jp._lib.load()  # builds the library if needed
    exe = str(tmp_path / "c_abi_smoke")
    subprocess.run(["gcc", "-std=gnu99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "c_abi_smoke.c"),
                    "-ldl", "-o", exe], check=True)
    return exe, jp._lib.LIB_PATH


def test_header_is_c_and_library_fails_loudly_without_a_gpu(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: see the gpu-marked test")
    exe, lib = _build(tmp_path)
    r = subprocess.run([exe, lib], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "no device: jp_init -> JP_CUDA_ERROR" in r.stdout


@pytest.mark.gpu
def test_c_program_runs_the_reference_2x2_apply(tmp_path):
    exe, lib = _build(tmp_path)
    r = subprocess.run([exe, lib], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "apply ok" in r.stdout, r.stdout + r.stderr
