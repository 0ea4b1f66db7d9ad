"""The scalar lean kernels never multiply or convert to compare `column <op> literal`: per CTA the comparison is
folded into one modular interval test on the field's mantissa, or on its digit bytes (cqg_lean2.cuh:
lean2_interval, lean2_code). tests/native/interval_check.cu checks that folding exhaustively against the scaled
integer comparison it replaces (all six operators, literals with 0..3 fraction digits, mantissas 0..12000 and a
sweep up to 10^7 - 1). Compiled for the HOST by nvcc: no GPU needed."""
import os
import shutil
import subprocess

import pytest

from conftest import ROOT


@pytest.mark.timeout(300)
def test_interval_folding_equals_scaled_integer_compare(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not found")
    exe = str(tmp_path / "interval_check")
    src = os.path.join(ROOT, "tests", "native", "interval_check.cu")
    subprocess.run([nvcc, "-std=c++17", "-O2", "-gencode", "arch=compute_100a,code=sm_100a", "-I" + os.path.join(ROOT, "include"),
                    "-I" + os.path.join(ROOT, "cq_b200", "csrc"), "-o", exe, src], check=True, capture_output=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:]
    assert r.stdout.startswith("ok ")


@pytest.mark.timeout(300)
def test_four_byte_decimal_decode_is_exact(tmp_path):
    """tests/native/decimal_check.cu: every field of 1..4 characters over digits, '.', signs, exponent letters,
    blanks, quotes, newline, a UTF-8 lead byte — with every kind of byte in front of it — decodes to the reference's
    value or is declined (lean2_dec4_word / lean2_dec4c_word, the kernels' own arithmetic compiled for the host)."""
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not found")
    exe = str(tmp_path / "decimal_check")
    src = os.path.join(ROOT, "tests", "native", "decimal_check.cu")
    subprocess.run([nvcc, "-std=c++17", "-O2", "-gencode", "arch=compute_100a,code=sm_100a", "-I" + os.path.join(ROOT, "include"),
                    "-I" + os.path.join(ROOT, "cq_b200", "csrc"), "-o", exe, src], check=True, capture_output=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:]
    assert r.stdout.startswith("ok ")


@pytest.mark.timeout(300)
def test_field_split_on_bitmasks_equals_byte_split(tmp_path):
    """tests/native/split_check.cu: Lean2Stops (the general lean kernels' field split on delimiter / terminator
    bitmasks) against a byte-by-byte split on 800 000 random rows, empty and missing fields included."""
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not found")
    exe = str(tmp_path / "split_check")
    src = os.path.join(ROOT, "tests", "native", "split_check.cu")
    subprocess.run([nvcc, "-std=c++17", "-O2", "-gencode", "arch=compute_100a,code=sm_100a", "-I" + os.path.join(ROOT, "include"),
                    "-I" + os.path.join(ROOT, "cq_b200", "csrc"), "-o", exe, src], check=True, capture_output=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:]
    assert r.stdout.startswith("ok ")


@pytest.mark.timeout(300)
def test_byte_classification_masks_are_exact(tmp_path):
    """tests/native/masks_check.cu: lean2_masks16 (phase 1 of the scalar lean kernel: SWAR classes + dot-product
    movemask) against a byte loop, all byte-value pairs at all positions and 1.8 M random chunks, six delimiters."""
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not found")
    exe = str(tmp_path / "masks_check")
    src = os.path.join(ROOT, "tests", "native", "masks_check.cu")
    subprocess.run([nvcc, "-std=c++17", "-O2", "-gencode", "arch=compute_100a,code=sm_100a", "-I" + os.path.join(ROOT, "include"),
                    "-I" + os.path.join(ROOT, "cq_b200", "csrc"), "-o", exe, src], check=True, capture_output=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:]
    assert r.stdout.startswith("ok ")
