"""csrc/stdsort.cuh reproduces the permutation of libstdc++'s std::sort (unstable above 16 elements) - the tie order the reference's
floating-point sums over sorted candidate lists depend on. Checked against std::sort itself, same number of comparisons."""
import ctypes as C
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    so = os.path.join(ROOT, "tests", "_stdsort_check.so")
    src = os.path.join(ROOT, "tests", "stdsort_check.cpp")
    hdr = os.path.join(ROOT, "mpeg-pcc-tmc2_b200", "csrc", "stdsort.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++", "-o", so, src])
    return C.CDLL(so)


@pytest.mark.parametrize("key_range", [1, 2, 5, 40, 100000])
def test_same_permutation_and_comparisons_as_std_sort(key_range, lib):
    cmp = (C.c_long * 2)()
    assert lib.stdsort_check(300, 12, key_range, 1234 + key_range, cmp) == 0
    assert cmp[0] == cmp[1] and cmp[0] > 0   # not only the same result: the same comparisons
