"""Known-answer tests of the Fortran-subset -> C translator (oracle/f90toc.py) that builds oracle/_ref.

The parity of the oracle is pinned against the reference's routines AS TRANSLATED, so the translator itself has to be
trusted: tests/golden/f90toc_kat.F90 (written for this test, not reference code) exercises the constructs the reference's
hot-path routines use -- integer powers and integer division, the numeric intrinsics, do loops with negative step /
cycle / exit, if / else if, select case, arrays with lower bounds /= 1 in column-major order, whole-array and section
assignment, pointer sections (lower bound 1) and whole-array pointers (bounds kept), assumed-shape dummies, optional
dummies with present(), contained subroutines reading host variables, module variables, #ifdef -- and the translated,
gcc-compiled result is compared with values computed independently in Python.  Inputs are dyadic rationals so that
every expected value is exact."""
import ctypes as C
import math
import os
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


@pytest.fixture(scope="module")
def kat():
    import f90toc

    work = os.path.join(ROOT, "oracle", "_ref", "_selftest")      # git-ignored build directory
    os.makedirs(work, exist_ok=True)
    env = f90toc.Env([], {}, [], {})
    code, _tr = f90toc.translate_module(os.path.join(HERE, "golden", "f90toc_kat.F90"),
                                        only={"scalars", "arrays", "optional_and_shape", "driver_shape", "alloc_case"}, env=env,
                                        rename_modules={}, patches=(), defined=(), tr=None, prefix="kat_")
    assert "broken" not in code                     # the #ifdef NEVER block is dropped
    with open(os.path.join(work, "kat.c"), "w") as f:
        f.write(code)
    for name, text in (("ref_env.h", "static const double zero = 0.0, one = 1.0;\n"), ("ref_protos.h", "")):
        with open(os.path.join(work, name), "w") as f:
            f.write(text)
    so = os.path.join(work, "kat.so")
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-std=gnu11", "-w", "-I" + work, "-shared", "-o", so,
                           os.path.join(work, "kat.c"), "-lm"])
    return C.CDLL(so)


def expected_scalars(n, x):
    r = np.zeros(12)
    r[0] = x ** 3 + 1 / (x * x)
    r[1] = float(n // 3) + n % 5
    r[2] = max(x, 2.5, -x) - min(x, 0.25)
    r[3] = math.copysign(3.0, -x) + abs(-x) + max(x - 1, 0) + max(1 - x, 0)
    r[4] = sum(i * 0.5 for i in range(n, 0, -2))
    r[5] = 1.0 if x > 2 else (2.0 if (x > 1 and not n == 3) else 3.0)
    r[6] = 10.0 if n in (1, 2) else (70.0 if n == 7 else -1.0)
    k = 0
    for i in range(1, 11):
        if i % 2 == 0:
            continue
        if i > 7:
            break
        k += i
    r[7] = k
    r[8] = math.sqrt(x) * math.exp(-x) + math.log(x + 1)
    r[9] = 1e-4 * x ** 2 + 5 * (1.0 / 3.0 + x * 0.0)
    r[10] = x ** 10
    r[11] = (x + 1) ** 2 / 2 + 1 / (x + 1)
    return r


@pytest.mark.parametrize("n,x", [(7, 1.5), (2, 2.25), (3, 1.25), (10, 0.5), (1, 4.0)])
def test_scalar_constructs(kat, n, x):
    res = (C.c_double * 12)()
    kat.kat_scalars(C.byref(C.c_int(n)), C.byref(C.c_double(x)), res)
    got, exp = np.array(res), expected_scalars(n, x)
    # libm calls (sqrt/exp/log) may differ in the last place between Python's and C's libm: 1 ulp; everything else exact
    exact = [q for q in range(12) if q != 8]
    assert np.array_equal(got[exact], exp[exact]), (got - exp)
    assert abs(got[8] - exp[8]) <= 2e-16 * abs(exp[8])


@pytest.mark.parametrize("n", [2, 4, 7])
def test_array_constructs_and_module_state(kat, n):
    a = np.asfortranarray(np.arange((n + 1) * 4, dtype=float).reshape(n + 1, 4, order="F") * 0.25 + 0.125)
    a0 = a.copy()
    acc0 = C.c_double.in_dll(kat, "kat_acc").value
    cnt0 = C.c_int.in_dll(kat, "kat_counter").value
    out = (C.c_double * 8)()
    kat.kat_arrays(C.byref(C.c_int(n)), a.ctypes.data_as(C.c_void_p), out)
    b = np.full((n + 1, 4), 1.5)
    for jj, j in enumerate(range(-1, 3)):          # declared a(0:n, -1:2): column jj <-> Fortran index j
        b[:, jj] += a0[:, jj] * j
    a1 = a0.copy()
    a1[:, 1] = b[:, 2] * 2
    exp = np.zeros(8)
    exp[0] = a1[n, 1]; exp[1] = b[0, 0] + b[n, 3]
    exp[2] = b[1, 1]                                # p => b(1:, 0:); p(1, 1)
    exp[3] = b[n, 3]                                # p(n, 3)
    exp[4] = b[2, 0] + b[2, 3]                      # q => b(2, :); q(1) + q(4)
    exp[5] = b[0, 0]                                # p => b keeps the bounds: p(0, -1)
    exp[6] = b[0, 2] + b[1, 2]
    exp[7] = (acc0 + exp[0]) + (cnt0 + 1)
    assert np.array_equal(np.array(out), exp), np.array(out) - exp
    assert np.array_equal(a, a1)
    assert C.c_int.in_dll(kat, "kat_counter").value == cnt0 + 1


def test_assumed_shape_and_optional(kat):
    res = (C.c_double * 2)()
    kat.kat_driver_shape(res)
    m = np.array([[i + 0.1 * j for j in (1, 2)] for i in (1, 2, 3)])
    wgt = np.array([[i + 10 * j for j in (1, 2)] for i in (1, 2, 3)])
    s = 0.0
    for j in range(2):                              # the Fortran loop order
        for i in range(3):
            s += m[i, j] * wgt[i, j]
    assert res[0] == s and res[1] == -s


@pytest.mark.parametrize("n", [3, 6])
def test_allocatable_named_constructs_and_integer_intrinsics(kat, n):
    res = (C.c_double * 4)()
    kat.kat_alloc_case(C.byref(C.c_int(n)), res)
    q = lambda i, j, k: 100.0 * k + 10 * j + i  # noqa: E731
    m = sum(j for j in range(2, n + 2) if j <= 3)
    # Fortran mod(-7, 3) = -1 (sign of the dividend)
    assert list(res) == [q(2, 2, 0) + q(n, n + 1, 1), float(m), float(max(n, 3) - min(n, 3) - 1), q(n, 2, 1) * 0.5]
