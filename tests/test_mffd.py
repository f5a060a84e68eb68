"""NK matrix-free residual-Jacobian product (a17): FormFunction_mf and the MFFD matvec."""
import numpy as np
import pytest

from adflow_b200.solver import ADFLOW_B200

from util import case, oracle_form_function, rel_l2

pytestmark = pytest.mark.gpu


def state_vec(hb):
    return np.transpose(hb.w[hb.d.owned()], (2, 1, 0, 3)).reshape(-1).copy()


def test_form_function_matches_oracle(cuda_lib):
    prm, hb = case(12, 10, 8)
    U = state_vec(hb)
    U[5::6][:7] = -1.0  # exercise the turbulence clip of setW
    ref = oracle_form_function(prm, hb, U)
    s = ADFLOW_B200(prm)
    try:
        s.addBlock(hb)
        r = s.formFunction(U)
    finally:
        s.close()
    assert rel_l2(r, ref) < 1e-12


def test_mffd_matvec_fixed_h_matches_oracle_difference(cuda_lib):
    prm, hb = case(11, 9, 8)
    U = state_vec(hb)
    rng = np.random.default_rng(314)  # getStatePerturbation(314)
    a = rng.standard_normal(U.size) * np.abs(U).clip(1e-6)
    h = 1e-6
    F0 = oracle_form_function(prm, hb, U)
    F1 = oracle_form_function(prm, hb, U + h * a)
    yref = (F1 - F0) / h
    s = ADFLOW_B200(prm)
    try:
        s.addBlock(hb)
        s.mffdSetBase(U)
        y = s.mffdApply(a, h)
        assert s.mffdLastH() == h
        # linearity / repeatability: the base is not disturbed by an apply
        y2 = s.mffdApply(a, h)
    finally:
        s.close()
    assert np.array_equal(y, y2)
    # both sides difference residuals that agree to ~1e-13: agreement of the quotient ~1e-13*|F|/(h*|Ja|)
    assert rel_l2(y, yref) < 1e-6


def test_mffd_walker_pernice_h_and_fd_consistency(cuda_lib):
    """h <= 0 -> h = sqrt(eps) sqrt(1+||U||)/||a||; the product is consistent with a centred
    difference of the oracle's F (reference invariant: tests/reg_tests/test_jacVecProdFWD.py:87-200)."""
    prm, hb = case(10, 8, 7)
    U = state_vec(hb)
    rng = np.random.default_rng(7)
    a = rng.standard_normal(U.size) * np.abs(U).clip(1e-6)
    s = ADFLOW_B200(prm)
    try:
        s.addBlock(hb)
        s.mffdSetBase(U)
        y = s.mffdApply(a, -1.0)
        h = s.mffdLastH()
    finally:
        s.close()
    hexp = np.sqrt(np.finfo(float).eps) * np.sqrt(1.0 + np.linalg.norm(U)) / np.linalg.norm(a)
    assert abs(h - hexp) < 1e-12 * hexp
    hc = 1e-6 / np.abs(a).max()
    yc = (oracle_form_function(prm, hb, U + hc * a) - oracle_form_function(prm, hb, U - hc * a)) / (2 * hc)
    assert rel_l2(y, yc) < 5e-3
