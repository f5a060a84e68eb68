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


def test_mffd_device_vectors_equal_host_vectors(cuda_lib):
    """adfb_mffd_apply_device (vectors resident on the GPU, the PETSc VECCUDA path) gives bit-identical results"""
    import torch

    prm, hb = case(11, 9, 8)
    s = ADFLOW_B200(prm)
    try:
        s.addBlock(hb)
        U = s.getStates()
        s.mffdSetBase(U)
        a = np.random.default_rng(3).standard_normal(U.size)
        y_host = s.mffdApply(a, 1e-7)
        da = torch.from_numpy(a).cuda()
        dy = torch.zeros_like(da)
        s.mffdApplyDevice(da.data_ptr(), dy.data_ptr(), da.numel(), 1e-7)
        assert np.array_equal(dy.cpu().numpy(), y_host)
        # Walker-Pernice h on the device, and the a == 0 shortcut
        s.mffdApplyDevice(da.data_ptr(), dy.data_ptr(), da.numel(), -1.0)
        assert s.mffdLastH() > 0 and np.array_equal(dy.cpu().numpy(), s.mffdApply(a, -1.0))
        dz = torch.zeros_like(da)
        s.mffdApplyDevice(dz.data_ptr(), dy.data_ptr(), da.numel(), -1.0)
        assert float(dy.abs().max()) == 0.0
        # host pointers are rejected loudly
        import ctypes as C
        assert s.L.adfb_mffd_apply_device(a.ctypes.data, a.ctypes.data, a.size, 1e-7) != 0
    finally:
        s.close()


def test_fused_product_is_bitwise_the_three_pass_product(cuda_lib):
    """The matrix-free product with the perturbation formed inside the state preparation and the difference quotient
    formed by the kernels that write dw (ADFB_MFFD_FUSED=1) against the default three-pass form: same bits."""
    import os

    prm, hb = case(14, 11, 9)
    U = state_vec(hb)
    a = np.random.default_rng(2).standard_normal(U.size) * np.abs(U).clip(1e-6)
    s = ADFLOW_B200(prm)
    try:
        s.addBlock(hb)
        s.mffdSetBase(U)
        y0 = s.mffdApply(a, 1e-7).copy()
        os.environ["ADFB_MFFD_FUSED"] = "1"
        try:
            y1 = s.mffdApply(a, 1e-7).copy()
            y2 = s.mffdApply(a, 3e-7).copy()      # the captured graph follows a new h
        finally:
            del os.environ["ADFB_MFFD_FUSED"]
        y3 = s.mffdApply(a, 3e-7).copy()
    finally:
        s.close()
    assert np.isfinite(y1).all() and np.abs(y1).max() > 0
    assert np.array_equal(y1, y0)
    assert np.array_equal(y2, y3) and not np.array_equal(y1, y2)
