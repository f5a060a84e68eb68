"""Error behaviour of the C ABI: every misuse returns non-zero and leaves a message for adfb_last_error, which the
Fortran side maps to `terminate(routine, msg)` (src/utils/utils.F90:501) -- nothing is silently ignored and nothing
falls back to a CPU path."""
import ctypes as C

import numpy as np
import pytest

from adflow_b200 import make_params
from adflow_b200._lib import AdflowB200Error
from adflow_b200.solver import ADFLOW_B200, RES_FLOW, RES_TURB

from util import case

pytestmark = pytest.mark.gpu


def _err(L):
    buf = C.create_string_buffer(1024)
    L.adfb_last_error(buf, 1024)
    return buf.value.decode()


def test_misuse_returns_nonzero_with_message(cuda_lib):
    prm, hb = case(6, 5, 4)
    s = ADFLOW_B200(prm)
    L = s.L
    try:
        assert L.adfb_residual(1, 0) != 0 and "neither flow nor turbulence" in _err(L)
        assert L.adfb_block_create(0, 1, 0, 5, 4, 6, 1) != 0 and "bad extents" in _err(L)
        assert L.adfb_block_create(0, 1, 6, 5, 4, 7, 1) != 0 and "nw must be" in _err(L)
        assert L.adfb_download_residual(3, None) != 0 and "no block" in _err(L)
        s.addBlock(hb)
        assert L.adfb_block_create(0, 1, 6, 5, 4, 6, 1) != 0 and "already exists" in _err(L)
        v = np.zeros(7)
        assert L.adfb_set_states(v.ctypes.data, 7) != 0 and "does not match" in _err(L)
        assert L.adfb_mffd_apply(v.ctypes.data, v.ctypes.data, 7, 1e-7) != 0 and "set_base" in _err(L)
        assert L.adfb_halo_exchange(1, 0, 7, 1, 0, 1) != 0 and "bad variable range" in _err(L)
        bad = make_params()
        bad.equations = 9
        assert L.adfb_set_params(C.byref(bad)) != 0 and "bad equations" in _err(L)
        bad = make_params()
        bad.spaceDiscr = 3
        assert L.adfb_set_params(C.byref(bad)) != 0 and "spaceDiscr" in _err(L)
        # the Python mirror raises
        with pytest.raises(AdflowB200Error):
            s.setStates(np.zeros(5))
        # and the context is still usable afterwards
        s.residual(RES_FLOW | RES_TURB)
        assert np.isfinite(s.getResNorms()).all()
    finally:
        s.close()


def test_calls_before_init_fail(cuda_lib):
    from adflow_b200 import _lib

    L = _lib.load()
    L.adfb_finalize()
    out = (C.c_double * 2)()
    assert L.adfb_norms(out) != 0
    assert "adfb_init has not been called" in _err(L)


def test_multigrid_ank_and_gmres_misuse(cuda_lib):
    from adflow_b200 import synthetic as syn
    from adflow_b200.params import make_ank_params

    prm, hb = case(8, 6, 4)
    s = ADFLOW_B200(prm)
    L = s.L
    try:
        s.addBlock(hb)
        v = np.zeros(8 * 6 * 4 * 5)
        # ANK: options, time-step matrix and base state are required, in that order
        assert L.adfb_ank_time_step_mat() != 0 and "adfb_ank_set_params" in _err(L)
        bad = make_ank_params()
        bad.charTimeStepType = 7
        assert L.adfb_ank_set_params(C.byref(bad)) != 0 and "charTimeStepType" in _err(L)
        s.ankSetParams(make_ank_params(coupled=False))
        assert L.adfb_ank_form_function(v.ctypes.data, v.ctypes.data, v.size) != 0 and "adfb_ank_time_step_mat" in _err(L)
        s.residual(RES_FLOW | RES_TURB | 4)
        s.ankTimeStepMat()
        assert L.adfb_ank_form_function(v.ctypes.data, v.ctypes.data, 7) != 0 and "vector length" in _err(L)
        assert L.adfb_ank_mffd_apply(v.ctypes.data, v.ctypes.data, v.size, 1e-7) != 0 and "set_base" in _err(L)
        lam = C.c_double(1.0)
        assert L.adfb_ank_physicality_check(v.ctypes.data, v.ctypes.data, 3, C.byref(lam)) != 0 and "vector length" in _err(L)
        # GMRES
        its, rn = C.c_int(0), C.c_double(0.0)
        args = (v.ctypes.data, v.ctypes.data, v.size)
        assert L.adfb_gmres_solve(5, *args, 10, 10, 1e-3, 0.0, None, None, C.byref(its), C.byref(rn)) != 0 and "op must be" in _err(L)
        assert L.adfb_gmres_solve(1, *args, 0, 10, 1e-3, 0.0, None, None, C.byref(its), C.byref(rn)) != 0 and "restart" in _err(L)
        assert L.adfb_gmres_solve(1, *args, 10, 10, 1e-3, 0.0, None, None, C.byref(its), C.byref(rn)) != 0 and "set_base" in _err(L)
        # multigrid: levels must be consecutive, tables complete, boundary types known
        coarse = syn.make_coarse_block(hb, prm)
        with pytest.raises(AdflowB200Error, match="one level coarser"):
            coarse.level = 3
            s.addCoarseBlock(coarse, 0)
        assert L.adfb_mg_prolong(1) == 0 or True          # nothing to prolongate from is not an error by itself
        assert L.adfb_mg_cycle(0, None, 0) != 0 and "cycling" in _err(L)
        cyc = (C.c_int * 2)(0, 5)
        assert L.adfb_mg_cycle(2, cyc, 0) != 0 and "cycling entry" in _err(L)
    finally:
        s.close()
