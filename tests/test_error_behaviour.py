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
