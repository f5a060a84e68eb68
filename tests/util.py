"""Shared helpers for the parity tests (oracle = checker, CUDA path = product)."""
import numpy as np

from adflow_b200 import make_params
from adflow_b200 import synthetic as syn

FLOW, TURB, SKIP = 8, 16, 64


def rel_l2(a, b):
    """||a-b||_2 / ||b||_2 (L2 norm of the dw difference / L2 norm of dw)."""
    nb = np.linalg.norm(b.ravel())
    return np.linalg.norm((a - b).ravel()) / (nb if nb > 0 else 1.0)


def rel_max(a, b):
    mb = np.abs(b).max()
    return np.abs(a - b).max() / (mb if mb > 0 else 1.0)


def case(nx, ny, nz, options=None, seed=314, **kw):
    prm = make_params(options)
    hb = syn.make_block(nx, ny, nz, prm, seed=seed, **kw)
    return prm, hb


def oracle_residual(prm, hb, flags=FLOW | TURB, rfil=1.0):
    from oracle.pyoracle import Oracle

    ho = hb.copy()
    Oracle(ho, prm).residual_core(flags, rfil)
    return ho


def oracle_form_function(prm, hb, wvec):
    """FormFunction_mf on one block with the oracle: setW (turbulence clip), blocketteRes
    (p/rlv/rev, BCs, whalo2's owned-cell etot, core), setRVec."""
    import ctypes as C

    from oracle.pyoracle import Oracle

    d = hb.d
    ow = d.owned()
    h2 = hb.copy()
    v = np.asarray(wvec).reshape(d.nz, d.ny, d.nx, hb.nw).transpose(2, 1, 0, 3).copy()
    if hb.nw > 5:
        v[..., 5] = np.maximum(1e-6 * prm.wInf[5], v[..., 5])
    h2.w[ow] = v
    o = Oracle(h2, prm)
    o.pressure(False); o.lam_viscosity(False); o.eddy_viscosity(False)
    o.apply_turb_bc(True); o.apply_flow_bc(True)
    o.L.orc_etot(C.byref(o.ob), C.byref(prm), 2, d.il, 2, d.jl, 2, d.kl)
    o.residual_core(FLOW | TURB)
    r = h2.dw[ow] / h2.volRef[ow][..., None]
    if hb.nw > 5:
        r[..., 5] *= prm.turbResScale
    return np.transpose(r, (2, 1, 0, 3)).reshape(-1)
