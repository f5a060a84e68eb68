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
