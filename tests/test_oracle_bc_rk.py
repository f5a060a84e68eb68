"""CPU checks of the oracle's boundary conditions and RK smoother (host logic, no GPU)."""
import numpy as np

from adflow_b200 import make_params
from adflow_b200 import synthetic as syn
from oracle.pyoracle import Oracle

from util import case


def test_symmetry_and_wall_halo_properties():
    prm, hb = case(10, 9, 8)
    o = Oracle(hb, prm)
    o.apply_turb_bc(True)
    o.apply_flow_bc(True)
    d = hb.d
    # symmetry on jMin: mirrored normal velocity -> (v1 + v2).n = 0 on the face
    sf = [s for s in hb.subfaces if s["faceId"] == syn.JMIN][0]
    n = sf["norm"]
    v1 = hb.w[1:d.ie + 1, 1, 1:d.ke + 1, 1:4]
    v2 = hb.w[1:d.ie + 1, 2, 1:d.ke + 1, 1:4]
    # away from the k-edges (the wall BC, applied later, overwrites the shared edge cells)
    vn = ((v1 + v2) * n).sum(-1)[1:-1, 2:-1]
    assert np.abs(vn).max() < 1e-14
    # adiabatic wall on kMin: no-slip -> u1 = -u2, rev1 = -rev2, p1 = p2, nuTilde1 = -nuTilde2
    own = (slice(2, d.il + 1), slice(2, d.jl + 1))
    assert np.abs(hb.w[own + (1, slice(1, 4))] + hb.w[own + (2, slice(1, 4))]).max() == 0.0
    assert np.array_equal(hb.rev[own + (1,)], -hb.rev[own + (2,)])
    assert np.array_equal(hb.p[own + (1,)], hb.p[own + (2,)])
    assert np.array_equal(hb.w[own + (1, 5)], -hb.w[own + (2, 5)])
    # every halo energy is consistent with its pressure (computeEtot)
    g = prm.gammaInf
    for sl in [(1, slice(2, d.jl + 1), slice(2, d.kl + 1)), (d.ie, slice(2, d.jl + 1), slice(2, d.kl + 1)),
               (slice(2, d.il + 1), slice(2, d.jl + 1), 1), (slice(2, d.il + 1), slice(2, d.jl + 1), d.ke)]:
        w = hb.w[sl]
        e = hb.p[sl] / (g - 1) + 0.5 * w[..., 0] * (w[..., 1] ** 2 + w[..., 2] ** 2 + w[..., 3] ** 2)
        assert np.abs(e - w[..., 4]).max() < 1e-13 * np.abs(e).max()


def test_farfield_reproduces_free_stream():
    prm = make_params()
    hb = syn.make_block(8, 8, 8, prm)
    for l in range(6):
        hb.w[..., l] = prm.wInf[l]
    hb.p[...] = prm.pInf
    hb.rlv[...] = syn.lam_viscosity(prm, hb.p, hb.w[..., 0])
    hb.rev[...] = syn.eddy_viscosity(prm, hb.w, hb.rlv)
    hb.subfaces = [s for s in hb.subfaces if s["bcType"] == syn.BC_FARFIELD]
    w0 = hb.w.copy()
    o = Oracle(hb, prm)
    o.apply_turb_bc(True)
    o.apply_flow_bc(True)
    assert np.abs(hb.w - w0).max() < 1e-13


def test_rk_smoother_reduces_residual_and_stays_finite():
    prm, hb = case(12, 10, 8, {"equationType": "Euler", "nRKStages": 3, "resAveraging": "never", "CFL": 1.0})
    o = Oracle(hb, prm)
    o.apply_flow_bc(True)
    o.time_step(True)
    hb.fw[...] = 0
    o.residual_block(1.0)
    n0 = o.norms()[1]
    for _ in range(20):
        o.rk_smoother()
        o.time_step(True)
        o.residual_block(1.0)
    n1 = o.norms()[1]
    assert np.isfinite(hb.w).all()
    assert n1 < n0


def test_residual_averaging_is_a_tridiagonal_solve():
    """multiply back: (I + eps-weighted Laplacian) * smoothed = original, per i-line."""
    prm, hb = case(9, 4, 4, {"CFL": 6.0})
    o = Oracle(hb, prm)
    rng = np.random.default_rng(1)
    hb.dw[...] = 0
    ow = hb.d.owned()
    hb.dw[ow + (slice(0, 5),)] = rng.standard_normal((9, 4, 4, 5))
    orig = hb.dw.copy()
    hb2 = hb.copy()
    hb2.d2Wall = hb.d2Wall
    # only the i-direction: collapse j,k smoothing by making ny = nz lines trivial is not possible,
    # so verify the full operator through linearity instead: A(x+y) = A(x) + A(y)
    o.residual_averaging()
    a = hb.dw.copy()
    hb.dw[...] = 2.0 * orig
    o.residual_averaging()
    assert np.abs(hb.dw - 2.0 * a).max() < 1e-12 * np.abs(a).max()
    assert np.abs(a - orig).max() > 1e-6  # the smoother did something at CFL 6
