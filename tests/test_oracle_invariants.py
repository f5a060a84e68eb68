"""Mesh-independent invariants that pin the CPU oracle.

Independent of (and older than) the bit-exact comparison with the translated reference
routines in tests/test_oracle_vs_reference*.py, the oracle is held to the invariants the
reference's own tests rely on: free-stream preservation
(getFreeStreamResidual, src/NKSolver/NKSolvers.F90:303-320), discrete conservation
of the telescoping face-flux scatter (src/solver/fluxes.F90:103-104) and agreement
of the independently written numpy metrics with the C restatement."""
import numpy as np
import pytest

from adflow_b200 import make_params
from adflow_b200 import synthetic as syn
from oracle.pyoracle import Oracle

from util import FLOW, TURB, case


def freestream_block(nx, ny, nz, options=None):
    prm = make_params(options)
    hb = syn.make_block(nx, ny, nz, prm)
    hb.porI[...] = 1
    hb.porJ[...] = 1
    hb.porK[...] = 1
    for l in range(hb.nw):
        hb.w[..., l] = prm.wInf[l]
    hb.p[...] = prm.pInf
    if prm.equations != 1:
        hb.rlv[...] = syn.lam_viscosity(prm, hb.p, hb.w[..., 0])
    if prm.equations == 3:
        hb.rev[...] = syn.eddy_viscosity(prm, hb.w, hb.rlv)
    return prm, hb


def test_metrics_and_volume_numpy_vs_c():
    prm, hb = case(12, 10, 9)
    ref = {n: getattr(hb, n).copy() for n in ("si", "sj", "sk", "vol")}
    hb.si[...] = 0; hb.sj[...] = 0; hb.sk[...] = 0; hb.vol[...] = 0
    o = Oracle(hb, prm)
    o.metrics()
    o.volume()
    for n in ("si", "sj", "sk"):
        assert np.abs(getattr(hb, n) - ref[n]).max() == 0.0
    assert np.abs(hb.vol - ref["vol"]).max() <= 1e-15 * ref["vol"].max()
    assert hb.vol[1:-1, 1:-1, 1:-1].min() > 0


def test_closed_cell_normals_sum_to_zero():
    prm, hb = case(8, 7, 6)
    d = hb.d
    ow = d.owned()
    im = (slice(1, d.il), ow[1], ow[2])
    jm = (ow[0], slice(1, d.jl), ow[2])
    km = (ow[0], ow[1], slice(1, d.kl))
    s = hb.si[ow] - hb.si[im] + hb.sj[ow] - hb.sj[jm] + hb.sk[ow] - hb.sk[km]
    assert np.abs(s).max() < 1e-15 * np.abs(hb.si[ow]).max() * 10


def test_free_stream_preservation_rans():
    prm, hb = freestream_block(16, 12, 8)
    Oracle(hb, prm).residual_core(FLOW | TURB)
    ow = hb.d.owned()
    scale = np.abs(hb.si[ow]).max() * prm.wInf[1] * prm.wInf[4]
    for l in range(5):
        assert np.abs(hb.dw[ow + (l,)]).max() < 1e-13 * scale


def test_free_stream_preservation_euler():
    prm, hb = freestream_block(10, 9, 8, {"equationType": "Euler"})
    Oracle(hb, prm).residual_core(FLOW)
    ow = hb.d.owned()
    scale = np.abs(hb.si[ow]).max() * prm.wInf[1] * prm.wInf[4]
    for l in range(5):
        assert np.abs(hb.dw[ow + (l,)]).max() < 1e-13 * scale


def test_discrete_conservation_inviscid():
    """Sum over owned cells of the inviscid+dissipative dw telescopes to the boundary faces."""
    prm, hb = case(9, 8, 7, {"equationType": "Euler"})
    o = Oracle(hb, prm)
    hb.dw[...] = 0
    o.call("orc_central_flux", None)
    d = hb.d
    ow = d.owned()
    tot = hb.dw[ow + (0,)].sum()
    # mass flux through the six boundary faces, recomputed independently in numpy
    def massflux(s, por, cm, cp):
        vnp = (hb.w[cp + (slice(1, 4),)] * s).sum(-1)
        vnm = (hb.w[cm + (slice(1, 4),)] * s).sum(-1)
        pv = np.where(por == 0, 0.0, 1.0) * np.where(por == -1, 0.0, 0.5)
        return vnp * pv * hb.w[cp + (0,)] + vnm * pv * hb.w[cm + (0,)]
    J, K, I = ow[1], ow[2], ow[0]
    f = 0.0
    f += massflux(hb.si[d.il, J, K], hb.porI[d.il, J, K], (d.il, J, K), (d.ie, J, K)).sum()
    f -= massflux(hb.si[1, J, K], hb.porI[1, J, K], (1, J, K), (2, J, K)).sum()
    f += massflux(hb.sj[I, d.jl, K], hb.porJ[I, d.jl, K], (I, d.jl, K), (I, d.je, K)).sum()
    f -= massflux(hb.sj[I, 1, K], hb.porJ[I, 1, K], (I, 1, K), (I, 2, K)).sum()
    f += massflux(hb.sk[I, J, d.kl], hb.porK[I, J, d.kl], (I, J, d.kl), (I, J, d.ke)).sum()
    f -= massflux(hb.sk[I, J, 1], hb.porK[I, J, 1], (I, J, 1), (I, J, 2)).sum()
    mag = np.abs(hb.dw[ow + (0,)]).sum()
    assert abs(tot - f) < 1e-12 * mag


def test_residual_is_finite_and_deterministic():
    prm, hb = case(11, 9, 7)
    a = hb.copy(); b = hb.copy()
    Oracle(a, prm).residual_core(FLOW | TURB)
    Oracle(b, prm).residual_core(FLOW | TURB)
    assert np.isfinite(a.dw).all()
    assert np.array_equal(a.dw, b.dw)


def _mg_levels(prm, fine, nlev):
    levels = [fine]
    for _ in range(nlev - 1):
        levels.append(syn.make_coarse_block(levels[-1], prm))
    return levels


def test_multigrid_restriction_is_conservative_and_prolongation_reproduces_constants():
    """transferToCoarseGrid: the restricted state is the volume-weighted mean (so volume integrals are kept exactly on
    regular coarsening) and wr sums the fine residuals; transferToFineGrid: a constant correction is interpolated to
    the same constant (the 27/9/3/1 weights sum to 64)"""
    prm, fine = case(8, 6, 4, {"equationType": "Euler"})
    coarse = syn.make_coarse_block(fine, prm)
    of, oc = Oracle(fine, prm), Oracle(coarse, prm)
    rng = np.random.default_rng(4)
    fine.dw[...] = rng.standard_normal(fine.dw.shape)
    oc.mg_restrict(of)
    ow, owc = fine.d.owned(), coarse.d.owned()
    for l in range(4):   # density and the primitive velocities: volume-weighted averages
        fi = (fine.vol[ow] * fine.w[ow + (l,)]).sum()
        cv = sum(fine.vol[ow][a::2, b::2, c::2] for a in (0, 1) for b in (0, 1) for c in (0, 1))
        assert abs((cv * coarse.w[owc + (l,)]).sum() - fi) <= 1e-13 * abs(fi), l
    for l in range(5):   # restricted residual = sum of the 8 fine residuals (weights 1 on regular coarsening)
        assert abs(coarse.wr[owc + (l,)].sum() - fine.dw[ow + (l,)].sum()) <= 1e-12 * np.abs(fine.dw[ow + (l,)]).sum(), l
    # constant correction
    prm.mgBoundCorr = 1   # bcNeumann0: the boundary halos of the corrections copy the interior value
    np.copyto(coarse.w1, coarse.w[..., :5]); np.copyto(coarse.p1, coarse.p)
    delta = np.array([1e-3, 2e-3, -1e-3, 5e-4])
    for l in range(4):
        coarse.w[..., l] += delta[l]
    coarse.p[...] += 7e-3
    for s_ in coarse.subfaces:    # no mirroring: treat every face like a generic boundary
        s_["bcType"] = 3
    w0, p0 = fine.w.copy(), fine.p.copy()
    of.mg_prolong(oc)
    for l in range(4):
        assert np.abs((fine.w[ow + (l,)] - w0[ow + (l,)]) - delta[l]).max() < 1e-15, l
    assert np.abs((fine.p[ow] - p0[ow]) - 7e-3).max() < 1e-14


@pytest.mark.parametrize("cycle,nlev", [("2v", 2), ("3w", 3)])
def test_multigrid_cycle_leaves_a_converged_solution_converged(cycle, nlev):
    """free stream on a warped mesh (far field all round): the fine residual is zero to round-off, so the residual
    forcing term cancels the coarse residual and every level's update is round-off: the cycle must not disturb it"""
    from adflow_b200.solver import ADFLOW_B200

    prm = make_params({"equationType": "Euler", "nRKStages": 3, "resAveraging": "never"})
    fine = syn.make_block(8, 8, 8, prm, physical_faces={f: 3 for f in range(1, 7)})
    for l in range(fine.nw):
        fine.w[..., l] = prm.wInf[l]
    fine.p[...] = prm.pInf
    levels = _mg_levels(prm, fine, nlev)
    o = Oracle(fine, prm)
    o.apply_flow_bc(True)
    o.time_step(True); fine.fw[...] = 0; o.residual_block(prm.cdisRK[0])
    w0 = fine.w.copy()
    lv = 0
    cyc = ADFLOW_B200.cycleStrategy(cycle)
    for n, c in enumerate(cyc):     # executeMGCycle with the oracle's pieces
        if c == -1:
            lv -= 1
            of, oc = Oracle(levels[lv], prm), Oracle(levels[lv + 1], prm)
            of.mg_prolong(oc); of.apply_flow_bc(lv == 0)
        elif c == 0:
            ol = Oracle(levels[lv], prm)
            if n > 0 and cyc[n - 1] != 1:
                ol.time_step(True); ol.residual_block(prm.cdisRK[0])
            ol.rk_smoother()
        else:
            of, oc = Oracle(levels[lv], prm), Oracle(levels[lv + 1], prm)
            of.time_step(False); of.residual_block(prm.cdisRK[0])
            oc.mg_restrict(of); oc.apply_flow_bc(False); oc.time_step(True); oc.mg_store_w1()
            oc.residual_block_coarse(prm.cdisRK[0], init=0); oc.mg_forcing()
            lv += 1
    ow = fine.d.owned()
    assert np.abs(fine.w[ow] - w0[ow]).max() < 1e-11 * np.abs(w0[ow]).max()
