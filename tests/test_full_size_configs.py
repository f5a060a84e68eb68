"""The CUDA path at the block sizes of the other BASELINE.json configurations (test_full_size.py covers configs[1]):

  C1  48 x 48 x 48 Euler block, 3-stage Runge-Kutta cycle          (configs[0])
  C3  one 128 x 128 x 64 RANS-SA block: residual + one 4W cycle     (configs[2], per-GPU share)
  C4  two blocks on one GPU joined by an overset pattern, DADI step (configs[3] shape, k reduced to keep the oracle fast)
  C5  one 160 x 160 x 144 block: F(U) of the NK solver              (configs[4], per-GPU share)

Each against the oracle on the same seeded input; tolerances as in test_full_size.py (1e-12 residuals, 1e-10 / 1e-9 on
state changes over a smoother cycle)."""
import ctypes as C
import time

import numpy as np
import pytest

from adflow_b200 import make_params
from adflow_b200 import synthetic as syn
from adflow_b200.halo import build_overset_pattern, exchange_numpy_overset
from adflow_b200.solver import ADFLOW_B200, RES_FLOW, RES_TURB
from oracle.pyoracle import Oracle

from util import case, rel_l2, rel_max

pytestmark = pytest.mark.gpu


def test_c1_euler_three_stage_rk_cycle(cuda_lib):
    prm, hb = case(48, 48, 48, {"equationType": "Euler", "nRKStages": 3, "resAveraging": "never"})
    ho = hb.copy()
    o = Oracle(ho, prm)
    o.apply_flow_bc(True)
    o.time_step(True)
    ho.fw[...] = 0
    o.residual_block(prm.cdisRK[0])
    o.rk_smoother()
    s = ADFLOW_B200(prm)
    try:
        s.addBlock(hb)
        s.applyBCs(True, False)
        s.timeStep(False)
        s.smootherResidual(0)
        s.rkCycle()
        w, p, _, _ = s.downloadState(0)
    finally:
        s.close()
    ow = hb.d.owned()
    for l in range(5):
        a, b = w[ow + (l,)] - hb.w[ow + (l,)], ho.w[ow + (l,)] - hb.w[ow + (l,)]
        assert np.abs(b).max() > 0
        assert rel_l2(a, b) < 1e-10, (l, rel_l2(a, b))
    assert rel_max(p, ho.p) < 1e-11


def test_c3_block_residual_and_4w_cycle(cuda_lib):
    from test_mg_gpu import device, make_levels, oracle_mg_cycle, prepare_fine

    shape = (128, 128, 64)
    prm, levels = make_levels(shape, {"nRKStages": 5, "resAveraging": "never"}, 4)
    fine = levels[0]
    # --- full residual (blocketteRes) of the block
    hr = fine.copy()
    orr = Oracle(hr, prm)
    orr.pressure(False); orr.lam_viscosity(False); orr.eddy_viscosity(False)
    orr.apply_turb_bc(True); orr.apply_flow_bc(True)
    orr.residual_core(RES_FLOW | RES_TURB)
    dev_levels = [l.copy() for l in levels]
    s = device(prm, dev_levels)
    try:
        s.residual(RES_FLOW | RES_TURB)
        dw = s.downloadResidual(0)
        ow = fine.d.owned()
        for l in range(6):
            assert rel_l2(dw[ow + (l,)], hr.dw[ow + (l,)]) < 1e-12, (l, rel_l2(dw[ow + (l,)], hr.dw[ow + (l,)]))
        # --- one 4W cycle (executeMGCycle, multiGrid.F90:825), 5-stage RK on every level
        s.uploadState(0, dev_levels[0])
        cyc = ADFLOW_B200.cycleStrategy("4w")
        prepare_fine(Oracle(fine, prm))
        f0 = fine.w.copy()
        t0 = time.time()
        oracle_mg_cycle(prm, levels, cyc)
        t_or = time.time() - t0
        s.applyBCs(True, True)
        s.timeStep(False)
        s.smootherResidual(0)
        s.mgCycle(cyc)
        w, p, rlv, rev = s.downloadState(0)
        for l in range(5):
            a, b = w[ow + (l,)] - f0[ow + (l,)], fine.w[ow + (l,)] - f0[ow + (l,)]
            assert np.abs(b).max() > 0
            assert rel_l2(a, b) < 1e-8, (l, rel_l2(a, b), t_or)
        assert rel_max(p, fine.p) < 1e-10
    finally:
        s.close()


def test_c4_two_blocks_overset_dadi_step(cuda_lib):
    from test_overset_host import VARS, overset_entries

    n0, n1 = (96, 96, 12), (96, 96, 10)
    opts = {"smoother": "DADI", "resAveraging": "never"}
    prm = make_params(opts)
    blocks = [syn.make_block(*n0, prm, seed=11), syn.make_block(*n1, prm, seed=12)]
    pat = build_overset_pattern(overset_entries(n0, n1))
    ref = [b.copy() for b in blocks]
    orcs = [Oracle(b, prm) for b in ref]
    for o, b in zip(orcs, ref):
        o.apply_turb_bc(True); o.apply_flow_bc(True)
        o.time_step(True)
        b.fw[...] = 0
        o.residual_block(1.0)
        o.dadi_step()
    # executeDADIStep ends with whalo2: here the overset interpolation + computeEtotBlock on the owned cells
    exchange_numpy_overset(ref, pat, VARS)
    for o, b in zip(orcs, ref):
        d = b.d
        o.L.orc_etot(C.byref(o.ob), C.byref(prm), 2, d.il, 2, d.jl, 2, d.kl)
    s = ADFLOW_B200(prm)
    try:
        for hb in blocks:
            s.addBlock(hb)
        s.setOversetPattern(pat)
        s.applyBCs(True, True)
        s.timeStep(False)
        s.smootherResidual(0)
        s.dadiStep()
        for q, b in enumerate(ref):
            w, p, _, _ = s.downloadState(q)
            ow = b.d.owned()
            for l in range(5):
                a, c = w[ow + (l,)] - blocks[q].w[ow + (l,)], b.w[ow + (l,)] - blocks[q].w[ow + (l,)]
                assert np.abs(c).max() > 0
                assert rel_l2(a, c) < 1e-9, (q, l, rel_l2(a, c))
            assert rel_max(p, b.p) < 1e-10
    finally:
        s.close()


def test_c5_block_form_function(cuda_lib):
    from test_mffd import state_vec
    from util import oracle_form_function

    prm, hb = case(160, 160, 144)
    U = state_vec(hb)
    F0 = oracle_form_function(prm, hb, U)
    s = ADFLOW_B200(prm)
    try:
        s.addBlock(hb)
        r = s.formFunction(U)
    finally:
        s.close()
    assert np.isfinite(r).all()
    assert rel_l2(r, F0) < 1e-12, rel_l2(r, F0)
