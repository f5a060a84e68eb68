"""Regenerates tests/golden/reference_golden.json: vectors produced by the REFERENCE's own routines (translated
Fortran -> C where the source lies, oracle/_ref/libblockette_ref.so -- only available where /root/reference was
present at build time) on seeded synthetic blocks.  Committed so that the oracle and the CUDA path can be checked
against reference outputs even where oracle/_ref is absent.

Per case: checksums (sum, L2 norm) and sampled entries of
  * dw of blocketteResCore (src/NKSolver/blockette.F90:299-753) after the reference's own BC routines,
  * the state after one RungeKuttaSmoother (src/solver/smoothers.F90:4-86),
  * wr / coarse w after transferToCoarseGrid (src/solver/multiGrid.F90:5-324),
  * the ANK time-step block of one cell (NKSolvers.F90:2116-2329),
  * the state after executeDADIStep (smoothers.F90:425-693), nuTilde after sa_block (sa.F90:16-86, RANS) and the wall
    forces of wallIntegrationFace (surfaceIntegrations.F90:406-881).

    python tests/golden/make_reference_golden.py
"""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CASES = {
    "rans": (None, (12, 8, 10)),
    "euler": ({"equationType": "Euler", "nRKStages": 3, "resAveraging": "never"}, (8, 6, 6)),
    "laminar_matrix_coarse": ({"equationType": "laminar NS", "coarseDiscretization": "central plus matrix dissipation"}, (10, 8, 6)),
}
SAMPLES = [(2, 2, 2), (3, 4, 2), (5, 3, 4)]


def stats(a):
    a = np.asarray(a, dtype=np.float64)
    return {"sum": float(a.sum()), "l2": float(np.sqrt((a * a).sum()))}


def setup(name):
    from adflow_b200 import synthetic as syn
    from util import case

    opts, shape = CASES[name]
    prm, fine = case(*shape, opts)
    fine.subfaces.sort(key=lambda s_: 0 if s_["bcType"] in (2, 6) else 1)   # the reference's subface order (viscous first)
    coarse = syn.make_coarse_block(fine, prm)
    return prm, fine, coarse


def entries(a, nvar):
    return [[float(a[i, j, k, l]) for l in range(nvar)] for (i, j, k) in SAMPLES]


def compute_reference(name):
    """the reference's routines (needs oracle/_ref)"""
    from oracle import refblockette as rb
    from adflow_b200.params import make_ank_params

    prm, fine, coarse = setup(name)
    out = {"shape": list(CASES[name][1]), "options": CASES[name][0] or {}}
    ow = fine.d.owned()
    # BCs (reference) -> blocketteResCore
    r = rb.call(fine, prm, "bcroutines_applyallbc_block", 1)
    hb = fine.copy()
    for k_, n_ in (("w", "w"), ("p", "p"), ("rlv", "rlv"), ("rev", "rev")):
        getattr(hb, n_)[...] = r.a[k_]
    flags = 8 | (16 if prm.equations == 3 else 0)
    rc = rb.residual_core(hb, prm, flags)
    nv = hb.nw if prm.equations == 3 else 5
    out["core_dw"] = [stats(rc.a["dw"][ow + (l,)]) for l in range(nv)]
    out["core_dw_samples"] = entries(rc.a["dw"], nv)
    # multigrid: transferToCoarseGrid, RungeKuttaSmoother on level 2, transferToFineGrid
    mg = rb.RefMG(hb.copy(), coarse.copy(), prm)
    try:
        mg.transfer_to_coarse()
        rcoarse, rfine = mg.lv[2].a, mg.lv[1].a
        owc = coarse.d.owned()
        out["mg_wr"] = [stats(rcoarse["wr"][owc + (l,)]) for l in range(5)]
        out["mg_coarse_w"] = [stats(rcoarse["w"][owc + (l,)]) for l in range(5)]
        mg.call(2, "smoothers_rungekuttasmoother")
        out["mg_coarse_w_after_rk"] = [stats(rcoarse["w"][owc + (l,)]) for l in range(5)]
        mg.transfer_to_fine()
        out["mg_fine_w_after_prolong"] = [stats(rfine["w"][ow + (l,)]) for l in range(5)]
        out["mg_fine_w_samples"] = entries(rfine["w"], 5)
    finally:
        mg.close()
    # a whole 2V cycle run by the reference's executeMGCycle (RK smoother), from the block-path residual of the fine level
    from adflow_b200.solver import ADFLOW_B200
    hm = hb.copy()
    rt = rb.call(hm, prm, "solverutils_timestep_block", 0)
    for k_, n_ in (("dtl", "dtl"), ("radi", "radI"), ("radj", "radJ"), ("radk", "radK")):
        getattr(hm, n_)[...] = rt.a[k_]
    hm.fw[...] = 0; hm.dw[...] = 0
    rr_ = rb.call(hm, prm, "residuals_residual_block", rkstage=0)
    for k_, n_ in (("dw", "dw"), ("fw", "fw"), ("aa", "aa")):
        getattr(hm, n_)[...] = rr_.a[k_]
    mg2 = rb.RefMG(hm, coarse.copy(), prm)
    try:
        mg2.execute_mg_cycle(ADFLOW_B200.cycleStrategy("2v"))
        out["mg_cycle_2v_w"] = [stats(mg2.lv[1].a["w"][ow + (l,)]) for l in range(hb.nw)]
    finally:
        mg2.close()
    # one DADI step from the block-path residual; SA solve; wall forces (all on the state after the reference's BCs)
    hd = hb.copy()
    r3 = rb.call(hd, prm, "solverutils_timestep_block", 0)
    for k_, n_ in (("dtl", "dtl"), ("radi", "radI"), ("radj", "radJ"), ("radk", "radK")):
        getattr(hd, n_)[...] = r3.a[k_]
    hd.fw[...] = 0
    hd.dw[...] = 0
    rb.set_int("smoother", 2)
    try:
        r4 = rb.call(hd, prm, "residuals_residual_block", rkstage=0)
        for k_, n_ in (("dw", "dw"), ("fw", "fw"), ("aa", "aa")):
            getattr(hd, n_)[...] = r4.a[k_]
        r5 = rb.call(hd, prm, "smoothers_executedadistep", rkstage=0)
    finally:
        rb.set_int("smoother", 1)
    out["dadi_w"] = [stats(r5.a["w"][ow + (l,)]) for l in range(5)]
    if prm.equations == 3:
        r6 = rb.call(hb.copy(), prm, "sa_sa_block", 0)
        out["sa_nutilde"] = stats(r6.a["w"][ow + (5,)])
        out["sa_nutilde_samples"] = [float(r6.a["w"][i, j, k, 5]) for (i, j, k) in SAMPLES]
    # ANK time-step block ('None' and 'VLR') of one cell, dtl from the reference's timeStep_block
    r2 = rb.call(hb, prm, "solverutils_timestep_block", 0)
    h3 = hb.copy()
    h3.dtl[...] = r2.a["dtl"]; h3.aa[...] = prm.gammaInf * h3.p / h3.w[..., 0]
    from test_oracle_vs_reference_ank import bind_ank
    for kind in ("None", "VLR"):
        ank = make_ank_params(cfl=4.0, coupled=False, char_time_step=kind, cflLimit=40.0)
        rb.set_params(prm, h3.nw)
        rr = rb.RefBlock(h3, prm)
        rr.bind()
        bind_ank(ank, 5)
        blk = np.zeros((5, 5), order="F")
        i, j, k = SAMPLES[1]
        rb.lib().anksolver_computetimestepblock(C.byref(C.c_int(i)), C.byref(C.c_int(j)), C.byref(C.c_int(k)), blk.ctypes.data_as(C.c_void_p))
        out["ank_block_" + kind] = [[float(v) for v in row] for row in blk]
    return out


if __name__ == "__main__":
    gold = {k: compute_reference(k) for k in CASES}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_golden.json")
    json.dump(gold, open(path, "w"), indent=1, sort_keys=True)
    print("wrote", path)
