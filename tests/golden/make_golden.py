"""Regenerates tests/golden/oracle_golden.json.

These are ORACLE-generated regression vectors (norms/checksums of the CPU restatement on the
seeded synthetic blocks), NOT outputs of the reference executable: the reference (Fortran + MPI + PETSc + CGNS) cannot be
built as a whole here and its regression meshes are absent.  The oracle that produces them is
itself pinned bit-exact against the reference's own routines (translated to C, oracle/_ref;
tests/test_oracle_vs_reference*.py).  They freeze the oracle so that an accidental change to it, to
the synthetic generator or to the option handling is caught by `-m "not gpu"`, and they give the
GPU tests a second, file-based comparison target.

    python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CASES = {
    "rans_scalar": (None, (12, 10, 8)),
    "euler_scalar": ({"equationType": "Euler"}, (12, 10, 8)),
    "laminar_scalar": ({"equationType": "laminar NS"}, (10, 9, 7)),
    "rans_matrix": ({"discretization": "central plus matrix dissipation"}, (11, 9, 8)),
    "rans_upwind": ({"discretization": "upwind"}, (11, 9, 8)),
    "euler_upwind_minmod": ({"discretization": "upwind", "limiter": "minmod", "equationType": "Euler"}, (9, 9, 9)),
    "rans_qcr_2ndturb": ({"useQCR": True, "turbulenceOrder": "second order"}, (10, 8, 9)),
}


def compute(name):
    from oracle.pyoracle import Oracle
    from util import FLOW, TURB, case

    opts, shape = CASES[name]
    prm, hb = case(*shape, opts)
    o = Oracle(hb, prm)
    o.pressure(False); o.lam_viscosity(False); o.eddy_viscosity(False)
    o.apply_turb_bc(True); o.apply_flow_bc(True)
    flags = FLOW | (TURB if prm.equations == 3 else 0)
    o.residual_core(flags)
    ow = hb.d.owned()
    out = {"shape": list(shape), "options": opts or {}, "nw": hb.nw}
    out["state_norm"] = float(np.linalg.norm(hb.w[ow]))                      # "Norm of state vector"
    out["res_norms"] = [float(x) for x in o.norms()]                          # sum (dw/vol)^2
    out["dw_l2"] = [float(np.linalg.norm(hb.dw[ow + (l,)])) for l in range(hb.nw)]
    out["dw_sum"] = [float(hb.dw[ow + (l,)].sum()) for l in range(hb.nw)]
    out["dtl_sum"] = float(hb.dtl[ow].sum())
    # one RK cycle / one DADI step / one SA solve on top
    if prm.spaceDiscr == 1:
        h2 = hb.copy()
        o2 = Oracle(h2, prm)
        h2.fw[...] = 0
        o2.residual_block(1.0)
        o2.dadi_step()
        out["dadi_state_norm"] = float(np.linalg.norm(h2.w[ow][..., :5]))
        if prm.equations == 3:
            o2.sa_block()
            out["sa_nutilde_norm"] = float(np.linalg.norm(h2.w[ow + (5,)]))
    return out


if __name__ == "__main__":
    gold = {k: compute(k) for k in CASES}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_golden.json")
    json.dump(gold, open(path, "w"), indent=1, sort_keys=True)
    print("wrote", path)
