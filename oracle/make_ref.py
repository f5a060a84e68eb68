#!/usr/bin/env python
"""make_ref.py -- generate oracle/_ref/blockette_ref.c from the reference's own Fortran source
(TEST INFRASTRUCTURE ONLY; output is git-ignored and never committed).

    python oracle/make_ref.py [/root/reference]

Reads `<ref>/src/NKSolver/blockette.F90` where it lies, translates blocketteResCore and every
routine it calls to C with oracle/f90toc.py, and writes oracle/_ref/blockette_ref.c.  The
few statements that touch data structures outside the hot path are neutralised by the textual
patches listed in PATCHES below (each one named, none of them changes arithmetic on the path
exercised by the parity tests: steady, non-rotating, no overset, no wall-tensor storage).
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import f90toc  # noqa: E402

ONLY = ["blockettecore_placeholder"]
ROUTINES = [
    "blocketterescore", "metrics", "initres", "sasource", "saviscous", "saadvection", "saresscale", "timestep",
    "inviscidcentralflux", "inviscibdissfluxmatrix", "inviscidissfluxmatrix", "inviscidDissFluxMatrix".lower(),
    "inviscidDissFluxScalar".lower(), "inviscidUpwindFlux".lower(), "inviscidDissFluxScalarApprox".lower(),
    "inviscidDissFluxMatrixApprox".lower(), "computeSpeedOfSoundSquared".lower(), "allNodalGradients".lower(),
    "viscousFlux".lower(), "viscousFluxApprox".lower(), "sumDwandFw".lower(), "resScale".lower(),
]

# (regex, replacement) applied to every pre-processed (lower-cased, continuation-joined) line
PATCHES = [
    # rotating-frame rates come from derived types outside the path: steady non-rotating => 0
    (r"sections\(sectionid\)%rotrate\(\d\)", "zero"),
    (r"sections\(sectionid\)%timeperiod", "one"),
    (r"cgnsdoms\(nbkglobal\)%rotrate\(\d\)", "zero"),
    # rotational periodicity matrices: not on the path (pointer never associated)
    (r"associated\(rotmatrix[ijk]\)", ".false."),
    (r"rotmatrix\([^()]*\)", "zero"),
    # wall stress tensor storage into viscSubface (storeWallTensor): surface-force path, out of scope
    (r"^viscsubface\(.*$", "continue"),
    (r"visc[ijk]m(in|ax)pointer\([^()]*\)", "0"),
    # sa_block: the turbulence BC treatment (bmt matrices in, halo values out) is done by the harness;
    # steady flow, so the unsteady term is identically absent (turbUtils.F90:456-460 returns at once)
    (r"^call bcturbtreatment$", "continue"),
    (r"^call applyallturbbcthisblock\(.*$", "continue"),
    (r"^call unsteadyturbterm\(.*$", "continue"),
    # saSolve: wall-function branch (wallFunctions = .false. on the path) uses BCData/viscSubface
    ("block", r"^testwallfunctions: if", r"^end if testwallfunctions"),
    # module-wide `use X` without only-list inside routines: names resolve through ref_env.h
]

# external (other-module) data the translated routines see; declared in oracle/ref_env.h
ENV_INTS = """nw nwf nt1 nt2 equations equationmode turbmodel spacediscr ransequations nsequations eulerequations
 steady unsteady timespectral spalartallmaras dissscalar dissmatrix upwind currentlevel groundlevel
 irho ivx ivy ivz irhoe itu1 itu2 imx imy imz viscous addgridvelocities oversetpresent blockismoving
 useft2sa userotationsa turbprod useqcr approxsa secondord orderturb limiter precond riemann
 firstorder secondorder nolimiter vanalbeda minmod noprecond turkel choimerkle roe vanleer ausmdv
 strain vorticity katolaunder kpresent eddymodel rotationalperiodic correctfork righthanded
 usedisscontinuation nbkglobal sectionid ntimeintervalsspectral normalflux boundflux internalflux
 lumpeddiss fullturb cpmodel rkstage resaveraging ndom exchangepressureearly lowspeedpreconditioner
 noresaveraging alwaysresaveraging alternateresaveraging turbrelax turbrelaximplicit turbrelaxexplicit
 bp_nx bp_ny bp_nz bp_il bp_jl bp_kl bp_ie bp_je bp_ke bp_ib bp_jb bp_kb bp_addgridvelocities
 bp_righthanded bp_sectionid bp_blockismoving bp_nbkglobal""".split()

BOX3 = [("0", "(bp_ib + 1)"), ("0", "(bp_jb + 1)"), ("0", "(bp_kb + 1)")]


def box(ncomp=None):
    return BOX3 + ([("1", str(ncomp))] if ncomp else [])


def env_arrays():
    A = f90toc.Array
    arrs = {}
    for n in ["p", "gamma", "radi", "radj", "radk", "ux", "uy", "uz", "vx", "vy", "vz", "wx", "wy", "wz", "qx", "qy",
              "qz", "rlv", "rev", "vol", "volref", "d2wall", "shocksensor", "sfacei", "sfacej", "sfacek", "dtl", "aa"]:
        arrs["bp_" + n] = A("bp_" + n, "double", box())
    arrs["bp_w"] = A("bp_w", "double", box("nw"))
    arrs["bp_dw"] = A("bp_dw", "double", box("nw"))
    arrs["bp_fw"] = A("bp_fw", "double", box("nwf"))
    arrs["bp_wn"] = A("bp_wn", "double", box("nwf"))
    arrs["bp_pn"] = A("bp_pn", "double", box())
    # turbulence BC matrices of the six block faces (block.F90: bmti1(je,ke,nt1:nt2,nt1:nt2) ...)
    for n, (a, b) in {"bmti1": ("bp_je", "bp_ke"), "bmti2": ("bp_je", "bp_ke"), "bmtj1": ("bp_ie", "bp_ke"),
                      "bmtj2": ("bp_ie", "bp_ke"), "bmtk1": ("bp_ie", "bp_je"), "bmtk2": ("bp_ie", "bp_je")}.items():
        arrs["bp_" + n] = A("bp_" + n, "double", [("1", a), ("1", b), ("nt1", "1"), ("nt1", "1")])
    arrs["bp_scratch"] = A("bp_scratch", "double", box(10))
    arrs["etark"] = A("etark", "double", [("1", "6")])
    arrs["cdisrk"] = A("cdisrk", "double", [("1", "6")])
    arrs["coeftime"] = A("coeftime", "double", [("0", "8")])
    for n in ["x", "si", "sj", "sk"]:
        arrs["bp_" + n] = A("bp_" + n, "double", box(3))
    arrs["bp_iblank"] = A("bp_iblank", "int", box())
    for n in ["pori", "porj", "pork"]:
        arrs["bp_" + n] = A("bp_" + n, "int", box())
    for n in ["rotmatrixi", "rotmatrixj", "rotmatrixk"]:
        arrs["bp_" + n] = A("bp_" + n, "double", [("1", "1")])
    arrs["turbresscale"] = A("turbresscale", "double", [("1", "4")])
    return arrs


ENV_SUBS = {
    # name -> [(argname, ctype, isarray)] for external subroutines called by reference
    "terminate": [("routine", "str", False), ("msg", "str", False)],
    # driver-level procedures outside the translated set: no-op stubs in ref_env.c (single block, BCs and
    # halo exchange are applied by the test harness around the translated routine)
    "setpointers": [("nn", "int", False), ("level", "int", False), ("sps", "int", False)],
    "whalo1": [(n, "int", False) for n in ("level", "start", "end", "commpressure", "commgamma", "commviscous")],
    "whalo2": [(n, "int", False) for n in ("level", "start", "end", "commpressure", "commgamma", "commviscous")],
    "applyallbc": [("secondhalo", "int", False)],
}


# (source file relative to <ref>/src, C prefix of its procedures, routines to translate)
UNITS = [
    # USE_TAPENADE (the reference's own switch for its AD-differentiable subset) drops the cp-curve-fit
    # branches of eint/computeEtotBlock, which need tables outside the path (cpModel == cpConstant here)
    ("utils/flowUtils.F90", "flowutils_", ["computeetotblock", "computelamviscosity", "computepressuresimple",
                                           "etot", "eint"], ("USE_TAPENADE",)),
    ("NKSolver/blockette.F90", "", ROUTINES, ()),
    ("turbulence/turbUtils.F90", "turbutils_", ["computeeddyviscosity", "saeddyviscosity", "turbadvection"], ()),
    ("turbulence/sa.F90", "sa_", ["sa_block", "sasource", "saviscous", "saresscale", "sasolve"], ()),
    ("solver/residuals.F90", "residuals_", ["residualaveraging", "computedwdadi", "tridiagsolve"], ()),
    ("solver/smoothers.F90", "smoothers_", ["executerkstage", "executedadistep"], ()),
]
RENAME_MODULES = {"blockpointers": "bp_", "flowutils": "flowutils_", "turbutils": "turbutils_",
                  "residuals": "residuals_", "smoothers": "smoothers_", "sa": "sa_"}


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    if not os.path.exists(os.path.join(ref, "src", "NKSolver", "blockette.F90")):
        print("make_ref: %s/src not found -- reference not present, nothing generated" % ref)
        return 0
    env = f90toc.Env(ENV_INTS, env_arrays(), ["getcorrectfork"], dict(ENV_SUBS))
    outdir = os.path.join(HERE, "_ref")
    os.makedirs(outdir, exist_ok=True)
    tr = None
    for rel, prefix, routines, defined in UNITS:
        src = os.path.join(ref, "src", rel)
        code, tr = f90toc.translate_module(src, only=set(routines), env=env, rename_modules=RENAME_MODULES,
                                           patches=PATCHES, defined=defined, tr=tr, prefix=prefix)
        out = os.path.join(outdir, os.path.basename(rel).replace(".F90", "").lower() + "_ref.c")
        with open(out, "w") as f:
            f.write(code)
        print("make_ref: wrote %s (%d lines)" % (out, code.count("\n")))
    with open(os.path.join(outdir, "ref_protos.h"), "w") as f:
        f.write("/* GENERATED by oracle/make_ref.py -- prototypes of the translated reference procedures */\n")
        f.write("\n".join(tr.all_protos) + "\n")
    consts = f90toc.translate_parameters(os.path.join(ref, "src", "modules", "constants.F90"))
    with open(os.path.join(outdir, "ref_constants.h"), "w") as f:
        f.write(consts)
    return 0


if __name__ == "__main__":
    sys.exit(main())
