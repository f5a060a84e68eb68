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
    # sa_block: steady flow, so the unsteady term is identically absent (turbUtils.F90:456-460 returns at once)
    (r"^call unsteadyturbterm\(.*$", "continue"),
    # saSolve: wall-function branch (wallFunctions = .false. on the path) uses BCData/viscSubface
    ("block", r"^testwallfunctions: if", r"^end if testwallfunctions"),
    # ALE (deforming-mesh unsteady) hooks of applyAllBC: steady path, both return at once in the reference
    (r"^call interplevelalebc_block$", "continue"),
    # actuator-zone source terms (sourceTerms): no actuator regions on the path
    (r"^call sourceterms\(\)$", "continue"),
    (r"^call recoverlevelalebc_block$", "continue"),
    (r"^call interplevelale_block$", "continue"),
    (r"^call recoverlevelale_block$", "continue"),
    # wallIntegrationFace: per-face output arrays of BCData (Fp, Fv, area: post-processing only) are not kept
    (r"^bcdata\(mm\)%(fp|fv|area)\b.*=.*$", "continue"),
    (r"^bcdata\(mm\)%fv = zero$", "continue"),
    # BCData(nn)%comp(...) -> accessor functions over the harness' subface table (oracle/ref_env.h)
    (r"bcdata\((\w+)\)%(\w+)\(", r"bcd_\2(\1, "),
    (r"bcdata\((\w+)\)%(\w+)", r"bcd_\2(\1)"),
    # viscSubface(mm)%tau(i,j,l) (read side, surface integration) -> accessor over the harness' wall-stress table
    (r"viscsubface\((\w+)\)%(\w+)\(", r"vsf_\2(\1, "),
    # module-wide `use X` without only-list inside routines: names resolve through ref_env.h
]

MG_PATCHES = [
    # executeMGCycle: the iteration-type label is output only
    (r"^itertype = .*$", "continue"),
    (r"^call terminate\(\"executemgcycle\".*$", "continue"),
    (r"flowdoms\(nn,\s*finelevel,\s*sps\)%(\w+)", r"fl_\1"),
    (r"flowdoms\(nn,\s*coarselevel,\s*(?:sps|1)\)%bcdata", r"cl_bcdata_unused"),
    (r"flowdoms\(nn,\s*coarselevel,\s*(?:sps|1)\)%(\w+)", r"cl_\1"),
    (r"^type\(bcdatatype\).*$", "integer(kind=inttype) :: bcdata_unused"),
    (r"^bcdata => cl_bcdata_unused$", "continue"),
    (r"bcdata\((\w+)\)%(\w+)\(", r"cbcd_\2(\1, "),
    (r"bcdata\((\w+)\)%(\w+)", r"cbcd_\2(\1)"),
]

# ANK pieces of NKSolvers.F90 (module ANKSolver): PETSc vectors -> harness arrays, MPI reduction -> copy (one rank),
# character option -> integer code, MATMUL/TRANSPOSE of the nState x nState blocks -> helpers in ref_env.c
ANK_PATCHES = [
    (r"^call vecgetarrayf90\(wvec, wvec_pointer, ierr\)$", "wvec_pointer => ank_wvec"),
    (r"^call vecgetarrayf90\(deltaw, dvec_pointer, ierr\)$", "dvec_pointer => ank_dvec"),
    (r"^call vecgetarrayf90\(wvecturb, wvec_pointer, ierr\)$", "wvec_pointer => ank_wvec"),
    (r"^call vecgetarrayf90\(deltawturb, dvec_pointer, ierr\)$", "dvec_pointer => ank_dvec"),
    (r"^call vecrestorearrayf90\(.*$", "continue"),
    (r"^call echk\(.*$", "continue"),
    (r"^call mpi_allreduce\(lambdal, lambdap_recv,.*$", "lambdap_recv = lambdal"),
    (r"myisnan\(lambdal\)", "(lambdal /= lambdal)"),
    (r"ank_chartimesteptype == '(?i:none)'", "ank_chartimestepcode == 0"),
    (r"ank_chartimesteptype == '(?i:vlr)'", "ank_chartimestepcode == 1"),
    (r"ank_chartimesteptype == '(?i:turkel)'", "ank_chartimestepcode == 2"),
    (r"^timestepblock = matmul\((\w+), transpose\((\w+)\)\)$", r"call ank_matmul_nt(\1, \2, timestepblock)"),
    (r"^timestepblock = matmul\((\w+), (\w+)\)$", r"call ank_matmul(\1, \2, timestepblock)"),
    # `use inputPhysics, only: machInf => mach`: the module's free-stream Mach number, NOT the local variable mach
    (r"machinf", "ank_machinf"),
    (r"ank_machinf => mach", "ank_machinf"),
]

# external (other-module) data the translated routines see; declared in oracle/ref_env.h
ENV_INTS = """nw nwf nt1 nt2 equations equationmode turbmodel spacediscr ransequations nsequations eulerequations
 steady unsteady timespectral spalartallmaras dissscalar dissmatrix upwind currentlevel groundlevel
 irho ivx ivy ivz irhoe itu1 itu2 imx imy imz viscous addgridvelocities oversetpresent blockismoving
 useft2sa userotationsa turbprod useqcr approxsa secondord orderturb limiter precond riemann
 firstorder secondorder nolimiter vanalbeda minmod noprecond turkel choimerkle roe vanleer ausmdv
 strain vorticity katolaunder kpresent eddymodel rotationalperiodic correctfork righthanded
 usedisscontinuation nbkglobal sectionid ntimeintervalsspectral normalflux boundflux internalflux
 lumpeddiss fullturb cpmodel rkstage resaveraging bp_ndom exchangepressureearly lowspeedpreconditioner
 noresaveraging alwaysresaveraging alternateresaveraging turbrelax turbrelaximplicit turbrelaxexplicit
 bp_nx bp_ny bp_nz bp_il bp_jl bp_kl bp_ie bp_je bp_ke bp_ib bp_jb bp_kb bp_addgridvelocities
 bp_righthanded bp_sectionid bp_blockismoving bp_nbkglobal bp_nbocos bp_nviscbocos
 viscwallbctreatment eulerwallbctreatment outflowtreatment wallfunctions
 spectralsol computesepsensorks computecavitation cavexponent rvfn hscalinginlet totalconditions massflow
 lumpeddiss viscpc spacediscrcoarse smoother rungekutta dadi nrkstages nsubiterations subit radiineededfine radiineededcoarse dirscaling
 symm symmpolar nswalladiabatic nswallisothermal farfield eulerwall extrap supersonicinflow supersonicoutflow
 subsonicinflow subsonicoutflow massbleedoutflow imin imax jmin jmax kmin kmax
 constantpressure linextrapolpressure quadextrapolpressure normalmomentum
 nstepscycling nlluggs nllusgsline approxtotalits ank_chartimestepcode ank_nvec sh_ib sh_jb sh_kb fl_ib fl_jb fl_kb cl_il cl_jl cl_kl cl_ie cl_je cl_ke cl_ib cl_jb cl_kb cl_nbocos mgboundcorr bcdirichlet0 bcneumann
 slidinginterface oversetouterbound domaininterfaceall domaininterfacerhouvw domaininterfacep domaininterfacerho
 domaininterfacetotal bp_norphans""".split()

BOX_STRIDES = ["1", "(bp_ib + 1)", "(bp_ib + 1) * (bp_jb + 1)", "(bp_ib + 1) * (bp_jb + 1) * (bp_kb + 1)"]


# setPointers (src/utils/utils.F90:3419-3477) points dw, fw, scratch, wn, pn, dtl, radI/J/K, gamma and rlv of EVERY grid
# level at the finest level's arrays: on a coarse level they are the fine arrays indexed with coarse indices.  Their
# strides therefore come from the dimensions of the array that is actually bound (sh_ib/jb/kb, set by the harness:
# the current block's for single-level use, the finest block's inside a multigrid cycle)
SH_STRIDES = ["1", "(sh_ib + 1)", "(sh_ib + 1) * (sh_jb + 1)", "(sh_ib + 1) * (sh_jb + 1) * (sh_kb + 1)"]
SHARED = {"dw", "fw", "scratch", "wn", "pn", "dtl", "radi", "radj", "radk", "gamma", "rlv"}


def refarr(name, ctype, lo, hi, ncomp=None, comp_lo="1"):
    """blockPointers array: DECLARED bounds are the reference's (src/modules/block.F90, allocation in
    src/initFlow/initializeFlow.F90:457-530,686-722 -- they decide what `a(:, j, k)` means), STORAGE is
    the uniform box (0:ib,0:jb,0:kb[,ncomp]) that ref_env.h exposes, hence explicit strides and origin 0."""
    bounds = [(l, "(%s) - (%s) + 1" % (h, l)) for l, h in zip(lo, hi)]
    base = ["0", "0", "0"]
    st = SH_STRIDES if name in SHARED else BOX_STRIDES
    strides = st[:3]
    if ncomp is not None:
        bounds.append((comp_lo, str(ncomp)))
        base.append(comp_lo)
        strides = st[:4]
    return f90toc.Array("bp_" + name, ctype, bounds, strides=strides, base=base)


def env_arrays():
    A = f90toc.Array
    c2 = (("0", "0", "0"), ("bp_ib", "bp_jb", "bp_kb"))      # double halo
    c1 = (("1", "1", "1"), ("bp_ie", "bp_je", "bp_ke"))      # single halo
    c0 = (("2", "2", "2"), ("bp_il", "bp_jl", "bp_kl"))      # owned cells
    nd = (("1", "1", "1"), ("bp_il", "bp_jl", "bp_kl"))      # nodes 1:il
    arrs = {}
    for n in ["p", "gamma", "rlv", "rev", "vol", "volref", "aa", "shocksensor"]:
        arrs["bp_" + n] = refarr(n, "double", *c2)
    for n in ["radi", "radj", "radk", "dtl"]:
        arrs["bp_" + n] = refarr(n, "double", *c1)
    for n in ["ux", "uy", "uz", "vx", "vy", "vz", "wx", "wy", "wz", "qx", "qy", "qz"]:
        arrs["bp_" + n] = refarr(n, "double", *nd)
    arrs["bp_d2wall"] = refarr("d2wall", "double", *c0)
    arrs["bp_pn"] = refarr("pn", "double", *c0)
    arrs["bp_wn"] = refarr("wn", "double", *c0, ncomp="nwf")
    arrs["bp_w"] = refarr("w", "double", *c2, ncomp="nw")
    arrs["bp_dw"] = refarr("dw", "double", *c2, ncomp="nw")
    arrs["bp_fw"] = refarr("fw", "double", *c2, ncomp="nwf")
    arrs["bp_scratch"] = refarr("scratch", "double", *c2, ncomp=10)
    arrs["bp_wr"] = refarr("wr", "double", *c0, ncomp="nwf")
    arrs["bp_x"] = refarr("x", "double", ("0", "0", "0"), ("bp_ie", "bp_je", "bp_ke"), ncomp=3)
    arrs["bp_si"] = refarr("si", "double", ("0", "1", "1"), ("bp_ie", "bp_je", "bp_ke"), ncomp=3)
    arrs["bp_sj"] = refarr("sj", "double", ("1", "0", "1"), ("bp_ie", "bp_je", "bp_ke"), ncomp=3)
    arrs["bp_sk"] = refarr("sk", "double", ("1", "1", "0"), ("bp_ie", "bp_je", "bp_ke"), ncomp=3)
    arrs["bp_s"] = refarr("s", "double", *c1, ncomp=3)
    arrs["bp_sfacei"] = refarr("sfacei", "double", ("0", "1", "1"), ("bp_ie", "bp_je", "bp_ke"))
    arrs["bp_sfacej"] = refarr("sfacej", "double", ("1", "0", "1"), ("bp_ie", "bp_je", "bp_ke"))
    arrs["bp_sfacek"] = refarr("sfacek", "double", ("1", "1", "0"), ("bp_ie", "bp_je", "bp_ke"))
    arrs["bp_iblank"] = refarr("iblank", "int", *c2)
    arrs["bp_orphans"] = A("bp_orphans", "int", [("1", "3"), ("1", "bp_norphans")])   # blockPointers orphans(3, nOrphans)
    arrs["bp_globalcell"] = refarr("globalcell", "int", *c2)
    arrs["bp_pori"] = refarr("pori", "int", ("1", "2", "2"), ("bp_il", "bp_jl", "bp_kl"))
    arrs["bp_porj"] = refarr("porj", "int", ("2", "1", "2"), ("bp_il", "bp_jl", "bp_kl"))
    arrs["bp_pork"] = refarr("pork", "int", ("2", "2", "1"), ("bp_il", "bp_jl", "bp_kl"))
    # turbulence BC matrices of the six block faces (block.F90: bmti1(je,ke,nt1:nt2,nt1:nt2) ...)
    for n, (a, b) in {"bmti1": ("bp_je", "bp_ke"), "bmti2": ("bp_je", "bp_ke"), "bmtj1": ("bp_ie", "bp_ke"),
                      "bmtj2": ("bp_ie", "bp_ke"), "bmtk1": ("bp_ie", "bp_je"), "bmtk2": ("bp_ie", "bp_je")}.items():
        arrs["bp_" + n] = A("bp_" + n, "double", [("1", a), ("1", b), ("nt1", "1"), ("nt1", "1")])
        arrs["bp_" + n.replace("bmt", "bvt")] = A("bp_" + n.replace("bmt", "bvt"), "double", [("1", a), ("1", b), ("nt1", "1")])
    for n in ["rotmatrixi", "rotmatrixj", "rotmatrixk"]:
        arrs["bp_" + n] = A("bp_" + n, "double", [("1", "1")])
    # multigrid: w1/p1 (solution at the start of the coarse-level visit), restriction / interpolation tables
    arrs["bp_w1"] = refarr("w1", "double", *c1, ncomp="nwf")
    arrs["bp_p1"] = refarr("p1", "double", *c1)
    for dname, e, l in (("i", "bp_ie", "bp_il"), ("j", "bp_je", "bp_jl"), ("k", "bp_ke", "bp_kl")):
        arrs["bp_mg%sfine" % dname] = A("bp_mg%sfine" % dname, "int", [("1", e), ("1", "2")])
        arrs["bp_mg%sweight" % dname] = A("bp_mg%sweight" % dname, "double", [("2", "(%s) - 1" % l)])
        arrs["bp_mg%scoarse" % dname] = A("bp_mg%scoarse" % dname, "int", [("2", "(%s) - 1" % l), ("1", "2")])
    # the other level's block (fine level in transferToCoarseGrid, coarse level in transferToFineGrid)
    for pre in ("fl", "cl"):
        st = ["1", "(%s_ib + 1)" % pre, "(%s_ib + 1) * (%s_jb + 1)" % (pre, pre),
              "(%s_ib + 1) * (%s_jb + 1) * (%s_kb + 1)" % (pre, pre, pre)]
        bnd = [("0", "%s_ib + 1" % pre), ("0", "%s_jb + 1" % pre), ("0", "%s_kb + 1" % pre)]
        for n in ("p", "vol", "rev", "p1"):
            arrs["%s_%s" % (pre, n)] = A("%s_%s" % (pre, n), "double", list(bnd), strides=st[:3], base=["0", "0", "0"])
        arrs["%s_iblank" % pre] = A("%s_iblank" % pre, "int", list(bnd), strides=st[:3], base=["0", "0", "0"])
        for n, nc in (("w", "nw"), ("w1", "nwf")):
            arrs["%s_%s" % (pre, n)] = A("%s_%s" % (pre, n), "double", list(bnd) + [("1", nc)], strides=st, base=["0", "0", "0", "1"])
    # ANK: the PETSc vectors wVec / deltaW as plain arrays (bound by the harness)
    arrs["ank_wvec"] = A("ank_wvec", "double", [("1", "ank_nvec")], pointer=True) if False else A("ank_wvec", "double", [("1", "ank_nvec")])
    arrs["ank_dvec"] = A("ank_dvec", "double", [("1", "ank_nvec")])
    arrs["cycling"] = A("cycling", "int", [("1", "256")])
    arrs["cl_bctype"] = A("cl_bctype", "int", [("1", "64")])
    arrs["cl_bcfaceid"] = A("cl_bcfaceid", "int", [("1", "64")])
    arrs["bp_bctype"] = A("bp_bctype", "int", [("1", "64")])
    arrs["bp_bcfaceid"] = A("bp_bcfaceid", "int", [("1", "64")])
    arrs["winf"] = A("winf", "double", [("1", "10")])
    arrs["monloc"] = A("monloc", "double", [("1", "16")])
    # inputPhysics / inputCostFunctions data of the surface integration
    arrs["veldirfreestream"] = A("veldirfreestream", "double", [("1", "3")])
    arrs["pointref"] = A("pointref", "double", [("1", "3")])
    arrs["momentaxis"] = A("momentaxis", "double", [("1", "3"), ("1", "2")])
    arrs["cpmin_family"] = A("cpmin_family", "double", [("1", "4")])
    arrs["sepsenmaxfamily"] = A("sepsenmaxfamily", "double", [("1", "4")])
    arrs["turbresscale"] = A("turbresscale", "double", [("1", "4")])
    arrs["etark"] = A("etark", "double", [("1", "6")])
    arrs["cdisrk"] = A("cdisrk", "double", [("1", "6")])
    arrs["coeftime"] = A("coeftime", "double", [("0", "8")])
    return arrs


ENV_SUBS = {
    # name -> [(argname, ctype, isarray)] for external subroutines called by reference
    "terminate": [("routine", "str", False), ("msg", "str", False)],
    # driver-level procedures outside the translated set: no-op stubs in ref_env.c (single block, BCs and
    # halo exchange are applied by the test harness around the translated routine)
    "setpointers": [("nn", "int", False), ("level", "int", False), ("sps", "int", False)],
    "whalo1": [(n, "int", False) for n in ("level", "start", "end", "commpressure", "commgamma", "commviscous")],
    "whalo2": [(n, "int", False) for n in ("level", "start", "end", "commpressure", "commgamma", "commviscous")],
    "computeutau": [],
    "turbapi_turbsolveddadi": [],
    "ank_matmul": [("a", "double", True), ("b", "double", True), ("c", "double", True)],
    "ank_matmul_nt": [("a", "double", True), ("b", "double", True), ("c", "double", True)],
}


# (source file relative to <ref>/src, C prefix of its procedures, routines to translate)
UNITS = [
    # USE_TAPENADE (the reference's own switch for its AD-differentiable subset) drops the cp-curve-fit
    # branches of eint/computeEtotBlock, which need tables outside the path (cpModel == cpConstant here)
    ("utils/flowUtils.F90", "flowutils_", ["computeetotblock", "computelamviscosity", "computepressuresimple",
                                           "etot", "eint", "computespeedofsoundsquared", "allnodalgradients"], ("USE_TAPENADE",)),
    ("NKSolver/blockette.F90", "", ROUTINES, ()),
    ("modules/BCPointers.F90", "bcpointers_", [], ("USE_TAPENADE",)),
    ("utils/utils.F90", "", ["setbcpointers", "sumresiduals", "sumallresiduals"], ()),
    ("adjoint/adjointUtils.F90", "adjointutils_", ["referenceshocksensor"], ()),
    ("solver/BCRoutines.F90", "bcroutines_", ["applyallbc", "applyallbc_block", "bcsymm1sthalo", "bcsymm2ndhalo", "bcsymmpolar1sthalo",
                                              "bcsymmpolar2ndhalo", "bcnswalladiabatic",
                                              "bcnswallisothermal", "bcfarfield", "bceulerwall", "bcextrap", "bcsubsonicoutflow",
                                              "bcsubsonicinflow", "bcsupersonicinflow",
                                              "computeetot", "extrapolate2ndhalo"], ()),
    ("turbulence/turbUtils.F90", "turbutils_", ["computeeddyviscosity", "saeddyviscosity", "turbadvection"], ()),
    ("adjoint/adjointExtra.F90", "adjointextra_", ["volume_block", "metric_block"], ()),
    ("solver/surfaceIntegrations.F90", "surfaceintegrations_", ["wallintegrationface", "ksaggregationfunction"], ()),
    ("turbulence/turbBCRoutines.F90", "turbbcroutines_",
     ["applyallturbbc", "applyallturbbcthisblock", "bceddynowall", "bceddywall", "bcturbfarfield", "bcturbinflow", "bcturbinterface",
      "bcturboutflow", "bcturbsymm", "bcturbtreatment", "bcturbwall", "turb2ndhalo"], ()),
    ("turbulence/sa.F90", "sa_", ["sa_block", "sasource", "saviscous", "saresscale", "sasolve"], ()),
    ("solver/residuals.F90", "residuals_", ["residualaveraging", "computedwdadi", "tridiagsolve", "initres", "residual"], ()),
    ("solver/smoothers.F90", "smoothers_", ["executerkstage", "executedadistep", "rungekuttasmoother", "dadismoother"], ()),
    # the block-path residual of the smoother loops: fluxes.F90 (block twins of the blockette routines) and
    # residual_block / initres_block; USE_TAPENADE drops the ALE hooks and the coarse-grid dissipation calls
    ("solver/solverUtils.F90", "solverutils_", ["timestep_block"], ("USE_TAPENADE",)),
    ("solver/fluxes.F90", "fluxes_", ["inviscidcentralflux", "invisciddissfluxscalar", "invisciddissfluxmatrix", "inviscidupwindflux",
                                      "viscousflux", "invisciddissfluxscalarapprox", "invisciddissfluxmatrixapprox",
                                      "viscousfluxapprox"], ("USE_TAPENADE",)),
    # coarse-level (first order) dissipation of the multigrid cycle and residual_block WITH its coarse-level branch
    ("solver/fluxes.F90", "fluxes_", ["invisciddissfluxscalarcoarse", "invisciddissfluxmatrixcoarse"], (), "fluxes_coarse_ref.c"),
    ("solver/residuals.F90", "residuals_", ["initres_block"], ("USE_TAPENADE",), "residuals_initres_ref.c"),
    ("solver/residuals.F90", "residuals_", ["residual_block"], (), "residuals_block_ref.c"),
    # multigrid transfer operators.  flowDoms(nn, fineLevel / coarseLevel, sps)%x (the OTHER level's block, next
    # to the blockPointers of the current one) -> fl_x / cl_x env data bound by the harness; the coarse block's
    # BCData -> accessor functions over a second subface table (cbcd)
    ("solver/multiGrid.F90", "multigrid_", ["transfertocoarsegrid", "transfertofinegrid", "setcornerrowhalos",
                                            "setcorrectionscoarsehalos", "executemgcycle", "extrapolatesolution",
                                            "extrapolateviscosities"], (), None, MG_PATCHES),
    # ANK: time-step block of the matrix-free operator and the physicality check of the update
    ("NKSolver/NKSolvers.F90", "anksolver_", ["computetimestepblock", "physicalitycheckank", "physicalitycheckankturb"], (), "anksolver_ref.c", ANK_PATCHES,
     "anksolver"),
    # overset orphans: average of the valid neighbours (called by whalo1 / whalo2 after wOverset)
    ("utils/haloExchange.F90", "haloexchange_", ["orphanaverage"], (), "haloexchange_ref.c"),
]
RENAME_MODULES = {"blockpointers": "bp_", "flowutils": "flowutils_", "turbutils": "turbutils_",
                  "residuals": "residuals_", "smoothers": "smoothers_", "sa": "sa_",
                  "bcpointers": "bcpointers_", "bcroutines": "bcroutines_",
                  "turbbcroutines": "turbbcroutines_", "surfaceintegrations": "surfaceintegrations_",
                  "fluxes": "fluxes_", "solverutils": "solverutils_", "multigrid": "multigrid_"}


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    if not os.path.exists(os.path.join(ref, "src", "NKSolver", "blockette.F90")):
        print("make_ref: %s/src not found -- reference not present, nothing generated" % ref)
        return 0
    env = f90toc.Env(ENV_INTS, env_arrays(), ["getcorrectfork", "bcd_icbeg", "bcd_icend", "bcd_jcbeg", "bcd_jcend",
                                              "bcd_inbeg", "bcd_inend", "bcd_jnbeg", "bcd_jnend", "bcd_iblank",
                                              "bcd_subsonicinlettreatment", "cbcd_icbeg", "cbcd_icend", "cbcd_jcbeg", "cbcd_jcend"], dict(ENV_SUBS))
    outdir = os.path.join(HERE, "_ref")
    os.makedirs(outdir, exist_ok=True)
    tr = None
    for unit in UNITS:
        rel, prefix, routines, defined = unit[:4]
        patches = (list(unit[5]) if len(unit) > 5 else []) + PATCHES
        src = os.path.join(ref, "src", rel)
        code, tr = f90toc.translate_module(src, only=set(routines), env=env, rename_modules=RENAME_MODULES,
                                           patches=patches, defined=defined, tr=tr, prefix=prefix,
                                           module=unit[6] if len(unit) > 6 else None)
        out = os.path.join(outdir, unit[4] if len(unit) > 4 and unit[4] else os.path.basename(rel).replace(".F90", "").lower() + "_ref.c")
        with open(out, "w") as f:
            f.write(code)
        print("make_ref: wrote %s (%d lines)" % (out, code.count("\n")))
    with open(os.path.join(outdir, "ref_protos.h"), "w") as f:
        f.write("/* GENERATED by oracle/make_ref.py -- prototypes of the translated reference procedures */\n")
        f.write("\n".join(tr.all_protos) + "\n")
    consts = f90toc.translate_parameters(os.path.join(ref, "src", "modules", "constants.F90"))
    consts += f90toc.translate_parameters(os.path.join(ref, "src", "modules", "paramTurb.F90"))
    with open(os.path.join(outdir, "ref_constants.h"), "w") as f:
        f.write(consts)
    return 0


if __name__ == "__main__":
    sys.exit(main())
