/* ref_env.c -- TEST INFRASTRUCTURE ONLY.  Definitions for ref_env.h: the module variables the
 * translated reference routines (oracle/_ref/blockette_ref.c) read, plus the three tiny
 * utility procedures they call.  The test harness sets the variables by symbol name through
 * ctypes (tests/refblockette.py). */
#include <stdio.h>
#include <stdlib.h>
#include "ref_env.h"

int nw = 6, nwf = 5, nt1 = 6, nt2 = 6, viscous = 1, kpresent = 0, eddymodel = 1;
double pinfcorr, rhoinf, gammainf, timeref = 1.0, rgas, tref;
int equations, equationmode, turbmodel, turbprod, useqcr, useft2sa, userotationsa;
double prandtl, prandtlturb;
int spacediscr, orderturb, limiter, precond, riemann, riemanncoarse, approxsa;
double vis2coarse = 0.5;
double vis2, vis4, sigma, adis, acousticscalefactor, kappacoef;
int usedisscontinuation = 0;
double disscontmagnitude, disscontmidpoint, disscontsharpness;
double turbresscale[4];
int currentlevel = 1, groundlevel = 1;
double rfil = 1.0, totalr0 = 1.0, totalr = 1.0;
int ntimeintervalsspectral = 1, oversetpresent = 0, secondord = 0;
double rsak, rsacb1, rsacb2, rsacb3, rsacv1, rsacw1, rsacw2, rsacw3, rsact1, rsact2, rsact3, rsact4, rsacrot;
int bp_nx, bp_ny, bp_nz, bp_il, bp_jl, bp_kl, bp_ie, bp_je, bp_ke, bp_ib, bp_jb, bp_kb;
int bp_addgridvelocities = 0, bp_righthanded = 1, bp_sectionid = 1, bp_blockismoving = 0, bp_nbkglobal = 1;
double *bp_w, *bp_p, *bp_gamma, *bp_rlv, *bp_rev, *bp_vol, *bp_volref, *bp_d2wall, *bp_shocksensor;
double *bp_x, *bp_si, *bp_sj, *bp_sk, *bp_sfacei, *bp_sfacej, *bp_sfacek;
double *bp_dw, *bp_fw, *bp_dtl, *bp_aa, *bp_radi, *bp_radj, *bp_radk;
double *bp_ux, *bp_uy, *bp_uz, *bp_vx, *bp_vy, *bp_vz, *bp_wx, *bp_wy, *bp_wz, *bp_qx, *bp_qy, *bp_qz;
int *bp_iblank, *bp_pori, *bp_porj, *bp_pork;
int bp_norphans, *bp_orphans;
double muinf, eddyvisinfratio;
double *bp_rotmatrixi = NULL, *bp_rotmatrixj = NULL, *bp_rotmatrixk = NULL;

int cpmodel = 1 /* cpConstant */, rkstage = 1, resaveraging = 0, bp_ndom = 1, exchangepressureearly = 0;
int lowspeedpreconditioner = 0;
double gammaconstant, musuthdim, tsuthdim, ssuthdim, muref = 1.0, pinf;
double cfl, cflcoarse, cfllimit, smoop, deltat = 1.0;
double etark[6], cdisrk[6], coeftime[8];
double *bp_wn, *bp_pn, *bp_scratch;
double *bp_bmti1, *bp_bmti2, *bp_bmtj1, *bp_bmtj2, *bp_bmtk1, *bp_bmtk2;
int turbrelax = 2 /* turbRelaxImplicit */;
double alfaturb;

int bp_nbocos = 0, bp_nviscbocos = 0, bp_bctype[64], bp_bcfaceid[64];
int viscwallbctreatment = 1, eulerwallbctreatment = 1, outflowtreatment = 1, wallfunctions = 0, hscalinginlet = 0;
double winf[10];
int lumpeddiss = 0, viscpc = 0, spacediscrcoarse = 1, smoother = 1 /* RungeKutta */, nrkstages = 5, nsubiterations = 1, subit = 0, radiineededfine = 1, radiineededcoarse = 1, dirscaling = 1;
double* bp_wr;
double monloc[16];
double *bp_s;
int *bp_globalcell;
double *bp_bvti1, *bp_bvti2, *bp_bvtj1, *bp_bvtj2, *bp_bvtk1, *bp_bvtk2;
RefSubface bcd[64];

int spectralsol = 1, computesepsensorks = 0, computecavitation = 0, cavexponent = 0, rvfn = 1;
double pref = 1.0, lref = 1.0, machcoef = 1.0, cpmin_rho = 1.0, cavitationnumber = 1.0, cavsensorsharpness = 10.0, cavsensoroffset = 0.0;
double sepsensorsharpness = 10.0, sepsensoroffset = 0.0, sepsensorkssharpness = 10.0, sepsensorksphi = 90.0, sepsensorksoffset = 0.0,
       sepsenmaxrho = 1.0;
double veldirfreestream[3] = {1.0, 0.0, 0.0}, pointref[3], momentaxis[6] = {0, 0, 0, 1, 0, 0}, cpmin_family[4], sepsenmaxfamily[4];

/* Driver-level procedures that are NOT part of the translated set.  One block, pointers bound by
   the harness; halo exchange (no neighbours) is a no-op. */
void (*setpointers_hook)(int level) = NULL;
void setpointers(int* nn, int* level, int* sps) { (void)nn; (void)sps; if (setpointers_hook) setpointers_hook(*level); }
/* timeStep (src/solver/solverUtils.F90:4-41): loop over the blocks of currentLevel around timeStep_block */
void solverutils_timestep_block(int* onlyradii);
void solverutils_timestep(int* onlyradii) {
    int one = 1;
    setpointers(&one, &currentlevel, &one);
    solverutils_timestep_block(onlyradii);
}

double *bp_w1, *bp_p1, *bp_mgiweight, *bp_mgjweight, *bp_mgkweight;
int *bp_mgifine, *bp_mgjfine, *bp_mgkfine, *bp_mgicoarse, *bp_mgjcoarse, *bp_mgkcoarse;
int sh_ib, sh_jb, sh_kb;
int fl_ib, fl_jb, fl_kb, cl_il, cl_jl, cl_kl, cl_ie, cl_je, cl_ke, cl_ib, cl_jb, cl_kb, cl_nbocos, mgboundcorr;
double *fl_w, *fl_p, *fl_vol, *fl_rev, *fl_w1, *fl_p1, *cl_w, *cl_p, *cl_vol, *cl_rev, *cl_w1, *cl_p1;
int *fl_iblank, *cl_iblank, cl_bctype[64], cl_bcfaceid[64];
double fcoll = 1.0;
RefSubface cbcd[64];
void whalo1(int* a, int* b, int* c, int* d, int* e, int* f) { (void)a; (void)b; (void)c; (void)d; (void)e; (void)f; }
void whalo2(int* a, int* b, int* c, int* d, int* e, int* f) { (void)a; (void)b; (void)c; (void)d; (void)e; (void)f; }

/* branches of other physical models (cp curve fits, k-omega, SST, k-tau, v2-f): never taken on the
   path (cpConstant, Spalart-Allmaras); reaching one is a harness error */
#define UNREACHABLE(name) void name() { terminate(#name, "model outside the translated hot path"); }
UNREACHABLE(flowutils_computeetotcellcpfit)
UNREACHABLE(turbutils_kweddyviscosity)
UNREACHABLE(turbutils_ssteddyviscosity)
UNREACHABLE(turbutils_kteddyviscosity)
UNREACHABLE(turbutils_vfeddyviscosity)


/* src/utils/utils.F90:486-500 */
int getcorrectfork(void) { return kpresent && currentlevel <= groundlevel; }

/* src/utils/utils.F90:501 -- fatal error */
void terminate(const char* routine, const char* msg) {
    fprintf(stderr, "reference terminate() in %s: %s\n", routine, msg);
    abort();
}

/* ANK helpers: MATMUL / TRANSPOSE of the nState x nState blocks (column-major), terms summed in index order */
int ank_chartimestepcode = 0, ank_nvec = 0;
double ank_machinf = 0.0, *ank_wvec, *ank_dvec;
extern int anksolver_nstate;
void ank_matmul(double* a, double* b, double* c) {
    const int n = anksolver_nstate;
    double t[64];
    for (int j = 0; j < n; j++) for (int i = 0; i < n; i++) {
        double s = 0.0;
        for (int k = 0; k < n; k++) s += a[i + n * k] * b[k + n * j];
        t[i + n * j] = s;
    }
    for (int q = 0; q < n * n; q++) c[q] = t[q];
}
void ank_matmul_nt(double* a, double* b, double* c) {
    const int n = anksolver_nstate;
    double t[64];
    for (int j = 0; j < n; j++) for (int i = 0; i < n; i++) {
        double s = 0.0;
        for (int k = 0; k < n; k++) s += a[i + n * k] * b[j + n * k];
        t[i + n * j] = s;
    }
    for (int q = 0; q < n * n; q++) c[q] = t[q];
}

/* executeMGCycle environment */
int nstepscycling = 0, cycling[256], approxtotalits = 0, nsubiterturb = 1;
void solverutils_computeutau(void) {}
void sa_sa_block(int* resonly);
/* turbSolveDDADI, src/turbulence/turbAPI.F90:4-95 (Spalart-Allmaras, steady, one block, no neighbours) */
void turbsolveddadi(void) {
    int one = 1, resonly = 0;
    for (int it = 1; it <= nsubiterturb; it++) {
        setpointers(&one, &currentlevel, &one);
        sa_sa_block(&resonly);
    }
}
