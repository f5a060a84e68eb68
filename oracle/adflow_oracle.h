/*
 * adflow_oracle.h -- CPU restatement of the ADflow hot path (TEST INFRASTRUCTURE ONLY).
 *
 * This is the parity oracle of SURVEY.md section 8c.  It is never linked into
 * or called from the product library (adflow_b200/csrc); only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load it.
 *
 * PARITY PINNED against the reference's own routines, bit for bit.  The reference
 * as a whole (Fortran + MPI + PETSc + CGNS) cannot be built here (no Fortran
 * compiler), but its hot-path routines are plain Fortran 90 loops: oracle/f90toc.py
 * translates them statement by statement to C FROM THE SOURCE WHERE IT LIES
 * (/root/reference, never copied), oracle/Makefile target `ref` compiles the result
 * with gcc -O2 -ffp-contract=off into oracle/_ref/libblockette_ref.so, and
 * tests/test_oracle_vs_reference*.py assert np.array_equal between that library and
 * this restatement on seeded blocks for: blocketteResCore and all its flux / SA /
 * time-step routines (exact and approximate variants, scalar/matrix/upwind), state
 * preparation, metrics and volumes, flow and turbulence BCs, residual averaging,
 * the RK stage, computeDwDADI + the DADI step, the SA DD-ADI block solve, the multigrid transfer operators with
 * the coarse-level branches of residual / time step / smoothers / BCs (multiGrid.F90), and the ANK time-step block
 * and physicality check (NKSolvers.F90, module ANKSolver).  tests/golden/reference_golden.json keeps outputs of the
 * translated reference for use where oracle/_ref is absent.
 * STILL UNPINNED (no reference arithmetic to run): PETSc's matrix-free differencing
 * parameter h (adfb_mffd_*), and the halo-exchange index lists, which are built by
 * the reference's preprocessing from CGNS connectivity (checked instead against a
 * numpy model and by partition independence).  The golden JSONs of the reference
 * (tests/reg_tests/refs/) need its CGNS meshes, which are not in the tree.
 *
 * Layout: every array lives in one uniform box (0:ib,0:jb,0:kb) (column-major,
 * i fastest) so that the Fortran index (i,j,k) IS the C offset
 * i + NI*(j + NJ*k); arrays that the reference allocates with tighter bounds
 * (1:ie, 2:il, node and face arrays) simply leave the unused entries untouched.
 * Multi-component arrays put the component slowest: w(i,j,k,l) -> (l-1)*N + idx.
 */
#ifndef ADFLOW_ORACLE_H
#define ADFLOW_ORACLE_H
#include <stdint.h>
#include "../include/adflow_b200.h"

typedef struct OrcBlock {
    int32_t nx, ny, nz, nw, rightHanded;
    int32_t level;          /* multigrid level, 0/1 = finest (ground level); > 1: coarse-level branches of the smoother path */
    /* state (cell, 2 halos) */
    double *w, *p, *rlv, *rev;
    /* geometry */
    double *x;              /* nodes (0:ie,0:je,0:ke,3) */
    double *si, *sj, *sk;   /* face normals, 3 comps each */
    double *vol, *volRef, *d2Wall;
    int8_t *porI, *porJ, *porK;
    int32_t *iblank;
    /* residual + work */
    double *dw, *fw;        /* nw and 5 comps */
    double *ss, *dss;       /* entropy, sensor (3 comps) */
    double *aa, *radI, *radJ, *radK, *dtl;
    double *grad;           /* 12 nodal gradient arrays ux,uy,uz,vx,..,qz */
    /* smoother storage */
    double *wn, *pn;        /* 5 comps / 1 */
    double *scratch;        /* 10 comps: DADI work, SA qq etc. */
    double *shock;          /* frozen shock sensor (referenceShockSensor) */
    double *wallTau;        /* [dir 0..2][tauxx,yy,zz,xy,xz,yz,qx,qy,qz][box]: viscous stress / heat flux of
                               every face (viscSubface%tau, %q are its boundary planes); NULL = not stored */
    double *wr, *w1, *p1;   /* multigrid: residual forcing term (5), solution at the start of the coarse visit (5 / 1) */
} OrcBlock;

#ifdef __cplusplus
extern "C" {
#endif
void orc_metrics(const OrcBlock* b);
void orc_volume(const OrcBlock* b);
void orc_pressure(const OrcBlock* b, const AdfbParams* prm, int includeHalos);
void orc_lam_viscosity(const OrcBlock* b, const AdfbParams* prm, int includeHalos);
void orc_eddy_viscosity(const OrcBlock* b, const AdfbParams* prm, int includeHalos);
void orc_etot(const OrcBlock* b, const AdfbParams* prm, int i0, int i1, int j0, int j1, int k0, int k1);
void orc_residual_core(const OrcBlock* b, const AdfbParams* prm, unsigned flags, double rFil);
/* individual operators, exposed for per-term tests */
void orc_time_step(const OrcBlock* b, const AdfbParams* prm, int updateDtl);
void orc_central_flux(const OrcBlock* b, const AdfbParams* prm);
void orc_diss_scalar(const OrcBlock* b, const AdfbParams* prm, double rFil);
void orc_speed_of_sound(const OrcBlock* b, const AdfbParams* prm);
void orc_nodal_gradients(const OrcBlock* b);
void orc_viscous_flux(const OrcBlock* b, const AdfbParams* prm, double rFil);
void orc_sa_source(const OrcBlock* b, const AdfbParams* prm);
void orc_sa_advection(const OrcBlock* b, const AdfbParams* prm);
void orc_sa_viscous(const OrcBlock* b, const AdfbParams* prm);
void orc_sa_res_scale(const OrcBlock* b);
void orc_sum_dw_fw(const OrcBlock* b);
void orc_norms(const OrcBlock* b, const AdfbParams* prm, double out[2]);
/* adflow_oracle_smooth.c : BCs + Runge-Kutta smoother (subface arrays are HOST pointers) */
void orc_apply_turb_bc(const OrcBlock* b, const AdfbParams* prm, int nSub, const AdfbSubface* sf, int secondHalo);
void orc_apply_flow_bc(const OrcBlock* b, const AdfbParams* prm, int nSub, const AdfbSubface* sf, int secondHalo);
void orc_residual_averaging(const OrcBlock* b, const AdfbParams* prm);
void orc_residual_block(const OrcBlock* b, const AdfbParams* prm, double rFil);
void orc_rk_stage(const OrcBlock* b, const AdfbParams* prm, int rkStage, int nSub, const AdfbSubface* sf);
void orc_compute_dw_dadi(const OrcBlock* b, const AdfbParams* prm);
void orc_dadi_step(const OrcBlock* b, const AdfbParams* prm, int nSub, const AdfbSubface* sf);
/* adflow_oracle_fluxes.c: alternative dissipation schemes (same calling convention as orc_diss_scalar) */
void orc_diss_matrix(const OrcBlock* b, const AdfbParams* prm, double rFil);
void orc_upwind_flux(const OrcBlock* b, const AdfbParams* prm, double rFil);
void orc_reference_shock_sensor(const OrcBlock* b, const AdfbParams* prm);
/* wallIntegrationFace (src/solver/surfaceIntegrations.F90:406-881), force and moment part: out = Fp(3), Fv(3),
   Mp(3), Mv(3) summed over the wall subfaces (viscous walls: pressure + viscous, Euler walls: pressure) */
void orc_wall_forces(const OrcBlock* b, const AdfbParams* prm, int nSub, const AdfbSubface* sf, const double refPoint[3],
                     double pRef, double out[12]);
void orc_diss_scalar_approx(const OrcBlock* b, const AdfbParams* prm);
void orc_diss_matrix_approx(const OrcBlock* b, const AdfbParams* prm, double rFil);
void orc_viscous_flux_approx(const OrcBlock* b, const AdfbParams* prm, double rFil);
/* adflow_oracle_sa.c: one sa_block(resOnly=.false.) = residual + DD-ADI solve + rev + turbulence BCs */
void orc_sa_block(const OrcBlock* b, const AdfbParams* prm, int nSub, const AdfbSubface* sf);
void orc_rk_smoother(const OrcBlock* b, const AdfbParams* prm, int nSub, const AdfbSubface* sf);
/* adflow_oracle_mg.c: multigrid transfer operators and the coarse-level residual (src/solver/multiGrid.F90) */
void orc_diss_scalar_coarse(const OrcBlock* b, const AdfbParams* prm, double rFil);
void orc_diss_matrix_coarse(const OrcBlock* b, const AdfbParams* prm, double rFil);
void orc_residual_block_coarse(const OrcBlock* b, const AdfbParams* prm, double rFil, int init);
void orc_mg_corner_row_halos(const OrcBlock* b, const AdfbParams* prm);
void orc_mg_restrict(const OrcBlock* coarse, const OrcBlock* fine, const AdfbParams* prm, const int32_t* mgIFine,
                     const int32_t* mgJFine, const int32_t* mgKFine, const double* mgIWeight, const double* mgJWeight,
                     const double* mgKWeight);
void orc_mg_store_w1(const OrcBlock* coarse);
void orc_mg_forcing(const OrcBlock* coarse, const AdfbParams* prm);
void orc_mg_prolong(const OrcBlock* fine, const OrcBlock* coarse, const AdfbParams* prm, int nSubCoarse, const AdfbSubface* sfCoarse,
                    const int32_t* mgICoarse, const int32_t* mgJCoarse, const int32_t* mgKCoarse);
/* full-multigrid start-up: transferToFineGrid(corrections = .false.) with extrapolateSolution / extrapolateViscosities */
void orc_mg_prolong_solution(const OrcBlock* fine, const OrcBlock* coarse, const AdfbParams* prm, int nSubCoarse,
                             const AdfbSubface* sfCoarse, const int32_t* mgICoarse, const int32_t* mgJCoarse, const int32_t* mgKCoarse);
/* adflow_oracle_ank.c: ANK pieces (module ANKSolver of src/NKSolver/NKSolvers.F90) */
/* turbulence KSP of the decoupled ANK: physicalityCheckANKTurb (NKSolvers.F90:3212-3335, pinned bit-exact against the translated
   routine) and the vector part of FormFunction_mf_turb (:2540-2612) */
double orc_ank_physicality_check_turb(const AdfbAnkParams* ank, long nCells, const double* wVec, double* dVec, double lambdaP);
void orc_ank_turb_rvec(const OrcBlock* b, const AdfbParams* prm, const AdfbAnkParams* ank, const double* inVec, double* rVec);
/* orphanAverage, src/utils/haloExchange.F90:201-354 (pinned bit-exact against the translated routine, tests/test_oracle_vs_reference_orphans.py) */
void orc_orphan_average(const OrcBlock* b, const AdfbParams* prm, int nOrphans, const int32_t* orphans, int wStart, int wEnd,
                        int calcPressure, int calcLamVis, int calcEddyVis, double muInf, double eddyVisInfRatio);
void orc_ank_time_step_block(const OrcBlock* b, const AdfbParams* prm, const AdfbAnkParams* ank, int i, int j, int k, double* blk);
double orc_ank_physicality_check(const AdfbAnkParams* ank, int nState, long nCells, const double* wVec, double* dVec, double lambdaP);
#ifdef __cplusplus
}
#endif
#endif
