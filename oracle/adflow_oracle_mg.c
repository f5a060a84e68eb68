/*
 * adflow_oracle_mg.c -- CPU restatement of the multigrid pieces of the smoother path (TEST INFRASTRUCTURE ONLY):
 * restriction / residual forcing term (transferToCoarseGrid, src/solver/multiGrid.F90:5-324), interpolation of
 * the corrections (transferToFineGrid(.true.), :326-654), setCornerRowHalos (:1032-1357),
 * setCorrectionsCoarseHalos (:1359-1503), the first-order coarse-level scalar dissipation
 * (inviscidDissFluxScalarCoarse, src/solver/fluxes.F90:4977-5203) and the coarse-level branch of
 * initRes_block / residual_block (src/solver/residuals.F90:40-135, 468-497).
 *
 * PARITY PINNED: tests/test_oracle_vs_reference_mg.py compares every function below bit for bit with the
 * translated reference routines (oracle/_ref, multigrid_ref.c / fluxes_coarse_ref.c / residuals_block_ref.c).
 *
 * The index tables are the reference's (src/preprocessing/coarseUtils.F90:254-420), passed with the Fortran
 * index as the C index: mgIFine[(i) + (ie+1)*(m-1)], i = 1..ie (coarse), m = 1,2; mgIWeight[i], i = 2..il (coarse);
 * mgICoarse[(i) + (ie+1)*(m-1)], i = 2..il (FINE block, ie = fine ie).
 */
#include "orc_internal.h"

#define WR(c, l) b->wr[(long)(l) * d.N + (c)]

/* inviscidDissFluxScalarCoarse, fluxes.F90:4977-5203.  The reference converts w to conservative variables in place
   over the first halos (w(ivx) = rho*u, w(irhoE) = rhoE + p) and converts back at the end (w(ivx) * (1/rho),
   (rhoE + p) - p): the round trip is NOT exact and is part of the reference's arithmetic, so it is kept. */
static void diss_coarse_dir(const OrcBlock* b, Dims d, const double* rad, const int8_t* por, long sd, int i0, int j0, int k0, double fis0) {
    for (int k = k0; k <= d.kl; k++) for (int j = j0; j <= d.jl; j++) for (int i = i0; i <= d.il; i++) {
        long c = IDX(i, j, k), cp = c + sd;
        double ppor = zero;
        if (por[c] == ADFB_NORMALFLUX) ppor = half;
        double dis0 = fis0 * ppor * (rad[c] + rad[cp]);
        for (int l = 0; l < 5; l++) {
            double fs = dis0 * (W(cp, l) - W(c, l));
            FW(cp, l) += fs; FW(c, l) -= fs;
        }
    }
}
void orc_diss_scalar_coarse(const OrcBlock* b, const AdfbParams* prm, double rFil) {
    Dims d = dims_of(b);
    if (fabs(rFil) < thresholdReal) return;
    double fis0 = rFil * prm->vis2Coarse, sfil = one - rFil;
    for (int k = 1; k <= d.ke; k++) for (int j = 1; j <= d.je; j++) for (int i = 1; i <= d.ie; i++) {
        long c = IDX(i, j, k);
        W(c, IVX) = W(c, IRHO) * W(c, IVX);
        W(c, IVY) = W(c, IRHO) * W(c, IVY);
        W(c, IVZ) = W(c, IRHO) * W(c, IVZ);
        W(c, IRHOE) = W(c, IRHOE) + b->p[c];
    }
    for (int k = 2; k <= d.kl; k++) for (int j = 2; j <= d.jl; j++) for (int i = 2; i <= d.il; i++) {
        long c = IDX(i, j, k);
        for (int l = 0; l < 5; l++) FW(c, l) = sfil * FW(c, l);
    }
    diss_coarse_dir(b, d, b->radI, b->porI, d.sI, 1, 2, 2, fis0);
    diss_coarse_dir(b, d, b->radJ, b->porJ, d.sJ, 2, 1, 2, fis0);
    diss_coarse_dir(b, d, b->radK, b->porK, d.sK, 2, 2, 1, fis0);
    for (int k = 1; k <= d.ke; k++) for (int j = 1; j <= d.je; j++) for (int i = 1; i <= d.ie; i++) {
        long c = IDX(i, j, k);
        double rhoi = one / W(c, IRHO);
        W(c, IVX) = W(c, IVX) * rhoi;
        W(c, IVY) = W(c, IVY) * rhoi;
        W(c, IVZ) = W(c, IVZ) * rhoi;
        W(c, IRHOE) = W(c, IRHOE) - b->p[c];
    }
}

/* initRes_block + residual_block on a coarse level (currentLevel > groundLevel): init = 1 starts from the residual
   forcing term wr (residuals.F90:485-497), init = 0 from zero (transferToCoarseGrid, multiGrid.F90:256-268) */
void orc_residual_block_coarse(const OrcBlock* b, const AdfbParams* prm, double rFil, int init) {
    Dims d = dims_of(b);
    int viscous = prm->equations != ADFB_EULER;
    if (init) {
        for (int l = 0; l < 5; l++)
            for (int k = 2; k <= d.kl; k++) for (int j = 2; j <= d.jl; j++) for (int i = 2; i <= d.il; i++) {
                long c = IDX(i, j, k);
                DW(c, l) = WR(c, l);
            }
    } else {
        memset(b->dw, 0, sizeof(double) * 5 * d.N);
    }
    orc_central_flux(b, prm);
    if (prm->spaceDiscrCoarse == ADFB_DISS_MATRIX) orc_diss_matrix_coarse(b, prm, rFil);
    else if (prm->spaceDiscrCoarse == ADFB_UPWIND) {   /* inviscidUpwindFlux(fineGrid = .false.): first order, fluxes.F90:1532 */
        AdfbParams p1 = *prm;
        p1.limiter = ADFB_LIM_FIRSTORDER;
        orc_upwind_flux(b, &p1, rFil);
    } else orc_diss_scalar_coarse(b, prm, rFil);
    if (viscous && fabs(rFil) > thresholdReal) {
        orc_speed_of_sound(b, prm);
        orc_nodal_gradients(b);
        orc_viscous_flux(b, prm, rFil);
    }
    orc_sum_dw_fw(b);
}

/* setCornerRowHalos(nVar = nwf), multiGrid.F90:1032-1357 */
static void crh_copy(const OrcBlock* b, Dims d, const AdfbParams* prm, long dst, long src) {
    for (int l = 0; l < 5; l++) W(dst, l) = W(src, l);
    b->p[dst] = b->p[src];
    if (prm->equations != ADFB_EULER) b->rlv[dst] = b->rlv[src];
    if (prm->equations == ADFB_RANS) b->rev[dst] = b->rev[src];
}
static int imin_(int a, int b_) { return a < b_ ? a : b_; }
static int imax_(int a, int b_) { return a > b_ ? a : b_; }
void orc_mg_corner_row_halos(const OrcBlock* b, const AdfbParams* prm) {
    Dims d = dims_of(b);
    int mm, ll;
    mm = imin_(3, d.jl); ll = imax_(2, d.ny);
    for (int k = 2; k <= d.kl; k++) {
        const int js[4] = {2, mm, d.jl, ll};
        for (int q = 0; q < 4; q++) { crh_copy(b, d, prm, IDX(1, js[q], k), IDX(2, js[q], k)); }
        for (int q = 0; q < 4; q++) { crh_copy(b, d, prm, IDX(d.ie, js[q], k), IDX(d.il, js[q], k)); }
    }
    mm = imin_(3, d.kl); ll = imax_(2, d.nz);
    for (int j = 3; j <= d.ny; j++) {
        const int ks[4] = {2, mm, d.kl, ll};
        for (int q = 0; q < 4; q++) { crh_copy(b, d, prm, IDX(1, j, ks[q]), IDX(2, j, ks[q])); }
        for (int q = 0; q < 4; q++) { crh_copy(b, d, prm, IDX(d.ie, j, ks[q]), IDX(d.il, j, ks[q])); }
    }
    mm = imin_(3, d.il); ll = imax_(2, d.nx);
    for (int k = 3; k <= d.nz; k++) {
        const int is[4] = {2, mm, d.il, ll};
        for (int q = 0; q < 4; q++) { crh_copy(b, d, prm, IDX(is[q], 1, k), IDX(is[q], 2, k)); }
        for (int q = 0; q < 4; q++) { crh_copy(b, d, prm, IDX(is[q], d.je, k), IDX(is[q], d.jl, k)); }
    }
    mm = imin_(3, d.kl); ll = imax_(2, d.nz);
    for (int i = 1; i <= d.ie; i++) {
        const int ks[4] = {2, mm, d.kl, ll};
        for (int q = 0; q < 4; q++) { crh_copy(b, d, prm, IDX(i, 1, ks[q]), IDX(i, 2, ks[q])); }
        for (int q = 0; q < 4; q++) { crh_copy(b, d, prm, IDX(i, d.je, ks[q]), IDX(i, d.jl, ks[q])); }
    }
    mm = imin_(3, d.il); ll = imax_(2, d.nx);
    for (int j = 1; j <= d.je; j++) {
        const int is[4] = {2, mm, d.il, ll};
        for (int q = 0; q < 4; q++) { crh_copy(b, d, prm, IDX(is[q], j, 1), IDX(is[q], j, 2)); }
        for (int q = 0; q < 4; q++) { crh_copy(b, d, prm, IDX(is[q], j, d.ke), IDX(is[q], j, d.kl)); }
    }
    mm = imin_(3, d.jl); ll = imax_(2, d.ny);
    for (int i = 1; i <= d.ie; i++) {
        const int js[4] = {2, mm, d.jl, ll};
        for (int q = 0; q < 4; q++) { crh_copy(b, d, prm, IDX(i, js[q], 1), IDX(i, js[q], 2)); }
        for (int q = 0; q < 4; q++) { crh_copy(b, d, prm, IDX(i, js[q], d.ke), IDX(i, js[q], d.kl)); }
    }
}

/* transferToCoarseGrid, restriction part (multiGrid.F90:88-225): `b` is the COARSE block, `f` its fine block.
   Restricted residual -> wr, volume-weighted state -> w, p, rev; then computeEtotBlock / computeLamViscosity /
   computeEddyViscosity on the owned cells and setCornerRowHalos. */
void orc_mg_restrict(const OrcBlock* b, const OrcBlock* f, const AdfbParams* prm, const int32_t* mgIFine, const int32_t* mgJFine,
                     const int32_t* mgKFine, const double* mgIWeight, const double* mgJWeight, const double* mgKWeight) {
    Dims d = dims_of(b);
    Dims df = dims_of(f);
    const double blankFact = one;
#define FIDX(i, j, k) ((long)(i) + df.NI * ((long)(j) + df.NJ * (long)(k)))
    for (int k = 2; k <= d.kl; k++) {
        int kk = mgKFine[k], kk1 = mgKFine[k + (d.ke + 1)];
        for (int j = 2; j <= d.jl; j++) {
            int jj = mgJFine[j], jj1 = mgJFine[j + (d.je + 1)];
            for (int i = 2; i <= d.il; i++) {
                int ii = mgIFine[i], ii1 = mgIFine[i + (d.ie + 1)];
                long c = IDX(i, j, k);
                double weigth = mgKWeight[k] * mgJWeight[j] * mgIWeight[i];
                /* the reference's two summation orders: volumes / residuals (ii1 before jj1) ... */
                const long a000 = FIDX(ii, jj, kk), a100 = FIDX(ii1, jj, kk), a010 = FIDX(ii, jj1, kk), a110 = FIDX(ii1, jj1, kk);
                const long a001 = FIDX(ii, jj, kk1), a101 = FIDX(ii1, jj, kk1), a011 = FIDX(ii, jj1, kk1), a111 = FIDX(ii1, jj1, kk1);
                const double* v = f->vol;
                double vola = v[a000] + v[a100] + v[a010] + v[a110] + v[a001] + v[a101] + v[a011] + v[a111];
                vola = one / vola;
                for (int l = 0; l < 5; l++) {
                    const double* r = f->dw + (long)l * df.N;
                    WR(c, l) = (r[a000] + r[a010] + r[a100] + r[a110] + r[a001] + r[a011] + r[a101] + r[a111]) * weigth * blankFact;
                }
                /* ... and the state (jj1 before ii1) */
                for (int l = 0; l < 4; l++) {
                    const double* s = f->w + (long)l * df.N;
                    W(c, l) = (v[a000] * s[a000] + v[a010] * s[a010] + v[a100] * s[a100] + v[a110] * s[a110] + v[a001] * s[a001] +
                               v[a011] * s[a011] + v[a101] * s[a101] + v[a111] * s[a111]) * vola;
                }
                {
                    const double* s = f->p;
                    b->p[c] = (v[a000] * s[a000] + v[a010] * s[a010] + v[a100] * s[a100] + v[a110] * s[a110] + v[a001] * s[a001] +
                               v[a011] * s[a011] + v[a101] * s[a101] + v[a111] * s[a111]) * vola;
                    s = f->rev;
                    b->rev[c] = (v[a000] * s[a000] + v[a010] * s[a010] + v[a100] * s[a100] + v[a110] * s[a110] + v[a001] * s[a001] +
                                 v[a011] * s[a011] + v[a101] * s[a101] + v[a111] * s[a111]) * vola;
                }
            }
        }
    }
#undef FIDX
    orc_etot(b, prm, 2, d.il, 2, d.jl, 2, d.kl);
    orc_lam_viscosity(b, prm, 0);
    orc_eddy_viscosity(b, prm, 0);
    orc_mg_corner_row_halos(b, prm);
}

/* transferToCoarseGrid :270-288: w1 / p1 = restricted solution incl. the first halos */
void orc_mg_store_w1(const OrcBlock* b) {
    Dims d = dims_of(b);
    for (int k = 1; k <= d.ke; k++) for (int j = 1; j <= d.je; j++) for (int i = 1; i <= d.ie; i++) {
        long c = IDX(i, j, k);
        for (int l = 0; l < 5; l++) b->w1[(long)l * d.N + c] = W(c, l);
        b->p1[c] = b->p[c];
    }
}

/* transferToCoarseGrid :296-322: residual forcing term */
void orc_mg_forcing(const OrcBlock* b, const AdfbParams* prm) {
    Dims d = dims_of(b);
    for (int l = 0; l < 5; l++)
        for (int k = 2; k <= d.kl; k++) for (int j = 2; j <= d.jl; j++) for (int i = 2; i <= d.il; i++) {
            long c = IDX(i, j, k);
            double tmp = prm->fcoll * WR(c, l);
            WR(c, l) = tmp - DW(c, l);
            DW(c, l) = tmp;
        }
}

/* setCorrectionsCoarseHalos, multiGrid.F90:1359-1503: `b` is the coarse block holding the corrections in w */
static void corr_halos(const OrcBlock* b, Dims d, int nSub, const AdfbSubface* sfs, double fact, int nVarInt) {
    for (int m = 0; m < nSub; m++) {
        const AdfbSubface* sf = &sfs[m];
        long o1, o2, sa, sb;
        switch (sf->faceId) {
            case ADFB_IMIN: o1 = 1; o2 = 2; sa = d.sJ; sb = d.sK; break;
            case ADFB_IMAX: o1 = d.ie; o2 = d.il; sa = d.sJ; sb = d.sK; break;
            case ADFB_JMIN: o1 = d.sJ; o2 = 2 * d.sJ; sa = 1; sb = d.sK; break;
            case ADFB_JMAX: o1 = d.je * d.sJ; o2 = d.jl * d.sJ; sa = 1; sb = d.sK; break;
            case ADFB_KMIN: o1 = d.sK; o2 = 2 * d.sK; sa = 1; sb = d.sJ; break;
            default: o1 = d.ke * d.sK; o2 = d.kl * d.sK; sa = 1; sb = d.sJ; break;
        }
        long na = sf->icEnd - sf->icBeg + 1, nb = sf->jcEnd - sf->jcBeg + 1;
        if (sf->bcType == ADFB_BC_SYMM) {
            for (int j = sf->jcBeg; j <= sf->jcEnd; j++) for (int i = sf->icBeg; i <= sf->icEnd; i++) {
                long q = i * sa + j * sb, c1 = o1 + q, c2 = o2 + q;
                long o = (i - sf->icBeg) + na * (j - sf->jcBeg);
                double nnx = sf->norm[o], nny = sf->norm[o + na * nb], nnz = sf->norm[o + 2 * na * nb];
                double vn = two * (W(c2, IVX) * nnx + W(c2, IVY) * nny + W(c2, IVZ) * nnz);
                W(c1, IRHO) = W(c2, IRHO);
                W(c1, IVX) = W(c2, IVX) - vn * nnx;
                W(c1, IVY) = W(c2, IVY) - vn * nny;
                W(c1, IVZ) = W(c2, IVZ) - vn * nnz;
                W(c1, IRHOE) = W(c2, IRHOE);
                for (int l = 5; l < nVarInt; l++) W(c1, l) = W(c2, l);
            }
        } else {
            for (int l = 0; l < nVarInt; l++)
                for (int j = sf->jcBeg; j <= sf->jcEnd; j++) for (int i = sf->icBeg; i <= sf->icEnd; i++) {
                    long q = i * sa + j * sb;
                    W(o1 + q, l) = fact * W(o2 + q, l);
                }
        }
    }
}

/* transferToFineGrid(corrections = .true.), multiGrid.F90:326-590 up to (not including) the BCs and the exchange:
   `b` is the FINE block, `cb` its coarse block (whose w is overwritten by the corrections, like the reference),
   sfs the COARSE block's subfaces.  mgICoarse etc. are the fine block's tables. */
void orc_mg_prolong(const OrcBlock* b, const OrcBlock* cb, const AdfbParams* prm, int nSubC, const AdfbSubface* sfs,
                    const int32_t* mgICoarse, const int32_t* mgJCoarse, const int32_t* mgKCoarse) {
    Dims d = dims_of(b);
    Dims dc = dims_of(cb);
    const int nVarInt = 5;
    {
        Dims d = dc;  /* shadow: W() on the coarse block */
        const OrcBlock* b = cb;
        for (int k = 1; k <= d.ke; k++) for (int j = 1; j <= d.je; j++) for (int i = 1; i <= d.ie; i++) {
            long c = IDX(i, j, k);
            W(c, IRHO) = W(c, IRHO) - b->w1[0 * d.N + c];
            W(c, IVX) = W(c, IVX) - b->w1[1 * d.N + c];
            W(c, IVY) = W(c, IVY) - b->w1[2 * d.N + c];
            W(c, IVZ) = W(c, IVZ) - b->w1[3 * d.N + c];
            W(c, IRHOE) = b->p[c] - b->p1[c];
        }
        double fact = prm->mgBoundCorr == 0 ? zero : one;
        corr_halos(b, d, nSubC, sfs, fact, nVarInt);
    }
#define CIDX(i, j, k) ((long)(i) + dc.NI * ((long)(j) + dc.NJ * (long)(k)))
    for (int k = 2; k <= d.kl; k++) {
        int kk = mgKCoarse[k], kk1 = mgKCoarse[k + (d.ke + 1)];
        for (int j = 2; j <= d.jl; j++) {
            int jj = mgJCoarse[j], jj1 = mgJCoarse[j + (d.je + 1)];
            for (int i = 2; i <= d.il; i++) {
                int ii = mgICoarse[i], ii1 = mgICoarse[i + (d.ie + 1)];
                long c = IDX(i, j, k);
                for (int l = 0; l < nVarInt; l++) {
                    const double* ww = cb->w + (long)l * dc.N;
                    DW(c, l) = 0.421875 * ww[CIDX(ii, jj, kk)] +
                               0.140625 * (ww[CIDX(ii1, jj, kk)] + ww[CIDX(ii, jj1, kk)] + ww[CIDX(ii, jj, kk1)]) +
                               0.046875 * (ww[CIDX(ii1, jj1, kk)] + ww[CIDX(ii1, jj, kk1)] + ww[CIDX(ii, jj1, kk1)]) +
                               0.015625 * ww[CIDX(ii1, jj1, kk1)];
                }
            }
        }
    }
#undef CIDX
    for (int k = 2; k <= d.kl; k++) for (int j = 2; j <= d.jl; j++) for (int i = 2; i <= d.il; i++) {
        long c = IDX(i, j, k);
        W(c, IRHO) = W(c, IRHO) + DW(c, IRHO);
        W(c, IVX) = W(c, IVX) + DW(c, IVX);
        W(c, IVY) = W(c, IVY) + DW(c, IVY);
        W(c, IVZ) = W(c, IVZ) + DW(c, IVZ);
        b->p[c] = b->p[c] + DW(c, IRHOE);
        W(c, IRHO) = dmax(W(c, IRHO), 1.e-4 * prm->rhoInf);
        b->p[c] = dmax(b->p[c], 1.e-4 * prm->pInfCorr);
    }
    orc_etot(b, prm, 2, d.il, 2, d.jl, 2, d.kl);
    orc_lam_viscosity(b, prm, 0);
    orc_eddy_viscosity(b, prm, 0);
}

/* extrapolateSolution / extrapolateViscosities, multiGrid.F90:656-737 / 739-823: constant extrapolation into the halos,
   i, then j (taking the i halos along), then k */
static void extrapolate_halos(Dims d, double* a) {
    for (int k = 2; k <= d.kl; k++) for (int j = 2; j <= d.jl; j++) {
        a[IDX(0, j, k)] = a[IDX(2, j, k)]; a[IDX(1, j, k)] = a[IDX(2, j, k)];
        a[IDX(d.ie, j, k)] = a[IDX(d.il, j, k)]; a[IDX(d.ib, j, k)] = a[IDX(d.il, j, k)];
    }
    for (int k = 2; k <= d.kl; k++) for (int i = 0; i <= d.ib; i++) {
        a[IDX(i, 0, k)] = a[IDX(i, 2, k)]; a[IDX(i, 1, k)] = a[IDX(i, 2, k)];
        a[IDX(i, d.je, k)] = a[IDX(i, d.jl, k)]; a[IDX(i, d.jb, k)] = a[IDX(i, d.jl, k)];
    }
    for (int j = 0; j <= d.jb; j++) for (int i = 0; i <= d.ib; i++) {
        a[IDX(i, j, 0)] = a[IDX(i, j, 2)]; a[IDX(i, j, 1)] = a[IDX(i, j, 2)];
        a[IDX(i, j, d.ke)] = a[IDX(i, j, d.kl)]; a[IDX(i, j, d.kb)] = a[IDX(i, j, d.kl)];
    }
}

/* transferToFineGrid(corrections = .false.), multiGrid.F90:326-590: the full-multigrid start-up step that interpolates
   the SOLUTION of the coarse block `cb` to the fine block `b` (all nw variables, the pressure in the place of the total
   energy), up to (not including) the boundary conditions and the exchanges.  Like the reference, the coarse block's
   rho*E is overwritten by its pressure and its boundary halos by setCorrectionsCoarseHalos(fact = 1). */
void orc_mg_prolong_solution(const OrcBlock* b, const OrcBlock* cb, const AdfbParams* prm, int nSubC, const AdfbSubface* sfs,
                             const int32_t* mgICoarse, const int32_t* mgJCoarse, const int32_t* mgKCoarse) {
    Dims d = dims_of(b);
    Dims dc = dims_of(cb);
    const int nVarInt = b->nw;
    {
        Dims d = dc;  /* shadow: W() on the coarse block */
        const OrcBlock* b = cb;
        for (int k = 1; k <= d.ke; k++) for (int j = 1; j <= d.je; j++) for (int i = 1; i <= d.ie; i++) {
            long c = IDX(i, j, k);
            W(c, IRHOE) = b->p[c];
        }
        corr_halos(b, d, nSubC, sfs, one, nVarInt);
    }
#define CIDX(i, j, k) ((long)(i) + dc.NI * ((long)(j) + dc.NJ * (long)(k)))
    for (int k = 2; k <= d.kl; k++) {
        int kk = mgKCoarse[k], kk1 = mgKCoarse[k + (d.ke + 1)];
        for (int j = 2; j <= d.jl; j++) {
            int jj = mgJCoarse[j], jj1 = mgJCoarse[j + (d.je + 1)];
            for (int i = 2; i <= d.il; i++) {
                int ii = mgICoarse[i], ii1 = mgICoarse[i + (d.ie + 1)];
                long c = IDX(i, j, k);
                for (int l = 0; l < nVarInt; l++) {
                    const double* ww = cb->w + (long)l * dc.N;
                    W(c, l) = 0.421875 * ww[CIDX(ii, jj, kk)] +
                              0.140625 * (ww[CIDX(ii1, jj, kk)] + ww[CIDX(ii, jj1, kk)] + ww[CIDX(ii, jj, kk1)]) +
                              0.046875 * (ww[CIDX(ii1, jj1, kk)] + ww[CIDX(ii1, jj, kk1)] + ww[CIDX(ii, jj1, kk1)]) +
                              0.015625 * ww[CIDX(ii1, jj1, kk1)];
                }
            }
        }
    }
#undef CIDX
    for (int k = 2; k <= d.kl; k++) for (int j = 2; j <= d.jl; j++) for (int i = 2; i <= d.il; i++) {
        long c = IDX(i, j, k);
        b->p[c] = W(c, IRHOE);
    }
    orc_etot(b, prm, 2, d.il, 2, d.jl, 2, d.kl);
    for (int l = 0; l < b->nw; l++) extrapolate_halos(d, b->w + (long)l * d.N);
    extrapolate_halos(d, b->p);
    orc_lam_viscosity(b, prm, 0);
    orc_eddy_viscosity(b, prm, 0);
    if (prm->equations != ADFB_EULER) {
        extrapolate_halos(d, b->rlv);
        if (prm->equations == ADFB_RANS) extrapolate_halos(d, b->rev);
    }
}
