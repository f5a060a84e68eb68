/*
 * adflow_oracle.c -- CPU restatement of the ADflow residual hot path.
 * TEST INFRASTRUCTURE ONLY; see adflow_oracle.h (how parity is pinned).
 *
 * Operator order and arithmetic follow src/NKSolver/blockette.F90 (the
 * reference's own self-contained statement of the residual) applied to the
 * whole block as one tile; the block twins in src/solver/fluxes.F90 etc. are
 * cross-cited.  Where the Fortran repeats the same code for the i, j and k
 * directions the restatement uses one stride-parametrised body and calls it in
 * the reference's direction order, keeping the left-to-right association of
 * every expression so that results agree with a -O0 Fortran build to round-off.
 *
 * Build: gcc -O2 -fno-fast-math -ffp-contract=off (see oracle/Makefile).
 */
#include "orc_internal.h"

/* ------------------------------------------------------------------------ */
/* metrics: src/NKSolver/blockette.F90:854-960 (== metric_block,
   src/adjoint/adjointExtra.F90:179 ff. for the face normals). */
void orc_metrics(const OrcBlock* b) {
    Dims d = dims_of(b);
    double fact = b->rightHanded ? half : -half;
    double v1[3], v2[3];
    int i, j, k, l, m, n, c3;
    for (k = 1; k <= d.ke; k++) { n = k - 1;
        for (j = 1; j <= d.je; j++) { m = j - 1;
            for (i = 0; i <= d.ie; i++) {
                for (c3 = 0; c3 < 3; c3++) {
                    v1[c3] = X(IDX(i, j, n), c3) - X(IDX(i, m, k), c3);
                    v2[c3] = X(IDX(i, j, k), c3) - X(IDX(i, m, n), c3);
                }
                long c = IDX(i, j, k);
                b->si[0 * d.N + c] = fact * (v1[1] * v2[2] - v1[2] * v2[1]);
                b->si[1 * d.N + c] = fact * (v1[2] * v2[0] - v1[0] * v2[2]);
                b->si[2 * d.N + c] = fact * (v1[0] * v2[1] - v1[1] * v2[0]);
            } } }
    for (k = 1; k <= d.ke; k++) { n = k - 1;
        for (j = 0; j <= d.je; j++) {
            for (i = 1; i <= d.ie; i++) { l = i - 1;
                for (c3 = 0; c3 < 3; c3++) {
                    v1[c3] = X(IDX(i, j, n), c3) - X(IDX(l, j, k), c3);
                    v2[c3] = X(IDX(l, j, n), c3) - X(IDX(i, j, k), c3);
                }
                long c = IDX(i, j, k);
                b->sj[0 * d.N + c] = fact * (v1[1] * v2[2] - v1[2] * v2[1]);
                b->sj[1 * d.N + c] = fact * (v1[2] * v2[0] - v1[0] * v2[2]);
                b->sj[2 * d.N + c] = fact * (v1[0] * v2[1] - v1[1] * v2[0]);
            } } }
    for (k = 0; k <= d.ke; k++) {
        for (j = 1; j <= d.je; j++) { m = j - 1;
            for (i = 1; i <= d.ie; i++) { l = i - 1;
                for (c3 = 0; c3 < 3; c3++) {
                    v1[c3] = X(IDX(i, j, k), c3) - X(IDX(l, m, k), c3);
                    v2[c3] = X(IDX(l, j, k), c3) - X(IDX(i, m, k), c3);
                }
                long c = IDX(i, j, k);
                b->sk[0 * d.N + c] = fact * (v1[1] * v2[2] - v1[2] * v2[1]);
                b->sk[1 * d.N + c] = fact * (v1[2] * v2[0] - v1[0] * v2[2]);
                b->sk[2 * d.N + c] = fact * (v1[0] * v2[1] - v1[1] * v2[0]);
            } } }
}

/* volume_block: src/adjoint/adjointExtra.F90:5-177 */
static double volpym(double xp, double yp, double zp, const double* a, const double* bb,
                     const double* c, const double* dd) {
    return (xp - fourth * (a[0] + bb[0] + c[0] + dd[0])) *
               ((a[1] - c[1]) * (bb[2] - dd[2]) - (a[2] - c[2]) * (bb[1] - dd[1])) +
           (yp - fourth * (a[1] + bb[1] + c[1] + dd[1])) *
               ((a[2] - c[2]) * (bb[0] - dd[0]) - (a[0] - c[0]) * (bb[2] - dd[2])) +
           (zp - fourth * (a[2] + bb[2] + c[2] + dd[2])) *
               ((a[0] - c[0]) * (bb[1] - dd[1]) - (a[1] - c[1]) * (bb[0] - dd[0]));
}
void orc_volume(const OrcBlock* b) {
    Dims d = dims_of(b);
    const double haloCellRatio = 1e-10;
    int i, j, k, l, m, n, q;
    memset(b->vol, 0, sizeof(double) * d.N);
    for (k = 1; k <= d.ke; k++) { n = k - 1;
        for (j = 1; j <= d.je; j++) { m = j - 1;
            for (i = 1; i <= d.ie; i++) { l = i - 1;
                double P[8][3]; /* ijk, imk, imn, ijn, ljk, lmk, lmn, ljn */
                long id[8] = {IDX(i, j, k), IDX(i, m, k), IDX(i, m, n), IDX(i, j, n),
                              IDX(l, j, k), IDX(l, m, k), IDX(l, m, n), IDX(l, j, n)};
                for (q = 0; q < 8; q++) { P[q][0] = X(id[q], 0); P[q][1] = X(id[q], 1); P[q][2] = X(id[q], 2); }
                double xp = eighth * (P[0][0] + P[1][0] + P[2][0] + P[3][0] + P[4][0] + P[5][0] + P[6][0] + P[7][0]);
                double yp = eighth * (P[0][1] + P[1][1] + P[2][1] + P[3][1] + P[4][1] + P[5][1] + P[6][1] + P[7][1]);
                double zp = eighth * (P[0][2] + P[1][2] + P[2][2] + P[3][2] + P[4][2] + P[5][2] + P[6][2] + P[7][2]);
                const double *ijk = P[0], *imk = P[1], *imn = P[2], *ijn = P[3];
                const double *ljk = P[4], *lmk = P[5], *lmn = P[6], *ljn = P[7];
                double vp1 = volpym(xp, yp, zp, ijk, ijn, imn, imk);
                double vp2 = volpym(xp, yp, zp, ljk, lmk, lmn, ljn);
                double vp3 = volpym(xp, yp, zp, ijk, ljk, ljn, ijn);
                double vp4 = volpym(xp, yp, zp, imk, imn, lmn, lmk);
                double vp5 = volpym(xp, yp, zp, ijk, imk, lmk, ljk);
                double vp6 = volpym(xp, yp, zp, ijn, ljn, lmn, imn);
                b->vol[IDX(i, j, k)] = fabs(sixth * (vp1 + vp2 + vp3 + vp4 + vp5 + vp6));
            } } }
    for (k = 2; k <= d.kl; k++) for (j = 2; j <= d.jl; j++) {
        if (b->vol[IDX(1, j, k)] / b->vol[IDX(2, j, k)] < haloCellRatio) b->vol[IDX(1, j, k)] = b->vol[IDX(2, j, k)];
        if (b->vol[IDX(d.ie, j, k)] / b->vol[IDX(d.il, j, k)] < haloCellRatio) b->vol[IDX(d.ie, j, k)] = b->vol[IDX(d.il, j, k)];
    }
    for (k = 2; k <= d.kl; k++) for (i = 1; i <= d.ie; i++) {
        if (b->vol[IDX(i, 1, k)] / b->vol[IDX(i, 2, k)] < haloCellRatio) b->vol[IDX(i, 1, k)] = b->vol[IDX(i, 2, k)];
        if (b->vol[IDX(i, d.je, k)] / b->vol[IDX(i, d.jl, k)] < haloCellRatio) b->vol[IDX(i, d.je, k)] = b->vol[IDX(i, d.jl, k)];
    }
    for (j = 1; j <= d.je; j++) for (i = 1; i <= d.ie; i++) {
        if (b->vol[IDX(i, j, 1)] / b->vol[IDX(i, j, 2)] < haloCellRatio) b->vol[IDX(i, j, 1)] = b->vol[IDX(i, j, 2)];
        if (b->vol[IDX(i, j, d.ke)] / b->vol[IDX(i, j, d.kl)] < haloCellRatio) b->vol[IDX(i, j, d.ke)] = b->vol[IDX(i, j, d.kl)];
    }
}

/* ------------------------------------------------------------------------ */
/* computePressureSimple: src/utils/flowUtils.F90:867-930 */
void orc_pressure(const OrcBlock* b, const AdfbParams* prm, int includeHalos) {
    Dims d = dims_of(b);
    double gm1 = prm->gammaInf - one;
    int i0 = includeHalos ? 0 : 2, i1 = includeHalos ? d.ib : d.il;
    int j0 = includeHalos ? 0 : 2, j1 = includeHalos ? d.jb : d.jl;
    int k0 = includeHalos ? 0 : 2, k1 = includeHalos ? d.kb : d.kl;
    for (int k = k0; k <= k1; k++) for (int j = j0; j <= j1; j++) for (int i = i0; i <= i1; i++) {
        long c = IDX(i, j, k);
        double v2 = W(c, IVX) * W(c, IVX) + W(c, IVY) * W(c, IVY) + W(c, IVZ) * W(c, IVZ);
        b->p[c] = gm1 * (W(c, IRHOE) - half * W(c, IRHO) * v2);
        b->p[c] = dmax(b->p[c], 1.e-4 * prm->pInfCorr);
    }
}
/* computeLamViscosity: src/utils/flowUtils.F90:1201-1323 (no k-correction: SA) */
void orc_lam_viscosity(const OrcBlock* b, const AdfbParams* prm, int includeHalos) {
    Dims d = dims_of(b);
    if (prm->equations == ADFB_EULER) return;
    int i0 = includeHalos ? 1 : 2, i1 = includeHalos ? d.ie : d.il;
    int j0 = includeHalos ? 1 : 2, j1 = includeHalos ? d.je : d.jl;
    int k0 = includeHalos ? 1 : 2, k1 = includeHalos ? d.ke : d.kl;
    for (int k = k0; k <= k1; k++) for (int j = j0; j <= j1; j++) for (int i = i0; i <= i1; i++) {
        long c = IDX(i, j, k);
        double T = b->p[c] / (prm->RGas * W(c, IRHO));
        b->rlv[c] = prm->muSuth * ((prm->TSuth + prm->SSuth) / (T + prm->SSuth)) * pow(T / prm->TSuth, 1.5);
    }
}
/* saEddyViscosity: src/turbulence/turbUtils.F90:657-712 */
void orc_eddy_viscosity(const OrcBlock* b, const AdfbParams* prm, int includeHalos) {
    Dims d = dims_of(b);
    if (prm->equations != ADFB_RANS) return;
    if (b->level > 1) return;  /* computeEddyViscosity returns at once on coarse levels, turbUtils.F90:606-616 */
    int i0 = includeHalos ? 1 : 2, i1 = includeHalos ? d.ie : d.il;
    int j0 = includeHalos ? 1 : 2, j1 = includeHalos ? d.je : d.jl;
    int k0 = includeHalos ? 1 : 2, k1 = includeHalos ? d.ke : d.kl;
    double cv13 = prm->rsaCv1 * prm->rsaCv1 * prm->rsaCv1;
    for (int k = k0; k <= k1; k++) for (int j = j0; j <= j1; j++) for (int i = i0; i <= i1; i++) {
        long c = IDX(i, j, k);
        double rnuSA = W(c, ITU1) * W(c, IRHO);
        double chi = rnuSA / b->rlv[c];
        double chi3 = chi * chi * chi;
        double fv1 = chi3 / (chi3 + cv13);
        b->rev[c] = fv1 * rnuSA;
    }
}
/* computeEtotBlock, cpConstant, no k: src/utils/flowUtils.F90:551-672 */
void orc_etot(const OrcBlock* b, const AdfbParams* prm, int i0, int i1, int j0, int j1, int k0, int k1) {
    Dims d = dims_of(b);
    double ovgm1 = one / (prm->gammaInf - one);
    for (int k = k0; k <= k1; k++) for (int j = j0; j <= j1; j++) for (int i = i0; i <= i1; i++) {
        long c = IDX(i, j, k);
        W(c, IRHOE) = ovgm1 * b->p[c] +
                      half * W(c, IRHO) * (W(c, IVX) * W(c, IVX) + W(c, IVY) * W(c, IVY) + W(c, IVZ) * W(c, IVZ));
    }
}

/* ------------------------------------------------------------------------ */
/* timeStep: src/NKSolver/blockette.F90:1899-2148 (block twin
   src/solver/solverUtils.F90:43-355; the tile version always applies
   directional scaling).  No moving grids: sFace == 0. */
void orc_time_step(const OrcBlock* b, const AdfbParams* prm, int updateDt) {
    Dims d = dims_of(b);
    const double bfac = 2.0;
    double plim = 0.001 * prm->pInfCorr;
    double clim2 = 0.000001 * prm->gammaInf * prm->pInfCorr / prm->rhoInf;
    double gam = prm->gammaInf, adis = prm->adis, asf = prm->acousticScaleFactor;
    int viscous = prm->equations != ADFB_EULER;
    const double *si = b->si, *sj = b->sj, *sk = b->sk;
    for (int k = 1; k <= d.ke; k++) for (int j = 1; j <= d.je; j++) for (int i = 1; i <= d.ie; i++) {
        long c = IDX(i, j, k);
        double uux = W(c, IVX), uuy = W(c, IVY), uuz = W(c, IVZ);
        double cc2 = gam * b->p[c] / W(c, IRHO);
        cc2 = dmax(cc2, clim2);
        double sFace = zero;
        double sx = si[c - 1] + si[c], sy = si[d.N + c - 1] + si[d.N + c], sz = si[2 * d.N + c - 1] + si[2 * d.N + c];
        double qsi = uux * sx + uuy * sy + uuz * sz - sFace;
        double ri = half * (fabs(qsi) + asf * sqrt(cc2 * (sx * sx + sy * sy + sz * sz)));
        sx = sj[c - d.sJ] + sj[c]; sy = sj[d.N + c - d.sJ] + sj[d.N + c]; sz = sj[2 * d.N + c - d.sJ] + sj[2 * d.N + c];
        double qsj = uux * sx + uuy * sy + uuz * sz - sFace;
        double rj = half * (fabs(qsj) + asf * sqrt(cc2 * (sx * sx + sy * sy + sz * sz)));
        sx = sk[c - d.sK] + sk[c]; sy = sk[d.N + c - d.sK] + sk[d.N + c]; sz = sk[2 * d.N + c - d.sK] + sk[2 * d.N + c];
        double qsk = uux * sx + uuy * sy + uuz * sz - sFace;
        double rk = half * (fabs(qsk) + asf * sqrt(cc2 * (sx * sx + sy * sy + sz * sz)));
        if (updateDt) b->dtl[c] = ri + rj + rk;
        if (b->level > 1) {  /* doScaling = dirScaling .and. currentLevel <= groundLevel, solverUtils.F90:106 */
            b->radI[c] = ri; b->radJ[c] = rj; b->radK[c] = rk;
            continue;
        }
        ri = dmax(ri, eps_); rj = dmax(rj, eps_); rk = dmax(rk, eps_);
        double rij = pow(ri / rj, adis), rjk = pow(rj / rk, adis), rki = pow(rk / ri, adis);
        b->radI[c] = ri * (one + one / rij + rki);
        b->radJ[c] = rj * (one + one / rjk + rij);
        b->radK[c] = rk * (one + one / rki + rjk);
    }
    if (!updateDt) return;
    if (viscous) {
        for (int k = 2; k <= d.kl; k++) for (int j = 2; j <= d.jl; j++) for (int i = 2; i <= d.il; i++) {
            long c = IDX(i, j, k);
            double rmu = b->rlv[c];
            rmu = rmu + b->rev[c];
            rmu = half * rmu / (W(c, IRHO) * b->vol[c]);
            double sx = si[c] + si[c - 1], sy = si[d.N + c] + si[d.N + c - 1], sz = si[2 * d.N + c] + si[2 * d.N + c - 1];
            double vsi = rmu * (sx * sx + sy * sy + sz * sz);
            b->dtl[c] = b->dtl[c] + vsi;
            sx = sj[c] + sj[c - d.sJ]; sy = sj[d.N + c] + sj[d.N + c - d.sJ]; sz = sj[2 * d.N + c] + sj[2 * d.N + c - d.sJ];
            double vsj = rmu * (sx * sx + sy * sy + sz * sz);
            b->dtl[c] = b->dtl[c] + vsj;
            sx = sk[c] + sk[c - d.sK]; sy = sk[d.N + c] + sk[d.N + c - d.sK]; sz = sk[2 * d.N + c] + sk[2 * d.N + c - d.sK];
            double vsk = rmu * (sx * sx + sy * sy + sz * sz);
            b->dtl[c] = b->dtl[c] + vsk;
        }
    }
    for (int k = 2; k <= d.kl; k++) for (int j = 2; j <= d.jl; j++) for (int i = 2; i <= d.il; i++) {
        long c = IDX(i, j, k);
        const double* p = b->p;
        double dpi = fabs(p[c + 1] - two * p[c] + p[c - 1]) / (p[c + 1] + two * p[c] + p[c - 1] + plim);
        double dpj = fabs(p[c + d.sJ] - two * p[c] + p[c - d.sJ]) / (p[c + d.sJ] + two * p[c] + p[c - d.sJ] + plim);
        double dpk = fabs(p[c + d.sK] - two * p[c] + p[c - d.sK]) / (p[c + d.sK] + two * p[c] + p[c - d.sK] + plim);
        double rfl = one / (one + bfac * (dpi + dpj + dpk));
        b->dtl[c] = rfl / b->dtl[c];
    }
}

/* ------------------------------------------------------------------------ */
/* inviscidCentralFlux: src/NKSolver/blockette.F90:2150-2455
   (block twin src/solver/fluxes.F90:4-401).  One body for the three face
   directions; called i, j, k like the reference.  No rotation source
   (blockIsMoving == false). */
static void central_dir(const OrcBlock* b, Dims d, const double* s, const int8_t* por, long sd,
                        int i0, int j0, int k0) {
    for (int k = k0; k <= d.kl; k++) for (int j = j0; j <= d.jl; j++) for (int i = i0; i <= d.il; i++) {
        long c = IDX(i, j, k), cp = c + sd;
        double s1 = s[c], s2 = s[d.N + c], s3 = s[2 * d.N + c];
        double sFace = zero;
        double vnp = W(cp, IVX) * s1 + W(cp, IVY) * s2 + W(cp, IVZ) * s3;
        double vnm = W(c, IVX) * s1 + W(c, IVY) * s2 + W(c, IVZ) * s3;
        double porVel = one, porFlux = half;
        if (por[c] == ADFB_NOFLUX) porFlux = zero;
        if (por[c] == ADFB_BOUNDFLUX) { porVel = zero; vnp = sFace; vnm = sFace; }
        porVel = porVel * porFlux;
        double qsp = (vnp - sFace) * porVel, qsm = (vnm - sFace) * porVel;
        double rqsp = qsp * W(cp, IRHO), rqsm = qsm * W(c, IRHO);
        double pa = porFlux * (b->p[cp] + b->p[c]);
        double fs = rqsp + rqsm;
        DW(cp, IRHO) -= fs; DW(c, IRHO) += fs;
        fs = rqsp * W(cp, IVX) + rqsm * W(c, IVX) + pa * s1;
        DW(cp, IMX) -= fs; DW(c, IMX) += fs;
        fs = rqsp * W(cp, IVY) + rqsm * W(c, IVY) + pa * s2;
        DW(cp, IMY) -= fs; DW(c, IMY) += fs;
        fs = rqsp * W(cp, IVZ) + rqsm * W(c, IVZ) + pa * s3;
        DW(cp, IMZ) -= fs; DW(c, IMZ) += fs;
        fs = qsp * W(cp, IRHOE) + qsm * W(c, IRHOE) + porFlux * (vnp * b->p[cp] + vnm * b->p[c]);
        DW(cp, IRHOE) -= fs; DW(c, IRHOE) += fs;
    }
}
void orc_central_flux(const OrcBlock* b, const AdfbParams* prm) {
    (void)prm;
    Dims d = dims_of(b);
    central_dir(b, d, b->si, b->porI, d.sI, 1, 2, 2);
    central_dir(b, d, b->sj, b->porJ, d.sJ, 2, 1, 2);
    central_dir(b, d, b->sk, b->porK, d.sK, 2, 2, 1);
}

/* ------------------------------------------------------------------------ */
/* inviscidDissFluxScalar: src/NKSolver/blockette.F90:3029-3339
   (block twin src/solver/fluxes.F90:1049-1436).  No dissipation continuation. */
static void diss_scalar_dir(const OrcBlock* b, Dims d, const double* rad, const double* dss,
                            const int8_t* por, long sd, int i0, int j0, int k0, double fis2, double fis4) {
    const double dssMax = 0.25;
    const double* p = b->p;
    for (int k = k0; k <= d.kl; k++) for (int j = j0; j <= d.jl; j++) for (int i = i0; i <= d.il; i++) {
        long c = IDX(i, j, k), cp = c + sd, cpp = c + 2 * sd, cm = c - sd;
        double ppor = zero;
        if (por[c] == ADFB_NORMALFLUX) ppor = half;
        double rrad = ppor * (rad[c] + rad[cp]);
        double dis2 = fis2 * rrad * dmin(dssMax, dmax(dss[c], dss[cp]));
        double dis4 = fdim_(fis4 * rrad, dis2);
        double ddw, fs;
        ddw = W(cp, IRHO) - W(c, IRHO);
        fs = dis2 * ddw - dis4 * (W(cpp, IRHO) - W(cm, IRHO) - three * ddw);
        FW(cp, IRHO) += fs; FW(c, IRHO) -= fs;
        for (int l = IVX; l <= IVZ; l++) {
            ddw = W(cp, l) * W(cp, IRHO) - W(c, l) * W(c, IRHO);
            fs = dis2 * ddw - dis4 * (W(cpp, l) * W(cpp, IRHO) - W(cm, l) * W(cm, IRHO) - three * ddw);
            FW(cp, l) += fs; FW(c, l) -= fs;
        }
        ddw = (W(cp, IRHOE) + p[cp]) - (W(c, IRHOE) + p[c]);
        fs = dis2 * ddw - dis4 * ((W(cpp, IRHOE) + p[cpp]) - (W(cm, IRHOE) + p[cm]) - three * ddw);
        FW(cp, IRHOE) += fs; FW(c, IRHOE) -= fs;
    }
}
void orc_diss_scalar(const OrcBlock* b, const AdfbParams* prm, double rFil) {
    Dims d = dims_of(b);
    double sslim;
    if (prm->equations == ADFB_EULER) {
        sslim = 0.001 * prm->pInfCorr;
        memcpy(b->ss, b->p, sizeof(double) * d.N);
    } else {
        sslim = 0.001 * prm->pInfCorr / pow(prm->rhoInf, prm->gammaInf);
        for (int k = 0; k <= d.kb; k++) for (int j = 0; j <= d.jb; j++) for (int i = 0; i <= d.ib; i++) {
            long c = IDX(i, j, k);
            b->ss[c] = b->p[c] / pow(W(c, IRHO), prm->gammaInf);
        }
    }
    const double* ss = b->ss;
    for (int k = 1; k <= d.ke; k++) for (int j = 1; j <= d.je; j++) for (int i = 1; i <= d.ie; i++) {
        long c = IDX(i, j, k);
        b->dss[0 * d.N + c] = fabs((ss[c + 1] - two * ss[c] + ss[c - 1]) / (ss[c + 1] + two * ss[c] + ss[c - 1] + sslim));
        b->dss[1 * d.N + c] = fabs((ss[c + d.sJ] - two * ss[c] + ss[c - d.sJ]) / (ss[c + d.sJ] + two * ss[c] + ss[c - d.sJ] + sslim));
        b->dss[2 * d.N + c] = fabs((ss[c + d.sK] - two * ss[c] + ss[c - d.sK]) / (ss[c + d.sK] + two * ss[c] + ss[c - d.sK] + sslim));
    }
    double fis2 = rFil * prm->vis2, fis4 = rFil * prm->vis4, sfil = one - rFil;
    for (long q = 0; q < 5 * d.N; q++) b->fw[q] = sfil * b->fw[q];
    diss_scalar_dir(b, d, b->radI, b->dss + 0 * d.N, b->porI, d.sI, 1, 2, 2, fis2, fis4);
    diss_scalar_dir(b, d, b->radJ, b->dss + 1 * d.N, b->porJ, d.sJ, 2, 1, 2, fis2, fis4);
    diss_scalar_dir(b, d, b->radK, b->dss + 2 * d.N, b->porK, d.sK, 2, 2, 1, fis2, fis4);
}

/* computeSpeedOfSoundSquared: src/NKSolver/blockette.F90:5168-5203 (no k) */
void orc_speed_of_sound(const OrcBlock* b, const AdfbParams* prm) {
    Dims d = dims_of(b);
    for (int k = 1; k <= d.ke; k++) for (int j = 1; j <= d.je; j++) for (int i = 1; i <= d.ie; i++) {
        long c = IDX(i, j, k);
        b->aa[c] = prm->gammaInf * b->p[c] / W(c, IRHO);
    }
}

/* ------------------------------------------------------------------------ */
/* allNodalGradients: src/NKSolver/blockette.F90:5205-5515
   (block twin src/utils/flowUtils.F90:1676-2026).  gradient arrays ordered
   ux,uy,uz,vx,vy,vz,wx,wy,wz,qx,qy,qz.  `sd` is the sweep direction, t1/t2 the
   transverse strides in the reference's order. */
static void nodal_dir(const OrcBlock* b, Dims d, const double* s, long sd, long t1, long t2,
                      int i1, int j1, int k1, int dirAxis) {
    for (int k = 1; k <= k1; k++) for (int j = 1; j <= j1; j++) for (int i = 1; i <= i1; i++) {
        long c = IDX(i, j, k);
        double sv[3];
        for (int m = 0; m < 3; m++) {
            const double* sm = s + (long)m * d.N;
            sv[m] = sm[c - sd] + sm[c - sd + t1] + sm[c - sd + t2] + sm[c - sd + t1 + t2] +
                    sm[c] + sm[c + t1] + sm[c + t2] + sm[c + t1 + t2];
        }
        double ubar = fourth * (W(c, IVX) + W(c + t1, IVX) + W(c + t2, IVX) + W(c + t1 + t2, IVX));
        double vbar = fourth * (W(c, IVY) + W(c + t1, IVY) + W(c + t2, IVY) + W(c + t1 + t2, IVY));
        double wbar = fourth * (W(c, IVZ) + W(c + t1, IVZ) + W(c + t2, IVZ) + W(c + t1 + t2, IVZ));
        double a2 = fourth * (b->aa[c] + b->aa[c + t1] + b->aa[c + t2] + b->aa[c + t1 + t2]);
        int idx = dirAxis == 0 ? i : (dirAxis == 1 ? j : k);
        int iend = dirAxis == 0 ? d.ie : (dirAxis == 1 ? d.je : d.ke);
        if (idx > 1) {
            long n = c - sd;
            for (int m = 0; m < 3; m++) {
                GR(n, 0 + m) += ubar * sv[m];
                GR(n, 3 + m) += vbar * sv[m];
                GR(n, 6 + m) += wbar * sv[m];
                GR(n, 9 + m) -= a2 * sv[m];
            }
        }
        if (idx < iend) {
            long n = c;
            for (int m = 0; m < 3; m++) {
                GR(n, 0 + m) -= ubar * sv[m];
                GR(n, 3 + m) -= vbar * sv[m];
                GR(n, 6 + m) -= wbar * sv[m];
                GR(n, 9 + m) += a2 * sv[m];
            }
        }
    }
}
void orc_nodal_gradients(const OrcBlock* b) {
    Dims d = dims_of(b);
    memset(b->grad, 0, sizeof(double) * 12 * d.N);
    nodal_dir(b, d, b->sk, d.sK, d.sI, d.sJ, d.il, d.jl, d.ke, 2);
    nodal_dir(b, d, b->sj, d.sJ, d.sI, d.sK, d.il, d.je, d.kl, 1);
    nodal_dir(b, d, b->si, d.sI, d.sJ, d.sK, d.ie, d.jl, d.kl, 0);
    const double* vol = b->vol;
    for (int k = 1; k <= d.kl; k++) for (int j = 1; j <= d.jl; j++) for (int i = 1; i <= d.il; i++) {
        long c = IDX(i, j, k);
        double oVol = one / (vol[c] + vol[c + d.sK] + vol[c + 1] + vol[c + 1 + d.sK] + vol[c + d.sJ] +
                             vol[c + d.sJ + d.sK] + vol[c + 1 + d.sJ] + vol[c + 1 + d.sJ + d.sK]);
        for (int m = 0; m < 12; m++) GR(c, m) = GR(c, m) * oVol;
    }
}

/* ------------------------------------------------------------------------ */
/* viscousFlux: src/NKSolver/blockette.F90:5517-6465
   (block twin src/solver/fluxes.F90:2534-3485; blockette QCR floor 1e-10).
   Wall tau/q storage (viscSubface) is not part of the residual and omitted. */
static void viscous_dir(const OrcBlock* b, const AdfbParams* prm, Dims d, const double* s, const int8_t* por,
                        long sd, long t1, long t2, int i0, int j0, int k0, double rFilv, int dir) {
    const double xminn = 1.e-10, twoThird = two * third, Ccr1 = 0.3;
    double gam = prm->gammaInf;
    for (int k = k0; k <= d.kl; k++) for (int j = j0; j <= d.jl; j++) for (int i = i0; i <= d.il; i++) {
        long c = IDX(i, j, k), cp = c + sd;
        double porv = half * rFilv;
        if (por[c] == ADFB_NOFLUX) porv = zero;
        double mul = porv * (b->rlv[c] + b->rlv[cp]);
        double mue = porv * (b->rev[c] + b->rev[cp]);
        double mut = mul + mue;
        double gm1 = half * (gam + gam) - one;
        double factLamHeat = one / (prm->prandtl * gm1);
        double factTurbHeat = one / (prm->prandtlTurb * gm1);
        double heatCoef = mul * factLamHeat + mue * factTurbHeat;
        double g[12];
        long n = c, n1 = c - t1 - t2, n2 = c - t2, n3 = c - t1;
        for (int m = 0; m < 12; m++) g[m] = fourth * (GR(n1, m) + GR(n2, m) + GR(n3, m) + GR(n, m));
        double ss3[3];
        for (int m = 0; m < 3; m++) {
            ss3[m] = eighth * (X(n1 + sd, m) - X(n1 - sd, m) + X(n3 + sd, m) - X(n3 - sd, m) +
                               X(n2 + sd, m) - X(n2 - sd, m) + X(n + sd, m) - X(n - sd, m));
        }
        double snrm = one / sqrt(ss3[0] * ss3[0] + ss3[1] * ss3[1] + ss3[2] * ss3[2]);
        double ssx = snrm * ss3[0], ssy = snrm * ss3[1], ssz = snrm * ss3[2];
        double corr;
        corr = g[0] * ssx + g[1] * ssy + g[2] * ssz - (W(cp, IVX) - W(c, IVX)) * snrm;
        double u_x = g[0] - corr * ssx, u_y = g[1] - corr * ssy, u_z = g[2] - corr * ssz;
        corr = g[3] * ssx + g[4] * ssy + g[5] * ssz - (W(cp, IVY) - W(c, IVY)) * snrm;
        double v_x = g[3] - corr * ssx, v_y = g[4] - corr * ssy, v_z = g[5] - corr * ssz;
        corr = g[6] * ssx + g[7] * ssy + g[8] * ssz - (W(cp, IVZ) - W(c, IVZ)) * snrm;
        double w_x = g[6] - corr * ssx, w_y = g[7] - corr * ssy, w_z = g[8] - corr * ssz;
        corr = g[9] * ssx + g[10] * ssy + g[11] * ssz + (b->aa[cp] - b->aa[c]) * snrm;
        double q_x = g[9] - corr * ssx, q_y = g[10] - corr * ssy, q_z = g[11] - corr * ssz;
        double fracDiv = twoThird * (u_x + v_y + w_z);
        double tauxxS = two * u_x - fracDiv, tauyyS = two * v_y - fracDiv, tauzzS = two * w_z - fracDiv;
        double tauxyS = u_y + v_x, tauxzS = u_z + w_x, tauyzS = v_z + w_y;
        q_x = heatCoef * q_x; q_y = heatCoef * q_y; q_z = heatCoef * q_z;
        double tauxx, tauyy, tauzz, tauxy, tauxz, tauyz;
        if (prm->useQCR) {
            double den = sqrt(u_x * u_x + u_y * u_y + u_z * u_z + v_x * v_x + v_y * v_y + v_z * v_z +
                              w_x * w_x + w_y * w_y + w_z * w_z);
            den = dmax(den, xminn);
            double fact = mue * Ccr1 / den;
            double Wxy = u_y - v_x, Wxz = u_z - w_x, Wyz = v_z - w_y;
            double Wyx = -Wxy, Wzx = -Wxz, Wzy = -Wyz;
            double exx = fact * (Wxy * tauxyS + Wxz * tauxzS) * two;
            double eyy = fact * (Wyx * tauxyS + Wyz * tauyzS) * two;
            double ezz = fact * (Wzx * tauxzS + Wzy * tauyzS) * two;
            double exy = fact * (Wxy * tauyyS + Wxz * tauyzS + Wyx * tauxxS + Wyz * tauxzS);
            double exz = fact * (Wxy * tauyzS + Wxz * tauzzS + Wzx * tauxxS + Wzy * tauxyS);
            double eyz = fact * (Wyx * tauxzS + Wyz * tauzzS + Wzx * tauxyS + Wzy * tauyyS);
            tauxx = mut * tauxxS - exx; tauyy = mut * tauyyS - eyy; tauzz = mut * tauzzS - ezz;
            tauxy = mut * tauxyS - exy; tauxz = mut * tauxzS - exz; tauyz = mut * tauyzS - eyz;
        } else {
            tauxx = mut * tauxxS; tauyy = mut * tauyyS; tauzz = mut * tauzzS;
            tauxy = mut * tauxyS; tauxz = mut * tauxzS; tauyz = mut * tauyzS;
        }
        double ubar = half * (W(c, IVX) + W(cp, IVX));
        double vbar = half * (W(c, IVY) + W(cp, IVY));
        double wbar = half * (W(c, IVZ) + W(cp, IVZ));
        double s1 = s[c], s2 = s[d.N + c], s3 = s[2 * d.N + c];
        double fmx = tauxx * s1 + tauxy * s2 + tauxz * s3;
        double fmy = tauxy * s1 + tauyy * s2 + tauyz * s3;
        double fmz = tauxz * s1 + tauyz * s2 + tauzz * s3;
        double frhoE = (ubar * tauxx + vbar * tauxy + wbar * tauxz) * s1 +
                       (ubar * tauxy + vbar * tauyy + wbar * tauyz) * s2 +
                       (ubar * tauxz + vbar * tauyz + wbar * tauzz) * s3 - q_x * s1 - q_y * s2 - q_z * s3;
        FW(c, IMX) -= fmx; FW(c, IMY) -= fmy; FW(c, IMZ) -= fmz; FW(c, IRHOE) -= frhoE;
        FW(cp, IMX) += fmx; FW(cp, IMY) += fmy; FW(cp, IMZ) += fmz; FW(cp, IRHOE) += frhoE;
        if (b->wallTau) { /* tmpStore / viscSubface%tau,%q (blockette.F90:5812-5838): kept for every face */
            double* t = b->wallTau + (long)dir * 9 * d.N;
            t[c] = tauxx; t[d.N + c] = tauyy; t[2 * d.N + c] = tauzz; t[3 * d.N + c] = tauxy; t[4 * d.N + c] = tauxz;
            t[5 * d.N + c] = tauyz; t[6 * d.N + c] = q_x; t[7 * d.N + c] = q_y; t[8 * d.N + c] = q_z;
        }
    }
}
void orc_viscous_flux(const OrcBlock* b, const AdfbParams* prm, double rFil) {
    Dims d = dims_of(b);
    viscous_dir(b, prm, d, b->sk, b->porK, d.sK, d.sI, d.sJ, 2, 2, 1, rFil, 2);
    viscous_dir(b, prm, d, b->sj, b->porJ, d.sJ, d.sI, d.sK, 2, 1, 2, rFil, 1);
    viscous_dir(b, prm, d, b->si, b->porI, d.sI, d.sJ, d.sK, 1, 2, 2, rFil, 0);
}

/* ------------------------------------------------------------------------ */
/* saSource: src/NKSolver/blockette.F90:976-1168 (block twin src/turbulence/sa.F90:89-344).
   Non-rotating frame (omega = 0). */
void orc_sa_source(const OrcBlock* b, const AdfbParams* prm) {
    Dims d = dims_of(b);
    const double xminn = 1.e-10, f23 = two * third;
    double cv13 = prm->rsaCv1 * prm->rsaCv1 * prm->rsaCv1;
    double kar2Inv = one / (prm->rsaK * prm->rsaK);
    double cw36 = pow(prm->rsaCw3, 6.0);
    double term1Fact = prm->approxSA ? zero : one;
    const double *si = b->si, *sj = b->sj, *sk = b->sk;
    for (int k = 2; k <= d.kl; k++) for (int j = 2; j <= d.jl; j++) for (int i = 2; i <= d.il; i++) {
        long c = IDX(i, j, k);
        double gradv[3][3]; /* [velocity comp][direction] */
        for (int v = 0; v < 3; v++) for (int m = 0; m < 3; m++) {
            long o = (long)m * d.N;
            gradv[v][m] = W(c + 1, IVX + v) * si[o + c] - W(c - 1, IVX + v) * si[o + c - 1] +
                          W(c + d.sJ, IVX + v) * sj[o + c] - W(c - d.sJ, IVX + v) * sj[o + c - d.sJ] +
                          W(c + d.sK, IVX + v) * sk[o + c] - W(c - d.sK, IVX + v) * sk[o + c - d.sK];
        }
        double uux = gradv[0][0], uuy = gradv[0][1], uuz = gradv[0][2];
        double vvx = gradv[1][0], vvy = gradv[1][1], vvz = gradv[1][2];
        double wwx = gradv[2][0], wwy = gradv[2][1], wwz = gradv[2][2];
        double fact = fourth / b->vol[c];
        double sxx = two * fact * uux, syy = two * fact * vvy, szz = two * fact * wwz;
        double sxy = fact * (uuy + vvx), sxz = fact * (uuz + wwx), syz = fact * (vvz + wwy);
        double div2 = f23 * ((sxx + syy + szz) * (sxx + syy + szz));
        double strainMag2 = two * (sxy * sxy + sxz * sxz + syz * syz) + sxx * sxx + syy * syy + szz * szz;
        double vortx = two * fact * (wwy - vvz) - two * zero;
        double vorty = two * fact * (uuz - wwx) - two * zero;
        double vortz = two * fact * (vvx - uuy) - two * zero;
        double sqrtProd;
        if (prm->turbProd == ADFB_PROD_STRAIN) sqrtProd = sqrt(dmax(two * strainMag2 - div2, eps_));
        else sqrtProd = sqrt(vortx * vortx + vorty * vorty + vortz * vortz);
        double nu = b->rlv[c] / W(c, IRHO);
        double dist2Inv = one / (b->d2Wall[c] * b->d2Wall[c]);
        double chi = W(c, ITU1) / nu, chi2 = chi * chi, chi3 = chi * chi2;
        double fv1 = chi3 / (chi3 + cv13);
        double fv2 = one - chi / (one + chi * fv1);
        double ft2 = zero;
        if (prm->useft2SA) ft2 = prm->rsaCt3 * exp(-prm->rsaCt4 * chi2);
        double sst = sqrtProd + W(c, ITU1) * fv2 * kar2Inv * dist2Inv;
        if (prm->useRotationSA) sst = sst + prm->rsaCrot * dmin(zero, sqrt(two * strainMag2));
        sst = dmax(sst, xminn);
        double rr = W(c, ITU1) * kar2Inv * dist2Inv / sst;
        rr = dmin(rr, 10.0);
        double rr2 = rr * rr, rr6 = rr2 * rr2 * rr2;
        double gg = rr + prm->rsaCw2 * (rr6 - rr);
        double gg2 = gg * gg, gg6 = gg2 * gg2 * gg2;
        double termFw = pow((one + cw36) / (gg6 + cw36), sixth);
        double fwSa = gg * termFw;
        double term1 = prm->rsaCb1 * (one - ft2) * sqrtProd * term1Fact;
        double term2 = dist2Inv * (kar2Inv * prm->rsaCb1 * ((one - ft2) * fv2 + ft2) - prm->rsaCw1 * fwSa);
        DW(c, ITU1) = DW(c, ITU1) + (term1 + term2 * W(c, ITU1)) * W(c, ITU1);
    }
}

/* saAdvection: src/NKSolver/blockette.F90:1392-1870 (block twin turbAdvection,
   src/turbulence/turbUtils.F90:828-1553); sweeps k, j, i; no grid velocity. */
static void sa_advection_dir(const OrcBlock* b, Dims d, const double* s, long sd, int secondOrd) {
    for (int k = 2; k <= d.kl; k++) for (int j = 2; j <= d.jl; j++) for (int i = 2; i <= d.il; i++) {
        long c = IDX(i, j, k);
        double voli = half / b->vol[c];
        double qs = (zero + zero) * voli;
        double xa = (s[c] + s[c - sd]) * voli;
        double ya = (s[d.N + c] + s[d.N + c - sd]) * voli;
        double za = (s[2 * d.N + c] + s[2 * d.N + c - sd]) * voli;
        double uu = xa * W(c, IVX) + ya * W(c, IVY) + za * W(c, IVZ) - qs;
        double dwtx, dwt, dwtm1, dwtp1;
        if (uu > zero) {
            if (secondOrd) {
                dwtm1 = W(c - sd, ITU1) - W(c - 2 * sd, ITU1);
                dwt = W(c, ITU1) - W(c - sd, ITU1);
                dwtp1 = W(c + sd, ITU1) - W(c, ITU1);
                dwtx = dwt;
                if (dwt * dwtp1 > zero) { if (fabs(dwt) < fabs(dwtp1)) dwtx = dwtx + half * dwt; else dwtx = dwtx + half * dwtp1; }
                if (dwt * dwtm1 > zero) { if (fabs(dwt) < fabs(dwtm1)) dwtx = dwtx - half * dwt; else dwtx = dwtx - half * dwtm1; }
            } else {
                dwtx = W(c, ITU1) - W(c - sd, ITU1);
            }
        } else {
            if (secondOrd) {
                dwtm1 = W(c, ITU1) - W(c - sd, ITU1);
                dwt = W(c + sd, ITU1) - W(c, ITU1);
                dwtp1 = W(c + 2 * sd, ITU1) - W(c + sd, ITU1);
                dwtx = dwt;
                if (dwt * dwtp1 > zero) { if (fabs(dwt) < fabs(dwtp1)) dwtx = dwtx - half * dwt; else dwtx = dwtx - half * dwtp1; }
                if (dwt * dwtm1 > zero) { if (fabs(dwt) < fabs(dwtm1)) dwtx = dwtx + half * dwt; else dwtx = dwtx + half * dwtm1; }
            } else {
                dwtx = W(c + sd, ITU1) - W(c, ITU1);
            }
        }
        DW(c, ITU1) = DW(c, ITU1) - uu * dwtx;
    }
}
void orc_sa_advection(const OrcBlock* b, const AdfbParams* prm) {
    Dims d = dims_of(b);
    sa_advection_dir(b, d, b->sk, d.sK, prm->secondOrdTurb);
    sa_advection_dir(b, d, b->sj, d.sJ, prm->secondOrdTurb);
    sa_advection_dir(b, d, b->si, d.sI, prm->secondOrdTurb);
}

/* saViscous: src/NKSolver/blockette.F90:1170-1390 (block twin src/turbulence/sa.F90:346-676); k, j, i */
static void sa_viscous_dir(const OrcBlock* b, const AdfbParams* prm, Dims d, const double* s, long sd) {
    double cb3Inv = one / prm->rsaCb3, cb2 = prm->rsaCb2;
    const double* vol = b->vol;
    for (int k = 2; k <= d.kl; k++) for (int j = 2; j <= d.jl; j++) for (int i = 2; i <= d.il; i++) {
        long c = IDX(i, j, k), cm = c - sd, cp = c + sd;
        double voli = one / vol[c];
        double volmi = two / (vol[c] + vol[cm]);
        double volpi = two / (vol[c] + vol[cp]);
        double xm = s[cm] * volmi, ym = s[d.N + cm] * volmi, zm = s[2 * d.N + cm] * volmi;
        double xp = s[c] * volpi, yp = s[d.N + c] * volpi, zp = s[2 * d.N + c] * volpi;
        double xa = half * (s[c] + s[cm]) * voli;
        double ya = half * (s[d.N + c] + s[d.N + cm]) * voli;
        double za = half * (s[2 * d.N + c] + s[2 * d.N + cm]) * voli;
        double ttm = xm * xa + ym * ya + zm * za;
        double ttp = xp * xa + yp * ya + zp * za;
        double cnud = -cb2 * W(c, ITU1) * cb3Inv;
        double cam = ttm * cnud, cap = ttp * cnud;
        double nutm = half * (W(cm, ITU1) + W(c, ITU1));
        double nutp = half * (W(cp, ITU1) + W(c, ITU1));
        double nu = b->rlv[c] / W(c, IRHO);
        double num = half * (b->rlv[cm] / W(cm, IRHO) + nu);
        double nup = half * (b->rlv[cp] / W(cp, IRHO) + nu);
        double cdm = (num + (one + cb2) * nutm) * ttm * cb3Inv;
        double cdp = (nup + (one + cb2) * nutp) * ttp * cb3Inv;
        double c1m = dmax(cdm + cam, zero), c1p = dmax(cdp + cap, zero);
        double c10 = c1m + c1p;
        DW(c, ITU1) = DW(c, ITU1) + c1m * W(cm, ITU1) - c10 * W(c, ITU1) + c1p * W(cp, ITU1);
    }
}
void orc_sa_viscous(const OrcBlock* b, const AdfbParams* prm) {
    Dims d = dims_of(b);
    sa_viscous_dir(b, prm, d, b->sk, d.sK);
    sa_viscous_dir(b, prm, d, b->sj, d.sJ);
    sa_viscous_dir(b, prm, d, b->si, d.sI);
}
/* saResScale: src/NKSolver/blockette.F90:1872-1897 */
void orc_sa_res_scale(const OrcBlock* b) {
    Dims d = dims_of(b);
    for (int k = 2; k <= d.kl; k++) for (int j = 2; j <= d.jl; j++) for (int i = 2; i <= d.il; i++) {
        long c = IDX(i, j, k);
        double rblank = dmax((double)b->iblank[c], zero);
        DW(c, ITU1) = -b->volRef[c] * DW(c, ITU1) * rblank;
    }
}
/* sumDwandFw: src/NKSolver/blockette.F90:6839-6864 */
void orc_sum_dw_fw(const OrcBlock* b) {
    Dims d = dims_of(b);
    for (int l = 0; l < 5; l++)
        for (int k = 2; k <= d.kl; k++) for (int j = 2; j <= d.jl; j++) for (int i = 2; i <= d.il; i++) {
            long c = IDX(i, j, k);
            double rblank = dmax((double)b->iblank[c], zero);
            DW(c, l) = (DW(c, l) + FW(c, l)) * rblank;
        }
}

/* ------------------------------------------------------------------------ */
/* blocketteResCore operator order: src/NKSolver/blockette.F90:607-667.
   `rFil` is 1 when called like blocketteRes (:270); the RK smoother path
   (src/solver/residuals.F90:61-65) passes cdisRK(stage+1) and relies on fw
   persisting between stages -- in that case the caller must NOT clear fw,
   signalled by rFil != 1 (fw is cleared here only for rFil == 1, :608). */
void orc_residual_core(const OrcBlock* b, const AdfbParams* prm, unsigned flags, double rFil) {
    Dims d = dims_of(b);
    int flowRes = (flags & ADFB_RES_FLOW) != 0, turbRes = (flags & ADFB_RES_TURB) != 0;
    int viscous = prm->equations != ADFB_EULER;
    int l0 = flowRes ? 0 : 5, l1 = turbRes && prm->equations == ADFB_RANS ? b->nw : 5;
    if (flags & ADFB_RES_FLOW) { if (rFil == one) memset(b->fw, 0, sizeof(double) * 5 * d.N); }
    /* metrics are resident (si,sj,sk); initRes :962-974 */
    for (int l = l0; l < l1; l++) memset(b->dw + (long)l * d.N, 0, sizeof(double) * d.N);
    if (prm->equations == ADFB_RANS && turbRes) {
        orc_sa_source(b, prm);
        orc_sa_advection(b, prm);
        orc_sa_viscous(b, prm);
        orc_sa_res_scale(b);
    }
    orc_time_step(b, prm, 1);
    if (flowRes) {
        orc_central_flux(b, prm);
        if (flags & ADFB_RES_DISS_APPROX) { /* blockette.F90:636-644 */
            if (prm->spaceDiscr == ADFB_DISS_SCALAR) orc_diss_scalar_approx(b, prm);
            else if (prm->spaceDiscr == ADFB_DISS_MATRIX) orc_diss_matrix_approx(b, prm, rFil);
            else {
                AdfbParams p1 = *prm;
                p1.limiter = ADFB_LIM_FIRSTORDER; /* inviscidUpwindFlux(.False.): first order */
                orc_upwind_flux(b, &p1, rFil);
            }
        } else {
            if (prm->spaceDiscr == ADFB_DISS_SCALAR) orc_diss_scalar(b, prm, rFil);
            else if (prm->spaceDiscr == ADFB_DISS_MATRIX) orc_diss_matrix(b, prm, rFil);
            else if (prm->spaceDiscr == ADFB_UPWIND) orc_upwind_flux(b, prm, rFil);
        }
        if (viscous && fabs(rFil) > thresholdReal) {
            orc_speed_of_sound(b, prm);
            if (flags & ADFB_RES_VISC_APPROX) orc_viscous_flux_approx(b, prm, rFil);
            else {
                orc_nodal_gradients(b);
                orc_viscous_flux(b, prm, rFil);
            }
        }
        orc_sum_dw_fw(b);
    }
}

/* sumResiduals / sumAllResiduals: src/utils/utils.F90:6364-6459 */
void orc_norms(const OrcBlock* b, const AdfbParams* prm, double out[2]) {
    Dims d = dims_of(b);
    out[0] = out[1] = 0.0;
    for (int k = 2; k <= d.kl; k++) for (int j = 2; j <= d.jl; j++) for (int i = 2; i <= d.il; i++) {
        long c = IDX(i, j, k);
        double q = DW(c, 0) / b->vol[c];
        out[0] += q * q;
        double state_sum = 0.0, ovv = one / b->vol[c];
        for (int l = 0; l < 5; l++) state_sum += (DW(c, l) * ovv) * (DW(c, l) * ovv);
        for (int l = 5; l < b->nw; l++)
            state_sum += (DW(c, l) * ovv * prm->turbResScale) * (DW(c, l) * ovv * prm->turbResScale);
        (void)four; (void)five;
        out[1] += state_sum;
    }
}
