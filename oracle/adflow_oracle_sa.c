/*
 * adflow_oracle_sa.c -- CPU restatement of the Spalart-Allmaras DD-ADI solve
 * (sa_block / saSolve, src/turbulence/sa.F90:16-86,717-1267 with the block-path
 * saSource :89-344, turbAdvection src/turbulence/turbUtils.F90:828-1553, saViscous
 * sa.F90:346-676, saResScale :678-714).  TEST INFRASTRUCTURE ONLY (pinned, see adflow_oracle.h).
 *
 * Work arrays: scratch slot 0 = dvt (idvt), slot 1 = qq (central jacobian), slot 2 = bmt
 * (turbulence BC matrix of the face a halo cell belongs to, bcTurbTreatment).
 */
#include "orc_internal.h"

#define DVT(c) b->scratch[(c)]
#define QQ(c) b->scratch[d.N + (c)]
#define BMT(c) b->scratch[2 * d.N + (c)]

/* bcTurbTreatment (src/turbulence/turbBCRoutines.F90:662-797) for the scalar SA variable:
   bmt of each boundary face cell, stored at the first-halo cell of that face */
static void sa_bmt(const OrcBlock* b, const AdfbParams* prm, int nSub, const AdfbSubface* sf) {
    Dims d = dims_of(b);
    memset(&BMT(0), 0, sizeof(double) * d.N);
    for (int n = 0; n < nSub; n++) {
        const AdfbSubface* s = &sf[n];
        long off1, sa, sb;
        switch (s->faceId) {
            case ADFB_IMIN: off1 = 1; sa = d.sJ; sb = d.sK; break;
            case ADFB_IMAX: off1 = d.ie; sa = d.sJ; sb = d.sK; break;
            case ADFB_JMIN: off1 = d.sJ; sa = 1; sb = d.sK; break;
            case ADFB_JMAX: off1 = d.je * d.sJ; sa = 1; sb = d.sK; break;
            case ADFB_KMIN: off1 = d.sK; sa = 1; sb = d.sJ; break;
            default: off1 = d.ke * d.sK; sa = 1; sb = d.sJ; break;
        }
        long na = s->icEnd - s->icBeg + 1, nb = s->jcEnd - s->jcBeg + 1;
        for (int jb_ = s->jcBeg; jb_ <= s->jcEnd; jb_++) for (int ia = s->icBeg; ia <= s->icEnd; ia++) {
            long o = (ia - s->icBeg) + na * (jb_ - s->jcBeg);
            double bmt = zero;
            if (s->bcType == ADFB_BC_NSWALL_ADIABATIC || s->bcType == ADFB_BC_NSWALL_ISOTHERMAL ||
                s->bcType == ADFB_BC_SUBSONIC_INFLOW || s->bcType == ADFB_BC_SUPERSONIC_INFLOW) bmt = one; /* bcTurbWall, bcTurbInflow */
            else if (s->bcType == ADFB_BC_FARFIELD) {
                double dot = s->norm[o] * prm->wInf[IVX] + s->norm[o + na * nb] * prm->wInf[IVY] + s->norm[o + 2 * na * nb] * prm->wInf[IVZ] -
                             (s->rface ? s->rface[o] : zero);
                if (dot > zero) bmt = -one;
            } else bmt = -one;
            BMT(off1 + ia * sa + jb_ * sb) = bmt;
        }
    }
}

/* block-path saSource incl. the implicit diagonal qq: src/turbulence/sa.F90:89-344 */
static void sa_source_block(const OrcBlock* b, const AdfbParams* prm) {
    Dims d = dims_of(b);
    const double xminn = 1.e-10, f23 = two * third;
    double cv13 = prm->rsaCv1 * prm->rsaCv1 * prm->rsaCv1;
    double kar2Inv = one / (prm->rsaK * prm->rsaK);
    double cw36 = pow(prm->rsaCw3, 6.0);
    const double *si = b->si, *sj = b->sj, *sk = b->sk;
    for (int k = 2; k <= d.kl; k++) for (int j = 2; j <= d.jl; j++) for (int i = 2; i <= d.il; i++) {
        long c = IDX(i, j, k);
        double gv[3][3];
        for (int v = 0; v < 3; v++) for (int m = 0; m < 3; m++) {
            long o = (long)m * d.N;
            gv[v][m] = W(c + 1, IVX + v) * si[o + c] - W(c - 1, IVX + v) * si[o + c - 1] +
                       W(c + d.sJ, IVX + v) * sj[o + c] - W(c - d.sJ, IVX + v) * sj[o + c - d.sJ] +
                       W(c + d.sK, IVX + v) * sk[o + c] - W(c - d.sK, IVX + v) * sk[o + c - d.sK];
        }
        double fact = fourth / b->vol[c];
        double ss, strainMag2 = zero;
        if (prm->turbProd == ADFB_PROD_STRAIN) {
            double sxx = two * fact * gv[0][0], syy = two * fact * gv[1][1], szz = two * fact * gv[2][2];
            double sxy = fact * (gv[0][1] + gv[1][0]), sxz = fact * (gv[0][2] + gv[2][0]), syz = fact * (gv[1][2] + gv[2][1]);
            double div2 = f23 * ((sxx + syy + szz) * (sxx + syy + szz));
            strainMag2 = two * (sxy * sxy + sxz * sxz + syz * syz) + sxx * sxx + syy * syy + szz * szz;
            ss = sqrt(two * strainMag2 - div2);
        } else {
            double vortx = two * fact * (gv[2][1] - gv[1][2]), vorty = two * fact * (gv[0][2] - gv[2][0]), vortz = two * fact * (gv[1][0] - gv[0][1]);
            ss = sqrt(vortx * vortx + vorty * vorty + vortz * vortz);
        }
        double nut = W(c, ITU1);
        double nu = b->rlv[c] / W(c, IRHO);
        double dist2Inv = one / (b->d2Wall[c] * b->d2Wall[c]);
        double chi = nut / nu, chi2 = chi * chi, chi3 = chi * chi2;
        double fv1 = chi3 / (chi3 + cv13);
        double fv2 = one - chi / (one + chi * fv1);
        double ft2 = prm->useft2SA ? prm->rsaCt3 * exp(-prm->rsaCt4 * chi2) : zero;
        double sst = ss + nut * fv2 * kar2Inv * dist2Inv;
        if (prm->useRotationSA) sst = sst + prm->rsaCrot * dmin(zero, sqrt(two * strainMag2));
        sst = dmax(sst, xminn);
        double rr = nut * kar2Inv * dist2Inv / sst;
        rr = dmin(rr, 10.0);
        double rr2 = rr * rr, rr6 = rr2 * rr2 * rr2;
        double gg = rr + prm->rsaCw2 * (rr6 - rr);
        double gg2 = gg * gg, gg6 = gg2 * gg2 * gg2;
        double termFw = pow((one + cw36) / (gg6 + cw36), sixth);
        double fwSa = gg * termFw;
        double term1 = prm->approxSA ? zero : prm->rsaCb1 * (one - ft2) * ss;
        double term2 = dist2Inv * (kar2Inv * prm->rsaCb1 * ((one - ft2) * fv2 + ft2) - prm->rsaCw1 * fwSa);
        DVT(c) = (term1 + term2 * nut) * nut;
        double dfv1 = three * chi2 * cv13 / ((chi3 + cv13) * (chi3 + cv13));
        double dfv2 = (chi2 * dfv1 - one) / (nu * ((one + chi * fv1) * (one + chi * fv1)));
        double dft2 = -two * prm->rsaCt4 * chi * ft2 / nu;
        double drr = (one - rr * (fv2 + nut * dfv2)) * kar2Inv * dist2Inv / sst;
        double rr5 = rr2 * rr2 * rr;
        double dgg = (one - prm->rsaCw2 + 6.0 * prm->rsaCw2 * rr5) * drr;
        double dfw = (cw36 / (gg6 + cw36)) * termFw * dgg;
        double q = -two * term2 * nut - dist2Inv * nut * nut * (prm->rsaCb1 * kar2Inv * (dfv2 - ft2 * dfv2 - fv2 * dft2 + dft2) - prm->rsaCw1 * dfw);
        QQ(c) = dmax(q, zero);
    }
}

/* turbAdvection for one direction incl. qq (src/turbulence/turbUtils.F90:905-1096 k, j, i) */
static void sa_advection_block_dir(const OrcBlock* b, const AdfbParams* prm, Dims d, const double* s, long sd, int axis) {
    int secondOrd = prm->secondOrdTurb;
    for (int k = 2; k <= d.kl; k++) for (int j = 2; j <= d.jl; j++) for (int i = 2; i <= d.il; i++) {
        long c = IDX(i, j, k);
        int idx = axis == 0 ? i : (axis == 1 ? j : k);
        int last = axis == 0 ? d.il : (axis == 1 ? d.jl : d.kl);
        double voli = half / b->vol[c];
        double xa = (s[c] + s[c - sd]) * voli, ya = (s[d.N + c] + s[d.N + c - sd]) * voli, za = (s[2 * d.N + c] + s[2 * d.N + c - sd]) * voli;
        double uu = xa * W(c, IVX) + ya * W(c, IVY) + za * W(c, IVZ) - zero;
        double dwtx;
        if (uu > zero) {
            if (secondOrd) {
                double dwtm1 = W(c - sd, ITU1) - W(c - 2 * sd, ITU1), dwt = W(c, ITU1) - W(c - sd, ITU1), dwtp1 = W(c + sd, ITU1) - W(c, ITU1);
                dwtx = dwt;
                if (dwt * dwtp1 > zero) dwtx = dwtx + half * (fabs(dwt) < fabs(dwtp1) ? dwt : dwtp1);
                if (dwt * dwtm1 > zero) dwtx = dwtx - half * (fabs(dwt) < fabs(dwtm1) ? dwt : dwtm1);
            } else dwtx = W(c, ITU1) - W(c - sd, ITU1);
            DVT(c) = DVT(c) - uu * dwtx;
            QQ(c) = QQ(c) + uu;
            if (idx == 2) QQ(c) = QQ(c) + uu * dmax(BMT(c - sd), zero);
        } else {
            if (secondOrd) {
                double dwtm1 = W(c, ITU1) - W(c - sd, ITU1), dwt = W(c + sd, ITU1) - W(c, ITU1), dwtp1 = W(c + 2 * sd, ITU1) - W(c + sd, ITU1);
                dwtx = dwt;
                if (dwt * dwtp1 > zero) dwtx = dwtx - half * (fabs(dwt) < fabs(dwtp1) ? dwt : dwtp1);
                if (dwt * dwtm1 > zero) dwtx = dwtx + half * (fabs(dwt) < fabs(dwtm1) ? dwt : dwtm1);
            } else dwtx = W(c + sd, ITU1) - W(c, ITU1);
            DVT(c) = DVT(c) - uu * dwtx;
            QQ(c) = QQ(c) - uu;
            if (idx == last) QQ(c) = QQ(c) - uu * dmax(BMT(c + sd), zero);
        }
    }
}

/* SA diffusion coefficients of cell c along sd (shared by saViscous and saSolve) */
static void sa_diff_coef(const OrcBlock* b, const AdfbParams* prm, Dims d, const double* s, long sd, long c,
                         double* c1m, double* c1p, double* xa, double* ya, double* za) {
    double cb3Inv = one / prm->rsaCb3, cb2 = prm->rsaCb2;
    const double* vol = b->vol;
    long cm = c - sd, cp = c + sd;
    double voli = one / vol[c];
    double volmi = two / (vol[c] + vol[cm]);
    double volpi = two / (vol[c] + vol[cp]);
    double xm = s[cm] * volmi, ym = s[d.N + cm] * volmi, zm = s[2 * d.N + cm] * volmi;
    double xp = s[c] * volpi, yp = s[d.N + c] * volpi, zp = s[2 * d.N + c] * volpi;
    *xa = half * (s[c] + s[cm]) * voli; *ya = half * (s[d.N + c] + s[d.N + cm]) * voli; *za = half * (s[2 * d.N + c] + s[2 * d.N + cm]) * voli;
    double ttm = xm * *xa + ym * *ya + zm * *za;
    double ttp = xp * *xa + yp * *ya + zp * *za;
    double cnud = -cb2 * W(c, ITU1) * cb3Inv;
    double cam = ttm * cnud, cap = ttp * cnud;
    double nutm = half * (W(cm, ITU1) + W(c, ITU1)), nutp = half * (W(cp, ITU1) + W(c, ITU1));
    double nu = b->rlv[c] / W(c, IRHO);
    double num = half * (b->rlv[cm] / W(cm, IRHO) + nu), nup = half * (b->rlv[cp] / W(cp, IRHO) + nu);
    double cdm = (num + (one + cb2) * nutm) * ttm * cb3Inv;
    double cdp = (nup + (one + cb2) * nutp) * ttp * cb3Inv;
    *c1m = dmax(cdm + cam, zero);
    *c1p = dmax(cdp + cap, zero);
}

/* saViscous incl. qq: src/turbulence/sa.F90:346-676 */
static void sa_viscous_block_dir(const OrcBlock* b, const AdfbParams* prm, Dims d, const double* s, long sd, int axis) {
    for (int k = 2; k <= d.kl; k++) for (int j = 2; j <= d.jl; j++) for (int i = 2; i <= d.il; i++) {
        long c = IDX(i, j, k);
        int idx = axis == 0 ? i : (axis == 1 ? j : k);
        int last = axis == 0 ? d.il : (axis == 1 ? d.jl : d.kl);
        double c1m, c1p, xa, ya, za;
        sa_diff_coef(b, prm, d, s, sd, c, &c1m, &c1p, &xa, &ya, &za);
        double c10 = c1m + c1p;
        DVT(c) = DVT(c) + c1m * W(c - sd, ITU1) - c10 * W(c, ITU1) + c1p * W(c + sd, ITU1);
        double b1 = -c1m, c1 = c10, d1 = -c1p;
        if (idx == 2) QQ(c) = QQ(c) + c1 - b1 * dmax(BMT(c - sd), zero);
        else if (idx == last) QQ(c) = QQ(c) + c1 - d1 * dmax(BMT(c + sd), zero);
        else QQ(c) = QQ(c) + c1;
    }
}

/* one dd-ADI sweep of saSolve (src/turbulence/sa.F90:877-1002 j, :1006-1125 i, :1129-1250 k) */
static void sa_solve_dir(const OrcBlock* b, const AdfbParams* prm, Dims d, const double* s, long sd, int nl, long s1, int n1,
                         long s2, int n2, int multiplyByQQ) {
    int l = nl + 1;
    double* bb = (double*)calloc(l + 3, sizeof(double)); double* cc = (double*)calloc(l + 3, sizeof(double));
    double* dd = (double*)calloc(l + 3, sizeof(double)); double* ff = (double*)calloc(l + 3, sizeof(double));
    for (int q2 = 2; q2 <= n2 + 1; q2++) for (int q1 = 2; q1 <= n1 + 1; q1++) {
        long base = q1 * s1 + q2 * s2;
        for (int m = 2; m <= l; m++) {
            long c = base + m * sd;
            double c1m, c1p, xa, ya, za;
            sa_diff_coef(b, prm, d, s, sd, c, &c1m, &c1p, &xa, &ya, &za);
            bb[m] = -c1m; dd[m] = -c1p;
            double uu = xa * W(c, IVX) + ya * W(c, IVY) + za * W(c, IVZ) - zero;
            double um = zero, up = zero;
            if (uu < zero) um = uu;
            if (uu > zero) up = uu;
            bb[m] = bb[m] - up;
            dd[m] = dd[m] + um;
            double rblank = dmax((double)b->iblank[c], zero);
            cc[m] = QQ(c);
            ff[m] = DVT(c) * rblank;
            bb[m] = bb[m] * rblank;
            dd[m] = dd[m] * rblank;
        }
        for (int m = nl; m >= 2; m--) {
            double f = dd[m] / cc[m + 1];
            cc[m] = cc[m] - f * bb[m + 1];
            ff[m] = ff[m] - f * ff[m + 1];
        }
        ff[2] = ff[2] / cc[2];
        for (int m = 3; m <= l; m++) { ff[m] = ff[m] - bb[m] * ff[m - 1]; ff[m] = ff[m] / cc[m]; }
        for (int m = 2; m <= l; m++) { long c = base + m * sd; DVT(c) = multiplyByQQ ? ff[m] * QQ(c) : ff[m]; }
    }
    free(bb); free(cc); free(dd); free(ff);
}

/* sa_block(resOnly = .false.): src/turbulence/sa.F90:16-86 */
void orc_sa_block(const OrcBlock* b, const AdfbParams* prm, int nSub, const AdfbSubface* sf) {
    Dims d = dims_of(b);
    sa_bmt(b, prm, nSub, sf);
    sa_source_block(b, prm);
    sa_advection_block_dir(b, prm, d, b->sk, d.sK, 2);
    sa_advection_block_dir(b, prm, d, b->sj, d.sJ, 1);
    sa_advection_block_dir(b, prm, d, b->si, d.sI, 0);
    sa_viscous_block_dir(b, prm, d, b->sk, d.sK, 2);
    sa_viscous_block_dir(b, prm, d, b->sj, d.sJ, 1);
    sa_viscous_block_dir(b, prm, d, b->si, d.sI, 0);
    for (int k = 2; k <= d.kl; k++) for (int j = 2; j <= d.jl; j++) for (int i = 2; i <= d.il; i++) {
        long c = IDX(i, j, k);
        DW(c, ITU1) = -b->volRef[c] * DVT(c) * dmax((double)b->iblank[c], zero); /* saResScale */
    }
    /* saSolve: implicit relaxation factor on qq (turbRelaxImplicit, :849-851) */
    double factor = one + (one - prm->alfaTurb) / prm->alfaTurb;
    for (int k = 2; k <= d.kl; k++) for (int j = 2; j <= d.jl; j++) for (int i = 2; i <= d.il; i++) { long c = IDX(i, j, k); QQ(c) = factor * QQ(c); }
    sa_solve_dir(b, prm, d, b->sj, d.sJ, d.ny, d.sI, d.nx, d.sK, d.nz, 1);
    sa_solve_dir(b, prm, d, b->si, d.sI, d.nx, d.sJ, d.ny, d.sK, d.nz, 1);
    sa_solve_dir(b, prm, d, b->sk, d.sK, d.nz, d.sI, d.nx, d.sJ, d.ny, 0);
    for (int k = 2; k <= d.kl; k++) for (int j = 2; j <= d.jl; j++) for (int i = 2; i <= d.il; i++) {
        long c = IDX(i, j, k);
        W(c, ITU1) = W(c, ITU1) + one * DVT(c);
        W(c, ITU1) = dmax(W(c, ITU1), zero);
    }
    orc_eddy_viscosity(b, prm, 0);
    orc_apply_turb_bc(b, prm, nSub, sf, 1);
}
