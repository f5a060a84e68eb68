/*
 * adflow_oracle_fluxes.c -- CPU restatement of the alternative inviscid dissipation schemes:
 * matrix dissipation (inviscidDissFluxMatrix, src/NKSolver/blockette.F90:2457-3027; block twin
 * src/solver/fluxes.F90:403-1047) and the upwind/Roe scheme (inviscidUpwindFlux,
 * blockette.F90:3341-4365; src/solver/fluxes.F90:1438-2532).  TEST INFRASTRUCTURE ONLY
 * (pinned bit-exact against oracle/_ref, see adflow_oracle.h).
 */
#include "orc_internal.h"

static void diss_matrix_dir(const OrcBlock* b, const AdfbParams* prm, Dims d, const double* s, const double* dss, const int8_t* por,
                            long sd, int i0, int j0, int k0, double fis2, double fis4) {
    const double dpMax = 0.25, epsAcoustic = 0.25, epsShear = 0.025;
    double gam = prm->gammaInf;
    const double* p = b->p;
    for (int k = k0; k <= d.kl; k++) for (int j = j0; j <= d.jl; j++) for (int i = i0; i <= d.il; i++) {
        long c = IDX(i, j, k), cp = c + sd, cpp = c + 2 * sd, cm = c - sd;
        double ppor = zero;
        if (por[c] == ADFB_NORMALFLUX) ppor = one;
        double dis2 = ppor * fis2 * dmin(dpMax, dmax(dss[c], dss[cp]));
        double dis4 = fdim_(ppor * fis4, dis2);
        double ddw1 = W(cp, IRHO) - W(c, IRHO);
        double dr = dis2 * ddw1 - dis4 * (W(cpp, IRHO) - W(cm, IRHO) - three * ddw1);
        double dru, drv, drw;
        double* dq[3] = {&dru, &drv, &drw};
        for (int l = 0; l < 3; l++) {
            double ddw = W(cp, IRHO) * W(cp, IVX + l) - W(c, IRHO) * W(c, IVX + l);
            *dq[l] = dis2 * ddw - dis4 * (W(cpp, IRHO) * W(cpp, IVX + l) - W(cm, IRHO) * W(cm, IVX + l) - three * ddw);
        }
        double ddw5 = W(cp, IRHOE) - W(c, IRHOE);
        double dre = dis2 * ddw5 - dis4 * (W(cpp, IRHOE) - W(cm, IRHOE) - three * ddw5);
        double drk = zero, kAvg = zero;
        double gammaAvg = half * (gam + gam);
        double gm1 = gammaAvg - one, ovgm1 = one / gm1, gm53 = gammaAvg - five * third;
        double uAvg = half * (W(cp, IVX) + W(c, IVX)), vAvg = half * (W(cp, IVY) + W(c, IVY)), wAvg = half * (W(cp, IVZ) + W(c, IVZ));
        double a2Avg = half * (gam * p[cp] / W(cp, IRHO) + gam * p[c] / W(c, IRHO));
        double s1 = s[c], s2 = s[d.N + c], s3 = s[2 * d.N + c];
        double area = sqrt(s1 * s1 + s2 * s2 + s3 * s3);
        double tmp = one / dmax(1.e-25, area);
        double sx = s1 * tmp, sy = s2 * tmp, sz = s3 * tmp;
        double alphaAvg = half * (uAvg * uAvg + vAvg * vAvg + wAvg * wAvg);
        double hAvg = alphaAvg + ovgm1 * (a2Avg - gm53 * kAvg);
        double aAvg = sqrt(a2Avg);
        double unAvg = uAvg * sx + vAvg * sy + wAvg * sz;
        double ovaAvg = one / aAvg, ova2Avg = one / a2Avg;
        double sface = zero * tmp;
        double lam1 = fabs(unAvg - sface + aAvg), lam2 = fabs(unAvg - sface - aAvg), lam3 = fabs(unAvg - sface);
        double rrad = lam3 + aAvg;
        lam1 = dmax(lam1, epsAcoustic * rrad) * area;
        lam2 = dmax(lam2, epsAcoustic * rrad) * area;
        lam3 = dmax(lam3, epsShear * rrad) * area;
        double abv1 = half * (lam1 + lam2), abv2 = half * (lam1 - lam2), abv3 = abv1 - lam3;
        double abv4 = gm1 * (alphaAvg * dr - uAvg * dru - vAvg * drv - wAvg * drw + dre) - gm53 * drk;
        double abv5 = sx * dru + sy * drv + sz * drw - unAvg * dr;
        double abv6 = abv3 * abv4 * ova2Avg + abv2 * abv5 * ovaAvg;
        double abv7 = abv2 * abv4 * ovaAvg + abv3 * abv5;
        double fs;
        fs = lam3 * dr + abv6; FW(cp, IRHO) += fs; FW(c, IRHO) -= fs;
        fs = lam3 * dru + uAvg * abv6 + sx * abv7; FW(cp, IMX) += fs; FW(c, IMX) -= fs;
        fs = lam3 * drv + vAvg * abv6 + sy * abv7; FW(cp, IMY) += fs; FW(c, IMY) -= fs;
        fs = lam3 * drw + wAvg * abv6 + sz * abv7; FW(cp, IMZ) += fs; FW(c, IMZ) -= fs;
        fs = lam3 * dre + hAvg * abv6 + unAvg * abv7; FW(cp, IRHOE) += fs; FW(c, IRHOE) -= fs;
    }
}

/* inviscidDissFluxMatrix: src/NKSolver/blockette.F90:2457-3027 */
void orc_diss_matrix(const OrcBlock* b, const AdfbParams* prm, double rFil) {
    Dims d = dims_of(b);
    const double omega = 0.5, oneMinOmega = one - omega;
    double plim = 0.001 * prm->pInfCorr;
    const double* p = b->p;
    double fis2 = rFil * prm->vis2, fis4 = rFil * prm->vis4, sfil = one - rFil;
    for (long q = 0; q < 5 * d.N; q++) b->fw[q] = sfil * b->fw[q];
    for (int k = 1; k <= d.ke; k++) for (int j = 1; j <= d.je; j++) for (int i = 1; i <= d.ie; i++) {
        long c = IDX(i, j, k);
        long st[3] = {1, d.sJ, d.sK};
        for (int m = 0; m < 3; m++) {
            long sd = st[m];
            b->dss[m * d.N + c] = fabs((p[c + sd] - two * p[c] + p[c - sd]) /
                                       (omega * (p[c + sd] + two * p[c] + p[c - sd]) +
                                        oneMinOmega * (fabs(p[c + sd] - p[c]) + fabs(p[c] - p[c - sd])) + plim));
        }
    }
    diss_matrix_dir(b, prm, d, b->si, b->dss, b->porI, d.sI, 1, 2, 2, fis2, fis4);
    diss_matrix_dir(b, prm, d, b->sj, b->dss + d.N, b->porJ, d.sJ, 2, 1, 2, fis2, fis4);
    diss_matrix_dir(b, prm, d, b->sk, b->dss + 2 * d.N, b->porK, d.sK, 2, 2, 1, fis2, fis4);
}

/* ------------------------------------------------------------------------ */
/* inviscidUpwindFlux: src/NKSolver/blockette.F90:3341-4365 (Roe flux-difference splitting,
   no preconditioner, MUSCL(kappa) reconstruction of rho,u,v,w,p with no limiter / van Albada /
   minmod; leftRightState :3936-4125, riemannFlux :4129-4363).  No rotational periodicity,
   no moving grid, no k equation. */
static void left_right_state(const AdfbParams* prm, const double du1[5], const double du2[5], const double du3[5],
                             double left[5], double right[5]) {
    const double epsLim = 1.e-10;
    double kappa = prm->kappaCoef;
    double omk = fourth * (one - kappa), opk = fourth * (one + kappa);
    double factMinmod = (three - kappa) / dmax(1.e-10, one - kappa);
    for (int l = 0; l < 5; l++) {
        if (prm->limiter == ADFB_LIM_NONE) {
            left[l] = omk * du1[l] + opk * du2[l];
            right[l] = -omk * du3[l] - opk * du2[l];
        } else {
            double tmp = one / copysign(dmax(fabs(du2[l]), epsLim), du2[l]);
            double rl1 = dmax(zero, du2[l] / copysign(dmax(fabs(du1[l]), epsLim), du1[l]));
            double rl2 = dmax(zero, du1[l] * tmp);
            double rr1 = dmax(zero, du3[l] * tmp);
            double rr2 = dmax(zero, du2[l] / copysign(dmax(fabs(du3[l]), epsLim), du3[l]));
            if (prm->limiter == ADFB_LIM_VANALBADA) {
                rl1 = rl1 * (rl1 + one) / (rl1 * rl1 + one);
                rl2 = rl2 * (rl2 + one) / (rl2 * rl2 + one);
                rr1 = rr1 * (rr1 + one) / (rr1 * rr1 + one);
                rr2 = rr2 * (rr2 + one) / (rr2 * rr2 + one);
            } else {
                rl1 = dmin(one, factMinmod * rl1); rl2 = dmin(one, factMinmod * rl2);
                rr1 = dmin(one, factMinmod * rr1); rr2 = dmin(one, factMinmod * rr2);
            }
            left[l] = omk * rl1 * du1[l] + opk * rl2 * du2[l];
            right[l] = -opk * rr1 * du2[l] - omk * rr2 * du3[l];
        }
    }
}

static void riemann_flux(const AdfbParams* prm, const double left[5], const double right[5], double sx, double sy, double sz,
                         int por, double rFil, double flux[5]) {
    double gammaFace = half * (prm->gammaInf + prm->gammaInf);
    double porFlux = half * rFil;
    if (por == ADFB_NOFLUX || por == ADFB_BOUNDFLUX) porFlux = zero;
    double gm1 = gammaFace - one, gm53 = gammaFace - five * third;
    double z1l = sqrt(left[IRHO]), z1r = sqrt(right[IRHO]);
    double tmp = one / (z1l + z1r);
    double drk = 0.0, kAvg = 0.0;
    double ovgm1 = one / (prm->gammaInf - one);
    /* etot, src/utils/flowUtils.F90:674-760 (left(irhoE) holds the pressure) */
    double Etl = left[IRHO] * (ovgm1 * left[IRHOE] / left[IRHO] + half * (left[IVX] * left[IVX] + left[IVY] * left[IVY] + left[IVZ] * left[IVZ]));
    double Etr = right[IRHO] * (ovgm1 * right[IRHOE] / right[IRHO] + half * (right[IVX] * right[IVX] + right[IVY] * right[IVY] + right[IVZ] * right[IVZ]));
    double dr = right[IRHO] - left[IRHO];
    double dru = right[IRHO] * right[IVX] - left[IRHO] * left[IVX];
    double drv = right[IRHO] * right[IVY] - left[IRHO] * left[IVY];
    double drw = right[IRHO] * right[IVZ] - left[IRHO] * left[IVZ];
    double drE = Etr - Etl;
    double uAvg = tmp * (z1l * left[IVX] + z1r * right[IVX]);
    double vAvg = tmp * (z1l * left[IVY] + z1r * right[IVY]);
    double wAvg = tmp * (z1l * left[IVZ] + z1r * right[IVZ]);
    double hAvg = tmp * ((Etl + left[IRHOE]) / z1l + (Etr + right[IRHOE]) / z1r);
    double area = sqrt(sx * sx + sy * sy + sz * sz);
    tmp = one / dmax(1.e-25, area);
    sx = sx * tmp; sy = sy * tmp; sz = sz * tmp;
    double rFace = zero * tmp;
    double alphaAvg = half * (uAvg * uAvg + vAvg * vAvg + wAvg * wAvg);
    double a2Avg = fabs(gm1 * (hAvg - alphaAvg) - gm53 * kAvg);
    double aAvg = sqrt(a2Avg);
    double unAvg = uAvg * sx + vAvg * sy + wAvg * sz;
    double ovaAvg = one / aAvg, ova2Avg = one / a2Avg;
    if (por == ADFB_BOUNDFLUX) unAvg = rFace;
    double eta = half * (fabs((left[IVX] - right[IVX]) * sx + (left[IVY] - right[IVY]) * sy + (left[IVZ] - right[IVZ]) * sz) +
                         fabs(sqrt(gammaFace * left[IRHOE] / left[IRHO]) - sqrt(gammaFace * right[IRHOE] / right[IRHO])));
    double lam1 = fabs(unAvg - rFace + aAvg), lam2 = fabs(unAvg - rFace - aAvg), lam3 = fabs(unAvg - rFace);
    tmp = two * eta;
    if (lam1 < tmp) lam1 = eta + fourth * lam1 * lam1 / eta;
    if (lam2 < tmp) lam2 = eta + fourth * lam2 * lam2 / eta;
    if (lam3 < tmp) lam3 = eta + fourth * lam3 * lam3 / eta;
    lam1 = lam1 * area; lam2 = lam2 * area; lam3 = lam3 * area;
    double abv1 = half * (lam1 + lam2), abv2 = half * (lam1 - lam2), abv3 = abv1 - lam3;
    double abv4 = gm1 * (alphaAvg * dr - uAvg * dru - vAvg * drv - wAvg * drw + drE) - gm53 * drk;
    double abv5 = sx * dru + sy * drv + sz * drw - unAvg * dr;
    double abv6 = abv3 * abv4 * ova2Avg + abv2 * abv5 * ovaAvg;
    double abv7 = abv2 * abv4 * ovaAvg + abv3 * abv5;
    flux[IRHO] = -porFlux * (lam3 * dr + abv6);
    flux[IMX] = -porFlux * (lam3 * dru + uAvg * abv6 + sx * abv7);
    flux[IMY] = -porFlux * (lam3 * drv + vAvg * abv6 + sy * abv7);
    flux[IMZ] = -porFlux * (lam3 * drw + wAvg * abv6 + sz * abv7);
    flux[IRHOE] = -porFlux * (lam3 * drE + hAvg * abv6 + unAvg * abv7);
}

static void upwind_dir(const OrcBlock* b, const AdfbParams* prm, Dims d, const double* s, const int8_t* por, long sd,
                       int i0, int j0, int k0, double rFil) {
    for (int k = k0; k <= d.kl; k++) for (int j = j0; j <= d.jl; j++) for (int i = i0; i <= d.il; i++) {
        long c = IDX(i, j, k), cp = c + sd, cpp = c + 2 * sd, cm = c - sd;
        double left[5], right[5], flux[5];
        if (prm->limiter == ADFB_LIM_FIRSTORDER) {
            for (int l = 0; l < 4; l++) { left[l] = W(c, l); right[l] = W(cp, l); }
            left[IRHOE] = b->p[c]; right[IRHOE] = b->p[cp];
        } else {
            double du1[5], du2[5], du3[5];
            for (int l = 0; l < 4; l++) { du1[l] = W(c, l) - W(cm, l); du2[l] = W(cp, l) - W(c, l); du3[l] = W(cpp, l) - W(cp, l); }
            du1[4] = b->p[c] - b->p[cm]; du2[4] = b->p[cp] - b->p[c]; du3[4] = b->p[cpp] - b->p[cp];
            left_right_state(prm, du1, du2, du3, left, right);
            for (int l = 0; l < 4; l++) { left[l] = left[l] + W(c, l); right[l] = right[l] + W(cp, l); }
            left[IRHOE] = left[IRHOE] + b->p[c]; right[IRHOE] = right[IRHOE] + b->p[cp];
        }
        riemann_flux(prm, left, right, s[c], s[d.N + c], s[2 * d.N + c], por[c], rFil, flux);
        for (int l = 0; l < 5; l++) { FW(c, l) = FW(c, l) + flux[l]; FW(cp, l) = FW(cp, l) - flux[l]; }
    }
}

void orc_upwind_flux(const OrcBlock* b, const AdfbParams* prm, double rFil) {
    Dims d = dims_of(b);
    double sFil = one - rFil;
    for (int l = 0; l < 5; l++)
        for (int k = 2; k <= d.kl; k++) for (int j = 2; j <= d.jl; j++) for (int i = 2; i <= d.il; i++) FW(IDX(i, j, k), l) = sFil * FW(IDX(i, j, k), l);
    upwind_dir(b, prm, d, b->si, b->porI, d.sI, 1, 2, 2, rFil);
    upwind_dir(b, prm, d, b->sj, b->porJ, d.sJ, 2, 1, 2, rFil);
    upwind_dir(b, prm, d, b->sk, b->porK, d.sK, 2, 2, 1, rFil);
}

/* ------------------------------------------------------------------------ */
/* Approximate (preconditioner / ANK) flux variants, src/NKSolver/blockette.F90:
   inviscidDissFluxScalarApprox :4367-4617, inviscidDissFluxMatrixApprox :4619-5166,
   viscousFluxApprox :6467-6837.  The sensor is the FROZEN shockSensor array
   (referenceShockSensor, src/adjoint/adjointUtils.F90:1909-1969), first-order differences,
   fourth-difference dissipation lumped into the second with `sigma`.
   The sensor is the pressure for Euler AND for matrix dissipation (:1930), the entropy otherwise; the
   reference fills only the cells the dissipation stencils read (no edge/corner halos, :1941-1965), here the
   whole box is filled -- the extra cells are never read. */
void orc_reference_shock_sensor(const OrcBlock* b, const AdfbParams* prm) {
    Dims d = dims_of(b);
    int pressure = prm->equations == ADFB_EULER || prm->spaceDiscr == ADFB_DISS_MATRIX;
    for (long c = 0; c < d.N; c++)
        b->shock[c] = pressure ? b->p[c] : b->p[c] / pow(W(c, IRHO), prm->gammaInf);
}

static void diss_scalar_approx_dir(const OrcBlock* b, const AdfbParams* prm, Dims d, const double* rad, const double* dss,
                                   const int8_t* por, long sd, int i0, int j0, int k0) {
    const double dssMax = 0.25;
    double fis2 = prm->vis2, fis4 = prm->vis4;
    const double* p = b->p;
    for (int k = k0; k <= d.kl; k++) for (int j = j0; j <= d.jl; j++) for (int i = i0; i <= d.il; i++) {
        long c = IDX(i, j, k), cp = c + sd;
        double ppor = zero;
        if (por[c] == ADFB_NORMALFLUX) ppor = half;
        double rrad = ppor * (rad[c] + rad[cp]);
        double dis2 = fis2 * rrad * dmin(dssMax, dmax(dss[c], dss[cp])) + prm->sigma * fis4 * rrad;
        double ddw, fs;
        ddw = W(cp, IRHO) - W(c, IRHO); fs = dis2 * ddw; FW(cp, IRHO) += fs; FW(c, IRHO) -= fs;
        for (int l = IVX; l <= IVZ; l++) { ddw = W(cp, l) * W(cp, IRHO) - W(c, l) * W(c, IRHO); fs = dis2 * ddw; FW(cp, l) += fs; FW(c, l) -= fs; }
        ddw = (W(cp, IRHOE) + p[cp]) - (W(c, IRHOE) + p[c]); fs = dis2 * ddw; FW(cp, IRHOE) += fs; FW(c, IRHOE) -= fs;
    }
}
void orc_diss_scalar_approx(const OrcBlock* b, const AdfbParams* prm) {
    Dims d = dims_of(b);
    double sslim = prm->equations == ADFB_EULER ? 0.001 * prm->pInfCorr : 0.001 * prm->pInfCorr / pow(prm->rhoInf, prm->gammaInf);
    const double* ss = b->shock;
    for (int k = 1; k <= d.ke; k++) for (int j = 1; j <= d.je; j++) for (int i = 1; i <= d.ie; i++) {
        long c = IDX(i, j, k);
        b->dss[0 * d.N + c] = fabs((ss[c + 1] - two * ss[c] + ss[c - 1]) / (ss[c + 1] + two * ss[c] + ss[c - 1] + sslim));
        b->dss[1 * d.N + c] = fabs((ss[c + d.sJ] - two * ss[c] + ss[c - d.sJ]) / (ss[c + d.sJ] + two * ss[c] + ss[c - d.sJ] + sslim));
        b->dss[2 * d.N + c] = fabs((ss[c + d.sK] - two * ss[c] + ss[c - d.sK]) / (ss[c + d.sK] + two * ss[c] + ss[c - d.sK] + sslim));
    }
    diss_scalar_approx_dir(b, prm, d, b->radI, b->dss, b->porI, d.sI, 1, 2, 2);
    diss_scalar_approx_dir(b, prm, d, b->radJ, b->dss + d.N, b->porJ, d.sJ, 2, 1, 2);
    diss_scalar_approx_dir(b, prm, d, b->radK, b->dss + 2 * d.N, b->porK, d.sK, 2, 2, 1);
}

static void diss_matrix_approx_dir(const OrcBlock* b, const AdfbParams* prm, Dims d, const double* s, const double* dss,
                                   const int8_t* por, long sd, int i0, int j0, int k0, double fis2, double fis4, double fis0) {
    /* fis0 >= 0: inviscidDissFluxMatrixCoarse (fluxes.F90:5205-5711), dis0 = fis0 * ppor instead of the sensor form */
    const double dpMax = 0.25, epsAcoustic = 0.25, epsShear = 0.025;
    double gam = prm->gammaInf;
    const double* p = b->p;
    for (int k = k0; k <= d.kl; k++) for (int j = j0; j <= d.jl; j++) for (int i = i0; i <= d.il; i++) {
        long c = IDX(i, j, k), cp = c + sd;
        double ppor = zero;
        if (por[c] == ADFB_NORMALFLUX) ppor = one;
        double dis2 = fis0 >= zero ? fis0 * ppor : fis2 * ppor * dmin(dpMax, dmax(dss[c], dss[cp])) + prm->sigma * fis4 * ppor;
        double dr = dis2 * (W(cp, IRHO) - W(c, IRHO));
        double dru = dis2 * (W(cp, IRHO) * W(cp, IVX) - W(c, IRHO) * W(c, IVX));
        double drv = dis2 * (W(cp, IRHO) * W(cp, IVY) - W(c, IRHO) * W(c, IVY));
        double drw = dis2 * (W(cp, IRHO) * W(cp, IVZ) - W(c, IRHO) * W(c, IVZ));
        double dre = dis2 * (W(cp, IRHOE) - W(c, IRHOE));
        double gm1 = gam - one, ovgm1 = one / gm1;
        double uAvg = half * (W(cp, IVX) + W(c, IVX)), vAvg = half * (W(cp, IVY) + W(c, IVY)), wAvg = half * (W(cp, IVZ) + W(c, IVZ));
        double a2Avg = half * (gam * p[cp] / W(cp, IRHO) + gam * p[c] / W(c, IRHO));
        double s1 = s[c], s2 = s[d.N + c], s3 = s[2 * d.N + c];
        double area = sqrt(s1 * s1 + s2 * s2 + s3 * s3);
        double tmp = one / dmax(1.e-25, area);
        double sx = s1 * tmp, sy = s2 * tmp, sz = s3 * tmp;
        double alphaAvg = half * (uAvg * uAvg + vAvg * vAvg + wAvg * wAvg);
        double hAvg = alphaAvg + ovgm1 * a2Avg;
        double aAvg = sqrt(a2Avg);
        double unAvg = uAvg * sx + vAvg * sy + wAvg * sz;
        double ovaAvg = one / aAvg, ova2Avg = one / a2Avg;
        double lam1 = fabs(unAvg + aAvg), lam2 = fabs(unAvg - aAvg), lam3 = fabs(unAvg);
        double rrad = lam3 + aAvg;
        lam1 = dmax(lam1, epsAcoustic * rrad) * area;
        lam2 = dmax(lam2, epsAcoustic * rrad) * area;
        lam3 = dmax(lam3, epsShear * rrad) * area;
        double abv1 = half * (lam1 + lam2), abv2 = half * (lam1 - lam2), abv3 = abv1 - lam3;
        double abv4 = gm1 * (alphaAvg * dr - uAvg * dru - vAvg * drv - wAvg * drw + dre);
        double abv5 = sx * dru + sy * drv + sz * drw - unAvg * dr;
        double abv6 = abv3 * abv4 * ova2Avg + abv2 * abv5 * ovaAvg;
        double abv7 = abv2 * abv4 * ovaAvg + abv3 * abv5;
        double fs;
        fs = lam3 * dr + abv6; FW(cp, IRHO) += fs; FW(c, IRHO) -= fs;
        fs = lam3 * dru + uAvg * abv6 + sx * abv7; FW(cp, IMX) += fs; FW(c, IMX) -= fs;
        fs = lam3 * drv + vAvg * abv6 + sy * abv7; FW(cp, IMY) += fs; FW(c, IMY) -= fs;
        fs = lam3 * drw + wAvg * abv6 + sz * abv7; FW(cp, IMZ) += fs; FW(c, IMZ) -= fs;
        fs = lam3 * dre + hAvg * abv6 + unAvg * abv7; FW(cp, IRHOE) += fs; FW(c, IRHOE) -= fs;
    }
}
void orc_diss_matrix_approx(const OrcBlock* b, const AdfbParams* prm, double rFil) {
    Dims d = dims_of(b);
    const double omega = 0.5, oneMinOmega = one - omega;
    double plim = 0.001 * prm->pInfCorr;
    const double* ss = b->shock;
    double fis2 = rFil * prm->vis2, fis4 = rFil * prm->vis4, sfil = one - rFil;
    for (long q = 0; q < 5 * d.N; q++) b->fw[q] = sfil * b->fw[q];
    for (int k = 1; k <= d.ke; k++) for (int j = 1; j <= d.je; j++) for (int i = 1; i <= d.ie; i++) {
        long c = IDX(i, j, k);
        long st[3] = {1, d.sJ, d.sK};
        for (int m = 0; m < 3; m++) {
            long sd = st[m];
            b->dss[m * d.N + c] = fabs((ss[c + sd] - two * ss[c] + ss[c - sd]) /
                                       (omega * (ss[c + sd] + two * ss[c] + ss[c - sd]) +
                                        oneMinOmega * (fabs(ss[c + sd] - ss[c]) + fabs(ss[c] - ss[c - sd])) + plim));
        }
    }
    diss_matrix_approx_dir(b, prm, d, b->si, b->dss, b->porI, d.sI, 1, 2, 2, fis2, fis4, -one);
    diss_matrix_approx_dir(b, prm, d, b->sj, b->dss + d.N, b->porJ, d.sJ, 2, 1, 2, fis2, fis4, -one);
    diss_matrix_approx_dir(b, prm, d, b->sk, b->dss + 2 * d.N, b->porK, d.sK, 2, 2, 1, fis2, fis4, -one);
}

/* inviscidDissFluxMatrixCoarse, fluxes.F90:5205-5711: first-order matrix dissipation of the coarse multigrid levels */
void orc_diss_matrix_coarse(const OrcBlock* b, const AdfbParams* prm, double rFil) {
    Dims d = dims_of(b);
    if (fabs(rFil) < thresholdReal) return;
    double fis0 = rFil * prm->vis2Coarse, sfil = one - rFil;
    for (int k = 2; k <= d.kl; k++) for (int j = 2; j <= d.jl; j++) for (int i = 2; i <= d.il; i++) {
        long c = IDX(i, j, k);
        for (int l = 0; l < 5; l++) FW(c, l) = sfil * FW(c, l);
    }
    diss_matrix_approx_dir(b, prm, d, b->si, b->dss, b->porI, d.sI, 1, 2, 2, zero, zero, fis0);
    diss_matrix_approx_dir(b, prm, d, b->sj, b->dss + d.N, b->porJ, d.sJ, 2, 1, 2, zero, zero, fis0);
    diss_matrix_approx_dir(b, prm, d, b->sk, b->dss + 2 * d.N, b->porK, d.sK, 2, 2, 1, zero, zero, fis0);
}

/* viscousFluxApprox: thin-layer viscous flux, sweeps i, j, k (blockette.F90:6467-6837) */
static void viscous_approx_dir(const OrcBlock* b, const AdfbParams* prm, Dims d, const double* s, const int8_t* por,
                               long sd, long t1, long t2, int i0, int j0, int k0, double rFilv) {
    const double twoThird = two * third;
    double gam = prm->gammaInf;
    for (int k = k0; k <= d.kl; k++) for (int j = j0; j <= d.jl; j++) for (int i = i0; i <= d.il; i++) {
        long c = IDX(i, j, k), cp = c + sd;
        long n = c, n1 = c - t1 - t2, n2 = c - t2, n3 = c - t1;
        double ss3[3];
        for (int m = 0; m < 3; m++)
            ss3[m] = eighth * (X(n1 + sd, m) - X(n1 - sd, m) + X(n3 + sd, m) - X(n3 - sd, m) + X(n2 + sd, m) - X(n2 - sd, m) + X(n + sd, m) - X(n - sd, m));
        double ss = one / (ss3[0] * ss3[0] + ss3[1] * ss3[1] + ss3[2] * ss3[2]);
        double ssx = ss * ss3[0], ssy = ss * ss3[1], ssz = ss * ss3[2];
        double dd = W(cp, IVX) - W(c, IVX);
        double u_x = dd * ssx, u_y = dd * ssy, u_z = dd * ssz;
        dd = W(cp, IVY) - W(c, IVY);
        double v_x = dd * ssx, v_y = dd * ssy, v_z = dd * ssz;
        dd = W(cp, IVZ) - W(c, IVZ);
        double w_x = dd * ssx, w_y = dd * ssy, w_z = dd * ssz;
        dd = b->aa[cp] - b->aa[c];
        double q_x = -dd * ssx, q_y = -dd * ssy, q_z = -dd * ssz;
        double porv = half * rFilv;
        if (por[c] == ADFB_NOFLUX) porv = zero;
        double mul = porv * (b->rlv[c] + b->rlv[cp]), mue = porv * (b->rev[c] + b->rev[cp]);
        double mut = mul + mue;
        double gm1 = half * (gam + gam) - one;
        double heatCoef = mul * (one / (prm->prandtl * gm1)) + mue * (one / (prm->prandtlTurb * gm1));
        double fracDiv = twoThird * (u_x + v_y + w_z);
        double tauxx = mut * (two * u_x - fracDiv), tauyy = mut * (two * v_y - fracDiv), tauzz = mut * (two * w_z - fracDiv);
        double tauxy = mut * (u_y + v_x), tauxz = mut * (u_z + w_x), tauyz = mut * (v_z + w_y);
        q_x = heatCoef * q_x; q_y = heatCoef * q_y; q_z = heatCoef * q_z;
        double ubar = half * (W(c, IVX) + W(cp, IVX)), vbar = half * (W(c, IVY) + W(cp, IVY)), wbar = half * (W(c, IVZ) + W(cp, IVZ));
        double s1 = s[c], s2 = s[d.N + c], s3 = s[2 * d.N + c];
        double fmx = tauxx * s1 + tauxy * s2 + tauxz * s3;
        double fmy = tauxy * s1 + tauyy * s2 + tauyz * s3;
        double fmz = tauxz * s1 + tauyz * s2 + tauzz * s3;
        double frhoE = (ubar * tauxx + vbar * tauxy + wbar * tauxz) * s1 + (ubar * tauxy + vbar * tauyy + wbar * tauyz) * s2 +
                       (ubar * tauxz + vbar * tauyz + wbar * tauzz) * s3 - q_x * s1 - q_y * s2 - q_z * s3;
        FW(cp, IMX) += fmx; FW(cp, IMY) += fmy; FW(cp, IMZ) += fmz; FW(cp, IRHOE) += frhoE;
        FW(c, IMX) -= fmx; FW(c, IMY) -= fmy; FW(c, IMZ) -= fmz; FW(c, IRHOE) -= frhoE;
    }
}
void orc_viscous_flux_approx(const OrcBlock* b, const AdfbParams* prm, double rFil) {
    Dims d = dims_of(b);
    viscous_approx_dir(b, prm, d, b->si, b->porI, d.sI, d.sJ, d.sK, 1, 2, 2, rFil);
    viscous_approx_dir(b, prm, d, b->sj, b->porJ, d.sJ, d.sI, d.sK, 2, 1, 2, rFil);
    viscous_approx_dir(b, prm, d, b->sk, b->porK, d.sK, d.sI, d.sJ, 2, 2, 1, rFil);
}
