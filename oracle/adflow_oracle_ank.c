/*
 * adflow_oracle_ank.c -- CPU restatement of the ANK pieces on the path (TEST INFRASTRUCTURE ONLY):
 * computeTimeStepBlock (src/NKSolver/NKSolvers.F90:2116-2329) and physicalityCheckANK (:3013-3210).
 * PARITY PINNED bit for bit against the translated reference routines (oracle/_ref/anksolver_ref.c,
 * tests/test_oracle_vs_reference_ank.py).  Blocks are column-major nState x nState like the reference's.
 */
#include "orc_internal.h"

#define B_(r, c) blk[((r) - 1) + n * ((c) - 1)]
#define M_(m, r, c) m[((r) - 1) + n * ((c) - 1)]

static void matmul_(int n, const double* a, const double* b, double* c) { /* c = MATMUL(a, b), may alias */
    double t[36];
    for (int j = 0; j < n; j++) for (int i = 0; i < n; i++) {
        double s = 0.0;
        for (int k = 0; k < n; k++) s += a[i + n * k] * b[k + n * j];
        t[i + n * j] = s;
    }
    memcpy(c, t, sizeof(double) * n * n);
}
static void matmul_nt_(int n, const double* a, const double* b, double* c) { /* c = MATMUL(a, TRANSPOSE(b)) */
    double t[36];
    for (int j = 0; j < n; j++) for (int i = 0; i < n; i++) {
        double s = 0.0;
        for (int k = 0; k < n; k++) s += a[i + n * k] * b[j + n * k];
        t[i + n * j] = s;
    }
    memcpy(c, t, sizeof(double) * n * n);
}

void orc_ank_time_step_block(const OrcBlock* b, const AdfbParams* prm, const AdfbAnkParams* ank, int i, int j, int k, double* blk) {
    Dims d = dims_of(b);
    const int n = ank->coupled ? b->nw : 5;
    long c = IDX(i, j, k);
    double stateToCons[36], streamToCart[36], symmToCons[36], consToSymm[36];
    memset(blk, 0, sizeof(double) * n * n);
    memset(stateToCons, 0, sizeof stateToCons); memset(streamToCart, 0, sizeof streamToCart);
    memset(symmToCons, 0, sizeof symmToCons); memset(consToSymm, 0, sizeof consToSymm);
    double rho = W(c, IRHO), velX = W(c, IVX), velY = W(c, IVY), velZ = W(c, IVZ);
    double dtInv = one / (ank->cfl * b->dtl[c] * b->volRef[c]);
    M_(stateToCons, 1, 1) = one;
    M_(stateToCons, 2, 1) = velX; M_(stateToCons, 2, 2) = rho;
    M_(stateToCons, 3, 1) = velY; M_(stateToCons, 3, 3) = rho;
    M_(stateToCons, 4, 1) = velZ; M_(stateToCons, 4, 4) = rho;
    M_(stateToCons, 5, 5) = one;
    if (ank->coupled) M_(stateToCons, 6, 6) = prm->turbResScale / ank->turbCFLScale;
    if (ank->charTimeStepType == 0) {
        for (int q = 0; q < n * n; q++) blk[q] = stateToCons[q] * dtInv;
        return;
    }
    if (ank->coupled) { B_(6, 6) = one; M_(streamToCart, 6, 6) = one; M_(symmToCons, 6, 6) = one; M_(consToSymm, 6, 6) = one; }
    double aa = b->aa[c];
    if (prm->equations == ADFB_EULER) { aa = prm->gammaInf * b->p[c] / rho; b->aa[c] = aa; }
    double speed = sqrt(velX * velX + velY * velY + velZ * velZ);
    double speedOfSound = sqrt(aa);
    double mach = speed / speedOfSound;
    double machSqr = mach * mach;
    double gm1 = prm->gammaInf - one;
    M_(symmToCons, 1, 1) = rho / speedOfSound;
    M_(symmToCons, 1, 5) = -one / aa;
    M_(symmToCons, 2, 1) = rho * velX / speedOfSound; M_(symmToCons, 2, 2) = rho; M_(symmToCons, 2, 5) = -velX / aa;
    M_(symmToCons, 3, 1) = rho * velY / speedOfSound; M_(symmToCons, 3, 3) = rho; M_(symmToCons, 3, 5) = -velY / aa;
    M_(symmToCons, 4, 1) = rho * velZ / speedOfSound; M_(symmToCons, 4, 4) = rho; M_(symmToCons, 4, 5) = -velZ / aa;
    M_(symmToCons, 5, 1) = rho * speedOfSound * (machSqr / 2 + 1 / gm1);
    M_(symmToCons, 5, 2) = rho * velX; M_(symmToCons, 5, 3) = rho * velY; M_(symmToCons, 5, 4) = rho * velZ;
    M_(symmToCons, 5, 5) = -machSqr / 2;
    M_(consToSymm, 1, 1) = gm1 / 2 * speedOfSound * machSqr / rho;
    M_(consToSymm, 1, 2) = -gm1 * velX / (rho * speedOfSound);
    M_(consToSymm, 1, 3) = -gm1 * velY / (rho * speedOfSound);
    M_(consToSymm, 1, 4) = -gm1 * velZ / (rho * speedOfSound);
    M_(consToSymm, 1, 5) = gm1 / (rho * speedOfSound);
    M_(consToSymm, 2, 1) = -velX / rho; M_(consToSymm, 2, 2) = one / rho;
    M_(consToSymm, 3, 1) = -velY / rho; M_(consToSymm, 3, 3) = one / rho;
    M_(consToSymm, 4, 1) = -velZ / rho; M_(consToSymm, 4, 4) = one / rho;
    M_(consToSymm, 5, 1) = aa * (gm1 / 2 * machSqr - one);
    M_(consToSymm, 5, 2) = -gm1 * velX; M_(consToSymm, 5, 3) = -gm1 * velY; M_(consToSymm, 5, 4) = -gm1 * velZ;
    M_(consToSymm, 5, 5) = gm1;
    double blend = ank->cfl / ank->cflLimit;
    if (ank->charTimeStepType == 1) {  /* VLR */
        double machSqrTrunc = dmax(machSqr, 1e-4 * (ank->machInf * ank->machInf));
        double beta, tau;
        if (mach < one) { beta = sqrt(one - machSqrTrunc); tau = beta; }
        else { beta = sqrt(machSqrTrunc - one); tau = sqrt(one - one / machSqrTrunc) + 1e-4; }
        B_(1, 1) = blend * (beta * beta + tau) / (machSqrTrunc * tau) + (one - blend) * one;
        B_(1, 2) = blend * one / mach;
        B_(2, 1) = blend * one / mach;
        B_(2, 2) = one;
        B_(3, 3) = blend * one / tau + (one - blend) * one;
        B_(4, 4) = blend * one / tau + (one - blend) * one;
        B_(5, 5) = one;
        double speedXY = sqrt(velX * velX + velY * velY);
        double sinTheta = velY / speedXY, cosTheta = velX / speedXY, sinAlpha = velZ / speed, cosAlpha = speedXY / speed;
        M_(streamToCart, 1, 1) = one;
        M_(streamToCart, 2, 2) = cosAlpha * cosTheta; M_(streamToCart, 2, 3) = -sinTheta; M_(streamToCart, 2, 4) = -sinAlpha * cosTheta;
        M_(streamToCart, 3, 2) = cosAlpha * sinTheta; M_(streamToCart, 3, 3) = cosTheta; M_(streamToCart, 3, 4) = -sinAlpha * sinTheta;
        M_(streamToCart, 4, 2) = sinAlpha; M_(streamToCart, 4, 4) = cosAlpha;
        M_(streamToCart, 5, 5) = one;
        matmul_(n, streamToCart, blk, blk);
        matmul_nt_(n, blk, streamToCart, blk);
        matmul_(n, symmToCons, blk, blk);
        matmul_(n, blk, consToSymm, blk);
        matmul_(n, blk, stateToCons, blk);
        for (int q = 0; q < n * n; q++) blk[q] = blk[q] * dtInv;
    } else {  /* Turkel */
        double machSqrTrunc = dmin(one, dmax(machSqr, 1e-4 * (ank->machInf * ank->machInf)));
        double m2 = machSqrTrunc * machSqrTrunc, m4 = m2 * m2, m8 = m4 * m4;   /* x**10 = ((x^2)^2)^2 * x^2, __powidf2 order */
        double alpha = one - m8 * m2;
        B_(1, 1) = blend * one / machSqrTrunc + (one - blend) * one;
        B_(2, 1) = blend * alpha * velX / speedOfSound / machSqrTrunc;
        B_(3, 1) = blend * alpha * velY / speedOfSound / machSqrTrunc;
        B_(4, 1) = blend * alpha * velZ / speedOfSound / machSqrTrunc;
        B_(2, 2) = one; B_(3, 3) = one; B_(4, 4) = one; B_(5, 5) = one;
        matmul_(n, symmToCons, blk, blk);
        matmul_(n, blk, consToSymm, blk);
        matmul_(n, blk, stateToCons, blk);
        for (int q = 0; q < n * n; q++) blk[q] = blk[q] * dtInv;
    }
}

/* physicalityCheckANK, NKSolvers.F90:3013-3210: vectors of nState entries per owned cell; deltaW may be clipped
   (turbulence, coupled); returns the new lambdaP */
double orc_ank_physicality_check(const AdfbAnkParams* ank, int nState, long nCells, const double* wVec, double* dVec, double lambdaP) {
    double lambdaL = lambdaP;
    long ii = 0;
    for (long q = 0; q < nCells; q++) {
        double ratio = fabs(wVec[ii] / (dVec[ii] + eps_)) * ank->physLSTol;
        lambdaL = dmin(lambdaL, ratio);
        ii += 4;
        ratio = fabs(wVec[ii] / (dVec[ii] + eps_)) * ank->physLSTol;
        lambdaL = dmin(lambdaL, ratio);
        ii += 1;
        if (ank->coupled) {
            ratio = (wVec[ii] / (dVec[ii] + eps_)) * ank->physLSTolTurb;
            if (ratio < ank->stepFactor * ank->stepMin) {
                if (ratio > zero) dVec[ii] = wVec[ii] * ank->physLSTolTurb;
                ratio = one;
            }
            lambdaL = dmin(lambdaL, ratio);
            ii += 1;
            ii += nState - 6;
        }
    }
    if (lambdaL != lambdaL) lambdaL = zero;
    return lambdaL;
}

/* physicalityCheckANKTurb, NKSolvers.F90:3212-3335 (turbulence KSP of the decoupled ANK): vectors of nt2-nt1+1 = 1 entry
   per owned cell; too-limiting updates are clipped in dVec; returns the new lambdaP */
double orc_ank_physicality_check_turb(const AdfbAnkParams* ank, long nCells, const double* wVec, double* dVec, double lambdaP) {
    double lambdaL = lambdaP;
    for (long ii = 0; ii < nCells; ii++) {
        double ratio = (wVec[ii] / (dVec[ii] + eps_)) * ank->physLSTolTurb;
        if (ratio < ank->stepFactor * ank->stepMin) {
            if (ratio > zero) dVec[ii] = wVec[ii] * ank->physLSTolTurb;
            ratio = one;
        }
        lambdaL = dmin(lambdaL, ratio);
    }
    if (lambdaL != lambdaL) lambdaL = zero;
    return lambdaL;
}

/* the vector part of FormFunction_mf_turb, NKSolvers.F90:2540-2612, after setWANK(inVec, nt1, nt2) and
   blocketteRes(useFlowRes = .false.): setRVecANKTurb (:2935-2973: dw(itu1) / volRef * turbResScale(1)) plus the time
   stepping term inVec / (ANK_CFL dtl volRef) * turbResScale / ANK_turbCFLScale; owned cells, k, j, i order */
void orc_ank_turb_rvec(const OrcBlock* b, const AdfbParams* prm, const AdfbAnkParams* ank, const double* inVec, double* rVec) {
    const int il = b->nx + 1, jl = b->ny + 1, kl = b->nz + 1;
    const long NI = b->nx + 4, NJ = b->ny + 4, N = NI * NJ * (b->nz + 4);
    const double* dw = (const double*)b->dw; const double* volRef = (const double*)b->volRef; const double* dtl = (const double*)b->dtl;
    long ii = 0;
    for (int k = 2; k <= kl; k++)
        for (int j = 2; j <= jl; j++)
            for (int i = 2; i <= il; i++) {
                const long c = i + NI * (j + NJ * (long)k);
                const double ovv = one / volRef[c];
                rVec[ii] = dw[5 * N + c] * ovv * prm->turbResScale;
                const double dtinv = one / (ank->cfl * dtl[c] * volRef[c]);
                rVec[ii] = rVec[ii] + inVec[ii] * dtinv * prm->turbResScale / ank->turbCFLScale;
                ii++;
            }
}
