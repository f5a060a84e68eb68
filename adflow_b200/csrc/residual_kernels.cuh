// residual_kernels.cuh -- fused residual kernels (gather form) for sm_100a
//
// Replaces the per-block residual core of the reference
// (blocketteResCore, src/NKSolver/blockette.F90:299-753 and its operator twins
// in src/solver/fluxes.F90, src/turbulence/sa.F90, src/utils/flowUtils.F90).
//
// The reference scatters every face flux to its two cells
// (`dw(i+1) -= fs; dw(i) += fs`); here every thread owns one cell and GATHERS
// its six faces, adding them in exactly the order in which the reference's
// i/j/k sweeps would have touched that cell, so no atomics and no halo writes
// are needed and results agree with the scatter form to round-off.
//
// Launch plan per residual evaluation:
//   k_prep   : box cells     -> ss (entropy), aa, radI/J/K, [dtl]
//   k_nodal  : cells 1:ie    -> dss(3) and the 12 nodal gradients at nodes 1:il
//   k_resid  : owned cells   -> SA source/advection/diffusion, central + JST
//                               + viscous fluxes, epilogue -> dw(1:nw)
#pragma once
#include "adfb_common.cuh"
#include <math.h>

#define IRHO 0
#define IVX 1
#define IVY 2
#define IVZ 3
#define IRHOE 4
#define ITU1 5

namespace {

__device__ __forceinline__ bool cell_index(const Dims& d, int& i, int& j, int& k, int i0, int j0, int k0) {
    i = blockIdx.x * blockDim.x + threadIdx.x + i0;
    j = blockIdx.y * blockDim.y + threadIdx.y + j0;
    k = blockIdx.z * blockDim.z + threadIdx.z + k0;
    return true;
}

// ---------------------------------------------------------------------------
// k_prep: entropy (inviscidDissFluxScalar, blockette.F90:3055-3089), speed of
// sound squared (:5168-5203), spectral radii and local time step (timeStep,
// :1899-2148).  One pass over the box; radii/aa only on cells 1:ie, dtl on owned.
__global__ void __launch_bounds__(256) k_prep(Dims d, BlockDev b, int updateDt, int doRad) {
    int i, j, k;
    cell_index(d, i, j, k, 0, 0, 0);
    if (i > d.ib || j > d.jb || k > d.kb) return;
    const long long c = ADFB_IDX(i, j, k);
    const long long N = d.N;
    const double gam = c_prm.gammaInf;
    const double rho = b.w[c], p = b.p[c];
    b.ss[c] = (c_prm.equations == ADFB_EULER) ? p : p / pow(rho, gam);
    if (i < 1 || i > d.ie || j < 1 || j > d.je || k < 1 || k > d.ke) return;
    const bool viscous = c_prm.equations != ADFB_EULER;
    if (viscous) b.aa[c] = gam * p / rho;
    if (!doRad) return;  // smoother path: radii/dtl are frozen between timeStep calls

    const double clim2 = 0.000001 * gam * c_prm.pInfCorr / c_prm.rhoInf;
    const double adis = c_prm.adis, asf = c_prm.acousticScaleFactor;
    const double ux = b.w[N + c], uy = b.w[2 * N + c], uz = b.w[3 * N + c];
    double cc2 = gam * p / rho;
    cc2 = dmax_(cc2, clim2);
    double sx, sy, sz, q;
    sx = b.si[c - 1] + b.si[c]; sy = b.si[N + c - 1] + b.si[N + c]; sz = b.si[2 * N + c - 1] + b.si[2 * N + c];
    const double sxi = sx, syi = sy, szi = sz;
    q = ux * sx + uy * sy + uz * sz;
    double ri = 0.5 * (fabs(q) + asf * sqrt(cc2 * (sx * sx + sy * sy + sz * sz)));
    sx = b.sj[c - d.sJ] + b.sj[c]; sy = b.sj[N + c - d.sJ] + b.sj[N + c]; sz = b.sj[2 * N + c - d.sJ] + b.sj[2 * N + c];
    const double sxj = sx, syj = sy, szj = sz;
    q = ux * sx + uy * sy + uz * sz;
    double rj = 0.5 * (fabs(q) + asf * sqrt(cc2 * (sx * sx + sy * sy + sz * sz)));
    sx = b.sk[c - d.sK] + b.sk[c]; sy = b.sk[N + c - d.sK] + b.sk[N + c]; sz = b.sk[2 * N + c - d.sK] + b.sk[2 * N + c];
    q = ux * sx + uy * sy + uz * sz;
    double rk = 0.5 * (fabs(q) + asf * sqrt(cc2 * (sx * sx + sy * sy + sz * sz)));
    double dt = ri + rj + rk;
    ri = dmax_(ri, 1.e-25); rj = dmax_(rj, 1.e-25); rk = dmax_(rk, 1.e-25);
    const double rij = pow(ri / rj, adis), rjk = pow(rj / rk, adis), rki = pow(rk / ri, adis);
    b.radI[c] = ri * (1.0 + 1.0 / rij + rki);
    b.radJ[c] = rj * (1.0 + 1.0 / rjk + rij);
    b.radK[c] = rk * (1.0 + 1.0 / rki + rjk);

    if (!updateDt) return;
    if (i < 2 || i > d.il || j < 2 || j > d.jl || k < 2 || k > d.kl) return;
    if (viscous) {
        double rmu = b.rlv[c];
        rmu = rmu + b.rev[c];
        rmu = 0.5 * rmu / (rho * b.vol[c]);
        dt = dt + rmu * (sxi * sxi + syi * syi + szi * szi);
        dt = dt + rmu * (sxj * sxj + syj * syj + szj * szj);
        dt = dt + rmu * (sx * sx + sy * sy + sz * sz);
    }
    const double plim = 0.001 * c_prm.pInfCorr;
    const double* pp = b.p;
    const double dpi = fabs(pp[c + 1] - 2.0 * p + pp[c - 1]) / (pp[c + 1] + 2.0 * p + pp[c - 1] + plim);
    const double dpj = fabs(pp[c + d.sJ] - 2.0 * p + pp[c - d.sJ]) / (pp[c + d.sJ] + 2.0 * p + pp[c - d.sJ] + plim);
    const double dpk = fabs(pp[c + d.sK] - 2.0 * p + pp[c - d.sK]) / (pp[c + d.sK] + 2.0 * p + pp[c - d.sK] + plim);
    const double rfl = 1.0 / (1.0 + 2.0 * (dpi + dpj + dpk));
    b.dtl[c] = rfl / dt;
}

// ---------------------------------------------------------------------------
// nodal gradient contribution of one sweep direction (allNodalGradients,
// blockette.F90:5235-5479): dual-face "upper" (cell layer c+sd, added) and
// "lower" (cell layer c, subtracted); the reference's scatter order for a node
// is: subtract lower first, then add upper.
__device__ __forceinline__ void nodal_face(const BlockDev& b, long long N, long long c, long long sd, long long t1,
                                           long long t2, const double* __restrict__ s, double sv[3], double& ubar,
                                           double& vbar, double& wbar, double& a2) {
#pragma unroll
    for (int m = 0; m < 3; m++) {
        const double* sm = s + m * N;
        sv[m] = sm[c - sd] + sm[c - sd + t1] + sm[c - sd + t2] + sm[c - sd + t1 + t2] + sm[c] + sm[c + t1] +
                sm[c + t2] + sm[c + t1 + t2];
    }
    const double* w = b.w;
    ubar = 0.25 * (w[N + c] + w[N + c + t1] + w[N + c + t2] + w[N + c + t1 + t2]);
    vbar = 0.25 * (w[2 * N + c] + w[2 * N + c + t1] + w[2 * N + c + t2] + w[2 * N + c + t1 + t2]);
    wbar = 0.25 * (w[3 * N + c] + w[3 * N + c + t1] + w[3 * N + c + t2] + w[3 * N + c + t1 + t2]);
    a2 = 0.25 * (b.aa[c] + b.aa[c + t1] + b.aa[c + t2] + b.aa[c + t1 + t2]);
}

__device__ __forceinline__ void nodal_dir(const BlockDev& b, long long N, long long c, long long sd, long long t1,
                                          long long t2, const double* __restrict__ s, double g[12]) {
    double sv[3], ub, vb, wb, a2;
    nodal_face(b, N, c, sd, t1, t2, s, sv, ub, vb, wb, a2);
#pragma unroll
    for (int m = 0; m < 3; m++) { g[m] -= ub * sv[m]; g[3 + m] -= vb * sv[m]; g[6 + m] -= wb * sv[m]; g[9 + m] += a2 * sv[m]; }
    nodal_face(b, N, c + sd, sd, t1, t2, s, sv, ub, vb, wb, a2);
#pragma unroll
    for (int m = 0; m < 3; m++) { g[m] += ub * sv[m]; g[3 + m] += vb * sv[m]; g[6 + m] += wb * sv[m]; g[9 + m] -= a2 * sv[m]; }
}

// k_nodal: shock sensor dss (blockette.F90:3091-3105) on cells 1:ie and the
// nodal gradients (:5205-5515) on nodes 1:il.
__global__ void __launch_bounds__(256) k_nodal(Dims d, BlockDev b, int doGrad) {
    int i, j, k;
    cell_index(d, i, j, k, 1, 1, 1);
    if (i > d.ie || j > d.je || k > d.ke) return;
    const long long c = ADFB_IDX(i, j, k);
    const long long N = d.N;
    {
        const double sslim = (c_prm.equations == ADFB_EULER) ? 0.001 * c_prm.pInfCorr
                                                            : 0.001 * c_prm.pInfCorr / pow(c_prm.rhoInf, c_prm.gammaInf);
        const double* ss = b.ss;
        const double s0 = ss[c];
        b.dss[c] = fabs((ss[c + 1] - 2.0 * s0 + ss[c - 1]) / (ss[c + 1] + 2.0 * s0 + ss[c - 1] + sslim));
        b.dss[N + c] = fabs((ss[c + d.sJ] - 2.0 * s0 + ss[c - d.sJ]) / (ss[c + d.sJ] + 2.0 * s0 + ss[c - d.sJ] + sslim));
        b.dss[2 * N + c] = fabs((ss[c + d.sK] - 2.0 * s0 + ss[c - d.sK]) / (ss[c + d.sK] + 2.0 * s0 + ss[c - d.sK] + sslim));
    }
    if (!doGrad || i > d.il || j > d.jl || k > d.kl) return;
    double g[12];
#pragma unroll
    for (int m = 0; m < 12; m++) g[m] = 0.0;
    nodal_dir(b, N, c, d.sK, 1, d.sJ, b.sk, g);
    nodal_dir(b, N, c, d.sJ, 1, d.sK, b.sj, g);
    nodal_dir(b, N, c, 1, d.sJ, d.sK, b.si, g);
    const double* vol = b.vol;
    const double oVol = 1.0 / (vol[c] + vol[c + d.sK] + vol[c + 1] + vol[c + 1 + d.sK] + vol[c + d.sJ] +
                               vol[c + d.sJ + d.sK] + vol[c + 1 + d.sJ] + vol[c + 1 + d.sJ + d.sK]);
#pragma unroll
    for (int m = 0; m < 12; m++) b.grad[m * N + c] = g[m] * oVol;
}

// ---------------------------------------------------------------------------
// face fluxes.  `c` is the cell on the low side of the face, cp = c + sd.

// inviscidCentralFlux, blockette.F90:2150-2428
__device__ __forceinline__ void central_face(const BlockDev& b, long long N, long long c, long long cp,
                                             const double* __restrict__ s, int8_t por, double f[5]) {
    const double* w = b.w;
    const double s1 = s[c], s2 = s[N + c], s3 = s[2 * N + c];
    const double rp = w[cp], up = w[N + cp], vp = w[2 * N + cp], wp = w[3 * N + cp], ep = w[4 * N + cp];
    const double rm = w[c], um = w[N + c], vm = w[2 * N + c], wm = w[3 * N + c], em = w[4 * N + c];
    const double pp = b.p[cp], pm = b.p[c];
    double vnp = up * s1 + vp * s2 + wp * s3;
    double vnm = um * s1 + vm * s2 + wm * s3;
    double porVel = 1.0, porFlux = 0.5;
    if (por == ADFB_NOFLUX) porFlux = 0.0;
    if (por == ADFB_BOUNDFLUX) { porVel = 0.0; vnp = 0.0; vnm = 0.0; }
    porVel = porVel * porFlux;
    const double qsp = vnp * porVel, qsm = vnm * porVel;
    const double rqsp = qsp * rp, rqsm = qsm * rm;
    const double pa = porFlux * (pp + pm);
    f[0] = rqsp + rqsm;
    f[1] = rqsp * up + rqsm * um + pa * s1;
    f[2] = rqsp * vp + rqsm * vm + pa * s2;
    f[3] = rqsp * wp + rqsm * wm + pa * s3;
    f[4] = qsp * ep + qsm * em + porFlux * (vnp * pp + vnm * pm);
}

// inviscidDissFluxScalar, blockette.F90:3133-3338
__device__ __forceinline__ void jst_face(const BlockDev& b, long long N, long long c, long long sd,
                                         const double* __restrict__ rad, const double* __restrict__ dss, int8_t por,
                                         double fis2, double fis4, double f[5]) {
    const double* w = b.w;
    const double* p = b.p;
    const long long cp = c + sd, cpp = c + 2 * sd, cm = c - sd;
    const double ppor = (por == ADFB_NORMALFLUX) ? 0.5 : 0.0;
    const double rrad = ppor * (rad[c] + rad[cp]);
    const double dis2 = fis2 * rrad * dmin_(0.25, dmax_(dss[c], dss[cp]));
    const double dis4 = dmax_(fis4 * rrad - dis2, 0.0);
    const double r0 = w[c], r1 = w[cp], r2 = w[cpp], rm = w[cm];
    double ddw = r1 - r0;
    f[0] = dis2 * ddw - dis4 * (r2 - rm - 3.0 * ddw);
#pragma unroll
    for (int l = 1; l <= 3; l++) {
        ddw = w[l * N + cp] * r1 - w[l * N + c] * r0;
        f[l] = dis2 * ddw - dis4 * (w[l * N + cpp] * r2 - w[l * N + cm] * rm - 3.0 * ddw);
    }
    ddw = (w[4 * N + cp] + p[cp]) - (w[4 * N + c] + p[c]);
    f[4] = dis2 * ddw - dis4 * ((w[4 * N + cpp] + p[cpp]) - (w[4 * N + cm] + p[cm]) - 3.0 * ddw);
}

// viscousFlux, one face: blockette.F90:5576-5808 (k), :5876-6110 (j), :6172-6400 (i).
// n is the node at the (+,+,+) corner of cell c; the face's four nodes are
// n-t1-t2, n-t2, n-t1, n (reference order of the 4-node average).
__device__ __forceinline__ void visc_face(const BlockDev& b, long long N, long long c, long long sd, long long t1,
                                          long long t2, const double* __restrict__ s, int8_t por, double rFilv,
                                          double f[4]) {
    const long long cp = c + sd;
    const double* w = b.w;
    double porv = 0.5 * rFilv;
    if (por == ADFB_NOFLUX) porv = 0.0;
    const double mul = porv * (b.rlv[c] + b.rlv[cp]);
    const double mue = porv * (b.rev[c] + b.rev[cp]);
    const double mut = mul + mue;
    const double gm1 = c_prm.gammaInf - 1.0;
    const double heatCoef = mul * (1.0 / (c_prm.prandtl * gm1)) + mue * (1.0 / (c_prm.prandtlTurb * gm1));
    const long long n = c, n1 = c - t1 - t2, n2 = c - t2, n3 = c - t1;
    double g[12];
#pragma unroll
    for (int m = 0; m < 12; m++) {
        const double* gm = b.grad + m * N;
        g[m] = 0.25 * (gm[n1] + gm[n2] + gm[n3] + gm[n]);
    }
    double ss3[3];
#pragma unroll
    for (int m = 0; m < 3; m++) {
        const double* xm = b.x + m * N;
        ss3[m] = 0.125 * (xm[n1 + sd] - xm[n1 - sd] + xm[n3 + sd] - xm[n3 - sd] + xm[n2 + sd] - xm[n2 - sd] +
                          xm[n + sd] - xm[n - sd]);
    }
    const double snrm = 1.0 / sqrt(ss3[0] * ss3[0] + ss3[1] * ss3[1] + ss3[2] * ss3[2]);
    const double ssx = snrm * ss3[0], ssy = snrm * ss3[1], ssz = snrm * ss3[2];
    const double u0 = w[N + c], v0 = w[2 * N + c], w0 = w[3 * N + c];
    const double u1 = w[N + cp], v1 = w[2 * N + cp], w1 = w[3 * N + cp];
    double corr;
    corr = g[0] * ssx + g[1] * ssy + g[2] * ssz - (u1 - u0) * snrm;
    const double u_x = g[0] - corr * ssx, u_y = g[1] - corr * ssy, u_z = g[2] - corr * ssz;
    corr = g[3] * ssx + g[4] * ssy + g[5] * ssz - (v1 - v0) * snrm;
    const double v_x = g[3] - corr * ssx, v_y = g[4] - corr * ssy, v_z = g[5] - corr * ssz;
    corr = g[6] * ssx + g[7] * ssy + g[8] * ssz - (w1 - w0) * snrm;
    const double w_x = g[6] - corr * ssx, w_y = g[7] - corr * ssy, w_z = g[8] - corr * ssz;
    corr = g[9] * ssx + g[10] * ssy + g[11] * ssz + (b.aa[cp] - b.aa[c]) * snrm;
    double q_x = g[9] - corr * ssx, q_y = g[10] - corr * ssy, q_z = g[11] - corr * ssz;
    const double fracDiv = (2.0 * (1.0 / 3.0)) * (u_x + v_y + w_z);
    const double tauxxS = 2.0 * u_x - fracDiv, tauyyS = 2.0 * v_y - fracDiv, tauzzS = 2.0 * w_z - fracDiv;
    const double tauxyS = u_y + v_x, tauxzS = u_z + w_x, tauyzS = v_z + w_y;
    q_x = heatCoef * q_x; q_y = heatCoef * q_y; q_z = heatCoef * q_z;
    double tauxx = mut * tauxxS, tauyy = mut * tauyyS, tauzz = mut * tauzzS;
    double tauxy = mut * tauxyS, tauxz = mut * tauxzS, tauyz = mut * tauyzS;
    if (c_prm.useQCR) {
        double den = sqrt(u_x * u_x + u_y * u_y + u_z * u_z + v_x * v_x + v_y * v_y + v_z * v_z + w_x * w_x +
                          w_y * w_y + w_z * w_z);
        den = dmax_(den, 1.e-10);
        const double fact = mue * 0.3 / den;
        const double Wxy = u_y - v_x, Wxz = u_z - w_x, Wyz = v_z - w_y;
        const double Wyx = -Wxy, Wzx = -Wxz, Wzy = -Wyz;
        tauxx -= fact * (Wxy * tauxyS + Wxz * tauxzS) * 2.0;
        tauyy -= fact * (Wyx * tauxyS + Wyz * tauyzS) * 2.0;
        tauzz -= fact * (Wzx * tauxzS + Wzy * tauyzS) * 2.0;
        tauxy -= fact * (Wxy * tauyyS + Wxz * tauyzS + Wyx * tauxxS + Wyz * tauxzS);
        tauxz -= fact * (Wxy * tauyzS + Wxz * tauzzS + Wzx * tauxxS + Wzy * tauxyS);
        tauyz -= fact * (Wyx * tauxzS + Wyz * tauzzS + Wzx * tauxyS + Wzy * tauyyS);
    }
    const double ubar = 0.5 * (u0 + u1), vbar = 0.5 * (v0 + v1), wbar = 0.5 * (w0 + w1);
    const double s1 = s[c], s2 = s[N + c], s3 = s[2 * N + c];
    f[0] = tauxx * s1 + tauxy * s2 + tauxz * s3;
    f[1] = tauxy * s1 + tauyy * s2 + tauyz * s3;
    f[2] = tauxz * s1 + tauyz * s2 + tauzz * s3;
    f[3] = (ubar * tauxx + vbar * tauxy + wbar * tauxz) * s1 + (ubar * tauxy + vbar * tauyy + wbar * tauyz) * s2 +
           (ubar * tauxz + vbar * tauyz + wbar * tauzz) * s3 - q_x * s1 - q_y * s2 - q_z * s3;
}

// ---------------------------------------------------------------------------
// SA residual pieces for one cell.
// saAdvection, one direction: blockette.F90:1415-1560 (k), j, i analogous
__device__ __forceinline__ double sa_adv_dir(const BlockDev& b, long long N, long long c, long long sd,
                                             const double* __restrict__ s, double voli2, double ux, double uy, double uz) {
    const double* nt = b.w + ITU1 * N;
    const double xa = (s[c] + s[c - sd]) * voli2;
    const double ya = (s[N + c] + s[N + c - sd]) * voli2;
    const double za = (s[2 * N + c] + s[2 * N + c - sd]) * voli2;
    const double uu = xa * ux + ya * uy + za * uz;
    double dwtx;
    if (uu > 0.0) {
        if (c_prm.secondOrdTurb) {
            const double dwtm1 = nt[c - sd] - nt[c - 2 * sd];
            const double dwt = nt[c] - nt[c - sd];
            const double dwtp1 = nt[c + sd] - nt[c];
            dwtx = dwt;
            if (dwt * dwtp1 > 0.0) dwtx = dwtx + 0.5 * ((fabs(dwt) < fabs(dwtp1)) ? dwt : dwtp1);
            if (dwt * dwtm1 > 0.0) dwtx = dwtx - 0.5 * ((fabs(dwt) < fabs(dwtm1)) ? dwt : dwtm1);
        } else {
            dwtx = nt[c] - nt[c - sd];
        }
    } else {
        if (c_prm.secondOrdTurb) {
            const double dwtm1 = nt[c] - nt[c - sd];
            const double dwt = nt[c + sd] - nt[c];
            const double dwtp1 = nt[c + 2 * sd] - nt[c + sd];
            dwtx = dwt;
            if (dwt * dwtp1 > 0.0) dwtx = dwtx - 0.5 * ((fabs(dwt) < fabs(dwtp1)) ? dwt : dwtp1);
            if (dwt * dwtm1 > 0.0) dwtx = dwtx + 0.5 * ((fabs(dwt) < fabs(dwtm1)) ? dwt : dwtm1);
        } else {
            dwtx = nt[c + sd] - nt[c];
        }
    }
    return uu * dwtx;
}

// saViscous, one direction: blockette.F90:1197-1258 (k), j, i analogous.
// returns c1m*nu(m) - c10*nu + c1p*nu(p) added left-to-right onto `acc`.
__device__ __forceinline__ double sa_visc_dir(const BlockDev& b, long long N, long long c, long long sd,
                                              const double* __restrict__ s, double acc) {
    const double* w = b.w;
    const double* vol = b.vol;
    const long long cm = c - sd, cp = c + sd;
    const double cb3Inv = 1.0 / c_prm.rsaCb3, cb2 = c_prm.rsaCb2;
    const double voli = 1.0 / vol[c];
    const double volmi = 2.0 / (vol[c] + vol[cm]);
    const double volpi = 2.0 / (vol[c] + vol[cp]);
    const double xm = s[cm] * volmi, ym = s[N + cm] * volmi, zm = s[2 * N + cm] * volmi;
    const double xp = s[c] * volpi, yp = s[N + c] * volpi, zp = s[2 * N + c] * volpi;
    const double xa = 0.5 * (s[c] + s[cm]) * voli;
    const double ya = 0.5 * (s[N + c] + s[N + cm]) * voli;
    const double za = 0.5 * (s[2 * N + c] + s[2 * N + cm]) * voli;
    const double ttm = xm * xa + ym * ya + zm * za;
    const double ttp = xp * xa + yp * ya + zp * za;
    const double nt0 = w[ITU1 * N + c], ntm = w[ITU1 * N + cm], ntp = w[ITU1 * N + cp];
    const double cnud = -cb2 * nt0 * cb3Inv;
    const double cam = ttm * cnud, cap = ttp * cnud;
    const double nutm = 0.5 * (ntm + nt0), nutp = 0.5 * (ntp + nt0);
    const double nu = b.rlv[c] / w[c];
    const double num = 0.5 * (b.rlv[cm] / w[cm] + nu);
    const double nup = 0.5 * (b.rlv[cp] / w[cp] + nu);
    const double cdm = (num + (1.0 + cb2) * nutm) * ttm * cb3Inv;
    const double cdp = (nup + (1.0 + cb2) * nutp) * ttp * cb3Inv;
    const double c1m = dmax_(cdm + cam, 0.0), c1p = dmax_(cdp + cap, 0.0);
    const double c10 = c1m + c1p;
    return acc + c1m * ntm - c10 * nt0 + c1p * ntp;
}

// saSource: blockette.F90:976-1168
__device__ __forceinline__ double sa_source(const BlockDev& b, const Dims& d, long long c) {
    const long long N = d.N;
    const double* w = b.w;
    double gv[3][3];
#pragma unroll
    for (int v = 0; v < 3; v++) {
        const double* q = w + (IVX + v) * N;
        const double qip = q[c + 1], qim = q[c - 1], qjp = q[c + d.sJ], qjm = q[c - d.sJ], qkp = q[c + d.sK], qkm = q[c - d.sK];
#pragma unroll
        for (int m = 0; m < 3; m++) {
            const long long o = m * N;
            gv[v][m] = qip * b.si[o + c] - qim * b.si[o + c - 1] + qjp * b.sj[o + c] - qjm * b.sj[o + c - d.sJ] +
                       qkp * b.sk[o + c] - qkm * b.sk[o + c - d.sK];
        }
    }
    const double fact = 0.25 / b.vol[c];
    const double sxx = 2.0 * fact * gv[0][0], syy = 2.0 * fact * gv[1][1], szz = 2.0 * fact * gv[2][2];
    const double sxy = fact * (gv[0][1] + gv[1][0]), sxz = fact * (gv[0][2] + gv[2][0]), syz = fact * (gv[1][2] + gv[2][1]);
    const double div2 = (2.0 * (1.0 / 3.0)) * ((sxx + syy + szz) * (sxx + syy + szz));
    const double strainMag2 = 2.0 * (sxy * sxy + sxz * sxz + syz * syz) + sxx * sxx + syy * syy + szz * szz;
    double sqrtProd;
    if (c_prm.turbProd == ADFB_PROD_STRAIN) {
        sqrtProd = sqrt(dmax_(2.0 * strainMag2 - div2, 1.e-25));
    } else {
        const double vortx = 2.0 * fact * (gv[2][1] - gv[1][2]);
        const double vorty = 2.0 * fact * (gv[0][2] - gv[2][0]);
        const double vortz = 2.0 * fact * (gv[1][0] - gv[0][1]);
        sqrtProd = sqrt(vortx * vortx + vorty * vorty + vortz * vortz);
    }
    const double cv13 = c_prm.rsaCv1 * c_prm.rsaCv1 * c_prm.rsaCv1;
    const double kar2Inv = 1.0 / (c_prm.rsaK * c_prm.rsaK);
    const double cw3 = c_prm.rsaCw3;
    const double cw36 = (cw3 * cw3 * cw3) * (cw3 * cw3 * cw3);
    const double nt = w[ITU1 * N + c];
    const double nu = b.rlv[c] / w[c];
    const double dw_ = b.d2Wall[c];
    const double dist2Inv = 1.0 / (dw_ * dw_);
    const double chi = nt / nu, chi2 = chi * chi, chi3 = chi * chi2;
    const double fv1 = chi3 / (chi3 + cv13);
    const double fv2 = 1.0 - chi / (1.0 + chi * fv1);
    double ft2 = 0.0;
    if (c_prm.useft2SA) ft2 = c_prm.rsaCt3 * exp(-c_prm.rsaCt4 * chi2);
    double sst = sqrtProd + nt * fv2 * kar2Inv * dist2Inv;
    if (c_prm.useRotationSA) sst = sst + c_prm.rsaCrot * dmin_(0.0, sqrt(2.0 * strainMag2));
    sst = dmax_(sst, 1.e-10);
    double rr = nt * kar2Inv * dist2Inv / sst;
    rr = dmin_(rr, 10.0);
    const double rr2 = rr * rr, rr6 = rr2 * rr2 * rr2;
    const double gg = rr + c_prm.rsaCw2 * (rr6 - rr);
    const double gg2 = gg * gg, gg6 = gg2 * gg2 * gg2;
    const double termFw = pow((1.0 + cw36) / (gg6 + cw36), 1.0 / 6.0);
    const double fwSa = gg * termFw;
    const double term1 = c_prm.rsaCb1 * (1.0 - ft2) * sqrtProd * (c_prm.approxSA ? 0.0 : 1.0);
    const double term2 = dist2Inv * (kar2Inv * c_prm.rsaCb1 * ((1.0 - ft2) * fv2 + ft2) - c_prm.rsaCw1 * fwSa);
    return (term1 + term2 * nt) * nt;
}

// ---------------------------------------------------------------------------
// k_resid: the fused residual for one owned cell.
//   flowRes / turbRes select rows; rFil and persistFw implement the RK
//   dissipation blending (fw = sfil*fw_old + ..., src/solver/residuals.F90:61-65,
//   fluxes.F90:1193); for blocketteRes rFil == 1 and fw is never stored.
template <bool VISCOUS>
__global__ void __launch_bounds__(128) k_resid(Dims d, BlockDev b, int flowRes, int turbRes, double rFil, int persistFw,
                                               int doVisc, int doDiss) {
    int i, j, k;
    cell_index(d, i, j, k, 2, 2, 2);
    if (i > d.il || j > d.jl || k > d.kl) return;
    const long long c = ADFB_IDX(i, j, k);
    const long long N = d.N;
    const double rblank = dmax_((double)b.iblank[c], 0.0);

    if (turbRes) {
        // order of accumulation: source, advection k/j/i, diffusion k/j/i (blockette.F90:623-627)
        double r = 0.0;
        r = r + sa_source(b, d, c);
        const double voli2 = 0.5 / b.vol[c];
        const double ux = b.w[N + c], uy = b.w[2 * N + c], uz = b.w[3 * N + c];
        r = r - sa_adv_dir(b, N, c, d.sK, b.sk, voli2, ux, uy, uz);
        r = r - sa_adv_dir(b, N, c, d.sJ, b.sj, voli2, ux, uy, uz);
        r = r - sa_adv_dir(b, N, c, 1, b.si, voli2, ux, uy, uz);
        r = sa_visc_dir(b, N, c, d.sK, b.sk, r);
        r = sa_visc_dir(b, N, c, d.sJ, b.sj, r);
        r = sa_visc_dir(b, N, c, 1, b.si, r);
        b.dw[ITU1 * N + c] = -b.volRef[c] * r * rblank;  // saResScale, :1872-1897
    }
    if (!flowRes) return;

    double dw[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
    double f[5];
    // central: i, j, k ; minus face first (dw(i+1) -= fs at loop i-1), then plus face
    central_face(b, N, c - 1, c, b.si, b.porI[c - 1], f);
#pragma unroll
    for (int l = 0; l < 5; l++) dw[l] -= f[l];
    central_face(b, N, c, c + 1, b.si, b.porI[c], f);
#pragma unroll
    for (int l = 0; l < 5; l++) dw[l] += f[l];
    central_face(b, N, c - d.sJ, c, b.sj, b.porJ[c - d.sJ], f);
#pragma unroll
    for (int l = 0; l < 5; l++) dw[l] -= f[l];
    central_face(b, N, c, c + d.sJ, b.sj, b.porJ[c], f);
#pragma unroll
    for (int l = 0; l < 5; l++) dw[l] += f[l];
    central_face(b, N, c - d.sK, c, b.sk, b.porK[c - d.sK], f);
#pragma unroll
    for (int l = 0; l < 5; l++) dw[l] -= f[l];
    central_face(b, N, c, c + d.sK, b.sk, b.porK[c], f);
#pragma unroll
    for (int l = 0; l < 5; l++) dw[l] += f[l];

    // dissipation (scalar JST): fw = sfil*fw ; i, j, k ; fw(i+1) += fs ; fw(i) -= fs
    double fw[5];
    const double sfil = 1.0 - rFil;
#pragma unroll
    for (int l = 0; l < 5; l++) fw[l] = persistFw ? sfil * b.fw[l * N + c] : 0.0;
    if (doDiss && c_prm.spaceDiscr == ADFB_DISS_SCALAR) {
        const double fis2 = rFil * c_prm.vis2, fis4 = rFil * c_prm.vis4;
        jst_face(b, N, c - 1, 1, b.radI, b.dss, b.porI[c - 1], fis2, fis4, f);
#pragma unroll
        for (int l = 0; l < 5; l++) fw[l] += f[l];
        jst_face(b, N, c, 1, b.radI, b.dss, b.porI[c], fis2, fis4, f);
#pragma unroll
        for (int l = 0; l < 5; l++) fw[l] -= f[l];
        jst_face(b, N, c - d.sJ, d.sJ, b.radJ, b.dss + N, b.porJ[c - d.sJ], fis2, fis4, f);
#pragma unroll
        for (int l = 0; l < 5; l++) fw[l] += f[l];
        jst_face(b, N, c, d.sJ, b.radJ, b.dss + N, b.porJ[c], fis2, fis4, f);
#pragma unroll
        for (int l = 0; l < 5; l++) fw[l] -= f[l];
        jst_face(b, N, c - d.sK, d.sK, b.radK, b.dss + 2 * N, b.porK[c - d.sK], fis2, fis4, f);
#pragma unroll
        for (int l = 0; l < 5; l++) fw[l] += f[l];
        jst_face(b, N, c, d.sK, b.radK, b.dss + 2 * N, b.porK[c], fis2, fis4, f);
#pragma unroll
        for (int l = 0; l < 5; l++) fw[l] -= f[l];
    }

    if (VISCOUS && doVisc) {
        // viscous: k, j, i ; fw(k+1) += f at loop k-1 (minus face), fw(k) -= f (plus face)
        double v[4];
        visc_face(b, N, c - d.sK, d.sK, 1, d.sJ, b.sk, b.porK[c - d.sK], rFil, v);
#pragma unroll
        for (int l = 0; l < 4; l++) fw[l + 1] += v[l];
        visc_face(b, N, c, d.sK, 1, d.sJ, b.sk, b.porK[c], rFil, v);
#pragma unroll
        for (int l = 0; l < 4; l++) fw[l + 1] -= v[l];
        visc_face(b, N, c - d.sJ, d.sJ, 1, d.sK, b.sj, b.porJ[c - d.sJ], rFil, v);
#pragma unroll
        for (int l = 0; l < 4; l++) fw[l + 1] += v[l];
        visc_face(b, N, c, d.sJ, 1, d.sK, b.sj, b.porJ[c], rFil, v);
#pragma unroll
        for (int l = 0; l < 4; l++) fw[l + 1] -= v[l];
        visc_face(b, N, c - 1, 1, d.sJ, d.sK, b.si, b.porI[c - 1], rFil, v);
#pragma unroll
        for (int l = 0; l < 4; l++) fw[l + 1] += v[l];
        visc_face(b, N, c, 1, d.sJ, d.sK, b.si, b.porI[c], rFil, v);
#pragma unroll
        for (int l = 0; l < 4; l++) fw[l + 1] -= v[l];
    }
    if (persistFw) {
#pragma unroll
        for (int l = 0; l < 5; l++) b.fw[l * N + c] = fw[l];
    }
    // sumDwandFw, blockette.F90:6839-6864
#pragma unroll
    for (int l = 0; l < 5; l++) b.dw[l * N + c] = (dw[l] + fw[l]) * rblank;
}

}  // namespace

// ---------------------------------------------------------------------------
// host-side launcher (called from adfb_api.cu)
// doRad: 1 = recompute spectral radii + dtl (blockette order), 0 = keep them (block/smoother path)
static int launch_residual_core(const Dims& d, const BlockDev& b, const AdfbParams& prm, unsigned flags, double rFil,
                                int persistFw, int doRad, cudaStream_t stream) {
    const int flowRes = (flags & ADFB_RES_FLOW) != 0;
    const int turbRes = ((flags & ADFB_RES_TURB) != 0) && prm.equations == ADFB_RANS;
    const int updateDt = 1;  // blockette timeStep always computes dtl (blockette.F90:1929-1932)
    const bool viscous = prm.equations != ADFB_EULER;
    const int doDiss = fabs(rFil) >= 1.e-10;  // fluxes.F90:1082 early return
    const int doVisc = viscous && doDiss;
    dim3 tb(32, 4, 2);
    if (doRad || doDiss) {
        dim3 g((d.NI + tb.x - 1) / tb.x, (d.NJ + tb.y - 1) / tb.y, (d.NK + tb.z - 1) / tb.z);
        KT_BEGIN(K_PREP, stream);
        k_prep<<<g, tb, 0, stream>>>(d, b, updateDt, doRad);
        KT_END(K_PREP, stream);
    }
    if (flowRes && doDiss) {
        dim3 g((d.ie + tb.x - 1) / tb.x, (d.je + tb.y - 1) / tb.y, (d.ke + tb.z - 1) / tb.z);
        KT_BEGIN(K_NODAL, stream);
        k_nodal<<<g, tb, 0, stream>>>(d, b, doVisc);
        KT_END(K_NODAL, stream);
    }
    {
        dim3 tr(32, 4, 1);
        dim3 g((d.nx + tr.x - 1) / tr.x, (d.ny + tr.y - 1) / tr.y, (d.nz + tr.z - 1) / tr.z);
        KT_BEGIN(K_RESID, stream);
        if (viscous)
            k_resid<true><<<g, tr, 0, stream>>>(d, b, flowRes, turbRes, rFil, persistFw, doVisc, doDiss);
        else
            k_resid<false><<<g, tr, 0, stream>>>(d, b, flowRes, turbRes, rFil, persistFw, doVisc, doDiss);
        KT_END(K_RESID, stream);
    }
    return (int)cudaGetLastError();
}

