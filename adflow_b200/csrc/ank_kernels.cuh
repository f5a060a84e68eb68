// ank_kernels.cuh -- ANK pieces (module ANKSolver, src/NKSolver/NKSolvers.F90)
//
//   k_ank_tsblock : computeTimeStepBlock (:2116-2329) for every owned cell: nState x nState block (column-major) of the
//                   pseudo-time term, 'None' (stateToCons * dtInv), 'VLR' and 'Turkel' characteristic time stepping
//   k_ankvec      : setWANK (:2975-3011), setRVecANK / setRVec (:2895, :1262) + MatMultAdd(timeStepMat, inVec, rVec)
//                   (:2516), and the finite-difference quotient of the matrix-free product
//   k_ank_phys    : physicalityCheckANK (:3013-3210): per-cell ratios, clipping of too-limiting turbulence updates, MIN
#pragma once
#include "adfb_common.cuh"

namespace {

template <int N>
__device__ __forceinline__ void mm_(const double* a, const double* b, double* c) {   // c = a b (may alias)
    double t[N * N];
#pragma unroll
    for (int j = 0; j < N; j++)
#pragma unroll
        for (int i = 0; i < N; i++) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < N; k++) s += a[i + N * k] * b[k + N * j];
            t[i + N * j] = s;
        }
#pragma unroll
    for (int q = 0; q < N * N; q++) c[q] = t[q];
}
template <int N>
__device__ __forceinline__ void mmt_(const double* a, const double* b, double* c) {  // c = a b^T
    double t[N * N];
#pragma unroll
    for (int j = 0; j < N; j++)
#pragma unroll
        for (int i = 0; i < N; i++) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < N; k++) s += a[i + N * k] * b[j + N * k];
            t[i + N * j] = s;
        }
#pragma unroll
    for (int q = 0; q < N * N; q++) c[q] = t[q];
}

#define BK(r, c) blk[((r) - 1) + N * ((c) - 1)]
#define MT(m, r, c) m[((r) - 1) + N * ((c) - 1)]
template <int N>
__global__ void __launch_bounds__(64) k_ank_tsblock(Dims d, BlockDev b, AdfbAnkParams ank, double* __restrict__ out) {
    const long long nOwned = (long long)d.nx * d.ny * d.nz;
    const long long cell = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= nOwned) return;
    const int i = (int)(cell % d.nx) + 2, j = (int)((cell / d.nx) % d.ny) + 2, k = (int)(cell / ((long long)d.nx * d.ny)) + 2;
    const long long c = ADFB_IDX(i, j, k), NN = d.N;
    double blk[N * N], stateToCons[N * N];
#pragma unroll
    for (int q = 0; q < N * N; q++) { blk[q] = 0.0; stateToCons[q] = 0.0; }
    const double rho = b.w[c], velX = b.w[NN + c], velY = b.w[2 * NN + c], velZ = b.w[3 * NN + c];
    const double dtInv = 1.0 / (ank.cfl * b.dtl[c] * b.volRef[c]);
    MT(stateToCons, 1, 1) = 1.0;
    MT(stateToCons, 2, 1) = velX; MT(stateToCons, 2, 2) = rho;
    MT(stateToCons, 3, 1) = velY; MT(stateToCons, 3, 3) = rho;
    MT(stateToCons, 4, 1) = velZ; MT(stateToCons, 4, 4) = rho;
    MT(stateToCons, 5, 5) = 1.0;
    if (N == 6) MT(stateToCons, N, N) = c_prm.turbResScale / ank.turbCFLScale;
    double* o = out + cell * (N * N);
    if (ank.charTimeStepType == 0) {
#pragma unroll
        for (int q = 0; q < N * N; q++) o[q] = stateToCons[q] * dtInv;
        return;
    }
    double streamToCart[N * N], symmToCons[N * N], consToSymm[N * N];
#pragma unroll
    for (int q = 0; q < N * N; q++) { streamToCart[q] = 0.0; symmToCons[q] = 0.0; consToSymm[q] = 0.0; }
    if (N == 6) { BK(N, N) = 1.0; MT(streamToCart, N, N) = 1.0; MT(symmToCons, N, N) = 1.0; MT(consToSymm, N, N) = 1.0; }
    double aa = b.aa[c];
    if (c_prm.equations == ADFB_EULER) { aa = c_prm.gammaInf * b.p[c] / rho; b.aa[c] = aa; }
    const double speed = sqrt(velX * velX + velY * velY + velZ * velZ);
    const double sos = sqrt(aa);
    const double mach = speed / sos, machSqr = mach * mach, gm1 = c_prm.gammaInf - 1.0;
    MT(symmToCons, 1, 1) = rho / sos; MT(symmToCons, 1, 5) = -1.0 / aa;
    MT(symmToCons, 2, 1) = rho * velX / sos; MT(symmToCons, 2, 2) = rho; MT(symmToCons, 2, 5) = -velX / aa;
    MT(symmToCons, 3, 1) = rho * velY / sos; MT(symmToCons, 3, 3) = rho; MT(symmToCons, 3, 5) = -velY / aa;
    MT(symmToCons, 4, 1) = rho * velZ / sos; MT(symmToCons, 4, 4) = rho; MT(symmToCons, 4, 5) = -velZ / aa;
    MT(symmToCons, 5, 1) = rho * sos * (machSqr / 2 + 1 / gm1);
    MT(symmToCons, 5, 2) = rho * velX; MT(symmToCons, 5, 3) = rho * velY; MT(symmToCons, 5, 4) = rho * velZ;
    MT(symmToCons, 5, 5) = -machSqr / 2;
    MT(consToSymm, 1, 1) = gm1 / 2 * sos * machSqr / rho;
    MT(consToSymm, 1, 2) = -gm1 * velX / (rho * sos); MT(consToSymm, 1, 3) = -gm1 * velY / (rho * sos);
    MT(consToSymm, 1, 4) = -gm1 * velZ / (rho * sos); MT(consToSymm, 1, 5) = gm1 / (rho * sos);
    MT(consToSymm, 2, 1) = -velX / rho; MT(consToSymm, 2, 2) = 1.0 / rho;
    MT(consToSymm, 3, 1) = -velY / rho; MT(consToSymm, 3, 3) = 1.0 / rho;
    MT(consToSymm, 4, 1) = -velZ / rho; MT(consToSymm, 4, 4) = 1.0 / rho;
    MT(consToSymm, 5, 1) = aa * (gm1 / 2 * machSqr - 1.0);
    MT(consToSymm, 5, 2) = -gm1 * velX; MT(consToSymm, 5, 3) = -gm1 * velY; MT(consToSymm, 5, 4) = -gm1 * velZ;
    MT(consToSymm, 5, 5) = gm1;
    const double blend = ank.cfl / ank.cflLimit;
    if (ank.charTimeStepType == 1) {   // VLR
        const double m2t = dmax_(machSqr, 1e-4 * (ank.machInf * ank.machInf));
        double beta, tau;
        if (mach < 1.0) { beta = sqrt(1.0 - m2t); tau = beta; }
        else { beta = sqrt(m2t - 1.0); tau = sqrt(1.0 - 1.0 / m2t) + 1e-4; }
        BK(1, 1) = blend * (beta * beta + tau) / (m2t * tau) + (1.0 - blend) * 1.0;
        BK(1, 2) = blend * 1.0 / mach; BK(2, 1) = blend * 1.0 / mach; BK(2, 2) = 1.0;
        BK(3, 3) = blend * 1.0 / tau + (1.0 - blend) * 1.0;
        BK(4, 4) = blend * 1.0 / tau + (1.0 - blend) * 1.0;
        BK(5, 5) = 1.0;
        const double speedXY = sqrt(velX * velX + velY * velY);
        const double sinT = velY / speedXY, cosT = velX / speedXY, sinA = velZ / speed, cosA = speedXY / speed;
        MT(streamToCart, 1, 1) = 1.0;
        MT(streamToCart, 2, 2) = cosA * cosT; MT(streamToCart, 2, 3) = -sinT; MT(streamToCart, 2, 4) = -sinA * cosT;
        MT(streamToCart, 3, 2) = cosA * sinT; MT(streamToCart, 3, 3) = cosT; MT(streamToCart, 3, 4) = -sinA * sinT;
        MT(streamToCart, 4, 2) = sinA; MT(streamToCart, 4, 4) = cosA;
        MT(streamToCart, 5, 5) = 1.0;
        mm_<N>(streamToCart, blk, blk);
        mmt_<N>(blk, streamToCart, blk);
    } else {   // Turkel
        const double m2t = dmin_(1.0, dmax_(machSqr, 1e-4 * (ank.machInf * ank.machInf)));
        const double q2 = m2t * m2t, q4 = q2 * q2, q8 = q4 * q4;
        const double alpha = 1.0 - q8 * q2;
        BK(1, 1) = blend * 1.0 / m2t + (1.0 - blend) * 1.0;
        BK(2, 1) = blend * alpha * velX / sos / m2t;
        BK(3, 1) = blend * alpha * velY / sos / m2t;
        BK(4, 1) = blend * alpha * velZ / sos / m2t;
        BK(2, 2) = 1.0; BK(3, 3) = 1.0; BK(4, 4) = 1.0; BK(5, 5) = 1.0;
    }
    mm_<N>(symmToCons, blk, blk);
    mm_<N>(blk, consToSymm, blk);
    mm_<N>(blk, stateToCons, blk);
#pragma unroll
    for (int q = 0; q < N * N; q++) o[q] = blk[q] * dtInv;
}
#undef BK
#undef MT

// vectors of ns entries per owned cell (ns = 5: flow variables, ns = nw: coupled)
//  mode 0: setWANK                       w(1:ns) <- vec
//  mode 1: perturbed setWANK             w(1:ns) <- base + h*vec;  pert <- base + h*vec (kept for the time-step term)
//  mode 2: F = setRVec(ANK) + T*in       out <- dw/volRef (* turbResScale on the turbulence row) + sum_c T(l,c)*in(c)
//  mode 3: (F - base) / h
//  mode 4: out <- T*in only (the block-diagonal time-step matrix as a linear operator)
__global__ void __launch_bounds__(256) k_ankvec(Dims d, BlockDev b, int ns, const double* __restrict__ vec, const double* __restrict__ base,
                                                double* __restrict__ out, double* __restrict__ pert, const double* __restrict__ T, double h,
                                                int mode) {
    const long long nOwned = (long long)d.nx * d.ny * d.nz;
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nOwned * ns) return;
    const int l = (int)(q % ns);
    const long long cell = q / ns;
    const int i = (int)(cell % d.nx) + 2, j = (int)((cell / d.nx) % d.ny) + 2, k = (int)(cell / ((long long)d.nx * d.ny)) + 2;
    const long long c = ADFB_IDX(i, j, k);
    if (mode <= 1) {
        const double v = mode == 0 ? vec[q] : base[q] + h * vec[q];
        if (mode == 1) pert[q] = v;
        b.w[l * d.N + c] = v;
        return;
    }
    const double ovv = 1.0 / b.volRef[c];
    double r = b.dw[l * d.N + c] * ovv;
    if (l >= 5) r = b.dw[l * d.N + c] * ovv * c_prm.turbResScale;
    if (mode == 4) r = 0.0;
    const double* Tc = T + cell * (ns * ns);
    const double* in = vec + cell * ns;
    double a = 0.0;
    for (int m = 0; m < ns; m++) a += Tc[l + ns * m] * in[m];
    r = r + a;
    out[q] = (mode == 2 || mode == 4) ? r : (r - base[q]) / h;
}

// one thread per cell; part[blockIdx.x] = min over the block; deltaW clipped in place
__global__ void __launch_bounds__(256) k_ank_phys(long long nCells, int ns, int coupled, AdfbAnkParams ank, const double* __restrict__ wv,
                                                  double* __restrict__ dv, double lambda0, double* __restrict__ part) {
    __shared__ double s[256];
    double lam = lambda0;
    for (long long cell = (long long)blockIdx.x * blockDim.x + threadIdx.x; cell < nCells; cell += (long long)gridDim.x * blockDim.x) {
        const long long ii = cell * ns;
        double ratio = fabs(wv[ii] / (dv[ii] + 1.e-25)) * ank.physLSTol;
        lam = dmin_(lam, ratio);
        ratio = fabs(wv[ii + 4] / (dv[ii + 4] + 1.e-25)) * ank.physLSTol;
        lam = dmin_(lam, ratio);
        if (coupled) {
            ratio = (wv[ii + 5] / (dv[ii + 5] + 1.e-25)) * ank.physLSTolTurb;
            if (ratio < ank.stepFactor * ank.stepMin) {
                if (ratio > 0.0) dv[ii + 5] = wv[ii + 5] * ank.physLSTolTurb;
                ratio = 1.0;
            }
            lam = dmin_(lam, ratio);
        }
        if (lam != lam) lam = 0.0;   // myisnan(lambdaL) -> 0 (min() would drop a NaN silently)
    }
    s[threadIdx.x] = lam;
    __syncthreads();
    for (int st = 128; st > 0; st >>= 1) {
        if (threadIdx.x < st) s[threadIdx.x] = dmin_(s[threadIdx.x], s[threadIdx.x + st]);
        __syncthreads();
    }
    if (threadIdx.x == 0) part[blockIdx.x] = s[0];
}
__global__ void k_min_final(double* part, int nPart) {
    __shared__ double s[256];
    double v = part[0];
    for (int q = threadIdx.x; q < nPart; q += blockDim.x) v = dmin_(v, part[q]);
    s[threadIdx.x] = v;
    __syncthreads();
    for (int st = 128; st > 0; st >>= 1) {
        if (threadIdx.x < st) s[threadIdx.x] = dmin_(s[threadIdx.x], s[threadIdx.x + st]);
        __syncthreads();
    }
    if (threadIdx.x == 0) part[nPart] = s[0];
}


// turbulence KSP of the decoupled ANK (one turbulence variable per owned cell):
//   mode 0 : setWANK(inVec, nt1, nt2) (NKSolvers.F90:2975-3011): w(itu1) of the owned cells <- vec
//   mode 1 : the same with vec = base + h * a (matrix-free product); the perturbed vector is kept in pert
//   mode 2 : setRVecANKTurb (:2935-2973) + the time-stepping term of FormFunction_mf_turb (:2540-2612):
//            out = dw(itu1) / volRef * turbResScale + vec / (ANK_CFL dtl volRef) * turbResScale / ANK_turbCFLScale
//   mode 3 : (that - base) / h
__global__ void __launch_bounds__(256) k_ankvec_turb(Dims d, BlockDev b, AdfbAnkParams ank, const double* __restrict__ dtl, const double* __restrict__ vec,
                                                     const double* __restrict__ base, double* __restrict__ out, double* __restrict__ pert,
                                                     double h, int mode) {
    const long long nOwned = (long long)d.nx * d.ny * d.nz;
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nOwned) return;
    const int i = (int)(q % d.nx) + 2, j = (int)((q / d.nx) % d.ny) + 2, k = (int)(q / ((long long)d.nx * d.ny)) + 2;
    const long long c = ADFB_IDX(i, j, k);
    if (mode <= 1) {
        const double v = mode == 0 ? vec[q] : base[q] + h * vec[q];
        if (mode == 1) pert[q] = v;
        b.w[5 * d.N + c] = v;
        return;
    }
    const double ovv = 1.0 / b.volRef[c];
    double r = b.dw[5 * d.N + c] * ovv * c_prm.turbResScale;
    const double dtinv = 1.0 / (ank.cfl * dtl[c] * b.volRef[c]);
    r = r + vec[q] * dtinv * c_prm.turbResScale / ank.turbCFLScale;
    out[q] = mode == 2 ? r : (r - base[q]) / h;
}
// physicalityCheckANKTurb (:3212-3335): clip of too-limiting updates + MIN over the cells
__global__ void __launch_bounds__(256) k_ank_phys_turb(long long nCells, AdfbAnkParams ank, const double* __restrict__ wv, double* __restrict__ dv,
                                                       double lambda0, double* __restrict__ part) {
    __shared__ double s[256];
    double lam = lambda0;
    for (long long ii = (long long)blockIdx.x * blockDim.x + threadIdx.x; ii < nCells; ii += (long long)gridDim.x * blockDim.x) {
        double ratio = (wv[ii] / (dv[ii] + 1.e-25)) * ank.physLSTolTurb;
        if (ratio < ank.stepFactor * ank.stepMin) {
            if (ratio > 0.0) dv[ii] = wv[ii] * ank.physLSTolTurb;
            ratio = 1.0;
        }
        lam = dmin_(lam, ratio);
        if (lam != lam) lam = 0.0;
    }
    s[threadIdx.x] = lam;
    __syncthreads();
    for (int st = 128; st > 0; st >>= 1) {
        if (threadIdx.x < st) s[threadIdx.x] = dmin_(s[threadIdx.x], s[threadIdx.x + st]);
        __syncthreads();
    }
    if (threadIdx.x == 0) part[blockIdx.x] = s[0];
}

}  // namespace
