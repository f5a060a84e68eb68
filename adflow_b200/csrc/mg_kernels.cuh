// mg_kernels.cuh -- multigrid transfer operators of the smoother path (src/solver/multiGrid.F90)
//
//   k_mg_restrict     : transferToCoarseGrid :88-225 -- restricted residual -> wr, volume-weighted state -> w, p, rev of
//                       the coarse block, fused with computeEtotBlock / computeLamViscosity of the owned cells
//   k_mg_corner_rows  : setCornerRowHalos :1032-1357 (six ordered groups of halo copies; one CTA, barrier between groups)
//   k_mg_store_w1     : :270-288  w1 = w, p1 = p on 1:ie
//   k_mg_forcing      : :296-322  wr = fcoll*wr - dw, dw = fcoll*wr
//   k_mg_corrections  : transferToFineGrid :419-436  w := w - w1, w(irhoE) := p - p1 on 1:ie of the coarse block
//   k_mg_corr_halos   : setCorrectionsCoarseHalos :1359-1503, one launch per subface in BCData order
//   k_mg_prolong      : transferToFineGrid :468-552 -- trilinear interpolation (27/9/3/1 over 64) into dw, state update
//                       with the positivity clips, fused with etot / rlv / rev of the owned cells
//   k_w_roundtrip     : the conservative <-> primitive round trip that inviscidDissFluxScalarCoarse (fluxes.F90:5015-5024,
//                       5185-5200) leaves on w(1:ie): (rho*u)*(1/rho) and (rhoE + p) - p
//
// Tables are the reference's (coarseUtils.F90:254-420) with the Fortran index as the offset: mgIFine[i + (ie+1)*(m-1)].
#pragma once
#include "adfb_common.cuh"
#include "smoother_kernels.cuh"

struct MgTables {
    const int *fI, *fJ, *fK;        // coarse block: mg{I,J,K}Fine  (1:ie, 2)
    const double *wI, *wJ, *wK;     // coarse block: mg{I,J,K}Weight (2:il)
    const int *cI, *cJ, *cK;        // fine block:   mg{I,J,K}Coarse (2:il, 2)
};

namespace {

__device__ __forceinline__ double lam_visc(double p, double rho) {
    const double T = p / (c_prm.RGas * rho);
    return c_prm.muSuth * ((c_prm.TSuth + c_prm.SSuth) / (T + c_prm.SSuth)) * pow(T / c_prm.TSuth, 1.5);
}

__global__ void __launch_bounds__(128) k_mg_restrict(Dims d, BlockDev b, Dims df, BlockDev f, MgTables t) {
    ADFB_PDL_SYNC();
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 2;
    const int j = blockIdx.y * blockDim.y + threadIdx.y + 2;
    const int k = blockIdx.z * blockDim.z + threadIdx.z + 2;
    if (i > d.il || j > d.jl || k > d.kl) return;
    const long long N = d.N, NF = df.N, c = ADFB_IDX(i, j, k);
    const int ii = t.fI[i], ii1 = t.fI[i + d.ie + 1], jj = t.fJ[j], jj1 = t.fJ[j + d.je + 1], kk = t.fK[k], kk1 = t.fK[k + d.ke + 1];
    const double weigth = t.wK[k] * t.wJ[j] * t.wI[i];
#define FI_(a, b_, c_) ((long long)(a) + df.sJ * (long long)(b_) + df.sK * (long long)(c_))
    const long long a000 = FI_(ii, jj, kk), a100 = FI_(ii1, jj, kk), a010 = FI_(ii, jj1, kk), a110 = FI_(ii1, jj1, kk);
    const long long a001 = FI_(ii, jj, kk1), a101 = FI_(ii1, jj, kk1), a011 = FI_(ii, jj1, kk1), a111 = FI_(ii1, jj1, kk1);
#undef FI_
    const double* v = f.vol;
    const double v0 = v[a000], v1 = v[a010], v2 = v[a100], v3 = v[a110], v4 = v[a001], v5 = v[a011], v6 = v[a101], v7 = v[a111];
    double vola = v0 + v2 + v1 + v3 + v4 + v6 + v5 + v7;   // the reference's order: ii1 before jj1 for the volumes
    vola = 1.0 / vola;
#pragma unroll
    for (int l = 0; l < 5; l++) {
        const double* r = f.dw + l * NF;
        b.wr[l * N + c] = (r[a000] + r[a010] + r[a100] + r[a110] + r[a001] + r[a011] + r[a101] + r[a111]) * weigth * 1.0;
    }
    double s[4];
#pragma unroll
    for (int l = 0; l < 4; l++) {
        const double* q = f.w + l * NF;
        s[l] = (v0 * q[a000] + v1 * q[a010] + v2 * q[a100] + v3 * q[a110] + v4 * q[a001] + v5 * q[a011] + v6 * q[a101] + v7 * q[a111]) * vola;
        b.w[l * N + c] = s[l];
    }
    const double* q = f.p;
    const double p = (v0 * q[a000] + v1 * q[a010] + v2 * q[a100] + v3 * q[a110] + v4 * q[a001] + v5 * q[a011] + v6 * q[a101] + v7 * q[a111]) * vola;
    b.p[c] = p;
    q = f.rev;
    b.rev[c] = (v0 * q[a000] + v1 * q[a010] + v2 * q[a100] + v3 * q[a110] + v4 * q[a001] + v5 * q[a011] + v6 * q[a101] + v7 * q[a111]) * vola;
    b.w[4 * N + c] = (1.0 / (c_prm.gammaInf - 1.0)) * p + 0.5 * s[0] * (s[1] * s[1] + s[2] * s[2] + s[3] * s[3]);
    if (c_prm.equations != ADFB_EULER) b.rlv[c] = lam_visc(p, s[0]);
}

__device__ __forceinline__ void crh_copy(const BlockDev& b, long long N, long long dst, long long src) {
#pragma unroll
    for (int l = 0; l < 5; l++) b.w[l * N + dst] = b.w[l * N + src];
    b.p[dst] = b.p[src];
    if (c_prm.equations != ADFB_EULER) b.rlv[dst] = b.rlv[src];
    if (c_prm.equations == ADFB_RANS) b.rev[dst] = b.rev[src];
}
// one CTA; group g copies are independent of each other, groups are ordered
__global__ void __launch_bounds__(256) k_mg_corner_rows(Dims d, BlockDev b) {
    ADFB_PDL_SYNC();
    const long long N = d.N;
    const int t = threadIdx.x, nt = blockDim.x;
    auto mn = [](int a, int c) { return a < c ? a : c; };
    auto mx = [](int a, int c) { return a > c ? a : c; };
    // each work item: (line index, which of the 4 rows, low/high side)
    {   // i planes, k rows
        const int r[4] = {2, mn(3, d.jl), d.jl, mx(2, d.ny)};
        for (int q = t; q < (d.kl - 1) * 8; q += nt) {
            const int k = 2 + q / 8, m = (q % 8) / 2, hi = q % 2;
            crh_copy(b, N, ADFB_IDX(hi ? d.ie : 1, r[m], k), ADFB_IDX(hi ? d.il : 2, r[m], k));
        }
    }
    __syncthreads();
    {   // i planes, j rows
        const int r[4] = {2, mn(3, d.kl), d.kl, mx(2, d.nz)};
        for (int q = t; q < mx(d.ny - 2, 0) * 8; q += nt) {
            const int j = 3 + q / 8, m = (q % 8) / 2, hi = q % 2;
            crh_copy(b, N, ADFB_IDX(hi ? d.ie : 1, j, r[m]), ADFB_IDX(hi ? d.il : 2, j, r[m]));
        }
    }
    __syncthreads();
    {   // j planes, k lines
        const int r[4] = {2, mn(3, d.il), d.il, mx(2, d.nx)};
        for (int q = t; q < mx(d.nz - 2, 0) * 8; q += nt) {
            const int k = 3 + q / 8, m = (q % 8) / 2, hi = q % 2;
            crh_copy(b, N, ADFB_IDX(r[m], hi ? d.je : 1, k), ADFB_IDX(r[m], hi ? d.jl : 2, k));
        }
    }
    __syncthreads();
    {   // j planes, i lines
        const int r[4] = {2, mn(3, d.kl), d.kl, mx(2, d.nz)};
        for (int q = t; q < d.ie * 8; q += nt) {
            const int i = 1 + q / 8, m = (q % 8) / 2, hi = q % 2;
            crh_copy(b, N, ADFB_IDX(i, hi ? d.je : 1, r[m]), ADFB_IDX(i, hi ? d.jl : 2, r[m]));
        }
    }
    __syncthreads();
    {   // k planes, j lines
        const int r[4] = {2, mn(3, d.il), d.il, mx(2, d.nx)};
        for (int q = t; q < d.je * 8; q += nt) {
            const int j = 1 + q / 8, m = (q % 8) / 2, hi = q % 2;
            crh_copy(b, N, ADFB_IDX(r[m], j, hi ? d.ke : 1), ADFB_IDX(r[m], j, hi ? d.kl : 2));
        }
    }
    __syncthreads();
    {   // k planes, i lines
        const int r[4] = {2, mn(3, d.jl), d.jl, mx(2, d.ny)};
        for (int q = t; q < d.ie * 8; q += nt) {
            const int i = 1 + q / 8, m = (q % 8) / 2, hi = q % 2;
            crh_copy(b, N, ADFB_IDX(i, r[m], hi ? d.ke : 1), ADFB_IDX(i, r[m], hi ? d.kl : 2));
        }
    }
}

// mode 0: w1 = w, p1 = p (1:ie);  mode 1: corrections w -= w1, w(irhoE) = p - p1 (1:ie);
// mode 2: the w round trip of inviscidDissFluxScalarCoarse (1:ie);  mode 3: w(irhoE) = p (1:ie), the solution
// transfer of the full-multigrid start-up (transferToFineGrid(.false.), multiGrid.F90:455-470)
__global__ void __launch_bounds__(256) k_mg_cells1(Dims d, BlockDev b, int mode) {
    ADFB_PDL_SYNC();
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;
    const int j = blockIdx.y * blockDim.y + threadIdx.y + 1;
    const int k = blockIdx.z * blockDim.z + threadIdx.z + 1;
    if (i > d.ie || j > d.je || k > d.ke) return;
    const long long N = d.N, c = ADFB_IDX(i, j, k);
    if (mode == 0) {
#pragma unroll
        for (int l = 0; l < 5; l++) b.w1[l * N + c] = b.w[l * N + c];
        b.p1[c] = b.p[c];
    } else if (mode == 1) {
#pragma unroll
        for (int l = 0; l < 4; l++) b.w[l * N + c] = b.w[l * N + c] - b.w1[l * N + c];
        b.w[4 * N + c] = b.p[c] - b.p1[c];
    } else if (mode == 3) {
        b.w[4 * N + c] = b.p[c];
    } else {
        const double rho = b.w[c];
        const double rhoi = 1.0 / rho;
#pragma unroll
        for (int l = 1; l < 4; l++) {
            const double m = rho * b.w[l * N + c];
            b.w[l * N + c] = m * rhoi;
        }
        const double p = b.p[c];
        // (rhoE + p) - p: two separate roundings (no contraction possible, no fast-math)
        const double h = __dadd_rn(b.w[4 * N + c], p);
        b.w[4 * N + c] = __dsub_rn(h, p);
    }
}

__global__ void __launch_bounds__(256) k_mg_forcing(Dims d, BlockDev b, double fcoll) {
    ADFB_PDL_SYNC();
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 2;
    const int j = blockIdx.y * blockDim.y + threadIdx.y + 2;
    const int k = blockIdx.z * blockDim.z + threadIdx.z + 2;
    if (i > d.il || j > d.jl || k > d.kl) return;
    const long long N = d.N, c = ADFB_IDX(i, j, k);
#pragma unroll
    for (int l = 0; l < 5; l++) {
        const double tmp = fcoll * b.wr[l * N + c];
        b.wr[l * N + c] = tmp - b.dw[l * N + c];
        b.dw[l * N + c] = tmp;
    }
}

__global__ void __launch_bounds__(128) k_mg_corr_halos(Dims d, BlockDev b, FaceDev f, double fact, int nVarInt) {
    ADFB_PDL_SYNC();
    const int ia = blockIdx.x * blockDim.x + threadIdx.x + f.icBeg;
    const int jb = blockIdx.y * blockDim.y + threadIdx.y + f.jcBeg;
    if (ia > f.icEnd || jb > f.jcEnd) return;
    const long long N = d.N;
    const long long q = ia * f.sa + jb * f.sb;
    const long long c1 = f.off[1] + q, c2 = f.off[2] + q;
    double* w = b.w;
    if (f.bcType == ADFB_BC_SYMM) {
        const long long na = f.icEnd - f.icBeg + 1, nb = f.jcEnd - f.jcBeg + 1;
        const long long o = (ia - f.icBeg) + na * (jb - f.jcBeg);
        const double nnx = f.norm[o], nny = f.norm[o + na * nb], nnz = f.norm[o + 2 * na * nb];
        const double u = w[N + c2], v = w[2 * N + c2], ww = w[3 * N + c2];
        const double vn = 2.0 * (u * nnx + v * nny + ww * nnz);
        w[c1] = w[c2];
        w[N + c1] = u - vn * nnx;
        w[2 * N + c1] = v - vn * nny;
        w[3 * N + c1] = ww - vn * nnz;
        w[4 * N + c1] = w[4 * N + c2];
        for (int l = 5; l < nVarInt; l++) w[l * N + c1] = w[l * N + c2];
    } else {
        for (int l = 0; l < nVarInt; l++) w[l * N + c1] = fact * w[l * N + c2];
    }
}

__global__ void __launch_bounds__(128) k_mg_prolong(Dims d, BlockDev b, Dims dc, BlockDev cb, MgTables t, int nw) {
    ADFB_PDL_SYNC();
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 2;
    const int j = blockIdx.y * blockDim.y + threadIdx.y + 2;
    const int k = blockIdx.z * blockDim.z + threadIdx.z + 2;
    if (i > d.il || j > d.jl || k > d.kl) return;
    const long long N = d.N, NC = dc.N, c = ADFB_IDX(i, j, k);
    const int ii = t.cI[i], ii1 = t.cI[i + d.ie + 1], jj = t.cJ[j], jj1 = t.cJ[j + d.je + 1], kk = t.cK[k], kk1 = t.cK[k + d.ke + 1];
#define CI_(a, b_, c_) ((long long)(a) + dc.sJ * (long long)(b_) + dc.sK * (long long)(c_))
    const long long a000 = CI_(ii, jj, kk), a100 = CI_(ii1, jj, kk), a010 = CI_(ii, jj1, kk), a001 = CI_(ii, jj, kk1);
    const long long a110 = CI_(ii1, jj1, kk), a101 = CI_(ii1, jj, kk1), a011 = CI_(ii, jj1, kk1), a111 = CI_(ii1, jj1, kk1);
#undef CI_
    double dwv[5];
#pragma unroll
    for (int l = 0; l < 5; l++) {
        const double* ww = cb.w + l * NC;
        // the reference's expression, term by term (no contraction across the products: each is rounded)
        const double t0 = __dmul_rn(0.421875, ww[a000]);
        const double t1 = __dmul_rn(0.140625, (ww[a100] + ww[a010]) + ww[a001]);
        const double t2 = __dmul_rn(0.046875, (ww[a110] + ww[a101]) + ww[a011]);
        const double t3 = __dmul_rn(0.015625, ww[a111]);
        dwv[l] = __dadd_rn(__dadd_rn(__dadd_rn(t0, t1), t2), t3);
        b.dw[l * N + c] = dwv[l];
    }
    double rho = b.w[c] + dwv[0];
    const double u = b.w[N + c] + dwv[1], v = b.w[2 * N + c] + dwv[2], w = b.w[3 * N + c] + dwv[3];
    double p = b.p[c] + dwv[4];
    rho = dmax_(rho, 1.e-4 * c_prm.rhoInf);
    p = dmax_(p, 1.e-4 * c_prm.pInfCorr);
    b.w[c] = rho; b.w[N + c] = u; b.w[2 * N + c] = v; b.w[3 * N + c] = w; b.p[c] = p;
    b.w[4 * N + c] = (1.0 / (c_prm.gammaInf - 1.0)) * p + 0.5 * rho * (u * u + v * v + w * w);
    if (c_prm.equations == ADFB_EULER) return;
    const double rlv = lam_visc(p, rho);
    b.rlv[c] = rlv;
    if (c_prm.equations != ADFB_RANS || nw < 6 || b.coarse) return;
    const double rnuSA = b.w[5 * N + c] * rho;
    const double chi = rnuSA / rlv;
    const double chi3 = chi * chi * chi;
    const double cv13 = c_prm.rsaCv1 * c_prm.rsaCv1 * c_prm.rsaCv1;
    b.rev[c] = chi3 / (chi3 + cv13) * rnuSA;
}

// transferToFineGrid(corrections = .false.), multiGrid.F90:468-552: the SOLUTION of the coarse block (all nw variables, the
// pressure in the place of rho*E) interpolated to the owned cells of the fine block; p, then computeEtotBlock /
// computeLamViscosity / computeEddyViscosity of the owned cells fused in
__global__ void __launch_bounds__(128) k_mg_prolong_solution(Dims d, BlockDev b, Dims dc, BlockDev cb, MgTables t, int nw) {
    ADFB_PDL_SYNC();
    const int i = blockIdx.x * blockDim.x + threadIdx.x + 2;
    const int j = blockIdx.y * blockDim.y + threadIdx.y + 2;
    const int k = blockIdx.z * blockDim.z + threadIdx.z + 2;
    if (i > d.il || j > d.jl || k > d.kl) return;
    const long long N = d.N, NC = dc.N, c = ADFB_IDX(i, j, k);
    const int ii = t.cI[i], ii1 = t.cI[i + d.ie + 1], jj = t.cJ[j], jj1 = t.cJ[j + d.je + 1], kk = t.cK[k], kk1 = t.cK[k + d.ke + 1];
#define CI_(a, b_, c_) ((long long)(a) + dc.sJ * (long long)(b_) + dc.sK * (long long)(c_))
    const long long a000 = CI_(ii, jj, kk), a100 = CI_(ii1, jj, kk), a010 = CI_(ii, jj1, kk), a001 = CI_(ii, jj, kk1);
    const long long a110 = CI_(ii1, jj1, kk), a101 = CI_(ii1, jj, kk1), a011 = CI_(ii, jj1, kk1), a111 = CI_(ii1, jj1, kk1);
#undef CI_
    double v[6];
    for (int l = 0; l < nw && l < 6; l++) {
        const double* ww = cb.w + l * NC;
        const double t0 = __dmul_rn(0.421875, ww[a000]);
        const double t1 = __dmul_rn(0.140625, (ww[a100] + ww[a010]) + ww[a001]);
        const double t2 = __dmul_rn(0.046875, (ww[a110] + ww[a101]) + ww[a011]);
        const double t3 = __dmul_rn(0.015625, ww[a111]);
        v[l] = __dadd_rn(__dadd_rn(__dadd_rn(t0, t1), t2), t3);
    }
    const double rho = v[0], u = v[1], vv = v[2], w = v[3], p = v[4];
    b.w[c] = rho; b.w[N + c] = u; b.w[2 * N + c] = vv; b.w[3 * N + c] = w; b.p[c] = p;
    b.w[4 * N + c] = (1.0 / (c_prm.gammaInf - 1.0)) * p + 0.5 * rho * (u * u + vv * vv + w * w);
    if (nw >= 6) b.w[5 * N + c] = v[5];
    if (c_prm.equations == ADFB_EULER) return;
    const double rlv = lam_visc(p, rho);
    b.rlv[c] = rlv;
    if (c_prm.equations != ADFB_RANS || nw < 6 || b.coarse) return;
    const double rnuSA = v[5] * rho;
    const double chi = rnuSA / rlv;
    const double chi3 = chi * chi * chi;
    const double cv13 = c_prm.rsaCv1 * c_prm.rsaCv1 * c_prm.rsaCv1;
    b.rev[c] = chi3 / (chi3 + cv13) * rnuSA;
}

// extrapolateSolution + extrapolateViscosities, multiGrid.F90:656-823: constant extrapolation into the halos, i then j then k
// with the earlier directions' halos taken along -- every halo cell receives the owned cell with the clamped indices
__global__ void __launch_bounds__(256) k_mg_extrapolate(Dims d, BlockDev b, int nw) {
    ADFB_PDL_SYNC();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y * blockDim.y + threadIdx.y;
    const int k = blockIdx.z * blockDim.z + threadIdx.z;
    if (i > d.ib || j > d.jb || k > d.kb) return;
    const int ic = i < 2 ? 2 : (i > d.il ? d.il : i), jc = j < 2 ? 2 : (j > d.jl ? d.jl : j), kc = k < 2 ? 2 : (k > d.kl ? d.kl : k);
    if (ic == i && jc == j && kc == k) return;
    const long long N = d.N, c = ADFB_IDX(i, j, k), s = ADFB_IDX(ic, jc, kc);
    for (int l = 0; l < nw; l++) b.w[l * N + c] = b.w[l * N + s];
    b.p[c] = b.p[s];
    if (c_prm.equations == ADFB_EULER) return;
    b.rlv[c] = b.rlv[s];
    if (c_prm.equations == ADFB_RANS) b.rev[c] = b.rev[s];
}

}  // namespace
