// krylov_kernels.cuh -- vector kernels of the device GMRES (adfb_gmres_solve): the Krylov vectors of the NK / ANK
// solves stay on the GPU, only the (restart+1) projection coefficients per iteration cross to the host.
//   k_multidot   : partial sums of V_i . w for i = 0..nv-1 in ONE pass over w (classical Gram-Schmidt, the
//                  orthogonalisation the reference selects: KSPGMRESClassicalGramSchmidt without refinement,
//                  NKSolvers.F90:432,2036), deterministic two-pass reduction
//   k_multidot_final, k_axpy_many (w -= sum h_i V_i  /  u = sum y_i V_i), k_scale_to
#pragma once
#include "adfb_common.cuh"

#define ADFB_GMRES_MAXV 128
#define ADFB_GMRES_PARTS 256

namespace {

// part[p * nv + i] = sum over the slice of block p of V_i[q] * w[q]
__global__ void __launch_bounds__(256) k_multidot(const double* __restrict__ V, long long ld, int nv, const double* __restrict__ w, long long n,
                                                  double* __restrict__ part) {
    __shared__ double s[256];
    for (int i = 0; i < nv; i++) {
        const double* v = V + (long long)i * ld;
        double a = 0.0;
        for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (long long)gridDim.x * blockDim.x) a += v[q] * w[q];
        s[threadIdx.x] = a;
        __syncthreads();
        for (int st = 128; st > 0; st >>= 1) {
            if ((int)threadIdx.x < st) s[threadIdx.x] += s[threadIdx.x + st];
            __syncthreads();
        }
        if (threadIdx.x == 0) part[(long long)blockIdx.x * nv + i] = s[0];
        __syncthreads();
    }
}
// out[i] = sum_p part[p * nv + i]
__global__ void __launch_bounds__(256) k_multidot_final(const double* __restrict__ part, int nParts, int nv, double* __restrict__ out) {
    __shared__ double s[256];
    const int i = blockIdx.x;
    double a = 0.0;
    for (int p = threadIdx.x; p < nParts; p += 256) a += part[(long long)p * nv + i];
    s[threadIdx.x] = a;
    __syncthreads();
    for (int st = 128; st > 0; st >>= 1) {
        if ((int)threadIdx.x < st) s[threadIdx.x] += s[threadIdx.x + st];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[i] = s[0];
}
struct GmresCoef { double c[ADFB_GMRES_MAXV]; };
// y[q] = beta * y[q] + sum_i coef[i] * V_i[q]
__global__ void __launch_bounds__(256) k_axpy_many(const double* __restrict__ V, long long ld, int nv, GmresCoef coef, double beta,
                                                   double* __restrict__ y, long long n) {
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    double a = beta == 0.0 ? 0.0 : beta * y[q];
    for (int i = 0; i < nv; i++) a += coef.c[i] * V[(long long)i * ld + q];
    y[q] = a;
}
__global__ void __launch_bounds__(256) k_scale_to(const double* __restrict__ x, double a, double* __restrict__ y, long long n) {
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q < n) y[q] = a * x[q];
}

}  // namespace
