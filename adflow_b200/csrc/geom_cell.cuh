// geom_cell.cuh -- geometry-derived static arrays of one cell (shared by k_geom and the CPU emulation of the tile kernel)
#pragma once
#include "adfb_common.cuh"
#include <math.h>

// ---------------------------------------------------------------------------
// k_geom: geometry-derived static arrays, once per mesh (adfb_block_set_geometry).
//   ssum[dir] = s(c-sd) + s(c)          (timeStep sx/sy/sz, blockette.F90:1976-2006; saAdvection/saViscous xa)
//   sv[dir]   = 8-face normal sum of the dual face at cell layer c (allNodalGradients, :5247-5258)
//   ovol      = 1 / (8-cell volume sum) at node c (:5489-5492)
//   vn[dir]   = unit vector + inverse length between cell centres across face c (viscousFlux, :5638-5657)
// (plain __host__ __device__ function of the cell index: tests/emul runs it on the CPU to feed the tile-kernel emulation)
__host__ __device__ inline void geom_cell(const Dims& d, const BlockDev& b, int i, int j, int k) {
    if (i < 1 || i > d.ie || j < 1 || j > d.je || k < 1 || k > d.ke) return;
    const int N = (int)d.N, sJ = (int)d.sJ, sK = (int)d.sK;
    const int c = i + sJ * j + sK * k;
    const double *si = b.si, *sj = b.sj, *sk = b.sk;
#pragma unroll
    for (int m = 0; m < 3; m++) {
        b.ssum[m * N + c] = si[m * N + c - 1] + si[m * N + c];
        b.ssum[(3 + m) * N + c] = sj[m * N + c - sJ] + sj[m * N + c];
        b.ssum[(6 + m) * N + c] = sk[m * N + c - sK] + sk[m * N + c];
    }
    // dual-face sums; reference order: layer c-sd: (0, t1, t2, t1+t2), then layer c
    if (i <= d.il && j <= d.jl) {  // K sweep: i 1:il, j 1:jl, k 1:ke ; t1 = I, t2 = J
#pragma unroll
        for (int m = 0; m < 3; m++) {
            const double* s = sk + m * N;
            b.sv[(6 + m) * N + c] = s[c - sK] + s[c - sK + 1] + s[c - sK + sJ] + s[c - sK + 1 + sJ] + s[c] + s[c + 1] + s[c + sJ] + s[c + 1 + sJ];
        }
    }
    if (i <= d.il && k <= d.kl) {  // J sweep: t1 = I, t2 = K
#pragma unroll
        for (int m = 0; m < 3; m++) {
            const double* s = sj + m * N;
            b.sv[(3 + m) * N + c] = s[c - sJ] + s[c - sJ + 1] + s[c - sJ + sK] + s[c - sJ + 1 + sK] + s[c] + s[c + 1] + s[c + sK] + s[c + 1 + sK];
        }
    }
    if (j <= d.jl && k <= d.kl) {  // I sweep: t1 = J, t2 = K
#pragma unroll
        for (int m = 0; m < 3; m++) {
            const double* s = si + m * N;
            b.sv[m * N + c] = s[c - 1] + s[c - 1 + sJ] + s[c - 1 + sK] + s[c - 1 + sJ + sK] + s[c] + s[c + sJ] + s[c + sK] + s[c + sJ + sK];
        }
    }
    if (i <= d.il && j <= d.jl && k <= d.kl) {
        const double* vol = b.vol;
        b.ovol[c] = 1.0 / (vol[c] + vol[c + sK] + vol[c + 1] + vol[c + 1 + sK] + vol[c + sJ] + vol[c + sJ + sK] + vol[c + 1 + sJ] + vol[c + 1 + sJ + sK]);
        // face-normal unit vectors for the viscous gradient correction; node n = c
        const double* x = b.x;
        const int sd[3] = {1, sJ, sK}, t1[3] = {sJ, 1, 1}, t2[3] = {sK, sK, sJ};
#pragma unroll
        for (int dir = 0; dir < 3; dir++) {
            // faces exist for the two transverse indices >= 2
            const bool ok = (dir == 0) ? (j >= 2 && k >= 2) : (dir == 1) ? (i >= 2 && k >= 2) : (i >= 2 && j >= 2);
            if (!ok) continue;
            const int n = c, n1 = c - t1[dir] - t2[dir], n2 = c - t2[dir], n3 = c - t1[dir], s = sd[dir];
            double v[3];
#pragma unroll
            for (int m = 0; m < 3; m++) {
                const double* xm = x + m * N;
                v[m] = 0.125 * (xm[n1 + s] - xm[n1 - s] + xm[n3 + s] - xm[n3 - s] + xm[n2 + s] - xm[n2 - s] + xm[n + s] - xm[n - s]);
            }
            const double snrm = 1.0 / sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
            b.vn[(4 * dir + 0) * N + c] = snrm * v[0];
            b.vn[(4 * dir + 1) * N + c] = snrm * v[1];
            b.vn[(4 * dir + 2) * N + c] = snrm * v[2];
            b.vn[(4 * dir + 3) * N + c] = snrm;
        }
    }
}
