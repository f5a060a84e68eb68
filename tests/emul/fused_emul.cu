// fused_emul.cu -- CPU emulation of the tile kernel k_flowres (adflow_b200/csrc/fused_kernels.cuh).
// TEST INFRASTRUCTURE ONLY: runs the kernel's own per-thread phase functions (ft_nodal, ft_faces, ft_div: plain
// __host__ __device__ code) thread by thread, phase by phase, with the shared-memory tiles in host memory, so that the
// tile logic (index maps, halo handling, k marching, flux exchange) can be checked against the oracle without a GPU.
// Built by tests/test_fused_emul.py with nvcc (host code only is executed).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../adflow_b200/csrc/adfb_common.cuh"
#include "../../adflow_b200/csrc/fused_kernels.cuh"

#include "../../adflow_b200/csrc/geom_cell.cuh"
extern "C" {

struct EmulArrays {
    double *w, *p, *rlv, *rev, *x, *si, *sj, *sk, *vol, *aa, *ss, *radI, *radJ, *radK, *dw, *fw;
    double *ssum, *sv, *ovol, *vn;   // outputs of the geometry pass (caller allocates: 9N, 9N, N, 12N)
    int8_t *porI, *porJ, *porK;
    int32_t* iblank;
};

// the tile of the host chooser for a block (for the tests): out = TX, TY, kChunk, nT
int emul_choose(int nx, int ny, int nz, int tma, int nSM, int* out) {
    const Dims d = make_dims(nx, ny, nz);
    const FTile t = ftile_choose(d, tma != 0, nSM);
    out[0] = t.TX; out[1] = t.TY; out[2] = t.kChunk; out[3] = t.nT;
    return ftile_fits(t) ? 0 : 1;
}

int emul_flowres(int nx, int ny, int nz, const AdfbParams* prm, const EmulArrays* a, int TX, int TY, int kChunk, double rFil, int doDiss,
                 int merged, int persistFw) {
    const Dims d = make_dims(nx, ny, nz);
    BlockDev b;
    memset(&b, 0, sizeof b);
    b.w = a->w; b.p = a->p; b.rlv = a->rlv; b.rev = a->rev; b.x = a->x; b.si = a->si; b.sj = a->sj; b.sk = a->sk; b.vol = a->vol;
    b.aa = a->aa; b.ss = a->ss; b.radI = a->radI; b.radJ = a->radJ; b.radK = a->radK; b.dw = a->dw; b.fw = a->fw;
    b.ssum = a->ssum; b.sv = a->sv; b.ovol = a->ovol; b.vn = a->vn;
    b.porI = a->porI; b.porJ = a->porJ; b.porK = a->porK; b.iblank = a->iblank;
    for (int k = 0; k <= d.kb; k++)
        for (int j = 0; j <= d.jb; j++)
            for (int i = 0; i <= d.ib; i++) geom_cell(d, b, i, j, k);
    const AdfbParams& P = *prm;
    const bool viscous = P.equations != ADFB_EULER;
    // TX == 0: the tile the library's host chooser picks for this block (ftile_choose with the TMA constraint: odd TX), on kChunk SMs
    FTile t = TX > 0 ? ftile_make(TX, TY, kChunk, false) : ftile_choose(d, true, kChunk > 0 ? kChunk : 148);
    t.useTma = 0;
    if (!ftile_fits(t)) return 2;
    if (TX == 0 && (!(t.TX & 1) || t.kChunk < 1)) return 3;
    const int nti = (d.nx + t.TX - 2) / (t.TX - 1), ntj = (d.ny + t.TY - 2) / (t.TY - 1), nkc = (d.nz + t.kChunk - 1) / t.kChunk;
    std::vector<double> smem(FT_SMEM_DOUBLES + (size_t)(FT_NFLUX_SPLIT - FT_NFLUX) * FT_S0);
    FSmem sm;
    sm.ring = smem.data();
    sm.G = sm.ring + (size_t)FT_NSLOT * FV_NUM * FT_S2;
    sm.EE = sm.G + (size_t)FT_GP * FT_S0;
    sm.FX = sm.EE + (size_t)FT_GP * FT_S0;
    const bool visc = viscous && doDiss;
    std::vector<FRegs> regs(t.nT);
    std::vector<FCtx> ctx(t.nT);
    std::vector<FStep> steps(t.nT);
    // every thread of the CTA runs the per-thread code between two synchronisation points, then the next stretch
#define ALL_THREADS(stmt) for (int tid = 0; tid < t.nT; tid++) { FCtx& x = ctx[tid]; FRegs& r = regs[tid]; FStep& st = steps[tid]; (void)x; (void)r; (void)st; stmt; }
#define DISPATCH(fn, ...)                                                         \
    do {                                                                          \
        if (viscous) { if (merged) fn<true, true>(__VA_ARGS__); else fn<true, false>(__VA_ARGS__); } \
        else { if (merged) fn<false, true>(__VA_ARGS__); else fn<false, false>(__VA_ARGS__); }       \
    } while (0)
    for (int bz = 0; bz < nkc; bz++)
        for (int by = 0; by < ntj; by++)
            for (int bx = 0; bx < nti; bx++) {
                // poison the shared arrays: anything read before it is written shows up as NaN
                for (double& v : smem) v = nan("");
                const int ka = 2 + bz * t.kChunk;
                const int kb = (ka + t.kChunk - 1 < d.kl) ? ka + t.kChunk - 1 : d.kl;
                const int gi0 = bx * (t.TX - 1), gj0 = by * (t.TY - 1);
                auto load_plane = [&](int kk) {
                    double* slot = sm.ring + (size_t)(kk % FT_NSLOT) * FV_NUM * FT_S2;
                    for (int e = 0; e < t.PX * t.PY; e++) {
                        const int py = e / t.PX, px = e - py * t.PX;
                        const int gi = gi0 + px, gj = gj0 + py;
                        const bool valid = gi <= d.ib && gj <= d.jb;
                        const long long go = valid ? ((long long)gi + d.sJ * gj + d.sK * kk) : 0;
                        for (int v = 0; v < FV_NUM; v++) {
                            if (!ft_var_used(v, viscous, doDiss)) continue;
                            slot[v * FT_S2 + e] = valid ? ft_var_ptr(d, b, v)[go] : 0.0;
                        }
                    }
                };
                for (int tid = 0; tid < t.nT; tid++) {
                    ctx[tid] = ft_ctx(d, t, tid, bx, by);
                    memset(&steps[tid], 0, sizeof(FStep));
                    ft_prologue_regs(P, d, b, ctx[tid], ka - 1, regs[tid], doDiss, viscous);
                    if ((FT_EARLY || (FT_AHEAD & (4 | 16))) && visc) ft_load_nodal(d, b, ctx[tid], ka - 1, true, steps[tid].gn);
                }
                load_plane(ka - 1); load_plane(ka);
                if (FT_NSLOT >= 3) load_plane(ka + 1);
                for (int k = ka - 1; k <= kb; k++) {
                    const double* A = sm.ring + (size_t)(k % FT_NSLOT) * FV_NUM * FT_S2;
                    const double* B = sm.ring + (size_t)((k + 1) % FT_NSLOT) * FV_NUM * FT_S2;
                    const bool doIJ = k >= ka;
                    ALL_THREADS(DISPATCH(ft_step_a, d, b, t, x, k, kb, A, B, sm, r, st, doDiss, doIJ));
                    ALL_THREADS(DISPATCH(ft_step_b, P, d, b, t, x, k, kb, A, B, sm, r, st, rFil, doDiss, doIJ, 2));
                    // (barrier) the slot of plane k is free now
                    if (k + FT_NSLOT <= kb + 1) {
                        double* slot = sm.ring + (size_t)(k % FT_NSLOT) * FV_NUM * FT_S2;
                        for (int q = 0; q < FV_NUM * FT_S2; q++) slot[q] = nan("");
                        load_plane(k + FT_NSLOT);
                    }
                    ALL_THREADS(if (doIJ) { if (merged) ft_div<true>(d, b, t, x, k, sm, r, st, rFil, persistFw, MffdEpi{nullptr, 0}, 1.0);
                                            else ft_div<false>(d, b, t, x, k, sm, r, st, rFil, persistFw, MffdEpi{nullptr, 0}, 1.0); }
                                for (int l = 0; l < 10; l++) r.kprev[l] = st.kp[l]);
                }
            }
    return 0;
}

}  // extern "C"
