// Run-phase glue around the RRTMG calls, on the device (sm_100a): what the two GEOS drivers do on the
// host before and after rrtmg_lw / rrtmg_sw, as four kernels, so that a caller hands over the native
// GEOS state (levels top-down, Pa, kg/kg) once and gets the native flux arrays back.
//
// Restates
//   GEOSirrad_GridComp/GEOS_IrradGridComp.F90  LW_Driver  :3237-3371 -> irrad_prepare_kernel
//                                                          :3486-3533 -> irrad_finish_kernel
//   GEOSsolar_GridComp/GEOS_SolarGridComp.F90  SORADCORE  :6113-6223 -> solar_prepare_kernel
//                                                          :6395-6447 -> solar_finish_kernel
// One thread per column; every array is (column, level) with the column fastest, so a warp reads and
// writes 256 contiguous bytes per level.  The vertical flip is an index reversal inside the thread;
// TLEV and the layer heights are carried down the column in registers.  Same expression order as the
// Fortran, no contraction (--fmad=false): the prepared inputs are bit-identical to the oracle's.
#pragma once
#include "common.cuh"
#include "../../include/rrtmgx.h"

namespace rrtmgx {

__device__ __forceinline__ double clampd(double x, double lo, double hi) { return fmin(fmax(x, lo), hi); }
// radius limits imposed before the call: IRR:3277-3296 (LW), SOL:6140-6170 (SW)
__device__ __forceinline__ double reliq_limit(double r, int liqflg, bool sw) {
    if (liqflg == 0) return sw ? clampd(r, 10.0, 30.0) : clampd(r, 5.0, 10.0);
    if (liqflg == 1) return clampd(r, 2.5, 60.0);
    return r;
}
__device__ __forceinline__ double reice_limit(double r, int iceflg) {
    if (iceflg == 0) return clampd(r, 10.0, 30.0);
    if (iceflg == 1) return clampd(r, 13.0, 130.0);
    if (iceflg == 2) return clampd(r, 5.0, 131.0);
    if (iceflg == 3) return clampd(r, 5.0, 140.0);
    if (iceflg == 4) return clampd(r * 2., 1.0, 200.0);
    return r;
}
__device__ __forceinline__ double nonneg(double x) { return x < 0. ? 0. : x; }

// Interface temperature TLEV(K), K = 1..LM+1, of the GEOS (top-down) column whose PLE values sit at
// ple[0..LM] and layer temperatures at t[0..LM-1] (stride ld), with DP(K) = ple[K] - ple[K-1]
// (IRR:3255-3262, SOL:6173-6176); `tbot` is TLEV(LM+1): T2M for the LW driver, TS for the SW one.
__device__ __forceinline__ double glue_tlev(const double *__restrict__ ple, const double *__restrict__ t, size_t ld,
                                            int LM, int K, double tbot) {
    if (K == LM + 1) return tbot;
    if (K == 1) K = 2;   // model top: TLEV(1) = TLEV(2)
    const double dpk = ple[ld * K] - ple[ld * (K - 1)];
    const double dpm = ple[ld * (K - 1)] - ple[ld * (K - 2)];
    return (t[ld * (K - 2)] * dpk + t[ld * (K - 1)] * dpm) / (dpm + dpk);
}

// `S` holds the native arrays with leading dimension lds, first column col0; `L` the rrtmg_lw
// arguments of this chunk (leading dimension nc, layer 1 at the surface).
__global__ void __launch_bounds__(128)
irrad_prepare_kernel(int nc, int lds, int col0, RrtmgxIrradArgs S, RrtmgxLwArgs L) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nc) return;
    const int LM = S.lm;
    const size_t s0 = (size_t)col0 + c, ld = (size_t)lds;
    auto NA = [&](const double *x, int k) { return x[s0 + ld * k]; };   // native, 0-based level
    auto W2 = [&](const double *x, int k) -> double & { return const_cast<double *>(x)[c + (size_t)nc * k]; };
    const double wq = S.airmw / S.h2omw, wo3 = S.airmw / S.o3mw;
    const double *ple = S.ple + s0, *t = S.t + s0;
    W2(L.tsfc, 0) = S.ts[s0];                                                     // :3249-3251
    const double em = S.emis[s0], t2m = S.t2m[s0];
    for (int b = 0; b < 16; ++b) W2(L.emis, b) = em;
    W2(L.alat, 0) = S.lats[s0];
    double zm = 0., pl_prev = 0.;
    for (int K = 1; K <= LM; ++K) {                                               // :3265-3337
        const int LV = LM - K + 1, k = K - 1, lv = LV - 1;
        const double xx = 1.02 * 100 * (ple[ld * LV] - ple[ld * (LV - 1)]);      // DP(LV)
        W2(L.clwp, k) = xx * NA(S.qliq, lv);
        W2(L.ciwp, k) = xx * NA(S.qice, lv);
        W2(L.rel, k) = reliq_limit(NA(S.rliq, lv), S.liqflg, false);
        W2(L.rei, k) = reice_limit(NA(S.rice, lv), S.iceflg);
        const double plev = ple[ld * LV] / 100.;                                  // PLE_R(K-1)
        const double tlev = glue_tlev(ple, t, ld, LM, LV + 1, t2m);               // TLEV_R(K-1)
        W2(L.plev, K - 1) = plev;
        W2(L.tlev, K - 1) = tlev;
        const double pl = NA(S.pl, lv) / 100.;
        W2(L.play, k) = pl;
        W2(L.tlay, k) = t[ld * lv];
        const double q = NA(S.q, lv);
        W2(L.h2ovmr, k) = nonneg(q / (1. - q) * wq);                              // clean-up of negatives :3360-3370
        W2(L.o3vmr, k) = nonneg(NA(S.o3, lv) * wo3);
        W2(L.ch4vmr, k) = nonneg(NA(S.ch4, lv));
        W2(L.n2ovmr, k) = nonneg(NA(S.n2o, lv));
        W2(L.co2vmr, k) = nonneg(S.co2 ? NA(S.co2, lv) : S.co2_fixed);
        W2(L.o2vmr, k) = nonneg(S.o2);
        W2(L.ccl4vmr, k) = nonneg(S.ccl4);
        W2(L.cfc11vmr, k) = nonneg(NA(S.cfc11, lv));
        W2(L.cfc12vmr, k) = nonneg(NA(S.cfc12, lv));
        W2(L.cfc22vmr, k) = nonneg(NA(S.hcfc22, lv));
        W2(L.cldf, k) = nonneg(NA(S.fcld, lv));
        for (int b = 0; b < 16; ++b)                                              // absorption optical depth, :3336
            W2(L.tauaer, k + LM * b) = S.taua ? fmax(NA(S.taua, lv + LM * b) - NA(S.ssaa, lv + LM * b), 0.) : 0.;
        // layer mid-point height, :3350-3355: the jump from layer K-1 to K is centred on level K-1
        if (K >= 2) zm = zm + S.rgas * tlev / S.grav * (pl_prev - pl) / plev;
        W2(L.zm, k) = zm;
        pl_prev = pl;
    }
    W2(L.plev, LM) = ple[0] / 100.;                                               // :3341-3342
    W2(L.tlev, LM) = glue_tlev(ple, t, ld, LM, 1, t2m);
}

__global__ void __launch_bounds__(128)
irrad_finish_kernel(int nc, int lds, int col0, RrtmgxIrradArgs S, RrtmgxLwArgs L, int band_mask) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nc) return;
    const int LM = S.lm;
    const size_t s0 = (size_t)col0 + c, ld = (size_t)lds;
    auto R2 = [&](const double *x, int k) { return x[c + (size_t)nc * k]; };
    auto OUT = [&](double *x, int k) -> double & { return x[s0 + ld * k]; };
    const double ng = (double)NGPTLW;
    if (S.cldtt) S.cldtt[s0] = 1.0 - L.clearCounts[c] / ng;                       // :3494-3505
    if (S.cldhi) S.cldhi[s0] = 1.0 - L.clearCounts[c + (size_t)nc] / ng;
    if (S.cldmd) S.cldmd[s0] = 1.0 - L.clearCounts[c + (size_t)nc * 2] / ng;
    if (S.cldlo) S.cldlo[s0] = 1.0 - L.clearCounts[c + (size_t)nc * 3] / ng;
    for (int K = 0; K <= LM; ++K) {                                               // upward negative, :3508-3516
        const int lv = LM - K;
        OUT(S.flxu, K) = -R2(L.uflx, lv);
        OUT(S.flxd, K) = R2(L.dflx, lv);
        OUT(S.flcu, K) = -R2(L.uflxc, lv);
        OUT(S.flcd, K) = R2(L.dflxc, lv);
        OUT(S.dfdts, K) = -R2(L.duflx_dTs, lv);
        OUT(S.dfdtsc, K) = -R2(L.duflxc_dTs, lv);
    }
    S.sfcem[s0] = -(R2(L.uflx, 0) - R2(L.dflx, 0) * (1. - S.emis[s0]));            // :3521
    if (S.olrb)                                                                    // band OLR, :3536-3547
        for (int b = 0; b < 16; ++b)
            if (band_mask & (1 << b)) {
                S.olrb[(size_t)16 * s0 + b] = L.olrb[(size_t)16 * c + b];
                if (S.dolrb_dts) S.dolrb_dts[(size_t)16 * s0 + b] = L.dolrb_dTs[(size_t)16 * c + b];
            }
}

// GEOS_IrradGridComp.F90 Update :3861, :3929-3990 (USE_RRTMG): one thread per (column, level)
__global__ void __launch_bounds__(256)
irrad_update_kernel(RrtmgxIrradUpdateArgs U) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int K = blockIdx.y;
    if (c >= U.ncol) return;
    const size_t i = (size_t)K * U.ncol + c;
    const double delt = U.tsinst[c] - U.ts_int[c];
    const double flx_int = U.flxd_int[i] + U.flxu_int[i];   // :3604
    const double flc_int = U.flcd_int[i] + U.flcu_int[i];   // :3606
    if (U.flx) U.flx[i] = flx_int + U.dfdts[i] * delt;
    if (U.flc) U.flc[i] = flc_int + U.dfdtsc[i] * delt;
    if (U.flxu) U.flxu[i] = U.flxu_int[i] + U.dfdts[i] * delt;
    if (U.flcu) U.flcu[i] = U.flcu_int[i] + U.dfdtsc[i] * delt;
    if (U.flxd) U.flxd[i] = U.flxd_int[i];
    if (U.flcd) U.flcd[i] = U.flcd_int[i];
    if (K == 0) {
        if (U.olr) U.olr[c] = -(flx_int + U.dfdts[i] * delt);
        if (U.olc) U.olc[c] = -(flc_int + U.dfdtsc[i] * delt);
    }
    if (K == U.lm) {
        if (U.sfcem) U.sfcem[c] = U.sfcem_int[c] - U.dfdts[i] * delt;
        if (U.lws) U.lws[c] = flx_int + U.sfcem_int[c];
        if (U.lcs) U.lcs[c] = flc_int + U.sfcem_int[c];
        if (U.flns) U.flns[c] = flx_int + U.dfdts[i] * delt;
        if (U.flnsc) U.flnsc[c] = flc_int + U.dfdtsc[i] * delt;
    }
}

__global__ void __launch_bounds__(128)
solar_prepare_kernel(int nc, int lds, int col0, RrtmgxSolarArgs S, RrtmgxSwArgs L, const int *__restrict__ lit) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nc) return;
    const int LM = S.lm;
    // lit: the native column of packed column c (RRTMGX_LIT_ONLY, PackIt SOL:7753-7773); else columns col0 + c
    const size_t s0 = lit ? (size_t)lit[c] : (size_t)col0 + c, ld = (size_t)lds;
    auto NA = [&](const double *x, int k) { return x[s0 + ld * k]; };
    auto W2 = [&](const double *x, int k) -> double & { return const_cast<double *>(x)[c + (size_t)nc * k]; };
    const double wq = S.airmw / S.h2omw, wo3 = S.airmw / S.o3mw;
    const double *ple = S.ple + s0, *t = S.t + s0;   // PLE(:,1:LM+1) at ple[0..LM]
    const double ts = S.ts[s0];
    W2(L.coszen, 0) = S.zt[s0];
    W2(L.alat, 0) = S.lats[s0];
    W2(L.asdir, 0) = S.albvr[s0]; W2(L.asdif, 0) = S.albvf[s0];                   // SOL:6345
    W2(L.aldir, 0) = S.albnr[s0]; W2(L.aldif, 0) = S.albnf[s0];
    double zm = 0., pl_prev = 0.;
    for (int K = 1; K <= LM; ++K) {
        const int LV = LM - K + 1, k = K - 1, lv = LV - 1;
        const double dpr = ple[ld * LV] - ple[ld * (LV - 1)];                     // DPR(LV), :6133
        W2(L.ciwp, k) = (1.02 * 100 * dpr) * NA(S.qice, lv);                      // :6136-6137
        W2(L.clwp, k) = (1.02 * 100 * dpr) * NA(S.qliq, lv);
        W2(L.rei, k) = reice_limit(NA(S.rice, lv), S.iceflg);                     // :6140-6170
        W2(L.rel, k) = reliq_limit(NA(S.rliq, lv), S.liqflg, true);
        const double plev = ple[ld * LV] / 100.;                                  // PLE_R(K), :6180
        W2(L.plev, K - 1) = plev;
        const double pl = NA(S.pl, lv) / 100.;                                    // :6183-6198
        W2(L.play, k) = pl;
        W2(L.tlay, k) = t[ld * lv];
        const double q = NA(S.q, lv);
        W2(L.h2ovmr, k) = nonneg(q / (1. - q) * wq);                              // clean-up :6201-6206
        W2(L.o3vmr, k) = nonneg(NA(S.o3, lv) * wo3);
        W2(L.ch4vmr, k) = nonneg(NA(S.ch4, lv));
        W2(L.co2vmr, k) = nonneg(S.co2);
        W2(L.o2vmr, k) = nonneg(S.o2);
        W2(L.cld, k) = nonneg(NA(S.cl, lv));
        // ZL_R(K) uses TLEV_R(K) = TLEV(LM+2-K) and PLE_R(K), the interface below layer K, :6212-6218
        if (K >= 2) zm = zm + S.rgas * glue_tlev(ple, t, ld, LM, LM + 2 - K, ts) / S.grav * (pl_prev - pl) / plev;
        W2(L.zm, k) = zm;
        pl_prev = pl;
        for (int b = 0; b < 14; ++b) {                                            // :6116-6126, :6221-6223
            double ta = 0., ss = 0., as = 0.;
            if (S.taua) {
                ta = NA(S.taua, lv + LM * b); ss = NA(S.ssaa, lv + LM * b); as = NA(S.asya, lv + LM * b);
                if (ta > 0. && ss > 0.) { as = as / ss; ss = ss / ta; }
                else { ta = 0.; ss = 0.; as = 0.; }
            }
            W2(L.tauaer, k + LM * b) = ta; W2(L.ssaaer, k + LM * b) = ss; W2(L.asmaer, k + LM * b) = as;
        }
    }
    W2(L.plev, LM) = ple[0] / 100.;
}

__global__ void __launch_bounds__(128)
solar_finish_kernel(int nc, int lds, int col0, RrtmgxSolarArgs S, RrtmgxSwArgs L, const int *__restrict__ lit) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nc) return;
    const int LM = S.lm;
    const size_t s0 = lit ? (size_t)lit[c] : (size_t)col0 + c, ld = (size_t)lds;   // UnPackIt SOL:7776-7790
    auto R2 = [&](const double *x, int k) { return x[c + (size_t)nc * k]; };
    auto OUT = [&](double *x, int k) -> double & { return x[s0 + ld * k]; };
    for (int K = 0; K <= LM; ++K) {                                               // unflip :6395-6398, fluxes :6441-6444
        const int lv = LM - K;
        const double u = R2(L.swuflx, lv), d = R2(L.swdflx, lv), uc = R2(L.swuflxc, lv), dc = R2(L.swdflxc, lv);
        OUT(S.fsw, K) = d - u;
        OUT(S.fsc, K) = dc - uc;
        OUT(S.fswu, K) = u;
        OUT(S.fscu, K) = uc;
    }
    const double ng = (double)NGPTSW;
    double *cld[4] = {S.cldts, S.cldhs, S.cldms, S.cldls};                        // :6407-6410
    for (int n = 0; n < 4; ++n)
        if (cld[n]) cld[n][s0] = 1. - L.clearCounts[c + (size_t)nc * n] / ng;
    double *cot[4] = {S.cottp, S.cothp, S.cotmp, S.cotlp};                        // :6417-6439
    const double *cn[4] = {L.cotntp, L.cotnhp, L.cotnmp, L.cotnlp}, *cd[4] = {L.cotdtp, L.cotdhp, L.cotdmp, L.cotdlp};
    for (int n = 0; n < 4; ++n)
        if (cot[n]) cot[n][s0] = (cn[n][c] > 0. && cd[n][c] > 0.) ? cn[n][c] / cd[n][c] : S.undef;
    // surface diagnostics pass through unchanged
    double *so[7] = {S.nirr, S.nirf, S.parr, S.parf, S.uvrr, S.uvrf, nullptr};
    const double *si[6] = {L.nirr, L.nirf, L.parr, L.parf, L.uvrr, L.uvrf};
    for (int n = 0; n < 6; ++n)
        if (so[n]) so[n][s0] = si[n][c];
    if (S.fswband)
        for (int b = 0; b < 14; ++b) S.fswband[s0 + ld * b] = L.fswband[c + (size_t)nc * b];
}


// RRTMGX_LIT_ONLY: what UnPackIt leaves in the night columns (SOL:7791-7794, DEFAULT of the internal specs):
// a dark sun in the fluxes and surface components, MAPL_UNDEF in the cloud fractions and optical thicknesses.
__global__ void __launch_bounds__(128)
solar_night_kernel(int n, int lds, RrtmgxSolarArgs S) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    if (S.zt[c] > 0.) return;   // daytime = ZTH > 0, SOL:3686
    const size_t ld = (size_t)lds;
    for (int K = 0; K <= S.lm; ++K) {
        S.fsw[c + ld * K] = 0.; S.fsc[c + ld * K] = 0.; S.fswu[c + ld * K] = 0.; S.fscu[c + ld * K] = 0.;
    }
    double *so[6] = {S.nirr, S.nirf, S.parr, S.parf, S.uvrr, S.uvrf};
    for (int k = 0; k < 6; ++k)
        if (so[k]) so[k][c] = 0.;
    if (S.fswband)
        for (int b = 0; b < 14; ++b) S.fswband[c + ld * b] = 0.;
    double *un[8] = {S.cldts, S.cldhs, S.cldms, S.cldls, S.cottp, S.cothp, S.cotmp, S.cotlp};
    for (int k = 0; k < 8; ++k)
        if (un[k]) un[k][c] = S.undef;
}

}  // namespace rrtmgx
