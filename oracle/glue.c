/* oracle/glue.c -- CPU restatement of the Run-phase glue around the RRTMG calls (test
 * infrastructure only; see oracle.h).  Follows, statement by statement,
 *   GEOSirrad_GridComp/GEOS_IrradGridComp.F90  LW_Driver  :3237-3371 (flip, units, TLEV, ZM, clean-up)
 *                                                          :3486-3533 (unflip, sign, SFCEM, cloud fractions)
 *   GEOSsolar_GridComp/GEOS_SolarGridComp.F90  SORADCORE  :6113-6212 (aerosol normalisation, flip, units, ZL)
 *                                                          :6395-6447 (unflip, cloud fractions, COT ratios, FSW)
 * with `real` promoted to 8 bytes.  GEOS arrays are (ncol,LM) with level 1 at the model top:
 * x(i,k) -> x[i + ncol*(k-1)]; PLE is (ncol,0:LM) in the LW driver and (ncol,LM+1) in the SW one,
 * the same LM+1 values in memory. */
#include "internal.h"

#define G2(x, i, k) (x)[(size_t)(i) + (size_t)ncol * (size_t)(k)]              /* 0-based level k */
#define G3(x, i, k, b, nk) (x)[(size_t)(i) + (size_t)ncol * ((size_t)(k) + (size_t)(nk) * (size_t)(b))]

static double clamp(double x, double lo, double hi) { return f_min(f_max(x, lo), hi); }

/* radius limits imposed before the call: IRR:3277-3296 (LW), SOL:6140-6170 (SW) */
static double reliq_limit(double r, int liqflg, int sw) {
    if (liqflg == 0) return sw ? clamp(r, 10.0, 30.0) : clamp(r, 5.0, 10.0);
    if (liqflg == 1) return clamp(r, 2.5, 60.0);
    return r;
}
static double reice_limit(double r, int iceflg) {
    if (iceflg == 0) return clamp(r, 10.0, 30.0);
    if (iceflg == 1) return clamp(r, 13.0, 130.0);
    if (iceflg == 2) return clamp(r, 5.0, 131.0);
    if (iceflg == 3) return clamp(r, 5.0, 140.0);
    if (iceflg == 4) return clamp(r * 2., 1.0, 200.0);
    return r;
}

int oracle_irrad_prepare(const OracleIrradState *s, OracleLwInputs *o) {
    const int ncol = s->ncol, LM = s->lm;
    if (ncol <= 0 || LM < 2) return -4;
    const double wq = s->airmw / s->h2omw, wo3 = s->airmw / s->o3mw;
#pragma omp parallel
    {
        double *DP = (double *)__builtin_alloca(sizeof(double) * (size_t)(LM + 2));
        double *TLEV = (double *)__builtin_alloca(sizeof(double) * (size_t)(LM + 2));
#pragma omp for schedule(static)
        for (int ij = 0; ij < ncol; ++ij) {
            o->tsfc[ij] = s->ts[ij];                                    /* :3249 */
            for (int b = 0; b < 16; ++b) o->emis[ij + (size_t)ncol * b] = s->emis[ij];
            o->alat[ij] = s->lats[ij];
            /* level temperature in model ordering, :3255-3262 (PLE(0:LM), TLEV(1:LM+1)) */
            DP[1] = G2(s->ple, ij, 1) - G2(s->ple, ij, 0);
            for (int K = 2; K <= LM; ++K) {
                DP[K] = G2(s->ple, ij, K) - G2(s->ple, ij, K - 1);
                TLEV[K] = (G2(s->t, ij, K - 2) * DP[K] + G2(s->t, ij, K - 1) * DP[K - 1]) / (DP[K - 1] + DP[K]);
            }
            TLEV[LM + 1] = s->t2m[ij];
            TLEV[1] = TLEV[2];
            for (int K = 1; K <= LM; ++K) {                             /* :3265-3337 */
                const int LV = LM - K + 1;
                const int k = K - 1, lv = LV - 1;                       /* 0-based */
                const double xx = 1.02 * 100 * DP[LV];
                G2(o->clwp, ij, k) = xx * G2(s->qliq, ij, lv);
                G2(o->ciwp, ij, k) = xx * G2(s->qice, ij, lv);
                G2(o->rel, ij, k) = reliq_limit(G2(s->rliq, ij, lv), s->liqflg, 0);
                G2(o->rei, ij, k) = reice_limit(G2(s->rice, ij, lv), s->iceflg);
                G2(o->plev, ij, K - 1) = G2(s->ple, ij, LV) / 100.;
                G2(o->tlev, ij, K - 1) = TLEV[LV + 1];
                G2(o->play, ij, k) = G2(s->pl, ij, lv) / 100.;
                G2(o->tlay, ij, k) = G2(s->t, ij, lv);
                const double q = G2(s->q, ij, lv);
                G2(o->h2ovmr, ij, k) = q / (1. - q) * wq;
                G2(o->o3vmr, ij, k) = G2(s->o3, ij, lv) * wo3;
                G2(o->ch4vmr, ij, k) = G2(s->ch4, ij, lv);
                G2(o->n2ovmr, ij, k) = G2(s->n2o, ij, lv);
                G2(o->co2vmr, ij, k) = s->co2 ? G2(s->co2, ij, lv) : s->co2_fixed;
                G2(o->o2vmr, ij, k) = s->o2;
                G2(o->ccl4vmr, ij, k) = s->ccl4;
                G2(o->cfc11vmr, ij, k) = G2(s->cfc11, ij, lv);
                G2(o->cfc12vmr, ij, k) = G2(s->cfc12, ij, lv);
                G2(o->cfc22vmr, ij, k) = G2(s->hcfc22, ij, lv);
                G2(o->cldf, ij, k) = G2(s->fcld, ij, lv);
                for (int b = 0; b < 16; ++b)                            /* absorption optical depth, :3336 */
                    G3(o->tauaer, ij, k, b, LM) =
                        s->taua ? f_max(G3(s->taua, ij, lv, b, LM) - G3(s->ssaa, ij, lv, b, LM), 0.) : 0.;
            }
            G2(o->plev, ij, LM) = G2(s->ple, ij, 0) / 100.;             /* :3341-3342 */
            G2(o->tlev, ij, LM) = TLEV[1];
            G2(o->zm, ij, 0) = 0.;                                      /* :3350-3355 */
            for (int K = 2; K <= LM; ++K)
                G2(o->zm, ij, K - 1) = G2(o->zm, ij, K - 2) + s->rgas * G2(o->tlev, ij, K - 1) / s->grav *
                                                               (G2(o->play, ij, K - 2) - G2(o->play, ij, K - 1)) /
                                                               G2(o->plev, ij, K - 1);
            for (int k = 0; k < LM; ++k) {                              /* clean up negatives, :3360-3370 */
                double *v[] = {&G2(o->h2ovmr, ij, k), &G2(o->o3vmr, ij, k), &G2(o->ch4vmr, ij, k),
                               &G2(o->n2ovmr, ij, k), &G2(o->co2vmr, ij, k), &G2(o->o2vmr, ij, k),
                               &G2(o->ccl4vmr, ij, k), &G2(o->cfc11vmr, ij, k), &G2(o->cfc12vmr, ij, k),
                               &G2(o->cfc22vmr, ij, k), &G2(o->cldf, ij, k)};
                for (size_t n = 0; n < sizeof v / sizeof v[0]; ++n)
                    if (*v[n] < 0.) *v[n] = 0.;
            }
        }
    }
    o->cloudMH = LM - s->lcldmh + 1;                                    /* :3239-3240 */
    o->cloudLM = LM - s->lcldlm + 1;
    return 0;
}

int oracle_irrad_finish(int ncol, int LM, const double *emis, const int *clearCounts, const double *uflx,
                        const double *dflx, const double *uflxc, const double *dflxc, const double *duflx_dTs,
                        const double *duflxc_dTs, OracleIrradFluxes *f) {
    for (int ij = 0; ij < ncol; ++ij) {
        /* super-layer clear counts -> cloud fractions, :3494-3505 */
        if (f->cldtt) f->cldtt[ij] = 1.0 - clearCounts[ij] / (double)NGPTLW;
        if (f->cldhi) f->cldhi[ij] = 1.0 - clearCounts[ij + (size_t)ncol] / (double)NGPTLW;
        if (f->cldmd) f->cldmd[ij] = 1.0 - clearCounts[ij + (size_t)ncol * 2] / (double)NGPTLW;
        if (f->cldlo) f->cldlo[ij] = 1.0 - clearCounts[ij + (size_t)ncol * 3] / (double)NGPTLW;
        for (int K = 0; K <= LM; ++K) {                                 /* upward negative, :3508-3516 */
            const int lv = LM - K;                                      /* LV = LM-K+1, 0-based */
            G2(f->flxu, ij, K) = -G2(uflx, ij, lv);
            G2(f->flxd, ij, K) = G2(dflx, ij, lv);
            G2(f->flcu, ij, K) = -G2(uflxc, ij, lv);
            G2(f->flcd, ij, K) = G2(dflxc, ij, lv);
            G2(f->dfdts, ij, K) = -G2(duflx_dTs, ij, lv);
            G2(f->dfdtsc, ij, K) = -G2(duflxc_dTs, ij, lv);
        }
        f->sfcem[ij] = -(uflx[ij] - dflx[ij] * (1. - emis[ij]));        /* :3521 */
    }
    return 0;
}

int oracle_solar_prepare(const OracleSolarState *s, OracleSwInputs *o) {
    const int ncol = s->ncol, LM = s->lm;
    if (ncol <= 0 || LM < 2) return -4;
    const double wq = s->airmw / s->h2omw, wo3 = s->airmw / s->o3mw;
#pragma omp parallel
    {
        double *DPR = (double *)__builtin_alloca(sizeof(double) * (size_t)(LM + 2));
        double *TLEV = (double *)__builtin_alloca(sizeof(double) * (size_t)(LM + 2));
        double *TLEV_R = (double *)__builtin_alloca(sizeof(double) * (size_t)(LM + 2));
#pragma omp for schedule(static)
        for (int ij = 0; ij < ncol; ++ij) {
            /* PLE(:,1:LM+1) here: 0-based level k = K-1 */
            for (int K = 1; K <= LM; ++K) DPR[K] = G2(s->ple, ij, K) - G2(s->ple, ij, K - 1);   /* :6133 */
            for (int K = 1; K <= LM; ++K) {
                const int LV = LM - K + 1, k = K - 1, lv = LV - 1;
                G2(o->ciwp, ij, k) = (1.02 * 100 * DPR[LV]) * G2(s->qice, ij, lv);              /* :6136-6137 */
                G2(o->clwp, ij, k) = (1.02 * 100 * DPR[LV]) * G2(s->qliq, ij, lv);
                G2(o->rei, ij, k) = reice_limit(G2(s->rice, ij, lv), s->iceflg);                /* :6140-6170 */
                G2(o->rel, ij, k) = reliq_limit(G2(s->rliq, ij, lv), s->liqflg, 1);
            }
            for (int K = 2; K <= LM; ++K)                                                        /* :6173-6176 */
                TLEV[K] = (G2(s->t, ij, K - 2) * DPR[K] + G2(s->t, ij, K - 1) * DPR[K - 1]) / (DPR[K - 1] + DPR[K]);
            TLEV[LM + 1] = s->ts[ij];
            TLEV[1] = TLEV[2];
            for (int K = 1; K <= LM + 1; ++K) {                                                  /* :6180-6181 */
                G2(o->plev, ij, K - 1) = G2(s->ple, ij, LM + 1 - K) / 100.;
                TLEV_R[K] = TLEV[LM + 2 - K];
            }
            for (int K = 1; K <= LM; ++K) {                                                      /* :6183-6198 */
                const int k = K - 1, lv = LM - K;
                G2(o->play, ij, k) = G2(s->pl, ij, lv) / 100.;
                G2(o->tlay, ij, k) = G2(s->t, ij, lv);
                const double q = G2(s->q, ij, lv);
                G2(o->h2ovmr, ij, k) = q / (1. - q) * wq;
                G2(o->o3vmr, ij, k) = G2(s->o3, ij, lv) * wo3;
                G2(o->ch4vmr, ij, k) = G2(s->ch4, ij, lv);
                G2(o->co2vmr, ij, k) = s->co2;
                G2(o->o2vmr, ij, k) = s->o2;
                G2(o->cld, ij, k) = G2(s->cl, ij, lv);
                double *v[] = {&G2(o->h2ovmr, ij, k), &G2(o->o3vmr, ij, k), &G2(o->ch4vmr, ij, k),
                               &G2(o->co2vmr, ij, k), &G2(o->o2vmr, ij, k), &G2(o->cld, ij, k)};      /* :6201-6206 */
                for (size_t n = 0; n < sizeof v / sizeof v[0]; ++n)
                    if (*v[n] < 0.) *v[n] = 0.;
            }
            G2(o->zm, ij, 0) = 0.;                                                               /* :6212-6218 */
            for (int K = 2; K <= LM; ++K)
                G2(o->zm, ij, K - 1) = G2(o->zm, ij, K - 2) + s->rgas * TLEV_R[K] / s->grav *
                                                               (G2(o->play, ij, K - 2) - G2(o->play, ij, K - 1)) /
                                                               G2(o->plev, ij, K - 1);
            /* aerosol normalisation (:6116-6126) and flip (:6221-6223) */
            for (int b = 0; b < 14; ++b)
                for (int K = 1; K <= LM; ++K) {
                    const int k = K - 1, lv = LM - K;
                    double ta = 0., ss = 0., as = 0.;
                    if (s->taua) {
                        ta = G3(s->taua, ij, lv, b, LM); ss = G3(s->ssaa, ij, lv, b, LM); as = G3(s->asya, ij, lv, b, LM);
                        if (ta > 0. && ss > 0.) { as = as / ss; ss = ss / ta; }
                        else { ta = 0.; ss = 0.; as = 0.; }
                    }
                    G3(o->tauaer, ij, k, b, LM) = ta;
                    G3(o->ssaaer, ij, k, b, LM) = ss;
                    G3(o->asmaer, ij, k, b, LM) = as;
                }
        }
    }
    o->cloudLM = LM - s->lcldlm + 1;                                                             /* :6340 */
    o->cloudMH = LM - s->lcldmh + 1;
    return 0;
}

int oracle_solar_finish(int ncol, int LM, double undef, const int *clearCounts, const double *swuflx,
                        const double *swdflx, const double *swuflxc, const double *swdflxc, const double *cotd[4],
                        const double *cotn[4], OracleSolarFluxes *f) {
    for (int ij = 0; ij < ncol; ++ij) {
        for (int K = 0; K <= LM; ++K) {                                 /* unflip, :6395-6398; fluxes :6441-6444 */
            const int lv = LM - K;
            const double u = G2(swuflx, ij, lv), d = G2(swdflx, ij, lv);
            const double uc = G2(swuflxc, ij, lv), dc = G2(swdflxc, ij, lv);
            G2(f->fsw, ij, K) = d - u;
            G2(f->fsc, ij, K) = dc - uc;
            G2(f->fswu, ij, K) = u;
            G2(f->fscu, ij, K) = uc;
        }
        double *cld[4] = {f->cldts, f->cldhs, f->cldms, f->cldls};      /* :6407-6410 */
        for (int n = 0; n < 4; ++n)
            if (cld[n]) cld[n][ij] = 1. - clearCounts[ij + (size_t)ncol * n] / (double)NGPTSW;
        double *cot[4] = {f->cottp, f->cothp, f->cotmp, f->cotlp};      /* :6417-6439 */
        for (int n = 0; n < 4; ++n)
            if (cot[n]) cot[n][ij] = (cotn[n][ij] > 0. && cotd[n][ij] > 0.) ? cotn[n][ij] / cotd[n][ij] : undef;
    }
    return 0;
}

/* GEOS_IrradGridComp.F90 Update: DELT :3861, FLX_INT = FLXD_INT + FLXU_INT :3604-3606, USE_RRTMG branch :3929-3990.
 * Any output pointer may be NULL (unassociated export). */
int oracle_irrad_update(int ncol, int LM, const double *flxu_int, const double *flxd_int, const double *flcu_int,
                        const double *flcd_int, const double *dfdts, const double *dfdtsc, const double *sfcem_int,
                        const double *ts_int, const double *tsinst, OracleIrradExports *e) {
    for (int ij = 0; ij < ncol; ++ij) {
        const double DELT = tsinst[ij] - ts_int[ij];
        for (int K = 0; K <= LM; ++K) {
            const double FLX_INT = G2(flxd_int, ij, K) + G2(flxu_int, ij, K);
            const double FLC_INT = G2(flcd_int, ij, K) + G2(flcu_int, ij, K);
            if (e->flx) G2(e->flx, ij, K) = FLX_INT + G2(dfdts, ij, K) * DELT;
            if (e->flc) G2(e->flc, ij, K) = FLC_INT + G2(dfdtsc, ij, K) * DELT;
            if (e->flxu) G2(e->flxu, ij, K) = G2(flxu_int, ij, K) + G2(dfdts, ij, K) * DELT;
            if (e->flcu) G2(e->flcu, ij, K) = G2(flcu_int, ij, K) + G2(dfdtsc, ij, K) * DELT;
            if (e->flxd) G2(e->flxd, ij, K) = G2(flxd_int, ij, K);
            if (e->flcd) G2(e->flcd, ij, K) = G2(flcd_int, ij, K);
        }
        const double FLX0 = G2(flxd_int, ij, 0) + G2(flxu_int, ij, 0), FLC0 = G2(flcd_int, ij, 0) + G2(flcu_int, ij, 0);
        const double FLXL = G2(flxd_int, ij, LM) + G2(flxu_int, ij, LM), FLCL = G2(flcd_int, ij, LM) + G2(flcu_int, ij, LM);
        if (e->olr) e->olr[ij] = -(FLX0 + G2(dfdts, ij, 0) * DELT);
        if (e->olc) e->olc[ij] = -(FLC0 + G2(dfdtsc, ij, 0) * DELT);
        if (e->sfcem) e->sfcem[ij] = sfcem_int[ij] - G2(dfdts, ij, LM) * DELT;
        if (e->lws) e->lws[ij] = FLXL + sfcem_int[ij];
        if (e->lcs) e->lcs[ij] = FLCL + sfcem_int[ij];
        if (e->flns) e->flns[ij] = FLXL + G2(dfdts, ij, LM) * DELT;
        if (e->flnsc) e->flnsc[ij] = FLCL + G2(dfdtsc, ij, LM) * DELT;
    }
    return 0;
}
