/* oracle/lw.c -- CPU restatement of the RRTMG_LW column path (test infrastructure only).
 *
 * Routine-by-routine restatement, `real` promoted to fp64, of
 *   LW/src/rrtmg_lw_rad.F90      rrtmg_lw :15-344, rrtmg_lw_part :348-610
 *   LW/src/rrtmg_lw_setcoef.F90  setcoef :52-584
 *   LW/src/rrtmg_lw_taumol.F90   taugb1..16 :191-3126, addAerosols :3130-3146
 *   LW/src/rrtmg_lw_cldprmc.F90  cldprmc :24-385
 *   LW/src/rrtmg_lw_rtrnmc.F90   rtrnmc :27-390
 * (LW/ = GEOSirrad_GridComp/RRTMG/rrtmg_lw/gcm_model/).  Same loop nests, same expression
 * order; build with -ffp-contract=off.  Columns are processed in partitions of psize
 * columns like the reference; OpenMP runs partitions concurrently with private scratch.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "internal.h"

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

#define NG NGPTLW
#define NB NBNDLW

/* per-partition state of setcoef (module variables, LW/src/rrtmg_lw_setcoef.F90:22-47),
 * each array (nlay,pncol) unless noted */
typedef struct {
    int nlay, pncol;
    int *laytrop; /* (pncol) */
    double *colh2o, *colco2, *colo3, *coln2o, *colch4, *colo2, *colco, *colbrd, *colcfc11,
        *colcfc12, *colcfc22, *colccl4, *coldry;
    double *pwvcm;              /* (pncol) */
    double *planklev, *planklay; /* (16,0:nlay,pncol), (16,nlay,pncol) */
    double *plankbnd, *dplankbnd_dTs; /* (16,pncol) */
    double *forfac, *forfrac, *selffac, *selffrac, *scaleminor, *scaleminorn2, *minorfrac;
    int *jp, *jt, *jt1, *indfor, *indself, *indminor;
    double *rat_h2oco2, *rat_h2oco2_1, *rat_h2oo3, *rat_h2oo3_1, *rat_h2on2o, *rat_h2on2o_1,
        *rat_h2och4, *rat_h2och4_1, *rat_n2oco2, *rat_n2oco2_1, *rat_o3co2, *rat_o3co2_1;
    double *fac00, *fac01, *fac10, *fac11;
    double *wbroad;
} SetCoef;

static void *zalloc(size_t n) { return calloc(n ? n : 1, 1); }

static void setcoef_alloc(SetCoef *s, int nlay, int pncol) {
    size_t n = (size_t)nlay * pncol;
    s->nlay = nlay; s->pncol = pncol;
    s->laytrop = (int *)zalloc(sizeof(int) * pncol);
    double **d[] = {&s->colh2o, &s->colco2, &s->colo3, &s->coln2o, &s->colch4, &s->colo2, &s->colco,
                    &s->colbrd, &s->colcfc11, &s->colcfc12, &s->colcfc22, &s->colccl4, &s->coldry,
                    &s->forfac, &s->forfrac, &s->selffac, &s->selffrac, &s->scaleminor,
                    &s->scaleminorn2, &s->minorfrac, &s->rat_h2oco2, &s->rat_h2oco2_1,
                    &s->rat_h2oo3, &s->rat_h2oo3_1, &s->rat_h2on2o, &s->rat_h2on2o_1,
                    &s->rat_h2och4, &s->rat_h2och4_1, &s->rat_n2oco2, &s->rat_n2oco2_1,
                    &s->rat_o3co2, &s->rat_o3co2_1, &s->fac00, &s->fac01, &s->fac10, &s->fac11,
                    &s->wbroad};
    for (size_t i = 0; i < sizeof d / sizeof d[0]; ++i) *d[i] = (double *)zalloc(sizeof(double) * n);
    int **iv[] = {&s->jp, &s->jt, &s->jt1, &s->indfor, &s->indself, &s->indminor};
    for (size_t i = 0; i < sizeof iv / sizeof iv[0]; ++i) *iv[i] = (int *)zalloc(sizeof(int) * n);
    s->pwvcm = (double *)zalloc(sizeof(double) * pncol);
    s->planklev = (double *)zalloc(sizeof(double) * 16 * (nlay + 1) * pncol);
    s->planklay = (double *)zalloc(sizeof(double) * 16 * nlay * pncol);
    s->plankbnd = (double *)zalloc(sizeof(double) * 16 * pncol);
    s->dplankbnd_dTs = (double *)zalloc(sizeof(double) * 16 * pncol);
}

static void setcoef_free(SetCoef *s) {
    free(s->laytrop);
    double *d[] = {s->colh2o, s->colco2, s->colo3, s->coln2o, s->colch4, s->colo2, s->colco,
                   s->colbrd, s->colcfc11, s->colcfc12, s->colcfc22, s->colccl4, s->coldry,
                   s->forfac, s->forfrac, s->selffac, s->selffrac, s->scaleminor, s->scaleminorn2,
                   s->minorfrac, s->rat_h2oco2, s->rat_h2oco2_1, s->rat_h2oo3, s->rat_h2oo3_1,
                   s->rat_h2on2o, s->rat_h2on2o_1, s->rat_h2och4, s->rat_h2och4_1, s->rat_n2oco2,
                   s->rat_n2oco2_1, s->rat_o3co2, s->rat_o3co2_1, s->fac00, s->fac01, s->fac10,
                   s->fac11, s->wbroad, s->pwvcm, s->planklev, s->planklay, s->plankbnd,
                   s->dplankbnd_dTs};
    for (size_t i = 0; i < sizeof d / sizeof d[0]; ++i) free(d[i]);
    int *iv[] = {s->jp, s->jt, s->jt1, s->indfor, s->indself, s->indminor};
    for (size_t i = 0; i < sizeof iv / sizeof iv[0]; ++i) free(iv[i]);
}

/* 1-based accessors in the (nlay,pncol) partition layout */
#define IX(lay, icol) ((size_t)((lay)-1) + (size_t)nlay * ((icol)-1))
#define IX0(lev, icol) ((size_t)(lev) + (size_t)(nlay + 1) * ((icol)-1)) /* (0:nlay,pncol) */
#define TOTPLNK(i, b) T->totplnk[((i)-1) + 181 * ((b)-1)]
#define TOTPLNKD(i, b) T->totplnkderiv[((i)-1) + 181 * ((b)-1)]
#define CHI(m, j) T->chi_mls[((m)-1) + 7 * ((j)-1)]
#define PLEV(b, lev, icol) s->planklev[((b)-1) + 16 * ((size_t)(lev) + (size_t)(nlay + 1) * ((icol)-1))]
#define PLAY(b, lay, icol) s->planklay[((b)-1) + 16 * ((size_t)((lay)-1) + (size_t)nlay * ((icol)-1))]
#define PBND(b, icol) s->plankbnd[((b)-1) + 16 * ((icol)-1)]
#define DPBND(b, icol) s->dplankbnd_dTs[((b)-1) + 16 * ((icol)-1)]

static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* LW/src/rrtmg_lw_setcoef.F90:52-584 */
static int setcoef(SetCoef *s, int ncol, int nlay, int istart, int dudTs,
                   const double *pavel, const double *tavel, const double *pz, const double *tz,
                   const double *tbound, const double *semiss,
                   const double *h2ovmr, const double *o3vmr, const double *co2vmr,
                   const double *ch4vmr, const double *n2ovmr, const double *o2vmr,
                   const double *covmr, const double *cfc11vmr, const double *cfc12vmr,
                   const double *cfc22vmr, const double *ccl4vmr) {
    const LwTables *T = &g_lw;
    const double amd = 28.9660, amw = 18.0160;
    const double stpfac = 296. / 1013.;
    const double grav = T->grav, avogad = T->avogad;

    for (int icol = 1; icol <= ncol; ++icol) {
        for (int lay = 1; lay <= nlay; ++lay) {
            double h2o = h2ovmr[IX(lay, icol)];
            double amm = (1. - h2o) * amd + h2o * amw;
            s->coldry[IX(lay, icol)] = (pz[IX0(lay - 1, icol)] - pz[IX0(lay, icol)]) * 1.e3 * avogad /
                                       (1.e2 * grav * amm * (1. + h2o));
        }
        double amttl = 0., wvttl = 0.;
        for (int lay = 1; lay <= nlay; ++lay) {
            size_t i = IX(lay, icol);
            double summol = co2vmr[i] + o3vmr[i] + n2ovmr[i] + ch4vmr[i] + o2vmr[i];
            s->wbroad[i] = s->coldry[i] * (1. - summol);
            double btemp = h2ovmr[i] * s->coldry[i];
            amttl = amttl + s->coldry[i] + btemp;
            wvttl = wvttl + btemp;
        }
        double wvsh = (amw * wvttl) / (amd * amttl);
        s->pwvcm[icol - 1] = wvsh * (1.e3 * pz[IX0(0, icol)]) / (1.e2 * grav);
    }

    for (int icol = 1; icol <= ncol; ++icol) {
        int indbound = clampi(f_int(tbound[icol - 1] - 159.), 1, 180);
        double tbndfrac = tbound[icol - 1] - 159. - (double)indbound;
        int indlev0 = clampi(f_int(tz[IX0(0, icol)] - 159.), 1, 180);
        double t0frac = tz[IX0(0, icol)] - 159. - (double)indlev0;
        int upper_atmosphere_found = 0;
        s->laytrop[icol - 1] = 0;

        for (int lay = 1; lay <= nlay; ++lay) {
            size_t i = IX(lay, icol);
            double lcoldry = s->coldry[i];
            double wv = h2ovmr[i] * lcoldry;
            int indlay = clampi(f_int(tavel[i] - 159.), 1, 180);
            double tlayfrac = tavel[i] - 159. - (double)indlay;
            int indlev = clampi(f_int(tz[IX0(lay, icol)] - 159.), 1, 180);
            double tlevfrac = tz[IX0(lay, icol)] - 159. - (double)indlev;

            for (int iband = 1; iband <= 16; ++iband) {
                /* band 16 with istart /= 16 uses totplnk like bands 1-15 (:356-394);
                 * istart == 16 (never used by GEOS) would switch to totplk16 */
                const double *tp = T->totplnk + 181 * (iband - 1);
                const double *tpd = T->totplnkderiv + 181 * (iband - 1);
                if (iband == 16 && istart == 16) { tp = T->totplk16; tpd = T->totplk16deriv; }
#define TP(k) tp[(k)-1]
#define TPD(k) tpd[(k)-1]
                double dbdtlev, dbdtlay;
                if (lay == 1) {
                    dbdtlev = TP(indbound + 1) - TP(indbound);
                    PBND(iband, icol) = semiss[(iband - 1) + 16 * (icol - 1)] * (TP(indbound) + tbndfrac * dbdtlev);
                    dbdtlev = TP(indlev0 + 1) - TP(indlev0);
                    PLEV(iband, 0, icol) = TP(indlev0) + t0frac * dbdtlev;
                    if (dudTs) {
                        dbdtlev = TPD(indbound + 1) - TPD(indbound);
                        DPBND(iband, icol) = semiss[(iband - 1) + 16 * (icol - 1)] * (TPD(indbound) + tbndfrac * dbdtlev);
                    }
                }
                dbdtlev = TP(indlev + 1) - TP(indlev);
                PLEV(iband, lay, icol) = TP(indlev) + tlevfrac * dbdtlev;
                dbdtlay = TP(indlay + 1) - TP(indlay);
                PLAY(iband, lay, icol) = TP(indlay) + tlayfrac * dbdtlay;
#undef TP
#undef TPD
            }

            double plog = log(pavel[i]);
            int jp = clampi(f_int(36. - 5. * (plog + 0.04)), 1, 58);
            s->jp[i] = jp;
            int jp1 = jp + 1;
            double fp = 5. * (T->preflog[jp - 1] - plog);

            int jt = clampi(f_int(3. + (tavel[i] - T->tref[jp - 1]) / 15.), 1, 4);
            s->jt[i] = jt;
            double ft = ((tavel[i] - T->tref[jp - 1]) / 15.) - (double)(jt - 3);
            int jt1 = clampi(f_int(3. + (tavel[i] - T->tref[jp1 - 1]) / 15.), 1, 4);
            s->jt1[i] = jt1;
            double ft1 = ((tavel[i] - T->tref[jp1 - 1]) / 15.) - (double)(jt1 - 3);

            double water = wv / lcoldry;
            double scalefac = pavel[i] * stpfac / tavel[i];
            double factor;

            if (plog > 4.56) {
                if (upper_atmosphere_found) return -31; /* 'RRTMG LW pressure misordering' */
                s->laytrop[icol - 1] += 1;
                s->forfac[i] = scalefac / (1. + water);
                factor = (332. - tavel[i]) / 36.;
                s->indfor[i] = clampi(f_int(factor), 1, 2);
                s->forfrac[i] = factor - (double)s->indfor[i];
                s->selffac[i] = water * s->forfac[i];
                factor = (tavel[i] - 188.) / 7.2;
                s->indself[i] = clampi(f_int(factor) - 7, 1, 9);
                s->selffrac[i] = factor - (double)(s->indself[i] + 7);
                s->scaleminor[i] = pavel[i] / tavel[i];
                s->scaleminorn2[i] = (pavel[i] / tavel[i]) * (s->wbroad[i] / (lcoldry + wv));
                factor = (tavel[i] - 180.8) / 7.2;
                s->indminor[i] = clampi(f_int(factor), 1, 18);
                s->minorfrac[i] = factor - (double)s->indminor[i];
                s->rat_h2oco2[i] = CHI(1, jp) / CHI(2, jp);
                s->rat_h2oco2_1[i] = CHI(1, jp + 1) / CHI(2, jp + 1);
                s->rat_h2oo3[i] = CHI(1, jp) / CHI(3, jp);
                s->rat_h2oo3_1[i] = CHI(1, jp + 1) / CHI(3, jp + 1);
                s->rat_h2on2o[i] = CHI(1, jp) / CHI(4, jp);
                s->rat_h2on2o_1[i] = CHI(1, jp + 1) / CHI(4, jp + 1);
                s->rat_h2och4[i] = CHI(1, jp) / CHI(6, jp);
                s->rat_h2och4_1[i] = CHI(1, jp + 1) / CHI(6, jp + 1);
                s->rat_n2oco2[i] = CHI(4, jp) / CHI(2, jp);
                s->rat_n2oco2_1[i] = CHI(4, jp + 1) / CHI(2, jp + 1);
            } else {
                upper_atmosphere_found = 1;
                s->forfac[i] = scalefac / (1. + water);
                factor = (tavel[i] - 188.) / 36.;
                s->indfor[i] = 3;
                s->forfrac[i] = factor - 1.;
                s->selffac[i] = 0.;
                s->scaleminor[i] = pavel[i] / tavel[i];
                s->scaleminorn2[i] = (pavel[i] / tavel[i]) * (s->wbroad[i] / (lcoldry + wv));
                factor = (tavel[i] - 180.8) / 7.2;
                s->indminor[i] = clampi(f_int(factor), 1, 18);
                s->minorfrac[i] = factor - (double)s->indminor[i];
                s->rat_h2oco2[i] = CHI(1, jp) / CHI(2, jp);
                s->rat_h2oco2_1[i] = CHI(1, jp + 1) / CHI(2, jp + 1);
                s->rat_o3co2[i] = CHI(3, jp) / CHI(2, jp);
                s->rat_o3co2_1[i] = CHI(3, jp + 1) / CHI(2, jp + 1);
            }

            s->colh2o[i] = 1.e-20 * h2ovmr[i] * lcoldry;
            s->colco2[i] = 1.e-20 * co2vmr[i] * lcoldry;
            s->colo3[i] = 1.e-20 * o3vmr[i] * lcoldry;
            s->coln2o[i] = 1.e-20 * n2ovmr[i] * lcoldry;
            s->colch4[i] = 1.e-20 * ch4vmr[i] * lcoldry;
            s->colo2[i] = 1.e-20 * o2vmr[i] * lcoldry;
            s->colco[i] = 1.e-20 * covmr[i] * lcoldry;
            s->colcfc11[i] = 1.e-20 * cfc11vmr[i] * lcoldry;
            s->colcfc12[i] = 1.e-20 * cfc12vmr[i] * lcoldry;
            s->colcfc22[i] = 1.e-20 * cfc22vmr[i] * lcoldry;
            s->colccl4[i] = 1.e-20 * ccl4vmr[i] * lcoldry;
            s->colbrd[i] = 1.e-20 * s->wbroad[i];
            if (s->colco2[i] == 0.) s->colco2[i] = 1.e-32 * lcoldry;
            if (s->colo3[i] == 0.) s->colo3[i] = 1.e-32 * lcoldry;
            if (s->coln2o[i] == 0.) s->coln2o[i] = 1.e-32 * lcoldry;
            if (s->colch4[i] == 0.) s->colch4[i] = 1.e-32 * lcoldry;
            if (s->colco[i] == 0.) s->colco[i] = 1.e-32 * lcoldry;

            double compfp = 1. - fp;
            s->fac10[i] = compfp * ft;
            s->fac00[i] = compfp * (1. - ft);
            s->fac11[i] = fp * ft1;
            s->fac01[i] = fp * (1. - ft1);
            s->selffac[i] = s->colh2o[i] * s->selffac[i];
            s->forfac[i] = s->colh2o[i] * s->forfac[i];
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * taumol, LW/src/rrtmg_lw_taumol.F90
 * ---------------------------------------------------------------------------------------- */

/* binary-species parameter: speccomb = a + rat*b; specparm = a/speccomb capped at oneminus;
 * specmult = mult*specparm; js = 1 + int(specmult); fs = mod(specmult,1.)  (e.g. :436-441) */
typedef struct { double speccomb, specparm, fs; int js; } Spec;
static inline Spec spec(double cola, double rat, double colb, double mult) {
    Spec r;
    r.speccomb = cola + rat * colb;
    r.specparm = cola / r.speccomb;
    if (r.specparm >= g_lw.oneminus) r.specparm = g_lw.oneminus;
    double specmult = mult * r.specparm;
    r.js = 1 + f_int(specmult);
    r.fs = fmod(specmult, 1.0);
    return r;
}

/* lower-atmosphere key-species interpolation weights and table offsets for one of the two
 * reference pressures (e.g. taugb3 :482-511 and :554-576): terms are summed left to right */
typedef struct { int n; int off[6]; double w[6]; } Stencil;
static inline Stencil stencil_lower(double specparm, double fs, double fa, double fb) {
    Stencil s;
    if (specparm < 0.125) {
        double p = fs - 1.;
        double p2 = p * p, p4 = p2 * p2; /* p**4 */
        double fk0 = p4, fk1 = 1. - p - 2.0 * p4, fk2 = p + p4;
        s.n = 6;
        s.w[0] = fk0 * fa; s.off[0] = 0;
        s.w[1] = fk1 * fa; s.off[1] = 1;
        s.w[2] = fk2 * fa; s.off[2] = 2;
        s.w[3] = fk0 * fb; s.off[3] = 9;
        s.w[4] = fk1 * fb; s.off[4] = 10;
        s.w[5] = fk2 * fb; s.off[5] = 11;
    } else if (specparm > 0.875) {
        double p = -fs;
        double p2 = p * p, p4 = p2 * p2;
        double fk0 = p4, fk1 = 1. - p - 2.0 * p4, fk2 = p + p4;
        s.n = 6;
        s.w[0] = fk2 * fa; s.off[0] = -1;
        s.w[1] = fk1 * fa; s.off[1] = 0;
        s.w[2] = fk0 * fa; s.off[2] = 1;
        s.w[3] = fk2 * fb; s.off[3] = 8;
        s.w[4] = fk1 * fb; s.off[4] = 9;
        s.w[5] = fk0 * fb; s.off[5] = 10;
    } else {
        s.n = 4;
        s.w[0] = (1. - fs) * fa; s.off[0] = 0;
        s.w[1] = fs * fa;        s.off[1] = 1;
        s.w[2] = (1. - fs) * fb; s.off[2] = 9;
        s.w[3] = fs * fb;        s.off[3] = 10;
    }
    return s;
}
static inline double stencil_sum(const Stencil *s, const double *abs_g, int ind) {
    /* abs_g points at absa(1,ig); ind is the 1-based ind0/ind1 */
    double acc = s->w[0] * abs_g[ind - 1 + s->off[0]];
    for (int k = 1; k < s->n; ++k) acc = acc + s->w[k] * abs_g[ind - 1 + s->off[k]];
    return acc;
}

/* t(i,ig) + f*(t(i+1,ig) - t(i,ig)) for a table t(lead,ng) */
static inline double lerp1(const double *t, int lead, int i, int ig, double f) {
    const double *c = t + (size_t)lead * (ig - 1);
    return c[i - 1] + f * (c[i] - c[i - 1]);
}
/* binary minor species k(jm,indm,ig) of shape (nj,19,ng), e.g. taugb3 :548-552 */
static inline double minor2(const double *k, int nj, int jm, int indm, int ig, double fm, double minorfrac) {
    const double *c = k + (size_t)nj * 19 * (ig - 1);
#define K(j, m) c[((j)-1) + nj * ((m)-1)]
    double m1 = K(jm, indm) + fm * (K(jm + 1, indm) - K(jm, indm));
    double m2 = K(jm, indm + 1) + fm * (K(jm + 1, indm + 1) - K(jm, indm + 1));
#undef K
    return m1 + minorfrac * (m2 - m1);
}

/* view of one column of the partition */
typedef struct {
    int nlay, laytrop;
    const int *jp, *jt, *jt1, *indfor, *indself, *indminor;
    const double *fac00, *fac01, *fac10, *fac11, *colh2o, *colco2, *colo3, *coln2o, *colch4,
        *colo2, *colco, *colbrd, *colcfc11, *colcfc12, *colcfc22, *colccl4, *coldry, *forfac,
        *forfrac, *selffac, *selffrac, *scaleminor, *scaleminorn2, *minorfrac, *rat_h2oco2,
        *rat_h2oco2_1, *rat_h2oo3, *rat_h2oo3_1, *rat_h2on2o, *rat_h2on2o_1, *rat_h2och4,
        *rat_h2och4_1, *rat_n2oco2, *rat_n2oco2_1, *rat_o3co2, *rat_o3co2_1, *pavel;
    double *taug, *pfracs; /* (nlay,ngptlw) of this column */
} Col;

#define A(x) (c->x[lay - 1])
#define TAUG(ig) c->taug[(lay - 1) + (size_t)c->nlay * ((ig)-1)]
#define PFR(ig) c->pfracs[(lay - 1) + (size_t)c->nlay * ((ig)-1)]
#define ABSA_G(B, ig) ((B)->absa + (size_t)65 * (B)->nspa * ((ig)-1))
/* absb(235*nspb, ng); band 16 has nspb = 0 in rrtmg_lw_init.F90:195 yet carries a 235-row kb,
 * so its ind0/ind1 collapse to 1 while the g stride stays 235 (taugb16 :3109-3110) */
#define ABSB_G(B, ig) ((B)->absb + (size_t)235 * ((B)->nspb ? (B)->nspb : 1) * ((ig)-1))
#define IND0_LO(nsp) (((A(jp) - 1) * 5 + (A(jt) - 1)) * (nsp))
#define IND1_LO(nsp) ((A(jp) * 5 + (A(jt1) - 1)) * (nsp))
#define IND0_UP(nsp) (((A(jp) - 13) * 5 + (A(jt) - 1)) * (nsp))
#define IND1_UP(nsp) (((A(jp) - 12) * 5 + (A(jt1) - 1)) * (nsp))
#define CHIM(m, j) g_lw.chi_mls[((m)-1) + 7 * ((j)-1)]
#define TAUSELF(B, ig) (A(selffac) * lerp1((B)->selfref, 10, A(indself), ig, A(selffrac)))
#define TAUFOR(B, ig) (A(forfac) * lerp1((B)->forref, 4, A(indfor), ig, A(forfrac)))
#define KEY4(absg, ind0, ind1) (A(fac00) * (absg)[(ind0)-1] + A(fac10) * (absg)[(ind0)] + \
                                A(fac01) * (absg)[(ind1)-1] + A(fac11) * (absg)[(ind1)])
/* Planck fraction interpolated in the binary-species parameter, e.g. :604-605 */
#define PFRAC2(fr, ng_, ig, jpl, fpl) ((fr)[((ig)-1) + (ng_) * ((jpl)-1)] + (fpl) * \
        ((fr)[((ig)-1) + (ng_) * (jpl)] - (fr)[((ig)-1) + (ng_) * ((jpl)-1)]))

/* adjusted minor column amount when the gas exceeds its reference abundance, e.g. :460-467 */
static inline double adjcol(double col, double coldry, double chi, double thresh, double base, double expo) {
    double chi_x = col / coldry;
    double rat = 1.e20 * chi_x / chi;
    if (rat > thresh) {
        double adjfac = base + pow(rat - base, expo);
        return adjfac * chi * coldry * 1.e-20;
    }
    return col;
}

/* band 1: :191-285 */
static void taugb1(Col *c) {
    const LwBand *B = &g_lw.b[0];
    const int ng = B->ng, gs = 0;
    for (int lay = 1; lay <= c->laytrop; ++lay) {
        int ind0 = IND0_LO(B->nspa) + 1, ind1 = IND1_LO(B->nspa) + 1;
        int indm = A(indminor);
        double pp = A(pavel);
        double corradj = 1.;
        if (pp < 250.) corradj = 1. - 0.15 * (250. - pp) / 154.4;
        double scalen2 = A(colbrd) * A(scaleminorn2);
        for (int ig = 1; ig <= ng; ++ig) {
            double tauself = TAUSELF(B, ig);
            double taufor = TAUFOR(B, ig);
            double taun2 = scalen2 * lerp1(B->ka_mn2, 19, indm, ig, A(minorfrac));
            TAUG(gs + ig) = corradj * (A(colh2o) * KEY4(ABSA_G(B, ig), ind0, ind1) + tauself + taufor + taun2);
            PFR(gs + ig) = B->fracrefa[ig - 1];
        }
    }
    for (int lay = c->laytrop + 1; lay <= c->nlay; ++lay) {
        int ind0 = IND0_UP(B->nspb) + 1, ind1 = IND1_UP(B->nspb) + 1;
        int indm = A(indminor);
        double pp = A(pavel);
        double corradj = 1. - 0.15 * (pp / 95.6);
        double scalen2 = A(colbrd) * A(scaleminorn2);
        for (int ig = 1; ig <= ng; ++ig) {
            double taufor = TAUFOR(B, ig);
            double taun2 = scalen2 * lerp1(B->kb_mn2, 19, indm, ig, A(minorfrac));
            TAUG(gs + ig) = corradj * (A(colh2o) * KEY4(ABSB_G(B, ig), ind0, ind1) + taufor + taun2);
            PFR(gs + ig) = B->fracrefb[ig - 1];
        }
    }
}

/* band 2: :289-363 */
static void taugb2(Col *c) {
    const LwBand *B = &g_lw.b[1];
    const int ng = B->ng, gs = g_lw.ngs[0];
    for (int lay = 1; lay <= c->laytrop; ++lay) {
        int ind0 = IND0_LO(B->nspa) + 1, ind1 = IND1_LO(B->nspa) + 1;
        double pp = A(pavel);
        double corradj = 1. - .05 * (pp - 100.) / 900.;
        for (int ig = 1; ig <= ng; ++ig) {
            double tauself = TAUSELF(B, ig);
            double taufor = TAUFOR(B, ig);
            TAUG(gs + ig) = corradj * (A(colh2o) * KEY4(ABSA_G(B, ig), ind0, ind1) + tauself + taufor);
            PFR(gs + ig) = B->fracrefa[ig - 1];
        }
    }
    for (int lay = c->laytrop + 1; lay <= c->nlay; ++lay) {
        int ind0 = IND0_UP(B->nspb) + 1, ind1 = IND1_UP(B->nspb) + 1;
        for (int ig = 1; ig <= ng; ++ig) {
            double taufor = TAUFOR(B, ig);
            TAUG(gs + ig) = A(colh2o) * KEY4(ABSB_G(B, ig), ind0, ind1) + taufor;
            PFR(gs + ig) = B->fracrefb[ig - 1];
        }
    }
}

/* upper-atmosphere binary key species with nspb = 5 (bands 3, 4, 5), e.g. :675-685 */
static inline double key_upper5(const double *absg, int ind0, int ind1, const Spec *s0,
                                const Spec *s1, double fac00, double fac10, double fac01,
                                double fac11) {
    double fac000 = (1. - s0->fs) * fac00, fac010 = (1. - s0->fs) * fac10;
    double fac100 = s0->fs * fac00, fac110 = s0->fs * fac10;
    double fac001 = (1. - s1->fs) * fac01, fac011 = (1. - s1->fs) * fac11;
    double fac101 = s1->fs * fac01, fac111 = s1->fs * fac11;
    return s0->speccomb * (fac000 * absg[ind0 - 1] + fac100 * absg[ind0] + fac010 * absg[ind0 + 4] +
                           fac110 * absg[ind0 + 5]) +
           s1->speccomb * (fac001 * absg[ind1 - 1] + fac101 * absg[ind1] + fac011 * absg[ind1 + 4] +
                           fac111 * absg[ind1 + 5]);
}

/* band 3: :367-695 */
static void taugb3(Col *c) {
    const LwBand *B = &g_lw.b[2];
    const int ng = B->ng, gs = g_lw.ngs[1];
    double refrat_planck_a = CHIM(1, 9) / CHIM(2, 9);
    double refrat_m_a = CHIM(1, 3) / CHIM(2, 3);
    double refrat_planck_b = CHIM(1, 13) / CHIM(2, 13);
    double refrat_m_b = refrat_planck_b;
    for (int lay = 1; lay <= c->laytrop; ++lay) {
        Spec s0 = spec(A(colh2o), A(rat_h2oco2), A(colco2), 8.);
        Spec s1 = spec(A(colh2o), A(rat_h2oco2_1), A(colco2), 8.);
        Spec sm = spec(A(colh2o), refrat_m_a, A(colco2), 8.);
        double adjcoln2o = adjcol(A(coln2o), A(coldry), CHIM(4, A(jp) + 1), 1.5, 0.5, 0.65);
        Spec sp = spec(A(colh2o), refrat_planck_a, A(colco2), 8.);
        int ind0 = IND0_LO(B->nspa) + s0.js, ind1 = IND1_LO(B->nspa) + s1.js;
        int indm = A(indminor);
        Stencil st0 = stencil_lower(s0.specparm, s0.fs, A(fac00), A(fac10));
        Stencil st1 = stencil_lower(s1.specparm, s1.fs, A(fac01), A(fac11));
        for (int ig = 1; ig <= ng; ++ig) {
            double tauself = TAUSELF(B, ig);
            double taufor = TAUFOR(B, ig);
            double absn2o = minor2(B->ka_mn2o, 9, sm.js, indm, ig, sm.fs, A(minorfrac));
            double tau_major = s0.speccomb * stencil_sum(&st0, ABSA_G(B, ig), ind0);
            double tau_major1 = s1.speccomb * stencil_sum(&st1, ABSA_G(B, ig), ind1);
            TAUG(gs + ig) = tau_major + tau_major1 + tauself + taufor + adjcoln2o * absn2o;
            PFR(gs + ig) = PFRAC2(B->fracrefa, ng, ig, sp.js, sp.fs);
        }
    }
    for (int lay = c->laytrop + 1; lay <= c->nlay; ++lay) {
        Spec s0 = spec(A(colh2o), A(rat_h2oco2), A(colco2), 4.);
        Spec s1 = spec(A(colh2o), A(rat_h2oco2_1), A(colco2), 4.);
        Spec sm = spec(A(colh2o), refrat_m_b, A(colco2), 4.);
        double adjcoln2o = adjcol(A(coln2o), A(coldry), CHIM(4, A(jp) + 1), 1.5, 0.5, 0.65);
        Spec sp = spec(A(colh2o), refrat_planck_b, A(colco2), 4.);
        int ind0 = IND0_UP(B->nspb) + s0.js, ind1 = IND1_UP(B->nspb) + s1.js;
        int indm = A(indminor);
        for (int ig = 1; ig <= ng; ++ig) {
            double taufor = TAUFOR(B, ig);
            double absn2o = minor2(B->kb_mn2o, 5, sm.js, indm, ig, sm.fs, A(minorfrac));
            TAUG(gs + ig) = key_upper5(ABSB_G(B, ig), ind0, ind1, &s0, &s1, A(fac00), A(fac10), A(fac01), A(fac11)) +
                            taufor + adjcoln2o * absn2o;
            PFR(gs + ig) = PFRAC2(B->fracrefb, ng, ig, sp.js, sp.fs);
        }
    }
}

/* band 4: :699-960 */
static void taugb4(Col *c) {
    const LwBand *B = &g_lw.b[3];
    const int ng = B->ng, gs = g_lw.ngs[2];
    double refrat_planck_a = CHIM(1, 11) / CHIM(2, 11);
    double refrat_planck_b = CHIM(3, 13) / CHIM(2, 13);
    for (int lay = 1; lay <= c->laytrop; ++lay) {
        Spec s0 = spec(A(colh2o), A(rat_h2oco2), A(colco2), 8.);
        Spec s1 = spec(A(colh2o), A(rat_h2oco2_1), A(colco2), 8.);
        Spec sp = spec(A(colh2o), refrat_planck_a, A(colco2), 8.);
        int ind0 = IND0_LO(B->nspa) + s0.js, ind1 = IND1_LO(B->nspa) + s1.js;
        Stencil st0 = stencil_lower(s0.specparm, s0.fs, A(fac00), A(fac10));
        Stencil st1 = stencil_lower(s1.specparm, s1.fs, A(fac01), A(fac11));
        for (int ig = 1; ig <= ng; ++ig) {
            double tauself = TAUSELF(B, ig);
            double taufor = TAUFOR(B, ig);
            double tau_major = s0.speccomb * stencil_sum(&st0, ABSA_G(B, ig), ind0);
            double tau_major1 = s1.speccomb * stencil_sum(&st1, ABSA_G(B, ig), ind1);
            TAUG(gs + ig) = tau_major + tau_major1 + tauself + taufor;
            PFR(gs + ig) = PFRAC2(B->fracrefa, ng, ig, sp.js, sp.fs);
        }
    }
    for (int lay = c->laytrop + 1; lay <= c->nlay; ++lay) {
        Spec s0 = spec(A(colo3), A(rat_o3co2), A(colco2), 4.);
        Spec s1 = spec(A(colo3), A(rat_o3co2_1), A(colco2), 4.);
        Spec sp = spec(A(colo3), refrat_planck_b, A(colco2), 4.);
        int ind0 = IND0_UP(B->nspb) + s0.js, ind1 = IND1_UP(B->nspb) + s1.js;
        for (int ig = 1; ig <= ng; ++ig) {
            TAUG(gs + ig) = key_upper5(ABSB_G(B, ig), ind0, ind1, &s0, &s1, A(fac00), A(fac10), A(fac01), A(fac11));
            PFR(gs + ig) = PFRAC2(B->fracrefb, ng, ig, sp.js, sp.fs);
        }
        /* empirical stratospheric CO2 cooling-rate fix, :948-954 */
        TAUG(gs + 8) = TAUG(gs + 8) * 0.92;
        TAUG(gs + 9) = TAUG(gs + 9) * 0.88;
        TAUG(gs + 10) = TAUG(gs + 10) * 1.07;
        TAUG(gs + 11) = TAUG(gs + 11) * 1.1;
        TAUG(gs + 12) = TAUG(gs + 12) * 0.99;
        TAUG(gs + 13) = TAUG(gs + 13) * 0.88;
        TAUG(gs + 14) = TAUG(gs + 14) * 0.943;
    }
}

/* band 5: :964-1239 */
static void taugb5(Col *c) {
    const LwBand *B = &g_lw.b[4];
    const int ng = B->ng, gs = g_lw.ngs[3];
    double refrat_planck_a = CHIM(1, 5) / CHIM(2, 5);
    double refrat_m_a = CHIM(1, 7) / CHIM(2, 7);
    double refrat_planck_b = CHIM(3, 43) / CHIM(2, 43);
    for (int lay = 1; lay <= c->laytrop; ++lay) {
        Spec s0 = spec(A(colh2o), A(rat_h2oco2), A(colco2), 8.);
        Spec s1 = spec(A(colh2o), A(rat_h2oco2_1), A(colco2), 8.);
        Spec sm = spec(A(colh2o), refrat_m_a, A(colco2), 8.);
        Spec sp = spec(A(colh2o), refrat_planck_a, A(colco2), 8.);
        int ind0 = IND0_LO(B->nspa) + s0.js, ind1 = IND1_LO(B->nspa) + s1.js;
        int indm = A(indminor);
        Stencil st0 = stencil_lower(s0.specparm, s0.fs, A(fac00), A(fac10));
        Stencil st1 = stencil_lower(s1.specparm, s1.fs, A(fac01), A(fac11));
        for (int ig = 1; ig <= ng; ++ig) {
            double tauself = TAUSELF(B, ig);
            double taufor = TAUFOR(B, ig);
            double abso3 = minor2(B->ka_mo3, 9, sm.js, indm, ig, sm.fs, A(minorfrac));
            double tau_major = s0.speccomb * stencil_sum(&st0, ABSA_G(B, ig), ind0);
            double tau_major1 = s1.speccomb * stencil_sum(&st1, ABSA_G(B, ig), ind1);
            TAUG(gs + ig) = tau_major + tau_major1 + tauself + taufor + abso3 * A(colo3) +
                            A(colccl4) * B->ccl4[ig - 1];
            PFR(gs + ig) = PFRAC2(B->fracrefa, ng, ig, sp.js, sp.fs);
        }
    }
    for (int lay = c->laytrop + 1; lay <= c->nlay; ++lay) {
        Spec s0 = spec(A(colo3), A(rat_o3co2), A(colco2), 4.);
        Spec s1 = spec(A(colo3), A(rat_o3co2_1), A(colco2), 4.);
        Spec sp = spec(A(colo3), refrat_planck_b, A(colco2), 4.);
        int ind0 = IND0_UP(B->nspb) + s0.js, ind1 = IND1_UP(B->nspb) + s1.js;
        for (int ig = 1; ig <= ng; ++ig) {
            TAUG(gs + ig) = key_upper5(ABSB_G(B, ig), ind0, ind1, &s0, &s1, A(fac00), A(fac10), A(fac01), A(fac11)) +
                            A(colccl4) * B->ccl4[ig - 1];
            PFR(gs + ig) = PFRAC2(B->fracrefb, ng, ig, sp.js, sp.fs);
        }
    }
}

/* band 6: :1243-1327 */
static void taugb6(Col *c) {
    const LwBand *B = &g_lw.b[5];
    const int ng = B->ng, gs = g_lw.ngs[4];
    for (int lay = 1; lay <= c->laytrop; ++lay) {
        double adjcolco2 = adjcol(A(colco2), A(coldry), CHIM(2, A(jp) + 1), 3.0, 2.0, 0.77);
        int ind0 = IND0_LO(B->nspa) + 1, ind1 = IND1_LO(B->nspa) + 1;
        int indm = A(indminor);
        for (int ig = 1; ig <= ng; ++ig) {
            double tauself = TAUSELF(B, ig);
            double taufor = TAUFOR(B, ig);
            double absco2 = lerp1(B->ka_mco2, 19, indm, ig, A(minorfrac));
            TAUG(gs + ig) = A(colh2o) * KEY4(ABSA_G(B, ig), ind0, ind1) + tauself + taufor +
                            adjcolco2 * absco2 + A(colcfc11) * B->cfc11adj[ig - 1] +
                            A(colcfc12) * B->cfc12[ig - 1];
            PFR(gs + ig) = B->fracrefa[ig - 1];
        }
    }
    for (int lay = c->laytrop + 1; lay <= c->nlay; ++lay) {
        for (int ig = 1; ig <= ng; ++ig) {
            TAUG(gs + ig) = 0.0 + A(colcfc11) * B->cfc11adj[ig - 1] + A(colcfc12) * B->cfc12[ig - 1];
            PFR(gs + ig) = B->fracrefa[ig - 1];
        }
    }
}

/* band 7: :1331-1603 */
static void taugb7(Col *c) {
    const LwBand *B = &g_lw.b[6];
    const int ng = B->ng, gs = g_lw.ngs[5];
    double refrat_planck_a = CHIM(1, 3) / CHIM(3, 3);
    double refrat_m_a = refrat_planck_a;
    for (int lay = 1; lay <= c->laytrop; ++lay) {
        Spec s0 = spec(A(colh2o), A(rat_h2oo3), A(colo3), 8.);
        Spec s1 = spec(A(colh2o), A(rat_h2oo3_1), A(colo3), 8.);
        Spec sm = spec(A(colh2o), refrat_m_a, A(colo3), 8.);
        double adjcolco2 = adjcol(A(colco2), A(coldry), CHIM(2, A(jp) + 1), 3.0, 3.0, 0.79);
        Spec sp = spec(A(colh2o), refrat_planck_a, A(colo3), 8.);
        int ind0 = IND0_LO(B->nspa) + s0.js, ind1 = IND1_LO(B->nspa) + s1.js;
        int indm = A(indminor);
        Stencil st0 = stencil_lower(s0.specparm, s0.fs, A(fac00), A(fac10));
        Stencil st1 = stencil_lower(s1.specparm, s1.fs, A(fac01), A(fac11));
        for (int ig = 1; ig <= ng; ++ig) {
            double tauself = TAUSELF(B, ig);
            double taufor = TAUFOR(B, ig);
            double absco2 = minor2(B->ka_mco2, 9, sm.js, indm, ig, sm.fs, A(minorfrac));
            double tau_major = s0.speccomb * stencil_sum(&st0, ABSA_G(B, ig), ind0);
            double tau_major1 = s1.speccomb * stencil_sum(&st1, ABSA_G(B, ig), ind1);
            TAUG(gs + ig) = tau_major + tau_major1 + tauself + taufor + adjcolco2 * absco2;
            PFR(gs + ig) = PFRAC2(B->fracrefa, ng, ig, sp.js, sp.fs);
        }
    }
    for (int lay = c->laytrop + 1; lay <= c->nlay; ++lay) {
        double adjcolco2 = adjcol(A(colco2), A(coldry), CHIM(2, A(jp) + 1), 3.0, 2.0, 0.79);
        int ind0 = IND0_UP(B->nspb) + 1, ind1 = IND1_UP(B->nspb) + 1;
        int indm = A(indminor);
        for (int ig = 1; ig <= ng; ++ig) {
            double absco2 = lerp1(B->kb_mco2, 19, indm, ig, A(minorfrac));
            TAUG(gs + ig) = A(colo3) * KEY4(ABSB_G(B, ig), ind0, ind1) + adjcolco2 * absco2;
            PFR(gs + ig) = B->fracrefb[ig - 1];
        }
        /* empirical stratospheric O3 cooling-rate fix, :1592-1597 */
        TAUG(gs + 6) = TAUG(gs + 6) * 0.92;
        TAUG(gs + 7) = TAUG(gs + 7) * 0.88;
        TAUG(gs + 8) = TAUG(gs + 8) * 1.07;
        TAUG(gs + 9) = TAUG(gs + 9) * 1.1;
        TAUG(gs + 10) = TAUG(gs + 10) * 0.99;
        TAUG(gs + 11) = TAUG(gs + 11) * 0.855;
    }
}

/* band 8: :1607-1728 */
static void taugb8(Col *c) {
    const LwBand *B = &g_lw.b[7];
    const int ng = B->ng, gs = g_lw.ngs[6];
    for (int lay = 1; lay <= c->laytrop; ++lay) {
        double adjcolco2 = adjcol(A(colco2), A(coldry), CHIM(2, A(jp) + 1), 3.0, 2.0, 0.65);
        int ind0 = IND0_LO(B->nspa) + 1, ind1 = IND1_LO(B->nspa) + 1;
        int indm = A(indminor);
        for (int ig = 1; ig <= ng; ++ig) {
            double tauself = TAUSELF(B, ig);
            double taufor = TAUFOR(B, ig);
            double absco2 = lerp1(B->ka_mco2, 19, indm, ig, A(minorfrac));
            double abso3 = lerp1(B->ka_mo3, 19, indm, ig, A(minorfrac));
            double absn2o = lerp1(B->ka_mn2o, 19, indm, ig, A(minorfrac));
            TAUG(gs + ig) = A(colh2o) * KEY4(ABSA_G(B, ig), ind0, ind1) + tauself + taufor +
                            adjcolco2 * absco2 + A(colo3) * abso3 + A(coln2o) * absn2o +
                            A(colcfc12) * B->cfc12[ig - 1] + A(colcfc22) * B->cfc22adj[ig - 1];
            PFR(gs + ig) = B->fracrefa[ig - 1];
        }
    }
    for (int lay = c->laytrop + 1; lay <= c->nlay; ++lay) {
        double adjcolco2 = adjcol(A(colco2), A(coldry), CHIM(2, A(jp) + 1), 3.0, 2.0, 0.65);
        int ind0 = IND0_UP(B->nspb) + 1, ind1 = IND1_UP(B->nspb) + 1;
        int indm = A(indminor);
        for (int ig = 1; ig <= ng; ++ig) {
            double absco2 = lerp1(B->kb_mco2, 19, indm, ig, A(minorfrac));
            double absn2o = lerp1(B->kb_mn2o, 19, indm, ig, A(minorfrac));
            TAUG(gs + ig) = A(colo3) * KEY4(ABSB_G(B, ig), ind0, ind1) + adjcolco2 * absco2 +
                            A(coln2o) * absn2o + A(colcfc12) * B->cfc12[ig - 1] +
                            A(colcfc22) * B->cfc22adj[ig - 1];
            PFR(gs + ig) = B->fracrefb[ig - 1];
        }
    }
}

/* band 9: :1732-1994 */
static void taugb9(Col *c) {
    const LwBand *B = &g_lw.b[8];
    const int ng = B->ng, gs = g_lw.ngs[7];
    double refrat_planck_a = CHIM(1, 9) / CHIM(6, 9);
    double refrat_m_a = CHIM(1, 3) / CHIM(6, 3);
    for (int lay = 1; lay <= c->laytrop; ++lay) {
        Spec s0 = spec(A(colh2o), A(rat_h2och4), A(colch4), 8.);
        Spec s1 = spec(A(colh2o), A(rat_h2och4_1), A(colch4), 8.);
        Spec sm = spec(A(colh2o), refrat_m_a, A(colch4), 8.);
        double adjcoln2o = adjcol(A(coln2o), A(coldry), CHIM(4, A(jp) + 1), 1.5, 0.5, 0.65);
        Spec sp = spec(A(colh2o), refrat_planck_a, A(colch4), 8.);
        int ind0 = IND0_LO(B->nspa) + s0.js, ind1 = IND1_LO(B->nspa) + s1.js;
        int indm = A(indminor);
        Stencil st0 = stencil_lower(s0.specparm, s0.fs, A(fac00), A(fac10));
        Stencil st1 = stencil_lower(s1.specparm, s1.fs, A(fac01), A(fac11));
        for (int ig = 1; ig <= ng; ++ig) {
            double tauself = TAUSELF(B, ig);
            double taufor = TAUFOR(B, ig);
            double absn2o = minor2(B->ka_mn2o, 9, sm.js, indm, ig, sm.fs, A(minorfrac));
            double tau_major = s0.speccomb * stencil_sum(&st0, ABSA_G(B, ig), ind0);
            double tau_major1 = s1.speccomb * stencil_sum(&st1, ABSA_G(B, ig), ind1);
            TAUG(gs + ig) = tau_major + tau_major1 + tauself + taufor + adjcoln2o * absn2o;
            PFR(gs + ig) = PFRAC2(B->fracrefa, ng, ig, sp.js, sp.fs);
        }
    }
    for (int lay = c->laytrop + 1; lay <= c->nlay; ++lay) {
        double adjcoln2o = adjcol(A(coln2o), A(coldry), CHIM(4, A(jp) + 1), 1.5, 0.5, 0.65);
        int ind0 = IND0_UP(B->nspb) + 1, ind1 = IND1_UP(B->nspb) + 1;
        int indm = A(indminor);
        for (int ig = 1; ig <= ng; ++ig) {
            double absn2o = lerp1(B->kb_mn2o, 19, indm, ig, A(minorfrac));
            TAUG(gs + ig) = A(colch4) * KEY4(ABSB_G(B, ig), ind0, ind1) + adjcoln2o * absn2o;
            PFR(gs + ig) = B->fracrefb[ig - 1];
        }
    }
}

/* band 10: :1998-2066 */
static void taugb10(Col *c) {
    const LwBand *B = &g_lw.b[9];
    const int ng = B->ng, gs = g_lw.ngs[8];
    for (int lay = 1; lay <= c->laytrop; ++lay) {
        int ind0 = IND0_LO(B->nspa) + 1, ind1 = IND1_LO(B->nspa) + 1;
        for (int ig = 1; ig <= ng; ++ig) {
            double tauself = TAUSELF(B, ig);
            double taufor = TAUFOR(B, ig);
            TAUG(gs + ig) = A(colh2o) * KEY4(ABSA_G(B, ig), ind0, ind1) + tauself + taufor;
            PFR(gs + ig) = B->fracrefa[ig - 1];
        }
    }
    for (int lay = c->laytrop + 1; lay <= c->nlay; ++lay) {
        int ind0 = IND0_UP(B->nspb) + 1, ind1 = IND1_UP(B->nspb) + 1;
        for (int ig = 1; ig <= ng; ++ig) {
            double taufor = TAUFOR(B, ig);
            TAUG(gs + ig) = A(colh2o) * KEY4(ABSB_G(B, ig), ind0, ind1) + taufor;
            PFR(gs + ig) = B->fracrefb[ig - 1];
        }
    }
}

/* band 11: :2070-2149 */
static void taugb11(Col *c) {
    const LwBand *B = &g_lw.b[10];
    const int ng = B->ng, gs = g_lw.ngs[9];
    for (int lay = 1; lay <= c->laytrop; ++lay) {
        int ind0 = IND0_LO(B->nspa) + 1, ind1 = IND1_LO(B->nspa) + 1;
        int indm = A(indminor);
        double scaleo2 = A(colo2) * A(scaleminor);
        for (int ig = 1; ig <= ng; ++ig) {
            double tauself = TAUSELF(B, ig);
            double taufor = TAUFOR(B, ig);
            double tauo2 = scaleo2 * lerp1(B->ka_mo2, 19, indm, ig, A(minorfrac));
            TAUG(gs + ig) = A(colh2o) * KEY4(ABSA_G(B, ig), ind0, ind1) + tauself + taufor + tauo2;
            PFR(gs + ig) = B->fracrefa[ig - 1];
        }
    }
    for (int lay = c->laytrop + 1; lay <= c->nlay; ++lay) {
        int ind0 = IND0_UP(B->nspb) + 1, ind1 = IND1_UP(B->nspb) + 1;
        int indm = A(indminor);
        double scaleo2 = A(colo2) * A(scaleminor);
        for (int ig = 1; ig <= ng; ++ig) {
            double taufor = TAUFOR(B, ig);
            double tauo2 = scaleo2 * lerp1(B->kb_mo2, 19, indm, ig, A(minorfrac));
            TAUG(gs + ig) = A(colh2o) * KEY4(ABSB_G(B, ig), ind0, ind1) + taufor + tauo2;
            PFR(gs + ig) = B->fracrefb[ig - 1];
        }
    }
}

/* band 12: :2153-2356 */
static void taugb12(Col *c) {
    const LwBand *B = &g_lw.b[11];
    const int ng = B->ng, gs = g_lw.ngs[10];
    double refrat_planck_a = CHIM(1, 10) / CHIM(2, 10);
    for (int lay = 1; lay <= c->laytrop; ++lay) {
        Spec s0 = spec(A(colh2o), A(rat_h2oco2), A(colco2), 8.);
        Spec s1 = spec(A(colh2o), A(rat_h2oco2_1), A(colco2), 8.);
        Spec sp = spec(A(colh2o), refrat_planck_a, A(colco2), 8.);
        int ind0 = IND0_LO(B->nspa) + s0.js, ind1 = IND1_LO(B->nspa) + s1.js;
        Stencil st0 = stencil_lower(s0.specparm, s0.fs, A(fac00), A(fac10));
        Stencil st1 = stencil_lower(s1.specparm, s1.fs, A(fac01), A(fac11));
        for (int ig = 1; ig <= ng; ++ig) {
            double tauself = TAUSELF(B, ig);
            double taufor = TAUFOR(B, ig);
            double tau_major = s0.speccomb * stencil_sum(&st0, ABSA_G(B, ig), ind0);
            double tau_major1 = s1.speccomb * stencil_sum(&st1, ABSA_G(B, ig), ind1);
            TAUG(gs + ig) = tau_major + tau_major1 + tauself + taufor;
            PFR(gs + ig) = PFRAC2(B->fracrefa, ng, ig, sp.js, sp.fs);
        }
    }
    for (int lay = c->laytrop + 1; lay <= c->nlay; ++lay)
        for (int ig = 1; ig <= ng; ++ig) { TAUG(gs + ig) = 0.0; PFR(gs + ig) = 0.0; }
}

/* band 13: :2360-2620 */
static void taugb13(Col *c) {
    const LwBand *B = &g_lw.b[12];
    const int ng = B->ng, gs = g_lw.ngs[11];
    double refrat_planck_a = CHIM(1, 5) / CHIM(4, 5);
    double refrat_m_a = CHIM(1, 1) / CHIM(4, 1);
    double refrat_m_a3 = CHIM(1, 3) / CHIM(4, 3);
    for (int lay = 1; lay <= c->laytrop; ++lay) {
        Spec s0 = spec(A(colh2o), A(rat_h2on2o), A(coln2o), 8.);
        Spec s1 = spec(A(colh2o), A(rat_h2on2o_1), A(coln2o), 8.);
        Spec smco2 = spec(A(colh2o), refrat_m_a, A(coln2o), 8.);
        double adjcolco2 = adjcol(A(colco2), A(coldry), 3.55e-4, 3.0, 2.0, 0.68);
        Spec smco = spec(A(colh2o), refrat_m_a3, A(coln2o), 8.);
        Spec sp = spec(A(colh2o), refrat_planck_a, A(coln2o), 8.);
        int ind0 = IND0_LO(B->nspa) + s0.js, ind1 = IND1_LO(B->nspa) + s1.js;
        int indm = A(indminor);
        Stencil st0 = stencil_lower(s0.specparm, s0.fs, A(fac00), A(fac10));
        Stencil st1 = stencil_lower(s1.specparm, s1.fs, A(fac01), A(fac11));
        for (int ig = 1; ig <= ng; ++ig) {
            double tauself = TAUSELF(B, ig);
            double taufor = TAUFOR(B, ig);
            double absco2 = minor2(B->ka_mco2, 9, smco2.js, indm, ig, smco2.fs, A(minorfrac));
            double absco = minor2(B->ka_mco, 9, smco.js, indm, ig, smco.fs, A(minorfrac));
            double tau_major = s0.speccomb * stencil_sum(&st0, ABSA_G(B, ig), ind0);
            double tau_major1 = s1.speccomb * stencil_sum(&st1, ABSA_G(B, ig), ind1);
            TAUG(gs + ig) = tau_major + tau_major1 + tauself + taufor + adjcolco2 * absco2 +
                            A(colco) * absco;
            PFR(gs + ig) = PFRAC2(B->fracrefa, ng, ig, sp.js, sp.fs);
        }
    }
    for (int lay = c->laytrop + 1; lay <= c->nlay; ++lay) {
        int indm = A(indminor);
        for (int ig = 1; ig <= ng; ++ig) {
            double abso3 = lerp1(B->kb_mo3, 19, indm, ig, A(minorfrac));
            TAUG(gs + ig) = A(colo3) * abso3;
            PFR(gs + ig) = B->fracrefb[ig - 1];
        }
    }
}

/* band 14: :2624-2686 */
static void taugb14(Col *c) {
    const LwBand *B = &g_lw.b[13];
    const int ng = B->ng, gs = g_lw.ngs[12];
    for (int lay = 1; lay <= c->laytrop; ++lay) {
        int ind0 = IND0_LO(B->nspa) + 1, ind1 = IND1_LO(B->nspa) + 1;
        for (int ig = 1; ig <= ng; ++ig) {
            double tauself = TAUSELF(B, ig);
            double taufor = TAUFOR(B, ig);
            TAUG(gs + ig) = A(colco2) * KEY4(ABSA_G(B, ig), ind0, ind1) + tauself + taufor;
            PFR(gs + ig) = B->fracrefa[ig - 1];
        }
    }
    for (int lay = c->laytrop + 1; lay <= c->nlay; ++lay) {
        int ind0 = IND0_UP(B->nspb) + 1, ind1 = IND1_UP(B->nspb) + 1;
        for (int ig = 1; ig <= ng; ++ig) {
            TAUG(gs + ig) = A(colco2) * KEY4(ABSB_G(B, ig), ind0, ind1);
            PFR(gs + ig) = B->fracrefb[ig - 1];
        }
    }
}

/* band 15: :2690-2913 */
static void taugb15(Col *c) {
    const LwBand *B = &g_lw.b[14];
    const int ng = B->ng, gs = g_lw.ngs[13];
    double refrat_planck_a = CHIM(4, 1) / CHIM(2, 1);
    double refrat_m_a = refrat_planck_a;
    for (int lay = 1; lay <= c->laytrop; ++lay) {
        Spec s0 = spec(A(coln2o), A(rat_n2oco2), A(colco2), 8.);
        Spec s1 = spec(A(coln2o), A(rat_n2oco2_1), A(colco2), 8.);
        Spec sm = spec(A(coln2o), refrat_m_a, A(colco2), 8.);
        Spec sp = spec(A(coln2o), refrat_planck_a, A(colco2), 8.);
        int ind0 = IND0_LO(B->nspa) + s0.js, ind1 = IND1_LO(B->nspa) + s1.js;
        int indm = A(indminor);
        double scalen2 = A(colbrd) * A(scaleminor);
        Stencil st0 = stencil_lower(s0.specparm, s0.fs, A(fac00), A(fac10));
        Stencil st1 = stencil_lower(s1.specparm, s1.fs, A(fac01), A(fac11));
        for (int ig = 1; ig <= ng; ++ig) {
            double tauself = TAUSELF(B, ig);
            double taufor = TAUFOR(B, ig);
            double taun2 = scalen2 * minor2(B->ka_mn2, 9, sm.js, indm, ig, sm.fs, A(minorfrac));
            double tau_major = s0.speccomb * stencil_sum(&st0, ABSA_G(B, ig), ind0);
            double tau_major1 = s1.speccomb * stencil_sum(&st1, ABSA_G(B, ig), ind1);
            TAUG(gs + ig) = tau_major + tau_major1 + tauself + taufor + taun2;
            PFR(gs + ig) = PFRAC2(B->fracrefa, ng, ig, sp.js, sp.fs);
        }
    }
    for (int lay = c->laytrop + 1; lay <= c->nlay; ++lay)
        for (int ig = 1; ig <= ng; ++ig) { TAUG(gs + ig) = 0.0; PFR(gs + ig) = 0.0; }
}

/* band 16: :2917-3126 */
static void taugb16(Col *c) {
    const LwBand *B = &g_lw.b[15];
    const int ng = B->ng, gs = g_lw.ngs[14];
    double refrat_planck_a = CHIM(1, 6) / CHIM(6, 6);
    for (int lay = 1; lay <= c->laytrop; ++lay) {
        Spec s0 = spec(A(colh2o), A(rat_h2och4), A(colch4), 8.);
        Spec s1 = spec(A(colh2o), A(rat_h2och4_1), A(colch4), 8.);
        Spec sp = spec(A(colh2o), refrat_planck_a, A(colch4), 8.);
        int ind0 = IND0_LO(B->nspa) + s0.js, ind1 = IND1_LO(B->nspa) + s1.js;
        Stencil st0 = stencil_lower(s0.specparm, s0.fs, A(fac00), A(fac10));
        Stencil st1 = stencil_lower(s1.specparm, s1.fs, A(fac01), A(fac11));
        for (int ig = 1; ig <= ng; ++ig) {
            double tauself = TAUSELF(B, ig);
            double taufor = TAUFOR(B, ig);
            double tau_major = s0.speccomb * stencil_sum(&st0, ABSA_G(B, ig), ind0);
            double tau_major1 = s1.speccomb * stencil_sum(&st1, ABSA_G(B, ig), ind1);
            TAUG(gs + ig) = tau_major + tau_major1 + tauself + taufor;
            PFR(gs + ig) = PFRAC2(B->fracrefa, ng, ig, sp.js, sp.fs);
        }
    }
    for (int lay = c->laytrop + 1; lay <= c->nlay; ++lay) {
        int ind0 = IND0_UP(B->nspb) + 1, ind1 = IND1_UP(B->nspb) + 1;
        for (int ig = 1; ig <= ng; ++ig) {
            TAUG(gs + ig) = A(colch4) * KEY4(ABSB_G(B, ig), ind0, ind1);
            PFR(gs + ig) = B->fracrefb[ig - 1];
        }
    }
}

/* taumol :155-187 + addAerosols :3130-3146 */
static void taumol(const SetCoef *s, int ncol, int nlay, const double *pavel, const double *taua,
                   double *taug, double *pfracs) {
    for (int icol = 1; icol <= ncol; ++icol) {
        size_t o = (size_t)nlay * (icol - 1);
        Col c;
        c.nlay = nlay; c.laytrop = s->laytrop[icol - 1];
        c.jp = s->jp + o; c.jt = s->jt + o; c.jt1 = s->jt1 + o; c.indfor = s->indfor + o;
        c.indself = s->indself + o; c.indminor = s->indminor + o;
        c.fac00 = s->fac00 + o; c.fac01 = s->fac01 + o; c.fac10 = s->fac10 + o; c.fac11 = s->fac11 + o;
        c.colh2o = s->colh2o + o; c.colco2 = s->colco2 + o; c.colo3 = s->colo3 + o;
        c.coln2o = s->coln2o + o; c.colch4 = s->colch4 + o; c.colo2 = s->colo2 + o;
        c.colco = s->colco + o; c.colbrd = s->colbrd + o; c.colcfc11 = s->colcfc11 + o;
        c.colcfc12 = s->colcfc12 + o; c.colcfc22 = s->colcfc22 + o; c.colccl4 = s->colccl4 + o;
        c.coldry = s->coldry + o; c.forfac = s->forfac + o; c.forfrac = s->forfrac + o;
        c.selffac = s->selffac + o; c.selffrac = s->selffrac + o; c.scaleminor = s->scaleminor + o;
        c.scaleminorn2 = s->scaleminorn2 + o; c.minorfrac = s->minorfrac + o;
        c.rat_h2oco2 = s->rat_h2oco2 + o; c.rat_h2oco2_1 = s->rat_h2oco2_1 + o;
        c.rat_h2oo3 = s->rat_h2oo3 + o; c.rat_h2oo3_1 = s->rat_h2oo3_1 + o;
        c.rat_h2on2o = s->rat_h2on2o + o; c.rat_h2on2o_1 = s->rat_h2on2o_1 + o;
        c.rat_h2och4 = s->rat_h2och4 + o; c.rat_h2och4_1 = s->rat_h2och4_1 + o;
        c.rat_n2oco2 = s->rat_n2oco2 + o; c.rat_n2oco2_1 = s->rat_n2oco2_1 + o;
        c.rat_o3co2 = s->rat_o3co2 + o; c.rat_o3co2_1 = s->rat_o3co2_1 + o;
        c.pavel = pavel + o;
        c.taug = taug + (size_t)nlay * NG * (icol - 1);
        c.pfracs = pfracs + (size_t)nlay * NG * (icol - 1);
        taugb1(&c); taugb2(&c); taugb3(&c); taugb4(&c); taugb5(&c); taugb6(&c); taugb7(&c);
        taugb8(&c); taugb9(&c); taugb10(&c); taugb11(&c); taugb12(&c); taugb13(&c); taugb14(&c);
        taugb15(&c); taugb16(&c);
        /* addAerosols: taua(nlay,nbndlw,ncol) */
        for (int ig = 1; ig <= NG; ++ig) {
            int ibnd = g_lw.ngb[ig - 1];
            for (int lay = 1; lay <= nlay; ++lay)
                c.taug[(lay - 1) + (size_t)nlay * (ig - 1)] =
                    c.taug[(lay - 1) + (size_t)nlay * (ig - 1)] +
                    taua[(lay - 1) + (size_t)nlay * ((ibnd - 1) + (size_t)NB * (icol - 1))];
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * cldprmc, LW/src/rrtmg_lw_cldprmc.F90:24-385
 * ---------------------------------------------------------------------------------------- */
#define G3(lay, ig, icol) ((size_t)((lay)-1) + (size_t)nlay * (((ig)-1) + (size_t)NG * ((icol)-1)))

static int lookup_index(double factor, int hi, int *index) {
    /* the index clamp / extrapolation traps shared by iceflag 2,3,4 and liqflag 1 */
    int idx = f_int(factor);
    if (idx >= hi) {
        if (idx == hi) idx = hi - 1;
        else return -1; /* 'excessive high-radius extrapolation forbidden!' */
    } else if (idx <= 0) {
        if (idx == 0) idx = 1;
        else return -2; /* 'excessive low-radius extrapolation forbidden!' */
    }
    *index = idx;
    return 0;
}

static int cldprmc(int ncol, int nlay, const unsigned char *cldymc, const double *ciwpmc,
                   const double *clwpmc, const double *reice, const double *reliq, int iceflag,
                   int liqflag, double *taucmc, unsigned char *cloudy) {
    const LwTables *T = &g_lw;
    static const int ice1b[16] = {1, 2, 3, 3, 3, 4, 4, 4, 5, 5, 5, 5, 5, 5, 5, 5};
    for (int icol = 1; icol <= ncol; ++icol)
        for (int ilay = 1; ilay <= nlay; ++ilay) {
            cloudy[IX(ilay, icol)] = 0;
            for (int ig = 1; ig <= NG; ++ig)
                if (cldymc[G3(ilay, ig, icol)]) { cloudy[IX(ilay, icol)] = 1; break; }
        }
    memset(taucmc, 0, sizeof(double) * (size_t)nlay * NG * ncol);

    if (iceflag < 0 || iceflag > 4) return -41; /* 'cldprmc: invalid iceflag' */
    for (int icol = 1; icol <= ncol; ++icol)
        for (int ilay = 1; ilay <= nlay; ++ilay) {
            if (!cloudy[IX(ilay, icol)]) continue;
            double re = reice[IX(ilay, icol)];
            int index = 0;
            double fint = 0., abscoice0 = 0.;
            const double *tab = NULL;
            int lead = 0;
            if (iceflag == 0) {
                abscoice0 = T->absice0[0] + T->absice0[1] / re;
            } else if (iceflag == 2) {
                double factor = (re - 2.) / 3.;
                int rc = lookup_index(factor, 43, &index);
                if (rc) return -42 + rc;
                fint = factor - (double)index; tab = T->absice2; lead = 43;
            } else if (iceflag == 3) {
                double factor = (re - 2.) / 3.;
                int rc = lookup_index(factor, 46, &index);
                if (rc) return -45 + rc;
                fint = factor - (double)index; tab = T->absice3; lead = 46;
            } else if (iceflag == 4) {
                double factor = re;
                int rc = lookup_index(factor, 200, &index);
                if (rc) return -48 + rc;
                fint = factor - (double)index; tab = T->absice4; lead = 200;
            }
            for (int ig = 1; ig <= NG; ++ig) {
                size_t k = G3(ilay, ig, icol);
                if (cldymc[k] && ciwpmc[k] > 0.) {
                    double abscoice;
                    if (iceflag == 0) {
                        abscoice = abscoice0;
                    } else if (iceflag == 1) {
                        int ib = ice1b[T->ngb[ig - 1] - 1];
                        abscoice = T->absice1[0 + 2 * (ib - 1)] + T->absice1[1 + 2 * (ib - 1)] / re;
                    } else {
                        int ib = T->ngb[ig - 1];
                        const double *cb = tab + (size_t)lead * (ib - 1);
                        abscoice = cb[index - 1] + fint * (cb[index] - (cb[index - 1]));
                    }
                    taucmc[k] = ciwpmc[k] * abscoice;
                }
            }
        }

    if (liqflag != 1) return -51; /* 'cldprmc: invalid liqflag' */
    for (int icol = 1; icol <= ncol; ++icol)
        for (int ilay = 1; ilay <= nlay; ++ilay) {
            if (!cloudy[IX(ilay, icol)]) continue;
            double factor = reliq[IX(ilay, icol)] - 1.5;
            int index;
            int rc = lookup_index(factor, 58, &index);
            if (rc) return -52 + rc;
            double fint = factor - (double)index;
            for (int ig = 1; ig <= NG; ++ig) {
                size_t k = G3(ilay, ig, icol);
                if (cldymc[k] && clwpmc[k] > 0.) {
                    int ib = T->ngb[ig - 1];
                    const double *cb = T->absliq1 + (size_t)58 * (ib - 1);
                    double abscoliq = cb[index - 1] + fint * (cb[index] - (cb[index - 1]));
                    taucmc[k] = taucmc[k] + clwpmc[k] * abscoliq;
                }
            }
        }

    /* refine cloudy flag to OPTICALLY cloudy, :371-383 */
    for (int icol = 1; icol <= ncol; ++icol)
        for (int ilay = 1; ilay <= nlay; ++ilay)
            if (cloudy[IX(ilay, icol)]) {
                int any = 0;
                for (int ig = 1; ig <= NG; ++ig)
                    if (taucmc[G3(ilay, ig, icol)] > 0.) { any = 1; break; }
                if (!any) cloudy[IX(ilay, icol)] = 0;
            }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * rtrnmc, LW/src/rrtmg_lw_rtrnmc.F90:27-390
 * ---------------------------------------------------------------------------------------- */
static void rtrnmc(const SetCoef *s, int ncol, int nlay, int dudTs, const double *semiss,
                   const double *taug, const double *pfracs, const unsigned char *cloudy,
                   const double *taucmc, double *totuflux, double *totdflux, double *totuclfl,
                   double *totdclfl, double *dtotuflux_dTs, double *dtotuclfl_dTs,
                   const int *band_output, double *olrb, double *dolrb_dTs) {
    const LwTables *T = &g_lw;
    const double wtdiff = 0.5;
    const double tblint = 10000.0;
    static const double a0[16] = {1.66, 1.55, 1.58, 1.66, 1.54, 1.454, 1.89, 1.33,
                                  1.668, 1.66, 1.66, 1.66, 1.66, 1.66, 1.66, 1.66};
    static const double a1[16] = {0.00, 0.25, 0.22, 0.00, 0.13, 0.446, -0.10, 0.40,
                                  -0.006, 0.00, 0.00, 0.00, 0.00, 0.00, 0.00, 0.00};
    static const double a2[16] = {0.00, -12.0, -11.7, 0.00, -0.72, -0.243, 0.19, -0.062,
                                  0.414, 0.00, 0.00, 0.00, 0.00, 0.00, 0.00, 0.00};
    double *agas = (double *)malloc(sizeof(double) * 4 * (size_t)nlay);
    double *atot = agas + nlay, *bbugas = agas + 2 * nlay, *bbutot = agas + 3 * nlay;
    size_t nlev = (size_t)(nlay + 1) * ncol;
    for (size_t i = 0; i < nlev; ++i) { totuflux[i] = 0.; totdflux[i] = 0.; totuclfl[i] = 0.; totdclfl[i] = 0.; }
    if (dudTs) for (size_t i = 0; i < nlev; ++i) { dtotuflux_dTs[i] = 0.; dtotuclfl_dTs[i] = 0.; }
    int any_bo = 0;
    for (int b = 0; b < 16; ++b) any_bo |= band_output[b];
    if (any_bo) {
        for (int i = 0; i < 16 * ncol; ++i) olrb[i] = 0.;
        if (dudTs) for (int i = 0; i < 16 * ncol; ++i) dolrb_dTs[i] = 0.;
    }

    for (int icol = 1; icol <= ncol; ++icol)
        for (int ig = 1; ig <= NG; ++ig) {
            int ibnd = T->ngb[ig - 1];
            double sumfac = wtdiff * T->delwave[ibnd - 1] * T->fluxfac;
            double secdiff;
            if (ibnd == 1 || ibnd == 4 || ibnd >= 10) {
                secdiff = 1.66;
            } else {
                secdiff = a0[ibnd - 1] + a1[ibnd - 1] * exp(a2[ibnd - 1] * s->pwvcm[icol - 1]);
                if (secdiff > 1.80) secdiff = 1.80;
                else if (secdiff < 1.50) secdiff = 1.50;
            }
            double radld = 0., radclrd = 0.;
            int down_streams_diverge = 0;
            for (int lev = nlay; lev >= 1; --lev) {
                size_t k = G3(lev, ig, icol);
                double plfrac = pfracs[k];
                double blay = PLAY(ibnd, lev, icol);
                double dplankup = PLEV(ibnd, lev, icol) - blay;
                double dplankdn = PLEV(ibnd, lev - 1, icol) - blay;
                double odepth = secdiff * taug[k];
                if (odepth < 0.) odepth = 0.;
                double tblind = odepth / (T->bpade + odepth);
                int itgas = f_int(tblint * tblind + 0.5);
                agas[lev - 1] = 1. - T->exp_tbl[itgas];
                double tfacgas = T->tfn_tbl[itgas];
                double bbdgas = plfrac * (blay + tfacgas * dplankdn);
                bbugas[lev - 1] = plfrac * (blay + tfacgas * dplankup);
                if (taucmc[k] <= 0.) {
                    radld = radld + (bbdgas - radld) * agas[lev - 1];
                } else {
                    double odcld = secdiff * taucmc[k];
                    odepth = T->tau_tbl[itgas];
                    double odtot = odepth + odcld;
                    tblind = odtot / (T->bpade + odtot);
                    int ittot = f_int(tblint * tblind + 0.5);
                    atot[lev - 1] = 1. - T->exp_tbl[ittot];
                    double tfactot = T->tfn_tbl[ittot];
                    double bbdtot = plfrac * (blay + tfactot * dplankdn);
                    bbutot[lev - 1] = plfrac * (blay + tfactot * dplankup);
                    radld = radld + (bbdtot - radld) * atot[lev - 1];
                }
                totdflux[IX0(lev - 1, icol)] = totdflux[IX0(lev - 1, icol)] + sumfac * radld;
                if (!down_streams_diverge)
                    if (cloudy[IX(lev, icol)]) down_streams_diverge = 1;
                if (down_streams_diverge)
                    radclrd = radclrd + (bbdgas - radclrd) * agas[lev - 1];
                else
                    radclrd = radld;
                totdclfl[IX0(lev - 1, icol)] = totdclfl[IX0(lev - 1, icol)] + sumfac * radclrd;
            }
            double rad0 = pfracs[G3(1, ig, icol)] * PBND(ibnd, icol);
            double d_rad0_dTs = 0., d_radlu_dTs = 0., d_radclru_dTs = 0.;
            if (dudTs) d_rad0_dTs = pfracs[G3(1, ig, icol)] * DPBND(ibnd, icol);
            double reflect = 1. - semiss[(ibnd - 1) + 16 * (icol - 1)];
            double radlu = rad0 + reflect * radld;
            double radclru = rad0 + reflect * radclrd;
            totuflux[IX0(0, icol)] = totuflux[IX0(0, icol)] + sumfac * radlu;
            totuclfl[IX0(0, icol)] = totuclfl[IX0(0, icol)] + sumfac * radclru;
            if (dudTs) {
                d_radlu_dTs = d_rad0_dTs;
                d_radclru_dTs = d_rad0_dTs;
                dtotuflux_dTs[IX0(0, icol)] = dtotuflux_dTs[IX0(0, icol)] + sumfac * d_radlu_dTs;
                dtotuclfl_dTs[IX0(0, icol)] = dtotuclfl_dTs[IX0(0, icol)] + sumfac * d_radclru_dTs;
            }
            double deluflux = 0., deluderiv = 0.;
            for (int lev = 1; lev <= nlay; ++lev) {
                size_t k = G3(lev, ig, icol);
                if (taucmc[k] <= 0.) {
                    radlu = radlu + (bbugas[lev - 1] - radlu) * agas[lev - 1];
                    if (dudTs) d_radlu_dTs = d_radlu_dTs - d_radlu_dTs * agas[lev - 1];
                } else {
                    radlu = radlu + (bbutot[lev - 1] - radlu) * atot[lev - 1];
                    if (dudTs) d_radlu_dTs = d_radlu_dTs - d_radlu_dTs * atot[lev - 1];
                }
                deluflux = sumfac * radlu;
                totuflux[IX0(lev, icol)] = totuflux[IX0(lev, icol)] + deluflux;
                if (down_streams_diverge)
                    radclru = radclru + (bbugas[lev - 1] - radclru) * agas[lev - 1];
                else
                    radclru = radlu;
                totuclfl[IX0(lev, icol)] = totuclfl[IX0(lev, icol)] + sumfac * radclru;
                if (dudTs) {
                    if (down_streams_diverge)
                        d_radclru_dTs = d_radclru_dTs - d_radclru_dTs * agas[lev - 1];
                    else
                        d_radclru_dTs = d_radlu_dTs;
                    deluderiv = sumfac * d_radlu_dTs;
                    dtotuflux_dTs[IX0(lev, icol)] = dtotuflux_dTs[IX0(lev, icol)] + deluderiv;
                    dtotuclfl_dTs[IX0(lev, icol)] = dtotuclfl_dTs[IX0(lev, icol)] + sumfac * d_radclru_dTs;
                }
            }
            if (band_output[ibnd - 1]) {
                olrb[(ibnd - 1) + 16 * (icol - 1)] = olrb[(ibnd - 1) + 16 * (icol - 1)] + deluflux;
                if (dudTs)
                    dolrb_dTs[(ibnd - 1) + 16 * (icol - 1)] = dolrb_dTs[(ibnd - 1) + 16 * (icol - 1)] + deluderiv;
            }
        }
    free(agas);
}

/* ------------------------------------------------------------------------------------------
 * rrtmg_lw_part, LW/src/rrtmg_lw_rad.F90:348-610
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    const double *play, *plev, *tlay, *tlev, *tsfc, *emis, *h2ovmr, *o3vmr, *co2vmr, *ch4vmr,
        *n2ovmr, *o2vmr, *cfc11vmr, *cfc12vmr, *cfc22vmr, *ccl4vmr, *cldf, *ciwp, *clwp, *rei,
        *rel, *tauaer, *zm, *alat;
} LwIn;

static int rrtmg_lw_part(int ncol, int colstart, int pncol, int nlay, int dudTs, const LwIn *in,
                         int iceflglw, int liqflglw, int dyofyr, int cloudLM, int cloudMH,
                         int *clearCounts, double *uflx, double *dflx, double *uflxc,
                         double *dflxc, double *duflx_dTs, double *duflxc_dTs,
                         const int *band_output, double *olrb, double *dolrb_dTs,
                         OracleTaps *taps) {
    size_t n2 = (size_t)nlay * pncol, n3 = (size_t)nlay * NG * pncol;
    double *buf = (double *)zalloc(sizeof(double) * (n2 * 21 + (size_t)(nlay + 1) * pncol * 8 +
                                                     (size_t)pncol * 2 + (size_t)16 * pncol * 3 +
                                                     (size_t)nlay * 16 * pncol + n3 * 5));
    double *p = buf;
#define TAKE(name, cnt) double *name = p; p += (cnt)
    TAKE(p_zm, n2); TAKE(p_play, n2); TAKE(p_tlay, n2); TAKE(p_cldf, n2); TAKE(p_ciwp, n2);
    TAKE(p_clwp, n2); TAKE(p_rei, n2); TAKE(p_rel, n2); TAKE(p_h2ovmr, n2); TAKE(p_o3vmr, n2);
    TAKE(p_co2vmr, n2); TAKE(p_ch4vmr, n2); TAKE(p_n2ovmr, n2); TAKE(p_o2vmr, n2);
    TAKE(p_covmr, n2); TAKE(p_cfc11vmr, n2); TAKE(p_cfc12vmr, n2); TAKE(p_cfc22vmr, n2);
    TAKE(p_ccl4vmr, n2); TAKE(spare1, n2); TAKE(spare2, n2);
    TAKE(p_plev, (size_t)(nlay + 1) * pncol); TAKE(p_tlev, (size_t)(nlay + 1) * pncol);
    TAKE(totuflux, (size_t)(nlay + 1) * pncol); TAKE(totdflux, (size_t)(nlay + 1) * pncol);
    TAKE(totuclfl, (size_t)(nlay + 1) * pncol); TAKE(totdclfl, (size_t)(nlay + 1) * pncol);
    TAKE(dtotuflux_dTs, (size_t)(nlay + 1) * pncol); TAKE(dtotuclfl_dTs, (size_t)(nlay + 1) * pncol);
    TAKE(p_alat, pncol); TAKE(p_tsfc, pncol);
    TAKE(p_emis, (size_t)16 * pncol); TAKE(p_olrb, (size_t)16 * pncol); TAKE(p_dolrb_dTs, (size_t)16 * pncol);
    TAKE(p_tauaer, (size_t)nlay * 16 * pncol);
    TAKE(taug, n3); TAKE(pfracs, n3); TAKE(ciwpmc, n3); TAKE(clwpmc, n3); TAKE(taucmc, n3);
#undef TAKE
    (void)spare1; (void)spare2;
    unsigned char *cldymc = (unsigned char *)zalloc(n3);
    unsigned char *cloudy = (unsigned char *)zalloc(n2);
    int *p_clearCounts = (int *)zalloc(sizeof(int) * 4 * pncol);
    int rc = 0;

    /* copy partition and reorder (:502-537); colstart is 1-based */
    for (int j = 0; j < pncol; ++j) {
        int gc = colstart - 1 + j;
        p_alat[j] = in->alat[gc];
        p_tsfc[j] = in->tsfc[gc];
        for (int l = 0; l < nlay; ++l) {
            size_t d = (size_t)l + (size_t)nlay * j, sidx = (size_t)gc + (size_t)ncol * l;
            p_zm[d] = in->zm[sidx]; p_play[d] = in->play[sidx]; p_tlay[d] = in->tlay[sidx];
            p_cldf[d] = in->cldf[sidx]; p_ciwp[d] = in->ciwp[sidx]; p_clwp[d] = in->clwp[sidx];
            p_rei[d] = in->rei[sidx]; p_rel[d] = in->rel[sidx]; p_h2ovmr[d] = in->h2ovmr[sidx];
            p_o3vmr[d] = in->o3vmr[sidx]; p_co2vmr[d] = in->co2vmr[sidx]; p_ch4vmr[d] = in->ch4vmr[sidx];
            p_n2ovmr[d] = in->n2ovmr[sidx]; p_o2vmr[d] = in->o2vmr[sidx]; p_covmr[d] = 0.;
            p_cfc11vmr[d] = in->cfc11vmr[sidx]; p_cfc12vmr[d] = in->cfc12vmr[sidx];
            p_cfc22vmr[d] = in->cfc22vmr[sidx]; p_ccl4vmr[d] = in->ccl4vmr[sidx];
        }
        for (int l = 0; l <= nlay; ++l) {
            p_plev[(size_t)l + (size_t)(nlay + 1) * j] = in->plev[(size_t)gc + (size_t)ncol * l];
            p_tlev[(size_t)l + (size_t)(nlay + 1) * j] = in->tlev[(size_t)gc + (size_t)ncol * l];
        }
        for (int b = 0; b < 16; ++b) {
            p_emis[b + 16 * j] = in->emis[(size_t)gc + (size_t)ncol * b];
            for (int l = 0; l < nlay; ++l)
                p_tauaer[(size_t)l + (size_t)nlay * (b + (size_t)16 * j)] =
                    in->tauaer[(size_t)gc + (size_t)ncol * (l + (size_t)nlay * b)];
        }
    }

    static const int seed_order[4] = {1, 2, 3, 4};
    rc = oracle_generate_stochastic_clouds(pncol, pncol, NG, nlay, p_zm, p_alat, dyofyr, p_play,
                                           p_cldf, p_ciwp, p_clwp, 1.e-20, cldymc, ciwpmc, clwpmc,
                                           seed_order);
    if (!rc) rc = oracle_clearCounts_threeBand(pncol, pncol, NG, nlay, cloudLM, cloudMH, cldymc, p_clearCounts);
    if (!rc)
        for (int n = 0; n < 4; ++n)
            for (int j = 0; j < pncol; ++j)
                clearCounts[(size_t)(colstart - 1 + j) + (size_t)ncol * n] = p_clearCounts[n + 4 * j];
    if (!rc) rc = cldprmc(pncol, nlay, cldymc, ciwpmc, clwpmc, p_rei, p_rel, iceflglw, liqflglw, taucmc, cloudy);

    SetCoef sc;
    setcoef_alloc(&sc, nlay, pncol);
    if (!rc)
        rc = setcoef(&sc, pncol, nlay, 1, dudTs, p_play, p_tlay, p_plev, p_tlev, p_tsfc, p_emis,
                     p_h2ovmr, p_o3vmr, p_co2vmr, p_ch4vmr, p_n2ovmr, p_o2vmr, p_covmr, p_cfc11vmr,
                     p_cfc12vmr, p_cfc22vmr, p_ccl4vmr);
    if (!rc) {
        taumol(&sc, pncol, nlay, p_play, p_tauaer, taug, pfracs);
        rtrnmc(&sc, pncol, nlay, dudTs, p_emis, taug, pfracs, cloudy, taucmc, totuflux, totdflux,
               totuclfl, totdclfl, dtotuflux_dTs, dtotuclfl_dTs, band_output, p_olrb, p_dolrb_dTs);
        /* copy the partitioned fluxes back (:586-605) */
        for (int j = 0; j < pncol; ++j) {
            int gc = colstart - 1 + j;
            for (int l = 0; l <= nlay; ++l) {
                size_t d = (size_t)gc + (size_t)ncol * l, sidx = (size_t)l + (size_t)(nlay + 1) * j;
                uflx[d] = totuflux[sidx]; dflx[d] = totdflux[sidx];
                uflxc[d] = totuclfl[sidx]; dflxc[d] = totdclfl[sidx];
                if (dudTs) { duflx_dTs[d] = dtotuflux_dTs[sidx]; duflxc_dTs[d] = dtotuclfl_dTs[sidx]; }
            }
            for (int b = 0; b < 16; ++b)
                if (band_output[b]) {
                    olrb[b + (size_t)16 * gc] = p_olrb[b + 16 * j];
                    if (dudTs) dolrb_dTs[b + (size_t)16 * gc] = p_dolrb_dTs[b + 16 * j];
                }
        }
        if (taps) {
            for (int j = 0; j < pncol; ++j) {
                int gc = colstart - 1 + j;
                if (taps->laytrop) taps->laytrop[gc] = sc.laytrop[j];
                if (taps->pwvcm) taps->pwvcm[gc] = sc.pwvcm[j];
                for (int l = 0; l < nlay; ++l) {
                    size_t d = (size_t)gc + (size_t)ncol * l, sidx = (size_t)l + (size_t)nlay * j;
                    if (taps->jp) taps->jp[d] = sc.jp[sidx];
                    if (taps->jt) taps->jt[d] = sc.jt[sidx];
                    if (taps->jt1) taps->jt1[d] = sc.jt1[sidx];
                    if (taps->indfor) taps->indfor[d] = sc.indfor[sidx];
                    if (taps->indself) taps->indself[d] = sc.indself[sidx];
                    if (taps->indminor) taps->indminor[d] = sc.indminor[sidx];
                    if (taps->fac00) taps->fac00[d] = sc.fac00[sidx];
                    if (taps->fac01) taps->fac01[d] = sc.fac01[sidx];
                    if (taps->fac10) taps->fac10[d] = sc.fac10[sidx];
                    if (taps->fac11) taps->fac11[d] = sc.fac11[sidx];
                }
                size_t o3 = (size_t)nlay * NG * gc, s3 = (size_t)nlay * NG * j, cnt = (size_t)nlay * NG;
                if (taps->cldymc) memcpy(taps->cldymc + o3, cldymc + s3, cnt);
                if (taps->ciwpmc) memcpy(taps->ciwpmc + o3, ciwpmc + s3, cnt * 8);
                if (taps->clwpmc) memcpy(taps->clwpmc + o3, clwpmc + s3, cnt * 8);
                if (taps->taug) memcpy(taps->taug + o3, taug + s3, cnt * 8);
                if (taps->pfracs) memcpy(taps->pfracs + o3, pfracs + s3, cnt * 8);
                if (taps->taucmc) memcpy(taps->taucmc + o3, taucmc + s3, cnt * 8);
            }
        }
    }
    setcoef_free(&sc);
    free(buf); free(cldymc); free(cloudy); free(p_clearCounts);
    return rc;
}

static int any_negative(const double *x, size_t n) {
    for (size_t i = 0; i < n; ++i)
        if (x[i] < 0.) return 1;
    return 0;
}

/* LW/src/rrtmg_lw_rad.F90:15-344 */
int oracle_rrtmg_lw(
    int ncol, int nlay, int psize, int dudTs,
    const double *play, const double *plev, const double *tlay, const double *tlev,
    const double *tsfc, const double *emis,
    const double *h2ovmr, const double *o3vmr, const double *co2vmr, const double *ch4vmr,
    const double *n2ovmr, const double *o2vmr, const double *cfc11vmr, const double *cfc12vmr,
    const double *cfc22vmr, const double *ccl4vmr,
    const double *cldf, const double *ciwp, const double *clwp, const double *rei,
    const double *rel, int iceflglw, int liqflglw,
    const double *tauaer, const double *zm, const double *alat, int dyofyr,
    int cloudLM, int cloudMH, int *clearCounts,
    double *uflx, double *dflx, double *uflxc, double *dflxc,
    double *duflx_dTs, double *duflxc_dTs,
    const int *band_output, double *olrb, double *dolrb_dTs, OracleTaps *taps) {
    size_t n2 = (size_t)ncol * nlay, n2p = (size_t)ncol * (nlay + 1);
    /* input traps, :209-318, in the reference's order; code = -(100 + position) */
    struct { const double *x; size_t n; } chk[] = {
        {play, n2}, {plev, n2p}, {tlay, n2}, {tlev, n2p}, {tsfc, (size_t)ncol}, {h2ovmr, n2},
        {o3vmr, n2}, {co2vmr, n2}, {ch4vmr, n2}, {n2ovmr, n2}, {o2vmr, n2}, {cfc11vmr, n2},
        {cfc12vmr, n2}, {cfc22vmr, n2}, {ccl4vmr, n2}, {emis, (size_t)ncol * 16}, {cldf, n2},
        {ciwp, n2}, {clwp, n2}, {rei, n2}, {rel, n2}, {tauaer, n2 * 16}};
    for (size_t i = 0; i < sizeof chk / sizeof chk[0]; ++i)
        if (any_negative(chk[i].x, chk[i].n)) return -(101 + (int)i);

    LwIn in = {play, plev, tlay, tlev, tsfc, emis, h2ovmr, o3vmr, co2vmr, ch4vmr, n2ovmr, o2vmr,
               cfc11vmr, cfc12vmr, cfc22vmr, ccl4vmr, cldf, ciwp, clwp, rei, rel, tauaer, zm, alat};
    int nparts = (ncol + psize - 1) / psize; /* ceiling(real(ncol)/real(psize)) */
    int rc_all = 0;
#pragma omp parallel for schedule(dynamic)
    for (int n = 0; n < nparts; ++n) {
        int colstart = n * psize + 1;
        int pn = psize < ncol - n * psize ? psize : ncol - n * psize;
        int rc = rrtmg_lw_part(ncol, colstart, pn, nlay, dudTs, &in, iceflglw, liqflglw, dyofyr,
                               cloudLM, cloudMH, clearCounts, uflx, dflx, uflxc, dflxc, duflx_dTs,
                               duflxc_dTs, band_output, olrb, dolrb_dTs, taps);
        if (rc) {
#pragma omp critical
            if (!rc_all) rc_all = rc;
        }
    }
    return rc_all;
}
