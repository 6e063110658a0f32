/* oracle/lw_init.c -- CPU restatement of rrtmg_lw_ini (test infrastructure only).
 *
 * Follows LW/src/rrtmg_lw_init.F90: lookup tables (:96-113), relative g-point weights
 * (:120-144), lwdatinit constants (:193-234), lwcmbdat maps (:269-324) and the sixteen
 * cmbgbN reductions (:329-1978), which are all instances of one pattern: absorption data
 * are summed over each group of original g-points with weights rwgt in ascending order;
 * Planck fractions are summed unweighted.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "internal.h"

LwTables g_lw;
static const int *s_ngn, *s_ngm;
static const double *s_wt;

/* weighted (or unweighted) reduction of an array whose LAST dimension is the 16 original
 * g-points: dst(lead, ngc) = sum_ipr src(lead, ipr) * rwgt(ipr + 16*band) */
static double *reduce_last(const double *src, int lead, int ib, int weighted) {
    if (!src) return NULL;
    int ngc = g_lw.ngc[ib];
    int g0 = ib == 0 ? 0 : g_lw.ngs[ib - 1];
    double *dst = (double *)malloc(sizeof(double) * (size_t)lead * ngc);
    for (int l = 0; l < lead; ++l) {
        int iprsm = 0;
        for (int igc = 0; igc < ngc; ++igc) {
            double sumk = 0.;
            for (int ipr = 0; ipr < s_ngn[g0 + igc]; ++ipr) {
                if (weighted)
                    sumk = sumk + src[l + (size_t)lead * iprsm] * g_lw.rwgt[iprsm + 16 * ib];
                else
                    sumk = sumk + src[l + (size_t)lead * iprsm];
                iprsm++;
            }
            dst[l + (size_t)lead * igc] = sumk;
        }
    }
    return dst;
}

/* unweighted reduction of fracref?o(16, nj) -> fracref?(ngc, nj) (g is the FIRST dimension) */
static double *reduce_first(const double *src, int nj, int ib) {
    if (!src) return NULL;
    int ngc = g_lw.ngc[ib];
    int g0 = ib == 0 ? 0 : g_lw.ngs[ib - 1];
    double *dst = (double *)malloc(sizeof(double) * (size_t)nj * ngc);
    for (int j = 0; j < nj; ++j) {
        int iprsm = 0;
        for (int igc = 0; igc < ngc; ++igc) {
            double sumf = 0.;
            for (int ipr = 0; ipr < s_ngn[g0 + igc]; ++ipr) {
                sumf = sumf + src[iprsm + 16 * j];
                iprsm++;
            }
            dst[igc + (size_t)ngc * j] = sumf;
        }
    }
    return dst;
}

static const double *band_tab(int band1, const char *name, int *n) {
    char key[64];
    snprintf(key, sizeof key, "lw.kg%02d.%s", band1, name);
    return blob_f64(key, n);
}

int lw_init(void) {
    int n;
    memset(&g_lw, 0, sizeof g_lw);
    const int *ngc = blob_i32("lw.wvn.ngc", &n);
    const int *ngs = blob_i32("lw.wvn.ngs", &n);
    const int *ngb = blob_i32("lw.wvn.ngb", &n);
    const int *nspa = blob_i32("lw.wvn.nspa", &n);
    const int *nspb = blob_i32("lw.wvn.nspb", &n);
    s_ngn = blob_i32("lw.wvn.ngn", &n);
    s_ngm = blob_i32("lw.wvn.ngm", &n);
    s_wt = blob_f64("lw.wvn.wt", &n);
    if (!ngc || !ngs || !ngb || !s_ngn || !s_ngm || !s_wt || !nspa || !nspb) return -1;
    for (int i = 0; i < NBNDLW; ++i) { g_lw.ngc[i] = ngc[i]; g_lw.ngs[i] = ngs[i]; }
    for (int i = 0; i < NGPTLW; ++i) g_lw.ngb[i] = ngb[i];

    /* lwdatinit (:214,222), rrlw_con.F90:37-38, rrlw_wvn.F90 delwave = wavenum2 - wavenum1 */
    g_lw.grav = 9.8066;
    g_lw.avogad = 6.02214199e+23;
    g_lw.oneminus = 1. - 1.e-6;
    g_lw.fluxfac = 3.14159265358979323846 * 2.e4;
    static const double wavenum1[16] = {10., 350., 500., 630., 700., 820., 980., 1080.,
                                        1180., 1390., 1480., 1800., 2080., 2250., 2380., 2600.};
    static const double wavenum2[16] = {350., 500., 630., 700., 820., 980., 1080., 1180.,
                                        1390., 1480., 1800., 2080., 2250., 2380., 2600., 3250.};
    for (int i = 0; i < 16; ++i) g_lw.delwave[i] = wavenum2[i] - wavenum1[i];

    /* lookup tables, rrtmg_lw_init.F90:96-113 */
    const int ntbl = 10000;
    const double pade = 0.278, expeps = 1.e-20;
    g_lw.tau_tbl[0] = 0.0;
    g_lw.tau_tbl[ntbl] = 1.e10;
    g_lw.exp_tbl[0] = 1.0;
    g_lw.exp_tbl[ntbl] = expeps;
    g_lw.tfn_tbl[0] = 0.0;
    g_lw.tfn_tbl[ntbl] = 1.0;
    g_lw.bpade = 1.0 / pade;
    for (int itr = 1; itr <= ntbl - 1; ++itr) {
        double tfn = (double)itr / (double)ntbl;
        g_lw.tau_tbl[itr] = g_lw.bpade * tfn / (1. - tfn);
        g_lw.exp_tbl[itr] = exp(-g_lw.tau_tbl[itr]);
        if (g_lw.exp_tbl[itr] <= expeps) g_lw.exp_tbl[itr] = expeps;
        if (g_lw.tau_tbl[itr] < 0.06)
            g_lw.tfn_tbl[itr] = g_lw.tau_tbl[itr] / 6.;
        else
            g_lw.tfn_tbl[itr] = 1. - 2. * ((1. / g_lw.tau_tbl[itr]) -
                                           (g_lw.exp_tbl[itr] / (1. - g_lw.exp_tbl[itr])));
    }

    /* relative weights, rrtmg_lw_init.F90:120-144 (ng(ibnd) = 16 = mg for every band) */
    int igcsm = 0;
    for (int ibnd = 1; ibnd <= NBNDLW; ++ibnd) {
        int iprsm = 0;
        double wtsm[16];
        if (ngc[ibnd - 1] < 16) {
            for (int igc = 1; igc <= ngc[ibnd - 1]; ++igc) {
                igcsm++;
                double wtsum = 0.;
                for (int ipr = 1; ipr <= s_ngn[igcsm - 1]; ++ipr) {
                    iprsm++;
                    wtsum = wtsum + s_wt[iprsm - 1];
                }
                wtsm[igc - 1] = wtsum;
            }
            for (int ig = 1; ig <= 16; ++ig) {
                int ind = (ibnd - 1) * 16 + ig;
                g_lw.rwgt[ind - 1] = s_wt[ig - 1] / wtsm[s_ngm[ind - 1] - 1];
            }
        } else {
            for (int ig = 1; ig <= 16; ++ig) {
                igcsm++;
                int ind = (ibnd - 1) * 16 + ig;
                g_lw.rwgt[ind - 1] = 1.0;
            }
        }
    }

    /* cmbgb1..16, rrtmg_lw_init.F90:329-1978 */
    for (int ib = 0; ib < NBNDLW; ++ib) {
        LwBand *B = &g_lw.b[ib];
        int b1 = ib + 1;
        B->ng = ngc[ib];
        B->nspa = nspa[ib];
        B->nspb = nspb[ib];
        B->absa = reduce_last(band_tab(b1, "kao", &n), 65 * B->nspa, ib, 1);
        B->absb = reduce_last(band_tab(b1, "kbo", &n), 235 * (B->nspb ? B->nspb : 1), ib, 1);
        B->selfref = reduce_last(band_tab(b1, "selfrefo", &n), 10, ib, 1);
        B->forref = reduce_last(band_tab(b1, "forrefo", &n), 4, ib, 1);
        const double *fa = band_tab(b1, "fracrefao", &n);
        B->fracrefa = fa ? (n == 16 ? reduce_last(fa, 1, ib, 0) : reduce_first(fa, n / 16, ib)) : NULL;
        const double *fb = band_tab(b1, "fracrefbo", &n);
        B->fracrefb = fb ? (n == 16 ? reduce_last(fb, 1, ib, 0) : reduce_first(fb, n / 16, ib)) : NULL;
        struct { const char *nm; double **dst; } minors[] = {
            {"kao_mn2", &B->ka_mn2}, {"kbo_mn2", &B->kb_mn2}, {"kao_mn2o", &B->ka_mn2o},
            {"kbo_mn2o", &B->kb_mn2o}, {"kao_mo3", &B->ka_mo3}, {"kbo_mo3", &B->kb_mo3},
            {"kao_mco2", &B->ka_mco2}, {"kbo_mco2", &B->kb_mco2}, {"kao_mco", &B->ka_mco},
            {"kao_mo2", &B->ka_mo2}, {"kbo_mo2", &B->kb_mo2}, {"ccl4o", &B->ccl4},
            {"cfc11adjo", &B->cfc11adj}, {"cfc12o", &B->cfc12}, {"cfc22adjo", &B->cfc22adj}};
        for (size_t m = 0; m < sizeof minors / sizeof minors[0]; ++m) {
            const double *src = band_tab(b1, minors[m].nm, &n);
            *minors[m].dst = src ? reduce_last(src, n / 16, ib, 1) : NULL;
        }
    }

    g_lw.pref = blob_f64("lw.ref.pref", &n);
    g_lw.preflog = blob_f64("lw.ref.preflog", &n);
    g_lw.tref = blob_f64("lw.ref.tref", &n);
    g_lw.chi_mls = blob_f64("lw.ref.chi_mls", &n);
    g_lw.totplnk = blob_f64("lw.wvn.totplnk", &n);
    g_lw.totplk16 = blob_f64("lw.wvn.totplk16", &n);
    g_lw.totplnkderiv = blob_f64("lw.wvn.totplnkderiv", &n);
    g_lw.totplk16deriv = blob_f64("lw.wvn.totplk16deriv", &n);
    g_lw.absice0 = blob_f64("lw.cld.absice0", &n);
    g_lw.absice1 = blob_f64("lw.cld.absice1", &n);
    g_lw.absice2 = blob_f64("lw.cld.absice2", &n);
    g_lw.absice3 = blob_f64("lw.cld.absice3", &n);
    g_lw.absice4 = blob_f64("lw.cld.absice4", &n);
    g_lw.absliq1 = blob_f64("lw.cld.absliq1", &n);
    if (!g_lw.pref || !g_lw.chi_mls || !g_lw.totplnk || !g_lw.absice3 || !g_lw.absliq1) return -2;
    return 0;
}

void lw_free(void) {
    for (int ib = 0; ib < NBNDLW; ++ib) {
        LwBand *B = &g_lw.b[ib];
        double **p[] = {&B->absa, &B->absb, &B->selfref, &B->forref, &B->fracrefa, &B->fracrefb,
                        &B->ka_mn2, &B->kb_mn2, &B->ka_mn2o, &B->kb_mn2o, &B->ka_mo3, &B->kb_mo3,
                        &B->ka_mco2, &B->kb_mco2, &B->ka_mco, &B->ka_mo2, &B->kb_mo2, &B->ccl4,
                        &B->cfc11adj, &B->cfc12, &B->cfc22adj};
        for (size_t i = 0; i < sizeof p / sizeof p[0]; ++i) { free(*p[i]); *p[i] = NULL; }
    }
}

/* access to reduced tables for the init tests */
const double *oracle_lw_table(const char *name, int band, int *n) {
    *n = 0;
    if (!strcmp(name, "tau_tbl")) { *n = 10001; return g_lw.tau_tbl; }
    if (!strcmp(name, "exp_tbl")) { *n = 10001; return g_lw.exp_tbl; }
    if (!strcmp(name, "tfn_tbl")) { *n = 10001; return g_lw.tfn_tbl; }
    if (!strcmp(name, "rwgt")) { *n = 256; return g_lw.rwgt; }
    if (band < 1 || band > NBNDLW) return NULL;
    LwBand *B = &g_lw.b[band - 1];
    int ng = B->ng;
    struct { const char *nm; double *p; int lead; } t[] = {
        {"absa", B->absa, 65 * B->nspa}, {"absb", B->absb, 235 * B->nspb},
        {"selfref", B->selfref, 10}, {"forref", B->forref, 4},
        {"fracrefa", B->fracrefa, B->nspa == 9 ? 9 : 1},
        {"fracrefb", B->fracrefb, B->nspb == 5 ? 5 : 1},
        {"ka_mn2", B->ka_mn2, band == 15 ? 171 : 19}, {"kb_mn2", B->kb_mn2, 19},
        {"ka_mn2o", B->ka_mn2o, band == 8 ? 19 : 171}, {"kb_mn2o", B->kb_mn2o, band == 3 ? 95 : 19},
        {"ka_mo3", B->ka_mo3, band == 5 ? 171 : 19}, {"kb_mo3", B->kb_mo3, 19},
        {"ka_mco2", B->ka_mco2, (band == 7 || band == 13) ? 171 : 19}, {"kb_mco2", B->kb_mco2, 19},
        {"ka_mco", B->ka_mco, 171}, {"ka_mo2", B->ka_mo2, 19}, {"kb_mo2", B->kb_mo2, 19},
        {"ccl4", B->ccl4, 1}, {"cfc11adj", B->cfc11adj, 1}, {"cfc12", B->cfc12, 1},
        {"cfc22adj", B->cfc22adj, 1}};
    for (size_t i = 0; i < sizeof t / sizeof t[0]; ++i)
        if (!strcmp(name, t[i].nm) && t[i].p) { *n = t[i].lead * ng; return t[i].p; }
    return NULL;
}
