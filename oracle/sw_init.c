/* oracle/sw_init.c -- CPU restatement of rrtmg_sw_ini (test infrastructure only).
 *
 * Follows SW/src/rrtmg_sw_init.F90: relative g-point weights (:128-152), swdatinit constants
 * (:195-221), swcmbdat maps (:257-355) and the cmbgb16s..cmbgb29 reductions (:463-1660):
 * absorption, Rayleigh and minor-absorber data are weight-summed over each group of original
 * g-points in ascending order; the solar source arrays (sfluxref, irradnce, facbrght,
 * snsptdrk) are summed unweighted.  The band-29 irradnce scaling that the reference applies
 * inside its data routine (SW/src/rrtmg_sw_k_g_29.F90:80-81) is applied here before reduction.
 * The SW exp_tbl (:113-121) is built by the reference but never read; it is not restated.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "internal.h"

SwTables g_sw;
static const int *s_ngn, *s_ngm;
static const double *s_wt;

static double *reduce_last(const double *src, int lead, int ib, int weighted) {
    if (!src) return NULL;
    int ngc = g_sw.ngc[ib];
    int g0 = ib == 0 ? 0 : g_sw.ngs[ib - 1];
    double *dst = (double *)malloc(sizeof(double) * (size_t)lead * ngc);
    for (int l = 0; l < lead; ++l) {
        int iprsm = 0;
        for (int igc = 0; igc < ngc; ++igc) {
            double sumk = 0.;
            for (int ipr = 0; ipr < s_ngn[g0 + igc]; ++ipr) {
                if (weighted)
                    sumk = sumk + src[l + (size_t)lead * iprsm] * g_sw.rwgt[iprsm + 16 * ib];
                else
                    sumk = sumk + src[l + (size_t)lead * iprsm];
                iprsm++;
            }
            dst[l + (size_t)lead * igc] = sumk;
        }
    }
    return dst;
}

/* src(16,nj) -> dst(ngc,nj) */
static double *reduce_first(const double *src, int nj, int ib, int weighted) {
    if (!src) return NULL;
    int ngc = g_sw.ngc[ib];
    int g0 = ib == 0 ? 0 : g_sw.ngs[ib - 1];
    double *dst = (double *)malloc(sizeof(double) * (size_t)nj * ngc);
    for (int j = 0; j < nj; ++j) {
        int iprsm = 0;
        for (int igc = 0; igc < ngc; ++igc) {
            double sumf = 0.;
            for (int ipr = 0; ipr < s_ngn[g0 + igc]; ++ipr) {
                if (weighted)
                    sumf = sumf + src[iprsm + 16 * j] * g_sw.rwgt[iprsm + 16 * ib];
                else
                    sumf = sumf + src[iprsm + 16 * j];
                iprsm++;
            }
            dst[igc + (size_t)ngc * j] = sumf;
        }
    }
    return dst;
}

static const double *band_tab(int band, const char *name, int *n) {
    char key[64];
    snprintf(key, sizeof key, "sw.kg%d.%s", band, name);
    return blob_f64(key, n);
}

int sw_init(void) {
    int n;
    memset(&g_sw, 0, sizeof g_sw);
    const int *ngc = blob_i32("sw.wvn.ngc", &n);
    const int *ngs = blob_i32("sw.wvn.ngs", &n);
    const int *ngb = blob_i32("sw.wvn.ngb", &n);
    const int *nspa = blob_i32("sw.wvn.nspa", &n);
    const int *nspb = blob_i32("sw.wvn.nspb", &n);
    const int *icxa = blob_i32("sw.wvn.icxa", &n);
    s_ngn = blob_i32("sw.wvn.ngn", &n);
    s_ngm = blob_i32("sw.wvn.ngm", &n);
    s_wt = blob_f64("sw.wvn.wt", &n);
    if (!ngc || !ngs || !ngb || !s_ngn || !s_ngm || !s_wt || !nspa || !nspb || !icxa) return -1;
    for (int i = 0; i < NBNDSW; ++i) { g_sw.ngc[i] = ngc[i]; g_sw.ngs[i] = ngs[i]; g_sw.icxa[i] = icxa[i]; }
    for (int i = 0; i < NGPTSW; ++i) g_sw.ngb[i] = ngb[i];
    g_sw.grav = 9.8066;            /* swdatinit :203 */
    g_sw.avogad = 6.02214199e+23;  /* :211 */
    g_sw.oneminus = 1. - 1.e-06;   /* rrsw_con.F90 */

    /* relative weights :128-152 */
    int igcsm = 0;
    for (int ibnd = 1; ibnd <= NBNDSW; ++ibnd) {
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
                g_sw.rwgt[ind - 1] = s_wt[ig - 1] / wtsm[s_ngm[ind - 1] - 1];
            }
        } else {
            for (int ig = 1; ig <= 16; ++ig) {
                igcsm++;
                g_sw.rwgt[(ibnd - 1) * 16 + ig - 1] = 1.0;
            }
        }
    }

    for (int ib = 0; ib < NBNDSW; ++ib) {
        SwBand *B = &g_sw.b[ib];
        int band = ib + 16;
        B->ng = ngc[ib];
        B->nspa = nspa[ib];
        B->nspb = nspb[ib];
        B->absa = reduce_last(band_tab(band, "kao", &n), 65 * (B->nspa ? B->nspa : 1), ib, 1);
        B->absb = reduce_last(band_tab(band, "kbo", &n), 235 * (B->nspb ? B->nspb : 1), ib, 1);
        B->selfref = reduce_last(band_tab(band, "selfrefo", &n), 10, ib, 1);
        const double *fr = band_tab(band, "forrefo", &n);
        B->nfor = fr ? n / 16 : 0;
        B->forref = reduce_last(fr, B->nfor, ib, 1);

        const double *sf = band_tab(band, "sfluxrefo", &n);
        B->nsrc = n / 16;
        const double *ir = band_tab(band, "irradnceo", &n);
        const double *fb = band_tab(band, "facbrghto", &n);
        const double *sd = band_tab(band, "snsptdrko", &n);
        double irs[16];
        if (band == 29) { /* SW/src/rrtmg_sw_k_g_29.F90:80-81 */
            double irradscl = 13.221 / (13.221 - 0.455);
            for (int i = 0; i < 16; ++i) irs[i] = irradscl * ir[i];
            ir = irs;
        }
        if (B->nsrc == 1) {
            B->sfluxref = reduce_last(sf, 1, ib, 0);
            B->irradnce = reduce_last(ir, 1, ib, 0);
            B->facbrght = reduce_last(fb, 1, ib, 0);
            B->snsptdrk = reduce_last(sd, 1, ib, 0);
        } else {
            B->sfluxref = reduce_first(sf, B->nsrc, ib, 0);
            B->irradnce = reduce_first(ir, B->nsrc, ib, 0);
            B->facbrght = reduce_first(fb, B->nsrc, ib, 0);
            B->snsptdrk = reduce_first(sd, B->nsrc, ib, 0);
        }
        const double *r = band_tab(band, "rayl", &n);
        B->rayl = r ? r[0] : 0.;
        B->raylv = reduce_last(band_tab(band, "raylo", &n), 1, ib, 1);
        B->rayla = reduce_first(band_tab(band, "raylao", &n), 9, ib, 1);
        B->raylb = reduce_last(band_tab(band, "raylbo", &n), 1, ib, 1);
        B->abso3a = reduce_last(band_tab(band, "abso3ao", &n), 1, ib, 1);
        B->abso3b = reduce_last(band_tab(band, "abso3bo", &n), 1, ib, 1);
        B->absch4 = reduce_last(band_tab(band, "absch4o", &n), 1, ib, 1);
        B->absco2 = reduce_last(band_tab(band, "absco2o", &n), 1, ib, 1);
        B->absh2o = reduce_last(band_tab(band, "absh2oo", &n), 1, ib, 1);
    }

    g_sw.pref = blob_f64("sw.ref.pref", &n);
    g_sw.preflog = blob_f64("sw.ref.preflog", &n);
    g_sw.tref = blob_f64("sw.ref.tref", &n);
    struct { const char *nm; const double **dst; } t[] = {
        {"extliq1", &g_sw.extliq1}, {"ssaliq1", &g_sw.ssaliq1}, {"asyliq1", &g_sw.asyliq1},
        {"extice2", &g_sw.extice2}, {"ssaice2", &g_sw.ssaice2}, {"asyice2", &g_sw.asyice2},
        {"extice3", &g_sw.extice3}, {"ssaice3", &g_sw.ssaice3}, {"asyice3", &g_sw.asyice3},
        {"fdlice3", &g_sw.fdlice3}, {"extice4", &g_sw.extice4}, {"ssaice4", &g_sw.ssaice4},
        {"asyice4", &g_sw.asyice4}, {"abari", &g_sw.abari}, {"bbari", &g_sw.bbari},
        {"cbari", &g_sw.cbari}, {"dbari", &g_sw.dbari}, {"ebari", &g_sw.ebari}, {"fbari", &g_sw.fbari}};
    for (size_t i = 0; i < sizeof t / sizeof t[0]; ++i) {
        char key[64];
        snprintf(key, sizeof key, "sw.cld.%s", t[i].nm);
        *t[i].dst = blob_f64(key, &n);
        if (!*t[i].dst) return -2;
    }
    g_sw.mgavgcyc = blob_f64("sw.nrlssi2.mgavgcyc", &n);
    g_sw.sbavgcyc = blob_f64("sw.nrlssi2.sbavgcyc", &n);
    if (!g_sw.pref || !g_sw.mgavgcyc || !g_sw.sbavgcyc) return -3;
    return 0;
}

void sw_free(void) {
    for (int ib = 0; ib < NBNDSW; ++ib) {
        SwBand *B = &g_sw.b[ib];
        double **p[] = {&B->absa, &B->absb, &B->selfref, &B->forref, &B->sfluxref, &B->irradnce,
                        &B->facbrght, &B->snsptdrk, &B->raylv, &B->rayla, &B->raylb, &B->abso3a,
                        &B->abso3b, &B->absch4, &B->absco2, &B->absh2o};
        for (size_t i = 0; i < sizeof p / sizeof p[0]; ++i) { free(*p[i]); *p[i] = NULL; }
    }
}

const double *oracle_sw_table(const char *name, int band, int *n) {
    *n = 0;
    if (!strcmp(name, "rwgt")) { *n = 224; return g_sw.rwgt; }
    if (band < 16 || band > 29) return NULL;
    SwBand *B = &g_sw.b[band - 16];
    int ng = B->ng;
    struct { const char *nm; double *p; int lead; } t[] = {
        {"absa", B->absa, 65 * B->nspa}, {"absb", B->absb, 235 * B->nspb},
        {"selfref", B->selfref, 10}, {"forref", B->forref, B->nfor},
        {"sfluxref", B->sfluxref, B->nsrc}, {"irradnce", B->irradnce, B->nsrc},
        {"facbrght", B->facbrght, B->nsrc}, {"snsptdrk", B->snsptdrk, B->nsrc},
        {"rayl", B->raylv, 1}, {"rayla", B->rayla, 9}, {"raylb", B->raylb, 1},
        {"abso3a", B->abso3a, 1}, {"abso3b", B->abso3b, 1}, {"absch4", B->absch4, 1},
        {"absco2", B->absco2, 1}, {"absh2o", B->absh2o, 1}};
    for (size_t i = 0; i < sizeof t / sizeof t[0]; ++i)
        if (!strcmp(name, t[i].nm) && t[i].p) { *n = t[i].lead * ng; return t[i].p; }
    if (!strcmp(name, "rayl") && !B->raylv) { *n = 1; return &B->rayl; }
    return NULL;
}
