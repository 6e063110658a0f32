/* oracle/mcica.c -- CPU restatement of the McICA subcolumn generator (test infrastructure only).
 *
 * Follows GEOS_RadiationShared/cloud_subcol_gen.F90 (generate_stochastic_clouds :132-487,
 * correlation_length :491-542, rng_kiss :546-607, clearCounts_threeBand :611-769) and
 * GEOS_RadiationShared/cloud_condensate_inhomogeneity.F90 (zcw_lookup :86-124), with `real`
 * promoted to fp64.  Arrays use the partition layout of the reference: x(nlay,dncol).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include "internal.h"

McicaState g_mcica = {0, NULL, 1.4315, 2.1219, 7., -25.584, 0.72192, 0.78996, 8.5, 40.404};

/* SH/cloud_subcol_gen.F90:108-129 + SH/cloud_condensate_inhomogeneity.F90:45-78 */
int oracle_set_mcica(int ih, const double *corr) {
    if (ih < 0 || ih > 2) return -1; /* 'unknown inhomogeneity type' */
    g_mcica.inhm = ih;
    g_mcica.xcw = NULL;
    if (ih == 1) g_mcica.xcw = blob_f64("mcica.xcw_beta", NULL);
    if (ih == 2) g_mcica.xcw = blob_f64("mcica.xcw_gamma", NULL);
    if (ih > 0 && !g_mcica.xcw) return -2;
    /* defaults: Oreopoulos et al. 2012, cloud_subcol_gen.F90:51-59 */
    g_mcica.aam1 = 1.4315; g_mcica.aam2 = 2.1219; g_mcica.aam30 = 7.; g_mcica.aam4 = -25.584;
    g_mcica.ram1 = 0.72192; g_mcica.ram2 = 0.78996; g_mcica.ram30 = 8.5; g_mcica.ram4 = 40.404;
    if (corr) {
        g_mcica.aam1 = corr[0]; g_mcica.aam2 = corr[1]; g_mcica.aam30 = corr[2]; g_mcica.aam4 = corr[3];
        g_mcica.ram1 = corr[4]; g_mcica.ram2 = corr[5]; g_mcica.ram30 = corr[6]; g_mcica.ram4 = corr[7];
    }
    return 0;
}

/* SH/cloud_condensate_inhomogeneity.F90:86-124 */
double zcw_lookup(double cdf, double sigma_qcw) {
    const int n1 = 1000, n2 = 140;
    if (g_mcica.inhm == 0) return 1.;
    double rind1 = cdf * (double)(n1 - 1) + 1.;
    int ind1 = f_int(rind1);
    if (ind1 > n1 - 1) ind1 = n1 - 1;
    if (ind1 < 1) ind1 = 1;
    rind1 = rind1 - (double)ind1;
    double rind2 = 40. * sigma_qcw - 3.;
    int ind2 = f_int(rind2);
    if (ind2 > n2 - 1) ind2 = n2 - 1;
    if (ind2 < 1) ind2 = 1;
    rind2 = rind2 - (double)ind2;
    const double *x = g_mcica.xcw;
#define XCW(i, j) x[((i)-1) + 1000 * ((j)-1)]
    double zcw = (1.0 - rind1) * (1.0 - rind2) * XCW(ind1, ind2)
               + (1.0 - rind1) * rind2 * XCW(ind1, ind2 + 1)
               + rind1 * (1.0 - rind2) * XCW(ind1 + 1, ind2)
               + rind1 * rind2 * XCW(ind1 + 1, ind2 + 1);
#undef XCW
    return zcw;
}

/* SH/cloud_subcol_gen.F90:546-607; int32 two's-complement wraparound and logical shifts */
void oracle_rng_kiss(int *seed1, int *seed2, int *seed3, int *seed4, double *ran_num) {
    uint32_t s1 = (uint32_t)*seed1, s2 = (uint32_t)*seed2, s3 = (uint32_t)*seed3, s4 = (uint32_t)*seed4;
    s1 = 69069u * s1 + 1327217885u;
    s2 = s2 ^ (s2 << 13);
    s2 = s2 ^ (s2 >> 17);
    s2 = s2 ^ (s2 << 5);
    s3 = 18000u * (s3 & 65535u) + (s3 >> 16);
    s4 = 30903u * (s4 & 65535u) + (s4 >> 16);
    uint32_t kiss = s1 + s2 + (s3 << 16) + s4;
    *seed1 = (int)s1; *seed2 = (int)s2; *seed3 = (int)s3; *seed4 = (int)s4;
    *ran_num = (double)(int32_t)kiss * 2.328306e-10 + 0.5;
}

/* SH/cloud_subcol_gen.F90:491-516 */
static void correlation_length(int ncol, double am1, double am2, double am30, double am4,
                               int doy, const double *alat, double *clength) {
    const double r2d = 180.0 / 3.14159265358979323846;
    double am3;
    if (doy > 181)
        am3 = -(4. * am30 / 365. * (double)(doy - 272));
    else
        am3 = 4. * am30 / 365. * (double)(doy - 91);
    for (int icol = 0; icol < ncol; ++icol) {
        double d = alat[icol] * r2d - am3;
        clength[icol] = (am1 + am2 * exp(-((d * d) / (am4 * am4)))) * 1.e3;
    }
}

/* SH/cloud_subcol_gen.F90:132-487 */
int oracle_generate_stochastic_clouds(
    int dncol, int ncol, int nsubcol, int nlay,
    const double *zmid, const double *alat, int doy,
    const double *play, const double *cldfrac, const double *ciwp, const double *clwp,
    double cwp_tiny, unsigned char *cldy_stoch, double *ciwp_stoch, double *clwp_stoch,
    const int *seed_order) {
    (void)dncol;
    int so[4] = {1, 2, 3, 4};
    if (seed_order) {
        for (int n = 0; n < 4; ++n) {
            so[n] = seed_order[n];
            if (so[n] < 1) return -11; /* 'seed_order element < 1' */
            if (so[n] > 4) return -12; /* 'seed_order element > 4' */
            /* the reference's repeat test (:288-293) indexes hit(n), which can never fire */
        }
    }
    const int maximo = 2147483647 - 1;
    int surface_at_one = play[0] > play[nlay - 1]; /* play(1,1) > play(nlay,1), :267 */
    int cond_inhomo = g_mcica.inhm > 0;

    double *adl = (double *)malloc(sizeof(double) * (size_t)ncol);
    double *rdl = (double *)malloc(sizeof(double) * (size_t)ncol);
    double *alpha = (double *)malloc(sizeof(double) * (size_t)nlay * 6);
    double *rcorr = alpha + nlay, *sigma_qcw = alpha + 2 * nlay, *cdf1 = alpha + 3 * nlay,
           *cdf2 = alpha + 4 * nlay, *cdf3 = alpha + 5 * nlay;
    correlation_length(ncol, g_mcica.aam1, g_mcica.aam2, g_mcica.aam30, g_mcica.aam4, doy, alat, adl);
    if (cond_inhomo)
        correlation_length(ncol, g_mcica.ram1, g_mcica.ram2, g_mcica.ram30, g_mcica.ram4, doy, alat, rdl);

    for (int icol = 0; icol < ncol; ++icol) {
        const double *zm = zmid + (size_t)nlay * icol;
        const double *pl = play + (size_t)nlay * icol;
        const double *cf = cldfrac + (size_t)nlay * icol;
        const double *ci = ciwp + (size_t)nlay * icol;
        const double *cl = clwp + (size_t)nlay * icol;
        for (int k = 1; k < nlay; ++k) alpha[k] = exp(-fabs(zm[k] - zm[k - 1]) / adl[icol]);
        if (cond_inhomo) {
            for (int k = 1; k < nlay; ++k) rcorr[k] = exp(-fabs(zm[k] - zm[k - 1]) / rdl[icol]);
            for (int k = 0; k < nlay; ++k) {
                if (cf[k] > 0.99) sigma_qcw[k] = 0.5;
                else if (cf[k] > 0.9) sigma_qcw[k] = 0.71;
                else sigma_qcw[k] = 1.0;
            }
        }
        double pseed[4];
        for (int n = 0; n < 4; ++n) pseed[n] = (surface_at_one ? pl[n] : pl[nlay - 1 - n]) * 100.;
        int seed[4];
        for (int n = 0; n < 4; ++n) {
            double p = pseed[so[n] - 1];
            seed[n] = f_int((p - (double)f_int(p)) * (double)maximo + 1.);
        }
        for (int isub = 0; isub < nsubcol; ++isub) {
            for (int k = 0; k < nlay; ++k) {
                oracle_rng_kiss(&seed[0], &seed[1], &seed[2], &seed[3], &cdf1[k]);
                oracle_rng_kiss(&seed[0], &seed[1], &seed[2], &seed[3], &cdf2[k]);
            }
            for (int k = 1; k < nlay; ++k)
                if (cdf2[k] < alpha[k]) cdf1[k] = cdf1[k - 1];
            if (cond_inhomo) {
                for (int k = 0; k < nlay; ++k) {
                    oracle_rng_kiss(&seed[0], &seed[1], &seed[2], &seed[3], &cdf2[k]);
                    oracle_rng_kiss(&seed[0], &seed[1], &seed[2], &seed[3], &cdf3[k]);
                }
                for (int k = 1; k < nlay; ++k)
                    if (cdf2[k] < rcorr[k]) cdf3[k] = cdf3[k - 1];
            }
            size_t base = (size_t)nlay * ((size_t)isub + (size_t)nsubcol * icol);
            for (int k = 0; k < nlay; ++k) {
                unsigned char cldy;
                double ciw, clw;
                if (cdf1[k] >= 1. - cf[k]) {
                    cldy = 1;
                    if (cond_inhomo) {
                        double zcw = zcw_lookup(cdf3[k], sigma_qcw[k]);
                        ciw = ci[k] * zcw;
                        clw = cl[k] * zcw;
                    } else {
                        ciw = ci[k];
                        clw = cl[k];
                    }
                    int ciwp_negligible = ciw <= cwp_tiny;
                    if (ciwp_negligible) ciw = 0.;
                    int clwp_negligible = clw <= cwp_tiny;
                    if (clwp_negligible) clw = 0.;
                    if (ciwp_negligible && clwp_negligible) cldy = 0;
                } else {
                    cldy = 0; ciw = 0.; clw = 0.;
                }
                cldy_stoch[base + k] = cldy;
                ciwp_stoch[base + k] = ciw;
                clwp_stoch[base + k] = clw;
            }
        }
    }
    free(adl); free(rdl); free(alpha);
    return 0;
}

/* SH/cloud_subcol_gen.F90:611-769 */
int oracle_clearCounts_threeBand(int dncol, int ncol, int nsubcol, int nlay, int cloudLM,
                                 int cloudMH, const unsigned char *cldy_stoch, int *clearCnts) {
    for (int i = 0; i < 4 * dncol; ++i) clearCnts[i] = 0;
    if (cloudLM == cloudMH) return -21; /* 'invalid pressure super-layers!' */
#define CLDY(l, s, c) cldy_stoch[((l)-1) + (size_t)nlay * (((s)-1) + (size_t)nsubcol * ((c)-1))]
#define ANY(l0, l1, found) { found = 0; for (int l_ = (l0); l_ <= (l1); ++l_) if (CLDY(l_, isub, icol)) { found = 1; break; } }
    for (int icol = 1; icol <= ncol; ++icol)
        for (int isub = 1; isub <= nsubcol; ++isub) {
            int found;
            ANY(1, nlay, found);
            if (!found) clearCnts[0 + 4 * (icol - 1)]++;
            if (cloudLM < cloudMH) { /* surface at level 1 */
                ANY(1, cloudLM, found);
                if (!found) clearCnts[3 + 4 * (icol - 1)]++;
                ANY(cloudLM + 1, cloudMH, found);
                if (!found) clearCnts[2 + 4 * (icol - 1)]++;
                ANY(cloudMH + 1, nlay, found);
                if (!found) clearCnts[1 + 4 * (icol - 1)]++;
            } else { /* TOA at level 1 */
                ANY(1, cloudMH - 1, found);
                if (!found) clearCnts[1 + 4 * (icol - 1)]++;
                ANY(cloudMH, cloudLM - 1, found);
                if (!found) clearCnts[2 + 4 * (icol - 1)]++;
                ANY(cloudLM, nlay, found);
                if (!found) clearCnts[3 + 4 * (icol - 1)]++;
            }
        }
#undef ANY
#undef CLDY
    return 0;
}
