/* oracle/sw.c -- CPU restatement of the RRTMG_SW column path (test infrastructure only).
 *
 * Routine-by-routine restatement, `real` promoted to fp64, of (SW/ =
 * GEOSsolar_GridComp/RRTMG/rrtmg_sw/gcm_model/, non-SOLAR_RADVAL build):
 *   SW/src/rrtmg_sw_rad.F90      rrtmg_sw :68-452, rrtmg_sw_sub :455-1801
 *   SW/src/NRLSSI2.F90           initialize_NRLSSI2, adjust_solcyc_amplitudes, interpolate_indices
 *   SW/src/rrtmg_sw_cldprmc.F90  cldprmc_sw :36-418
 *   SW/src/rrtmg_sw_setcoef.F90  setcoef_sw :23-241
 *   SW/src/rrtmg_sw_taumol.F90   taumol16..29 :213-2084
 *   SW/src/rrtmg_sw_spcvmc.F90   spcvmc_sw :34-1112, reftra_sw :1115-1370, vrtqdr_sw :1374-1588
 * Same loop nests, same expression order; build with -ffp-contract=off.  Columns are split
 * into clear and cloudy sets and processed in partitions of pncol columns like the reference;
 * OpenMP runs partitions concurrently with private scratch.
 *
 * Deviations that only make undefined reference behaviour defined: iceflag outside 1..4 and
 * liqflag /= 1 (which leave the reference's cloud coefficients unset) return -41/-51, and an
 * effective radius outside the table range (an out-of-bounds table read in the reference; GEOS
 * clamps radii before the call, GEOS_SolarGridComp.F90:6127-6221) returns -42/-52.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "internal.h"

#define NG NGPTSW
#define NB NBNDSW

static void *zalloc(size_t n) { return calloc(n ? n : 1, 1); }

/* ------------------------------------------------------------------------------------------
 * solar variability scalars, SW/src/rrtmg_sw_rad.F90:889-1127 + SW/src/NRLSSI2.F90
 * ---------------------------------------------------------------------------------------- */
#define NSOLFRAC 134
static const double Iint = 1360.37, Fint = 0.996047, Sint = -0.511590;
static const double Mg_avg = 0.1567652, SB_avg = 909.71260, Mg_0 = 0.14959542, SB_0 = 0.00066696;
static const double rrsw_scon = 1368.22; /* SW/modules/parrrsw.F90:111 */

static int adjust_solcyc_amplitudes(double solcycfr, const double *indsolvar, double *scl) {
    const double solcycfrac_min = 0.0189, solcycfrac_max = 0.3750;
    const double fracdiff_min2max = solcycfrac_max - solcycfrac_min;
    const double fracdiff_max2min = 1. - fracdiff_min2max;
    double wgt;
    if (solcycfr >= 0. && solcycfr < solcycfrac_min) {
        wgt = (solcycfr + 1. - solcycfrac_max) / fracdiff_max2min;
        scl[0] = indsolvar[0] + wgt * (1. - indsolvar[0]);
        scl[1] = indsolvar[1] + wgt * (1. - indsolvar[1]);
    } else if (solcycfr >= solcycfrac_min && solcycfr <= solcycfrac_max) {
        wgt = (solcycfr - solcycfrac_min) / fracdiff_min2max;
        scl[0] = 1. + wgt * (indsolvar[0] - 1.);
        scl[1] = 1. + wgt * (indsolvar[1] - 1.);
    } else if (solcycfr > solcycfrac_max && solcycfr <= 1.) {
        wgt = (solcycfr - solcycfrac_max) / fracdiff_max2min;
        scl[0] = indsolvar[0] + wgt * (1. - indsolvar[0]);
        scl[1] = indsolvar[1] + wgt * (1. - indsolvar[1]);
    } else
        return -61;
    return 0;
}

static int interpolate_indices(double solcycfr, double *Mg, double *SB) {
    const double *mg = g_sw.mgavgcyc, *sb = g_sw.sbavgcyc; /* (134) */
    const double intrvl_len = 1.0 / (NSOLFRAC - 2);
    const double intrvl_len_hf = 0.5 * intrvl_len;
    if (solcycfr > 0. && solcycfr < 1.) {
        int sfid = 1;
        double fraclo = 0., frachi = intrvl_len_hf;
        if (solcycfr <= intrvl_len_hf) {
            sfid = 1; fraclo = 0.; frachi = intrvl_len_hf;
        } else if (solcycfr > intrvl_len_hf && solcycfr < 1. - intrvl_len_hf) {
            sfid = (int)floor((solcycfr - intrvl_len_hf) * (NSOLFRAC - 2)) + 2;
            fraclo = (sfid - 2) * intrvl_len + intrvl_len_hf;
            frachi = fraclo + intrvl_len;
        } else if (solcycfr >= 1. - intrvl_len_hf) {
            sfid = (NSOLFRAC - 2) + 1;
            fraclo = 1. - intrvl_len_hf;
            frachi = 1.;
        }
        double intfrac = (solcycfr - fraclo) / (frachi - fraclo);
        *Mg = mg[sfid - 1] + intfrac * (mg[sfid] - mg[sfid - 1]);
        *SB = sb[sfid - 1] + intfrac * (sb[sfid] - sb[sfid - 1]);
    } else if (solcycfr == 0.) {
        *Mg = mg[0]; *SB = sb[0];
    } else if (solcycfr == 1.) {
        *Mg = mg[NSOLFRAC - 1]; *SB = sb[NSOLFRAC - 1];
    } else
        return -61;
    return 0;
}

typedef struct {
    int isolvar;
    double adjflux[NB];
    double svar_f, svar_s, svar_i;
    double svar_f_bnd[NB], svar_s_bnd[NB], svar_i_bnd[NB];
} Solar;

static int solar_setup(Solar *S, int isolvar, double scon, double adjes, const double *bndscl,
                       const double *indsolvar, const double *solcycfrac) {
    double solvar[NB], indsolvar_scl[2] = {1., 1.}, indsolvar_ndx[2] = {Mg_avg, SB_avg};
    double solcycfr = 0., Mg_now = 0., SB_now = 0.;
    double mean_svar_f = 1., mean_svar_s = 1.; /* isolvar_1_mean_svar_f/s */
    S->isolvar = isolvar;
    for (int b = 0; b < NB; ++b) {
        solvar[b] = 1.; S->adjflux[b] = 1.;
        S->svar_f_bnd[b] = 1.; S->svar_s_bnd[b] = 1.; S->svar_i_bnd[b] = 1.;
    }
    S->svar_f = 1.; S->svar_s = 1.; S->svar_i = 1.;

    if (isolvar == 1) {
        if (!solcycfrac) return -61;
        solcycfr = *solcycfrac;
        if (indsolvar && (indsolvar[0] != 1. || indsolvar[1] != 1.)) {
            int rc = adjust_solcyc_amplitudes(solcycfr, indsolvar, indsolvar_scl);
            if (rc) return rc;
        }
    }
    if (isolvar == 2 && indsolvar) { indsolvar_ndx[0] = indsolvar[0]; indsolvar_ndx[1] = indsolvar[1]; }

    /* initialize_NRLSSI2 (NRLSSI2.F90): cycle means of the scaled multipliers for isolvar = 1 */
    if (isolvar == 1) {
        double ind[2] = {1., 1.};
        if (indsolvar) { ind[0] = indsolvar[0]; ind[1] = indsolvar[1]; }
        int scl1 = ind[0] != 1., scl2 = ind[1] != 1.;
        if (scl1 || scl2) {
            const double intrvl_len = 1.0 / (NSOLFRAC - 2);
            const double intrvl_len_hf = 0.5 * intrvl_len;
            double iscl1_mean = 0., iscl2_mean = 0., iscl1_Mg_mean = 0., iscl2_SB_mean = 0.;
            if (scl1) iscl1_mean = (1. + ind[0]) / 2.;
            if (scl2) iscl2_mean = (1. + ind[1]) / 2.;
            double fr = intrvl_len_hf, scl[2];
            for (int n = 2; n <= NSOLFRAC - 1; ++n) {
                int rc = adjust_solcyc_amplitudes(fr, ind, scl);
                if (rc) return rc;
                if (scl1) iscl1_Mg_mean = iscl1_Mg_mean + scl[0] * g_sw.mgavgcyc[n - 1];
                if (scl2) iscl2_SB_mean = iscl2_SB_mean + scl[1] * g_sw.sbavgcyc[n - 1];
                fr = fr + intrvl_len;
            }
            if (scl1) iscl1_Mg_mean = iscl1_Mg_mean / (NSOLFRAC - 2);
            if (scl2) iscl2_SB_mean = iscl2_SB_mean / (NSOLFRAC - 2);
            if (scl1) mean_svar_f = (iscl1_Mg_mean - iscl1_mean * Mg_0) / (Mg_avg - Mg_0);
            if (scl2) mean_svar_s = (iscl2_SB_mean - iscl2_mean * SB_0) / (SB_avg - SB_0);
        }
    }

    if (scon == 0.) {
        if (isolvar == -1) {
            if (bndscl) for (int b = 0; b < NB; ++b) solvar[b] = bndscl[b];
        } else if (isolvar == 0) {
            /* defaults */
        } else if (isolvar == 1) {
            int rc = interpolate_indices(solcycfr, &Mg_now, &SB_now);
            if (rc) return rc;
            S->svar_f = indsolvar_scl[0] * (Mg_now - Mg_0) / (Mg_avg - Mg_0);
            S->svar_s = indsolvar_scl[1] * (SB_now - SB_0) / (SB_avg - SB_0);
            S->svar_i = 1.;
        } else if (isolvar == 2) {
            S->svar_f = (indsolvar_ndx[0] - Mg_0) / (Mg_avg - Mg_0);
            S->svar_s = (indsolvar_ndx[1] - SB_0) / (SB_avg - SB_0);
            S->svar_i = 1.;
        } else if (isolvar == 3) {
            if (bndscl) for (int b = 0; b < NB; ++b) solvar[b] = bndscl[b];
            for (int b = 0; b < NB; ++b)
                S->svar_f_bnd[b] = S->svar_s_bnd[b] = S->svar_i_bnd[b] = solvar[b];
        } else
            return -61;
    } else if (scon > 0.) {
        if (isolvar == -1) {
            for (int b = 0; b < NB; ++b) solvar[b] = scon / rrsw_scon;
            if (bndscl) for (int b = 0; b < NB; ++b) solvar[b] = solvar[b] * bndscl[b];
        } else if (isolvar == 0) {
            double scon_int = Fint + Sint + Iint;
            double svar_r = scon / scon_int;
            S->svar_f = svar_r; S->svar_s = svar_r; S->svar_i = svar_r;
        } else if (isolvar == 1) {
            int rc = interpolate_indices(solcycfr, &Mg_now, &SB_now);
            if (rc) return rc;
            S->svar_f = indsolvar_scl[0] * (Mg_now - Mg_0) / (Mg_avg - Mg_0);
            S->svar_s = indsolvar_scl[1] * (SB_now - SB_0) / (SB_avg - SB_0);
            S->svar_i = (scon - (mean_svar_f * Fint + mean_svar_s * Sint)) / Iint;
        } else if (isolvar == 2) {
            S->svar_f = (indsolvar_ndx[0] - Mg_0) / (Mg_avg - Mg_0);
            S->svar_s = (indsolvar_ndx[1] - SB_0) / (SB_avg - SB_0);
            S->svar_i = (scon - (S->svar_f * Fint + S->svar_s * Sint)) / Iint;
        } else if (isolvar == 3) {
            double scon_int = Fint + Sint + Iint;
            for (int b = 0; b < NB; ++b) solvar[b] = scon / scon_int;
            if (bndscl) for (int b = 0; b < NB; ++b) solvar[b] = solvar[b] * bndscl[b];
            for (int b = 0; b < NB; ++b)
                S->svar_f_bnd[b] = S->svar_s_bnd[b] = S->svar_i_bnd[b] = solvar[b];
        } else
            return -61;
    } else
        return -61;

    for (int b = 0; b < NB; ++b) S->adjflux[b] = adjes;
    if (isolvar < 0)
        for (int b = 0; b < NB; ++b) S->adjflux[b] = S->adjflux[b] * solvar[b];
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * partition-local arrays: (nlay,pncol), (nlay,ngpt,pncol), (nlay+1,ngpt,pncol)
 * ---------------------------------------------------------------------------------------- */
#define I2(lay, icol) ((size_t)((lay)-1) + (size_t)nlay * ((icol)-1))
#define I2P(lev, icol) ((size_t)((lev)-1) + (size_t)(nlay + 1) * ((icol)-1))
#define I3(lay, ig, icol) ((size_t)((lay)-1) + (size_t)nlay * (((ig)-1) + (size_t)NG * ((icol)-1)))
#define I3P(lev, ig, icol) ((size_t)((lev)-1) + (size_t)(nlay + 1) * (((ig)-1) + (size_t)NG * ((icol)-1)))
#define IB(lay, ib, icol) ((size_t)((lay)-1) + (size_t)nlay * (((ib)-1) + (size_t)NB * ((icol)-1)))
#define IGC(ig, icol) ((size_t)((ig)-1) + (size_t)NG * ((icol)-1))

/* VAR(i,ib) of a cloud table with leading dimension lead and bands 16:29 */
#define CT(tab, lead, i, ib) (tab)[((i)-1) + (size_t)(lead) * ((ib)-16)]
#define LIN_CT(tab, lead, i, ib, f) (CT(tab, lead, i, ib) + (f) * (CT(tab, lead, (i) + 1, ib) - CT(tab, lead, i, ib)))

/* The phase-split McICA cloud optical properties cldprmc_sw hands out in the SOLAR_RADVAL build
 * (SW/src/rrtmg_sw_cldprmc.F90:38-47, 84-92): liquid / ice, original ("or") and delta-scaled ("c") optical depth,
 * single-scattering albedo, asymmetry, and the forward-scattering fractions. */
enum { RV_LTAOR, RV_LOMOR, RV_LASOR, RV_LTAUC, RV_LOMGC, RV_LASYC, RV_ITAOR, RV_IOMOR, RV_IASOR, RV_ITAUC, RV_IOMGC,
       RV_IASYC, RV_FORWL, RV_FORWI, RV_CELL_COUNT };
enum { RV_NOUT = 120 };   /* the SOLAR_RADVAL dummies of rrtmg_sw, SW/src/rrtmg_sw_rad.F90:85-122 */

/* SW/src/rrtmg_sw_cldprmc.F90:36-418 */
static int cldprmc_sw(int ncol, int nlay, int iceflag, int liqflag, const unsigned char *cldymc,
                      const double *ciwpmc, const double *clwpmc, const double *reicmc,
                      const double *relqmc, double *taormc, double *taucmc, double *ssacmc,
                      double *asmcmc, double *const *rv /* SOLAR_RADVAL (:38-47, :84-92): NULL or RV_CELL_COUNT arrays */) {
    const SwTables *T = &g_sw;
    const double epsg = 1.e-06, cldmin = 1.e-20;
    if (iceflag < 1 || iceflag > 4) return -41;
    if (liqflag != 1) return -51;
    for (int icol = 1; icol <= ncol; ++icol)
        for (int ig = 1; ig <= NG; ++ig) {
            const int ib = T->ngb[ig - 1]; /* 16..29 */
            for (int lay = 1; lay <= nlay; ++lay) {
                const size_t k = I3(lay, ig, icol);
                if (!cldymc[k]) {
                    taormc[k] = 0.; taucmc[k] = 0.; ssacmc[k] = 1.; asmcmc[k] = 0.;
                    if (rv) { /* :394-412; forwliq / forwice are left unset by the reference (0 here) */
                        rv[RV_LTAOR][k] = 0.; rv[RV_LOMOR][k] = 1.; rv[RV_LASOR][k] = 0.;
                        rv[RV_LTAUC][k] = 0.; rv[RV_LOMGC][k] = 1.; rv[RV_LASYC][k] = 0.;
                        rv[RV_ITAOR][k] = 0.; rv[RV_IOMOR][k] = 1.; rv[RV_IASOR][k] = 0.;
                        rv[RV_ITAUC][k] = 0.; rv[RV_IOMGC][k] = 1.; rv[RV_IASYC][k] = 0.;
                    }
                    continue;
                }
                double extcoice, ssacoice, gice, forwice;
                if (ciwpmc[k] == 0.) {
                    extcoice = 0.; ssacoice = 0.; gice = 0.; forwice = 0.;
                } else {
                    const double radice = reicmc[I2(lay, icol)];
                    if (iceflag == 1) {
                        const int ibx = T->icxa[ib - 16]; /* 1..5 */
                        extcoice = T->abari[ibx - 1] + T->bbari[ibx - 1] / radice;
                        ssacoice = 1. - T->cbari[ibx - 1] - T->dbari[ibx - 1] * radice;
                        gice = T->ebari[ibx - 1] + T->fbari[ibx - 1] * radice;
                        gice = f_min(gice, 1. - epsg);
                        forwice = gice * gice;
                    } else if (iceflag == 2) {
                        const double factor = (radice - 2.) / 3.;
                        int index = f_int(factor);
                        if (index == 43) index = 42;
                        if (index < 1 || index > 42) return -42;
                        const double fint = factor - (double)index;
                        extcoice = LIN_CT(T->extice2, 43, index, ib, fint);
                        ssacoice = LIN_CT(T->ssaice2, 43, index, ib, fint);
                        gice = LIN_CT(T->asyice2, 43, index, ib, fint);
                        forwice = gice * gice;
                    } else if (iceflag == 3) {
                        const double factor = (radice - 2.) / 3.;
                        int index = f_int(factor);
                        if (index == 46) index = 45;
                        if (index < 1 || index > 45) return -42;
                        const double fint = factor - (double)index;
                        extcoice = LIN_CT(T->extice3, 46, index, ib, fint);
                        ssacoice = LIN_CT(T->ssaice3, 46, index, ib, fint);
                        gice = LIN_CT(T->asyice3, 46, index, ib, fint);
                        const double fdelta = LIN_CT(T->fdlice3, 46, index, ib, fint);
                        forwice = fdelta + 0.5 / ssacoice;
                        if (forwice > gice) forwice = gice;
                    } else {
                        const double factor = radice;
                        const int index = f_int(factor);
                        if (index < 1 || index > 199) return -42;
                        const double fint = factor - (double)index;
                        extcoice = LIN_CT(T->extice4, 200, index, ib, fint);
                        ssacoice = LIN_CT(T->ssaice4, 200, index, ib, fint);
                        gice = LIN_CT(T->asyice4, 200, index, ib, fint);
                        forwice = gice * gice;
                    }
                }
                double extcoliq, ssacoliq, gliq, forwliq;
                if (clwpmc[k] == 0.) {
                    extcoliq = 0.; ssacoliq = 0.; gliq = 0.; forwliq = 0.;
                } else {
                    const double radliq = relqmc[I2(lay, icol)];
                    int index = f_int(radliq - 1.5);
                    if (index == 0) index = 1;
                    if (index == 58) index = 57;
                    if (index < 1 || index > 57) return -52;
                    const double fint = radliq - 1.5 - (double)index;
                    extcoliq = LIN_CT(T->extliq1, 58, index, ib, fint);
                    ssacoliq = LIN_CT(T->ssaliq1, 58, index, ib, fint);
                    if (fint < 0. && ssacoliq > 1.) ssacoliq = CT(T->ssaliq1, 58, index, ib);
                    gliq = LIN_CT(T->asyliq1, 58, index, ib, fint);
                    forwliq = gliq * gliq;
                }
                const double tauliqorig = clwpmc[k] * extcoliq;
                const double tauiceorig = ciwpmc[k] * extcoice;
                taormc[k] = tauliqorig + tauiceorig;
                const double ssaliq = ssacoliq * (1. - forwliq) / (1. - forwliq * ssacoliq);
                const double ssaice = ssacoice * (1. - forwice) / (1. - forwice * ssacoice);
                const double tauliq = (1. - forwliq * ssacoliq) * tauliqorig;
                const double tauice = (1. - forwice * ssacoice) * tauiceorig;
                const double scatliq = ssaliq * tauliq;
                double scatice = ssaice * tauice;
                taucmc[k] = tauliq + tauice;
                if (rv) { /* phase-split properties, original and delta-scaled, :321-328, :342-351 */
                    rv[RV_LTAOR][k] = tauliqorig; rv[RV_ITAOR][k] = tauiceorig;
                    rv[RV_LOMOR][k] = ssacoliq; rv[RV_IOMOR][k] = ssacoice;
                    rv[RV_LASOR][k] = gliq; rv[RV_IASOR][k] = gice;
                    rv[RV_LTAUC][k] = tauliq; rv[RV_ITAUC][k] = tauice;
                    rv[RV_LOMGC][k] = ssaliq; rv[RV_IOMGC][k] = ssaice;
                    rv[RV_LASYC][k] = (gliq - forwliq) / (1. - forwliq);
                    rv[RV_IASYC][k] = (gice - forwice) / (1. - forwice);
                    rv[RV_FORWL][k] = forwliq; rv[RV_FORWI][k] = forwice;
                }
                if (taucmc[k] == 0.) taucmc[k] = cldmin;
                if (scatice == 0.) scatice = cldmin;
                ssacmc[k] = (scatliq + scatice) / taucmc[k];
                if (iceflag == 3) {
                    asmcmc[k] = (1. / (scatliq + scatice)) *
                                (scatliq * (gliq - forwliq) / (1. - forwliq) +
                                 scatice * ((gice - forwice) / (1. - forwice)));
                } else {
                    asmcmc[k] = (scatliq * (gliq - forwliq) / (1. - forwliq) +
                                 scatice * (gice - forwice) / (1. - forwice)) /
                                (scatliq + scatice);
                }
            }
        }
    return 0;
}

/* per-partition setcoef_sw outputs */
typedef struct {
    int *laytrop, *jp, *jt, *jt1, *indself, *indfor;
    double *colh2o, *colco2, *colo3, *colch4, *colo2, *colmol, *coldry;
    double *selffac, *selffrac, *forfac, *forfrac, *fac00, *fac01, *fac10, *fac11;
} SwCoef;

static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* SW/src/rrtmg_sw_setcoef.F90:23-241 */
static void setcoef_sw(SwCoef *s, int ncol, int nlay, const double *pavel, const double *tavel) {
    const SwTables *T = &g_sw;
    const double stpfac = 296. / 1013.;
    for (int icol = 1; icol <= ncol; ++icol) {
        s->laytrop[icol - 1] = 0;
        for (int lay = 1; lay <= nlay; ++lay) {
            const double plog = log(pavel[I2(lay, icol)]);
            if (plog >= 4.56) s->laytrop[icol - 1] = s->laytrop[icol - 1] + 1;
        }
    }
    for (int icol = 1; icol <= ncol; ++icol)
        for (int lay = 1; lay <= nlay; ++lay) {
            const size_t k = I2(lay, icol);
            const double plog = log(pavel[k]);
            const int jp = clampi(f_int(36. - 5 * (plog + 0.04)), 1, 58);
            s->jp[k] = jp;
            const int jp1 = jp + 1;
            const double fp = 5. * (T->preflog[jp - 1] - plog);
            const int jt = clampi(f_int(3. + (tavel[k] - T->tref[jp - 1]) / 15.), 1, 4);
            s->jt[k] = jt;
            const double ft = ((tavel[k] - T->tref[jp - 1]) / 15.) - (double)(jt - 3);
            const int jt1 = clampi(f_int(3. + (tavel[k] - T->tref[jp1 - 1]) / 15.), 1, 4);
            s->jt1[k] = jt1;
            const double ft1 = ((tavel[k] - T->tref[jp1 - 1]) / 15.) - (double)(jt1 - 3);
            const double water = s->colh2o[k] / s->coldry[k];
            const double scalefac = pavel[k] * stpfac / tavel[k];
            if (plog <= 4.56) {
                s->forfac[k] = scalefac / (1. + water);
                const double factor = (tavel[k] - 188.) / 36.;
                s->indfor[k] = 3;
                s->forfrac[k] = factor - 1.;
                s->selffac[k] = 0.; s->selffrac[k] = 0.; s->indself[k] = 0;
            } else {
                s->forfac[k] = scalefac / (1. + water);
                double factor = (332. - tavel[k]) / 36.;
                s->indfor[k] = clampi(f_int(factor), 1, 2);
                s->forfrac[k] = factor - (double)s->indfor[k];
                s->selffac[k] = water * s->forfac[k];
                factor = (tavel[k] - 188.) / 7.2;
                s->indself[k] = clampi(f_int(factor) - 7, 1, 9);
                s->selffrac[k] = factor - (double)(s->indself[k] + 7);
            }
            s->colh2o[k] = 1.e-20 * s->colh2o[k];
            s->colco2[k] = 1.e-20 * s->colco2[k];
            s->colo3[k] = 1.e-20 * s->colo3[k];
            s->colch4[k] = 1.e-20 * s->colch4[k];
            s->colo2[k] = 1.e-20 * s->colo2[k];
            s->colmol[k] = 1.e-20 * s->coldry[k] + s->colh2o[k];
            if (s->colco2[k] == 0.) s->colco2[k] = 1.e-32 * s->coldry[k];
            if (s->colch4[k] == 0.) s->colch4[k] = 1.e-32 * s->coldry[k];
            if (s->colo2[k] == 0.) s->colo2[k] = 1.e-32 * s->coldry[k];
            const double compfp = 1. - fp;
            s->fac10[k] = compfp * ft;
            s->fac00[k] = compfp * (1. - ft);
            s->fac11[k] = fp * ft1;
            s->fac01[k] = fp * (1. - ft1);
        }
}

/* ------------------------------------------------------------------------------------------
 * taumol_sw, SW/src/rrtmg_sw_taumol.F90
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    const SwCoef *s;
    const Solar *S;
    int nlay, icol;
    double *ssi, *sfluxzen; /* (ngpt,pncol) */
    double *taug, *taur;    /* (nlay,ngpt,pncol) */
} TCol;

#define A2(f) (c->s->f[I2(lay, icol)])
#define ABSA(ind, ig) B->absa[((ind)-1) + (size_t)leada * ((ig)-1)]
#define ABSB(ind, ig) B->absb[((ind)-1) + (size_t)leadb * ((ig)-1)]
#define SELF(i, ig) B->selfref[((i)-1) + 10 * ((ig)-1)]
#define FORR(i, ig) B->forref[((i)-1) + (size_t)B->nfor * ((ig)-1)]
#define LIN_SELF(ig) (SELF(inds, ig) + A2(selffrac) * (SELF(inds + 1, ig) - SELF(inds, ig)))
#define LIN_FOR(ig) (FORR(indf, ig) + A2(forfrac) * (FORR(indf + 1, ig) - FORR(indf, ig)))
#define SRC2(t, ig, js) (t)[((ig)-1) + (size_t)ng * ((js)-1)]
#define LIN_SRC(t, ig) (SRC2(t, ig, js) + fs * (SRC2(t, ig, js + 1) - SRC2(t, ig, js)))
#define TAUG(ig) c->taug[I3(lay, gs + (ig), icol)]
#define TAUR(ig) c->taur[I3(lay, gs + (ig), icol)]

#define BAND_PROLOGUE(band)                                                     \
    const SwBand *B = &g_sw.b[(band)-16];                                       \
    const int nlay = c->nlay, icol = c->icol, ng = B->ng;                       \
    (void)nlay;                                                                 \
    const int gs = (band) == 16 ? 0 : g_sw.ngs[(band)-17];                      \
    const int laytrop = c->s->laytrop[icol - 1];                                \
    const int nspa = B->nspa, nspb = B->nspb;                                   \
    const size_t leada = (size_t)65 * (nspa ? nspa : 1), leadb = (size_t)235 * (nspb ? nspb : 1); \
    const int ibnd = (band)-16;                                                 \
    (void)ng; (void)gs; (void)laytrop; (void)nspa; (void)nspb; (void)leada; (void)leadb; (void)ibnd

#define SPEC(cola, strrat, colb, mult)                                          \
    const double speccomb = (cola) + (strrat) * (colb);                         \
    double specparm = (cola) / speccomb;                                        \
    if (specparm >= g_sw.oneminus) specparm = g_sw.oneminus;                    \
    const double specmult = (mult) * specparm;                                  \
    const int js = 1 + f_int(specmult);                                         \
    const double fs = specmult - (double)f_int(specmult) /* mod(specmult,1.) */

#define FAC8                                                                    \
    const double fac000 = (1. - fs) * A2(fac00), fac010 = (1. - fs) * A2(fac10); \
    const double fac100 = fs * A2(fac00), fac110 = fs * A2(fac10);              \
    const double fac001 = (1. - fs) * A2(fac01), fac011 = (1. - fs) * A2(fac11); \
    const double fac101 = fs * A2(fac01), fac111 = fs * A2(fac11)

#define KEY8A(ig)                                                               \
    (fac000 * ABSA(ind0, ig) + fac100 * ABSA(ind0 + 1, ig) + fac010 * ABSA(ind0 + 9, ig) + \
     fac110 * ABSA(ind0 + 10, ig) + fac001 * ABSA(ind1, ig) + fac101 * ABSA(ind1 + 1, ig) + \
     fac011 * ABSA(ind1 + 9, ig) + fac111 * ABSA(ind1 + 10, ig))
#define KEY8B(ig)                                                               \
    (fac000 * ABSB(ind0, ig) + fac100 * ABSB(ind0 + 1, ig) + fac010 * ABSB(ind0 + 5, ig) + \
     fac110 * ABSB(ind0 + 6, ig) + fac001 * ABSB(ind1, ig) + fac101 * ABSB(ind1 + 1, ig) + \
     fac011 * ABSB(ind1 + 5, ig) + fac111 * ABSB(ind1 + 6, ig))
#define KEY4A(ig)                                                               \
    (A2(fac00) * ABSA(ind0, ig) + A2(fac10) * ABSA(ind0 + 1, ig) + A2(fac01) * ABSA(ind1, ig) + \
     A2(fac11) * ABSA(ind1 + 1, ig))
#define KEY4B(ig)                                                               \
    (A2(fac00) * ABSB(ind0, ig) + A2(fac10) * ABSB(ind0 + 1, ig) + A2(fac01) * ABSB(ind1, ig) + \
     A2(fac11) * ABSB(ind1 + 1, ig))
#define IND_LO(js_)                                                             \
    const int ind0 = ((A2(jp) - 1) * 5 + (A2(jt) - 1)) * nspa + (js_);          \
    const int ind1 = (A2(jp) * 5 + (A2(jt1) - 1)) * nspa + (js_)
#define IND_UP(js_)                                                             \
    const int ind0 = ((A2(jp) - 13) * 5 + (A2(jt) - 1)) * nspb + (js_);         \
    const int ind1 = ((A2(jp) - 12) * 5 + (A2(jt1) - 1)) * nspb + (js_)

/* solar source, constant in the band (e.g. taumol16 :322-347) */
static void src_const(TCol *c, int band) {
    BAND_PROLOGUE(band);
    const Solar *S = c->S;
    for (int ig = 1; ig <= ng; ++ig) {
        if (S->isolvar < 0)
            c->sfluxzen[IGC(gs + ig, icol)] = B->sfluxref[ig - 1];
        else if (S->isolvar <= 2)
            c->ssi[IGC(gs + ig, icol)] = S->svar_f * B->facbrght[ig - 1] + S->svar_s * B->snsptdrk[ig - 1] +
                                         S->svar_i * B->irradnce[ig - 1];
        else if (S->isolvar == 3)
            c->ssi[IGC(gs + ig, icol)] = S->svar_f_bnd[ibnd] * B->facbrght[ig - 1] +
                                         S->svar_s_bnd[ibnd] * B->snsptdrk[ig - 1] +
                                         S->svar_i_bnd[ibnd] * B->irradnce[ig - 1];
    }
}

/* solar source interpolated in the binary-species parameter at layer laysolfr (e.g. :504-522) */
static void src_interp(TCol *c, int band, int js, double fs) {
    BAND_PROLOGUE(band);
    const Solar *S = c->S;
    for (int ig = 1; ig <= ng; ++ig) {
        if (S->isolvar < 0)
            c->sfluxzen[IGC(gs + ig, icol)] = LIN_SRC(B->sfluxref, ig);
        else if (S->isolvar <= 2)
            c->ssi[IGC(gs + ig, icol)] = S->svar_f * LIN_SRC(B->facbrght, ig) + S->svar_s * LIN_SRC(B->snsptdrk, ig) +
                                         S->svar_i * LIN_SRC(B->irradnce, ig);
        else if (S->isolvar == 3)
            c->ssi[IGC(gs + ig, icol)] = S->svar_f_bnd[ibnd] * LIN_SRC(B->facbrght, ig) +
                                         S->svar_s_bnd[ibnd] * LIN_SRC(B->snsptdrk, ig) +
                                         S->svar_i_bnd[ibnd] * LIN_SRC(B->irradnce, ig);
    }
}

/* laysolfr search below the tropopause (e.g. taumol18 :571-607): the layer just above the one
 * where jp crosses layreffr, capped at laytrop */
#define SOLFR_LOWER(band, layreffr, cola, strrat, colb)                          \
    {                                                                            \
        int laysolfr = laytrop;                                                  \
        for (int lay = 1; lay <= laytrop; ++lay) {                               \
            if (lay < nlay && c->s->jp[I2(lay, icol)] < (layreffr) && c->s->jp[I2(lay + 1, icol)] >= (layreffr)) \
                laysolfr = (lay + 1 < laytrop) ? lay + 1 : laytrop;              \
            if (lay == laysolfr) {                                               \
                SPEC(A2(cola), strrat, A2(colb), 8.);                            \
                src_interp(c, band, js, fs);                                     \
                break;                                                           \
            }                                                                    \
        }                                                                        \
    }
/* ... and above it (taumol17 :488-527, taumol28 :1930-1971) */
#define SOLFR_UPPER(band, layreffr, cola, strrat, colb)                          \
    {                                                                            \
        int laysolfr = nlay;                                                     \
        for (int lay = laytrop + 1; lay <= nlay; ++lay) {                        \
            if (c->s->jp[I2(lay - 1, icol)] < (layreffr) && c->s->jp[I2(lay, icol)] >= (layreffr)) \
                laysolfr = lay;                                                  \
            if (lay == laysolfr) {                                               \
                SPEC(A2(cola), strrat, A2(colb), 4.);                            \
                src_interp(c, band, js, fs);                                     \
                break;                                                           \
            }                                                                    \
        }                                                                        \
    }

static void taumol16(TCol *c) { /* :213-348 */
    BAND_PROLOGUE(16);
    const double strrat1 = 252.131;
    for (int lay = 1; lay <= laytrop; ++lay) {
        SPEC(A2(colh2o), strrat1, A2(colch4), 8.);
        FAC8;
        IND_LO(js);
        const int inds = A2(indself), indf = A2(indfor);
        const double tauray = A2(colmol) * B->rayl;
        for (int ig = 1; ig <= ng; ++ig) {
            TAUG(ig) = speccomb * KEY8A(ig) + A2(colh2o) * (A2(selffac) * LIN_SELF(ig) + A2(forfac) * LIN_FOR(ig));
            TAUR(ig) = tauray;
        }
    }
    for (int lay = laytrop + 1; lay <= nlay; ++lay) {
        IND_UP(1);
        const double tauray = A2(colmol) * B->rayl;
        for (int ig = 1; ig <= ng; ++ig) {
            TAUG(ig) = A2(colch4) * KEY4B(ig);
            TAUR(ig) = tauray;
        }
    }
    src_const(c, 16);
}

static void taumol17(TCol *c) { /* :352-527 */
    BAND_PROLOGUE(17);
    const double strrat = 0.364641;
    for (int lay = 1; lay <= laytrop; ++lay) {
        SPEC(A2(colh2o), strrat, A2(colco2), 8.);
        FAC8;
        IND_LO(js);
        const int inds = A2(indself), indf = A2(indfor);
        const double tauray = A2(colmol) * B->rayl;
        for (int ig = 1; ig <= ng; ++ig) {
            TAUG(ig) = speccomb * KEY8A(ig) + A2(colh2o) * (A2(selffac) * LIN_SELF(ig) + A2(forfac) * LIN_FOR(ig));
            TAUR(ig) = tauray;
        }
    }
    for (int lay = laytrop + 1; lay <= nlay; ++lay) {
        SPEC(A2(colh2o), strrat, A2(colco2), 4.);
        FAC8;
        IND_UP(js);
        const int indf = A2(indfor);
        const double tauray = A2(colmol) * B->rayl;
        for (int ig = 1; ig <= ng; ++ig) {
            TAUG(ig) = speccomb * KEY8B(ig) + A2(colh2o) * A2(forfac) * LIN_FOR(ig);
            TAUR(ig) = tauray;
        }
    }
    SOLFR_UPPER(17, 30, colh2o, strrat, colco2);
}

/* bands 18, 19, 21 (lower part), 22, 24 share the lower-atmosphere shape; written out per band */
static void taumol18(TCol *c) { /* :531-685 */
    BAND_PROLOGUE(18);
    const double strrat = 38.9589;
    SOLFR_LOWER(18, 6, colh2o, strrat, colch4);
    for (int lay = 1; lay <= laytrop; ++lay) {
        SPEC(A2(colh2o), strrat, A2(colch4), 8.);
        FAC8;
        IND_LO(js);
        const int inds = A2(indself), indf = A2(indfor);
        const double tauray = A2(colmol) * B->rayl;
        for (int ig = 1; ig <= ng; ++ig) {
            TAUG(ig) = speccomb * KEY8A(ig) + A2(colh2o) * (A2(selffac) * LIN_SELF(ig) + A2(forfac) * LIN_FOR(ig));
            TAUR(ig) = tauray;
        }
    }
    for (int lay = laytrop + 1; lay <= nlay; ++lay) {
        IND_UP(1);
        const double tauray = A2(colmol) * B->rayl;
        for (int ig = 1; ig <= ng; ++ig) {
            TAUG(ig) = A2(colch4) * KEY4B(ig);
            TAUR(ig) = tauray;
        }
    }
}

static void taumol19(TCol *c) { /* :689-826 */
    BAND_PROLOGUE(19);
    const double strrat = 5.49281;
    SOLFR_LOWER(19, 3, colh2o, strrat, colco2);
    for (int lay = 1; lay <= laytrop; ++lay) {
        SPEC(A2(colh2o), strrat, A2(colco2), 8.);
        FAC8;
        IND_LO(js);
        const int inds = A2(indself), indf = A2(indfor);
        const double tauray = A2(colmol) * B->rayl;
        for (int ig = 1; ig <= ng; ++ig) {
            TAUG(ig) = speccomb * KEY8A(ig) + A2(colh2o) * (A2(selffac) * LIN_SELF(ig) + A2(forfac) * LIN_FOR(ig));
            TAUR(ig) = tauray;
        }
    }
    for (int lay = laytrop + 1; lay <= nlay; ++lay) {
        IND_UP(1);
        const double tauray = A2(colmol) * B->rayl;
        for (int ig = 1; ig <= ng; ++ig) {
            TAUG(ig) = A2(colco2) * KEY4B(ig);
            TAUR(ig) = tauray;
        }
    }
}

static void taumol20(TCol *c) { /* :830-942 */
    BAND_PROLOGUE(20);
    for (int lay = 1; lay <= laytrop; ++lay) {
        IND_LO(1);
        const int inds = A2(indself), indf = A2(indfor);
        const double tauray = A2(colmol) * B->rayl;
        for (int ig = 1; ig <= ng; ++ig) {
            TAUG(ig) = A2(colh2o) * (KEY4A(ig) + A2(selffac) * LIN_SELF(ig) + A2(forfac) * LIN_FOR(ig)) +
                       A2(colch4) * B->absch4[ig - 1];
            TAUR(ig) = tauray;
        }
    }
    for (int lay = laytrop + 1; lay <= nlay; ++lay) {
        IND_UP(1);
        const int indf = A2(indfor);
        const double tauray = A2(colmol) * B->rayl;
        for (int ig = 1; ig <= ng; ++ig) {
            TAUG(ig) = A2(colh2o) * (A2(fac00) * ABSB(ind0, ig) + A2(fac10) * ABSB(ind0 + 1, ig) +
                                     A2(fac01) * ABSB(ind1, ig) + A2(fac11) * ABSB(ind1 + 1, ig) +
                                     A2(forfac) * LIN_FOR(ig)) +
                       A2(colch4) * B->absch4[ig - 1];
            TAUR(ig) = tauray;
        }
    }
    src_const(c, 20);
}

static void taumol21(TCol *c) { /* :946-1104 */
    BAND_PROLOGUE(21);
    const double strrat = 0.0045321;
    SOLFR_LOWER(21, 8, colh2o, strrat, colco2);
    for (int lay = 1; lay <= laytrop; ++lay) {
        SPEC(A2(colh2o), strrat, A2(colco2), 8.);
        FAC8;
        IND_LO(js);
        const int inds = A2(indself), indf = A2(indfor);
        const double tauray = A2(colmol) * B->rayl;
        for (int ig = 1; ig <= ng; ++ig) {
            TAUG(ig) = speccomb * KEY8A(ig) + A2(colh2o) * (A2(selffac) * LIN_SELF(ig) + A2(forfac) * LIN_FOR(ig));
            TAUR(ig) = tauray;
        }
    }
    for (int lay = laytrop + 1; lay <= nlay; ++lay) {
        SPEC(A2(colh2o), strrat, A2(colco2), 4.);
        FAC8;
        IND_UP(js);
        const int indf = A2(indfor);
        const double tauray = A2(colmol) * B->rayl;
        for (int ig = 1; ig <= ng; ++ig) {
            TAUG(ig) = speccomb * KEY8B(ig) + A2(colh2o) * A2(forfac) * LIN_FOR(ig);
            TAUR(ig) = tauray;
        }
    }
}

static void taumol22(TCol *c) { /* :1108-1254 */
    BAND_PROLOGUE(22);
    const double o2adj = 1.6, strrat = 0.022708;
    {
        int laysolfr = laytrop;
        for (int lay = 1; lay <= laytrop; ++lay) {
            if (lay < nlay && c->s->jp[I2(lay, icol)] < 2 && c->s->jp[I2(lay + 1, icol)] >= 2)
                laysolfr = (lay + 1 < laytrop) ? lay + 1 : laytrop;
            if (lay == laysolfr) {
                SPEC(A2(colh2o), o2adj * strrat, A2(colo2), 8.);
                src_interp(c, 22, js, fs);
                break;
            }
        }
    }
    for (int lay = 1; lay <= laytrop; ++lay) {
        const double o2cont = 4.35e-4 * A2(colo2) / (350.0 * 2.0);
        SPEC(A2(colh2o), o2adj * strrat, A2(colo2), 8.);
        FAC8;
        IND_LO(js);
        const int inds = A2(indself), indf = A2(indfor);
        const double tauray = A2(colmol) * B->rayl;
        for (int ig = 1; ig <= ng; ++ig) {
            TAUG(ig) = speccomb * KEY8A(ig) + A2(colh2o) * (A2(selffac) * LIN_SELF(ig) + A2(forfac) * LIN_FOR(ig)) +
                       o2cont;
            TAUR(ig) = tauray;
        }
    }
    for (int lay = laytrop + 1; lay <= nlay; ++lay) {
        const double o2cont = 4.35e-4 * A2(colo2) / (350. * 2.);
        IND_UP(1);
        const double tauray = A2(colmol) * B->rayl;
        for (int ig = 1; ig <= ng; ++ig) {
            TAUG(ig) = A2(colo2) * o2adj * KEY4B(ig) + o2cont;
            TAUR(ig) = tauray;
        }
    }
}

static void taumol23(TCol *c) { /* :1258-1360 */
    BAND_PROLOGUE(23);
    const double givfac = 1.029;
    for (int lay = 1; lay <= laytrop; ++lay) {
        IND_LO(1);
        const int inds = A2(indself), indf = A2(indfor);
        for (int ig = 1; ig <= ng; ++ig) {
            const double tauray = A2(colmol) * B->raylv[ig - 1];
            TAUG(ig) = A2(colh2o) * (givfac * KEY4A(ig) + A2(selffac) * LIN_SELF(ig) + A2(forfac) * LIN_FOR(ig));
            TAUR(ig) = tauray;
        }
    }
    for (int lay = laytrop + 1; lay <= nlay; ++lay)
        for (int ig = 1; ig <= ng; ++ig) {
            TAUG(ig) = 0.;
            TAUR(ig) = A2(colmol) * B->raylv[ig - 1];
        }
    src_const(c, 23);
}

static void taumol24(TCol *c) { /* :1364-1503 */
    BAND_PROLOGUE(24);
    const double strrat = 0.124692;
    SOLFR_LOWER(24, 1, colh2o, strrat, colo2);
    for (int lay = 1; lay <= laytrop; ++lay) {
        SPEC(A2(colh2o), strrat, A2(colo2), 8.);
        FAC8;
        IND_LO(js);
        const int inds = A2(indself), indf = A2(indfor);
        for (int ig = 1; ig <= ng; ++ig) {
            const double tauray = A2(colmol) * LIN_SRC(B->rayla, ig);
            TAUG(ig) = speccomb * KEY8A(ig) + A2(colo3) * B->abso3a[ig - 1] +
                       A2(colh2o) * (A2(selffac) * LIN_SELF(ig) + A2(forfac) * LIN_FOR(ig));
            TAUR(ig) = tauray;
        }
    }
    for (int lay = laytrop + 1; lay <= nlay; ++lay) {
        IND_UP(1);
        for (int ig = 1; ig <= ng; ++ig) {
            const double tauray = A2(colmol) * B->raylb[ig - 1];
            TAUG(ig) = A2(colo2) * KEY4B(ig) + A2(colo3) * B->abso3b[ig - 1];
            TAUR(ig) = tauray;
        }
    }
}

static void taumol25(TCol *c) { /* :1507-1604 */
    BAND_PROLOGUE(25);
    for (int lay = 1; lay <= laytrop; ++lay) {
        IND_LO(1);
        for (int ig = 1; ig <= ng; ++ig) {
            const double tauray = A2(colmol) * B->raylv[ig - 1];
            TAUG(ig) = A2(colh2o) * KEY4A(ig) + A2(colo3) * B->abso3a[ig - 1];
            TAUR(ig) = tauray;
        }
    }
    for (int lay = laytrop + 1; lay <= nlay; ++lay)
        for (int ig = 1; ig <= ng; ++ig) {
            const double tauray = A2(colmol) * B->raylv[ig - 1];
            TAUG(ig) = A2(colo3) * B->abso3b[ig - 1];
            TAUR(ig) = tauray;
        }
    src_const(c, 25);
}

static void taumol26(TCol *c) { /* :1608-1685 */
    BAND_PROLOGUE(26);
    for (int lay = 1; lay <= nlay; ++lay)
        for (int ig = 1; ig <= ng; ++ig) {
            TAUG(ig) = 0.;
            TAUR(ig) = A2(colmol) * B->raylv[ig - 1];
        }
    src_const(c, 26);
}

static void taumol27(TCol *c) { /* :1689-1799 */
    BAND_PROLOGUE(27);
    for (int lay = 1; lay <= laytrop; ++lay) {
        IND_LO(1);
        for (int ig = 1; ig <= ng; ++ig) {
            const double tauray = A2(colmol) * B->raylv[ig - 1];
            TAUG(ig) = A2(colo3) * KEY4A(ig);
            TAUR(ig) = tauray;
        }
    }
    for (int lay = laytrop + 1; lay <= nlay; ++lay) {
        IND_UP(1);
        for (int ig = 1; ig <= ng; ++ig) {
            const double tauray = A2(colmol) * B->raylv[ig - 1];
            TAUG(ig) = A2(colo3) * KEY4B(ig);
            TAUR(ig) = tauray;
        }
    }
    src_const(c, 27);
}

static void taumol28(TCol *c) { /* :1803-1971 */
    BAND_PROLOGUE(28);
    const double strrat = 6.67029e-07;
    for (int lay = 1; lay <= laytrop; ++lay) {
        SPEC(A2(colo3), strrat, A2(colo2), 8.);
        FAC8;
        IND_LO(js);
        const double tauray = A2(colmol) * B->rayl;
        for (int ig = 1; ig <= ng; ++ig) {
            TAUG(ig) = speccomb * KEY8A(ig);
            TAUR(ig) = tauray;
        }
    }
    for (int lay = laytrop + 1; lay <= nlay; ++lay) {
        SPEC(A2(colo3), strrat, A2(colo2), 4.);
        FAC8;
        IND_UP(js);
        const double tauray = A2(colmol) * B->rayl;
        for (int ig = 1; ig <= ng; ++ig) {
            TAUG(ig) = speccomb * KEY8B(ig);
            TAUR(ig) = tauray;
        }
    }
    SOLFR_UPPER(28, 42, colo3, strrat, colo2);
}

static void taumol29(TCol *c) { /* :1975-2084 */
    BAND_PROLOGUE(29);
    for (int lay = 1; lay <= laytrop; ++lay) {
        IND_LO(1);
        const int inds = A2(indself), indf = A2(indfor);
        const double tauray = A2(colmol) * B->rayl;
        for (int ig = 1; ig <= ng; ++ig) {
            TAUG(ig) = A2(colh2o) * (KEY4A(ig) + A2(selffac) * LIN_SELF(ig) + A2(forfac) * LIN_FOR(ig)) +
                       A2(colco2) * B->absco2[ig - 1];
            TAUR(ig) = tauray;
        }
    }
    for (int lay = laytrop + 1; lay <= nlay; ++lay) {
        IND_UP(1);
        const double tauray = A2(colmol) * B->rayl;
        for (int ig = 1; ig <= ng; ++ig) {
            TAUG(ig) = A2(colco2) * KEY4B(ig) + A2(colh2o) * B->absh2o[ig - 1];
            TAUR(ig) = tauray;
        }
    }
    src_const(c, 29);
}

static void taumol_sw(const SwCoef *s, const Solar *S, int ncol, int nlay, double *ssi,
                      double *sfluxzen, double *taug, double *taur) {
    static void (*const bands[NB])(TCol *) = {taumol16, taumol17, taumol18, taumol19, taumol20,
                                              taumol21, taumol22, taumol23, taumol24, taumol25,
                                              taumol26, taumol27, taumol28, taumol29};
    for (int b = 0; b < NB; ++b)
        for (int icol = 1; icol <= ncol; ++icol) {
            TCol c = {s, S, nlay, icol, ssi, sfluxzen, taug, taur};
            bands[b](&c);
        }
}

/* ------------------------------------------------------------------------------------------
 * reftra_sw, SW/src/rrtmg_sw_spcvmc.F90:1115-1370
 * ---------------------------------------------------------------------------------------- */
static void reftra_sw(int ncol, int nlay, const unsigned char *cloudy, const double *pgg,
                      const double *prmuzl, const double *ptau, const double *pw, double *pref,
                      double *prefd, double *ptra, double *ptrad, int update_cloudy_cells_only) {
    const double eps = 1.e-08, od_lo = 0.06; /* rrsw_tbl.F90:32 */
    const double zwcrit = 0.9999995;
    for (int icol = 1; icol <= ncol; ++icol) {
        const double prmuz = prmuzl[icol - 1];
        for (int iw = 1; iw <= NG; ++iw)
            for (int jk = 1; jk <= nlay; ++jk) {
                if (update_cloudy_cells_only && !cloudy[I3(nlay + 1 - jk, iw, icol)]) continue;
                const size_t k = I3(jk, iw, icol), kp = I3P(jk, iw, icol);
                const double zto1 = ptau[k], zw = pw[k], zg = pgg[k];
                const double zw8 = zw, zg8 = zg;
                const double zg3 = 3. * zg;
                const double zgamma1 = (8. - zw * (5. + zg3)) * 0.25;
                const double zgamma2 = 3. * (zw * (1. - zg)) * 0.25;
                const double zgamma3 = (2. - zg3 * prmuz) * 0.25;
                const double zgamma4 = 1. - zgamma3;
                const double r8 = zg8 / (1.0 - zg8);
                const double zwo8 = zw8 / (1.0 - (1.0 - zw8) * (r8 * r8));
                const double zwo = zwo8;
                if (zwo >= zwcrit) {
                    const double za = zgamma1 * prmuz;
                    const double za1 = za - zgamma3;
                    const double zgt = zgamma1 * zto1;
                    const double ze1 = f_min(zto1 / prmuz, 500.);
                    const double ze2 = exp(-ze1);
                    pref[kp] = (zgt - za1 * (1. - ze2)) / (1. + zgt);
                    ptra[kp] = 1. - pref[kp];
                    prefd[kp] = zgt / (1. + zgt);
                    ptrad[kp] = 1. - prefd[kp];
                    if (ze2 == 1.) {
                        pref[kp] = 0.; ptra[kp] = 1.; prefd[kp] = 0.; ptrad[kp] = 1.;
                    }
                } else {
                    const double za1 = zgamma1 * zgamma4 + zgamma2 * zgamma3;
                    const double za2 = zgamma1 * zgamma3 + zgamma2 * zgamma4;
                    const double zrk = sqrt(zgamma1 * zgamma1 - zgamma2 * zgamma2);
                    const double zrp = zrk * prmuz;
                    const double zrp1 = 1. + zrp;
                    const double zrm1 = 1. - zrp;
                    const double zrk2 = 2. * zrk;
                    const double zrpp = 1. - zrp * zrp;
                    const double zrkg = zrk + zgamma1;
                    const double zr1 = zrm1 * (za2 + zrk * zgamma3);
                    const double zr2 = zrp1 * (za2 - zrk * zgamma3);
                    const double zr3 = zrk2 * (zgamma3 - za2 * prmuz);
                    const double zr4 = zrpp * zrkg;
                    const double zr5 = zrpp * (zrk - zgamma1);
                    const double zt1 = zrp1 * (za1 + zrk * zgamma4);
                    const double zt2 = zrm1 * (za1 - zrk * zgamma4);
                    const double zt3 = zrk2 * (zgamma4 + za1 * prmuz);
                    const double zt4 = zr4, zt5 = zr5;
                    const double zbeta = (zgamma1 - zrk) / zrkg;
                    const double ze1 = f_min(zrk * zto1, 5.);
                    const double ze2 = f_min(zto1 / prmuz, 5.);
                    double zem1, zem2;
                    if (ze1 <= od_lo) zem1 = 1. - ze1 + 0.5 * ze1 * ze1; else zem1 = exp(-ze1);
                    const double zep1 = 1. / zem1;
                    if (ze2 <= od_lo) zem2 = 1. - ze2 + 0.5 * ze2 * ze2; else zem2 = exp(-ze2);
                    const double zep2 = 1. / zem2;
                    const double zdenr = zr4 * zep1 + zr5 * zem1;
                    const double zdent = zt4 * zep1 + zt5 * zem1;
                    if (zdenr >= -eps && zdenr <= eps) {
                        pref[kp] = eps;
                        ptra[kp] = zem2;
                    } else {
                        pref[kp] = zw * (zr1 * zep1 - zr2 * zem1 - zr3 * zem2) / zdenr;
                        ptra[kp] = zem2 - zem2 * zw * (zt1 * zep1 - zt2 * zem1 - zt3 * zep2) / zdent;
                    }
                    const double zemm = zem1 * zem1;
                    const double zdend = 1. / ((1. - zbeta * zemm) * zrkg);
                    prefd[kp] = zgamma2 * (1. - zemm) * zdend;
                    ptrad[kp] = zrk2 * zem1 * zdend;
                }
            }
    }
}

/* vrtqdr_sw, SW/src/rrtmg_sw_spcvmc.F90:1374-1588; work arrays (nlay+1,ngpt,pncol) supplied */
static void vrtqdr_sw(int ncol, int nlay, const double *pref, const double *prefd,
                      const double *ptra, const double *ptrad, const double *pdbt,
                      const double *ptdbt, double *pfd, double *pfu, double *ztdn, double *prup,
                      double *prupd, double *prdnd) {
#define P(a, lev) a[I3P(lev, iw, icol)]
#define D(lay) pdbt[I3(lay, iw, icol)]
    for (int icol = 1; icol <= ncol; ++icol)
        for (int iw = 1; iw <= NG; ++iw) {
            P(prup, nlay + 1) = P(pref, nlay + 1);
            P(prupd, nlay + 1) = P(prefd, nlay + 1);
            double zreflect = 1. / (1. - P(prefd, nlay + 1) * P(prefd, nlay));
            P(prup, nlay) = P(pref, nlay) + (P(ptrad, nlay) * ((P(ptra, nlay) - D(nlay)) * P(prefd, nlay + 1) +
                                                                D(nlay) * P(pref, nlay + 1))) * zreflect;
            P(prupd, nlay) = P(prefd, nlay) + P(ptrad, nlay) * P(ptrad, nlay) * P(prefd, nlay + 1) * zreflect;
            for (int jk = 1; jk <= nlay - 1; ++jk) {
                const int ikp = nlay + 1 - jk, ikx = ikp - 1;
                const double zreflectj = 1. / (1. - P(prupd, ikp) * P(prefd, ikx));
                P(prup, ikx) = P(pref, ikx) + (P(ptrad, ikx) * ((P(ptra, ikx) - D(ikx)) * P(prupd, ikp) +
                                                                D(ikx) * P(prup, ikp))) * zreflectj;
                P(prupd, ikx) = P(prefd, ikx) + P(ptrad, ikx) * P(ptrad, ikx) * P(prupd, ikp) * zreflectj;
            }
            P(ztdn, 1) = 1.;
            P(prdnd, 1) = 0.;
            P(ztdn, 2) = P(ptra, 1);
            P(prdnd, 2) = P(prefd, 1);
            for (int jk = 2; jk <= nlay; ++jk) {
                const int ikp = jk + 1;
                zreflect = 1. / (1. - P(prefd, jk) * P(prdnd, jk));
                P(ztdn, ikp) = P(ptdbt, jk) * P(ptra, jk) +
                               (P(ptrad, jk) * ((P(ztdn, jk) - P(ptdbt, jk)) +
                                                P(ptdbt, jk) * P(pref, jk) * P(prdnd, jk))) * zreflect;
                P(prdnd, ikp) = P(prefd, jk) + P(ptrad, jk) * P(ptrad, jk) * P(prdnd, jk) * zreflect;
            }
            for (int jk = 1; jk <= nlay + 1; ++jk) {
                zreflect = 1. / (1. - P(prdnd, jk) * P(prupd, jk));
                P(pfu, jk) = (P(ptdbt, jk) * P(prup, jk) + (P(ztdn, jk) - P(ptdbt, jk)) * P(prupd, jk)) * zreflect;
                P(pfd, jk) = P(ptdbt, jk) + (P(ztdn, jk) - P(ptdbt, jk) +
                                             P(ptdbt, jk) * P(prup, jk) * P(prdnd, jk)) * zreflect;
            }
        }
#undef P
#undef D
}

/* ------------------------------------------------------------------------------------------
 * spcvmc_sw, SW/src/rrtmg_sw_spcvmc.F90:34-1112 (non-RADVAL).  Only the outputs that
 * rrtmg_sw_sub hands back are formed (the uv/nir/direct flux profiles it discards are not).
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    double *zgco, *zomco, *ztauo, *zdbt, *ztaur, *ztaug;                /* (nlay,ngpt,pncol) */
    double *ztdbt, *zfd, *zfu, *zref, *zrefd, *ztra, *ztrad, *w1, *w2, *w3, *w4; /* (nlay+1,..) */
    double *zsflxzen, *ssi;                                              /* (ngpt,pncol) */
} SpcWork;

/* The SOLAR_RADVAL part of the PAR super-layer diagnostics for one (column, g-point), SW/src/rrtmg_sw_spcvmc.F90:
 * low layer :784-868, mid :872-956, high :960-1044, whole subcolumn :1048-1105.  Per super-layer fifteen layer sums:
 *   0 stau                                   sum(ptaucmc)
 *   1 sltao   2 sltaossa   3 sltaossag       liquid, original:      tau, tau*om, tau*om*as
 *   4 sltau   5 sltaussa   6 sltaussag  7 sltaussaf   liquid, delta-scaled: tau, tau*om, tau*om*as, tau*om*forw
 *   8 sitao   9 sitaossa  10 sitaossag;     11 sitau 12 sitaussa 13 sitaussag 14 sitaussaf    the same for ice
 * (the weighted sums of a family are formed only where its optical depth sum is > 0, else 0), and fifteen output
 * families in the order of the dummy list, each {d,n}{t,h,m,l}p: (test sum, what the "d" member accumulates beside
 * wgt, what the "n" member accumulates).  z addresses column i of zrv; rows are pncol apart. */
static const struct { int test, d, n; } rv_family[15] = {
    {0, -1, 0},                                            /* cds                      */
    {1, -1, 1}, {4, -1, 4}, {8, -1, 8}, {11, -1, 11},      /* cotl, cdsl, coti, cdsi   */
    {1, 1, 2}, {4, 4, 5}, {8, 8, 9}, {11, 11, 12},         /* ssal, sdsl, ssai, sdsi   */
    {1, 2, 3}, {4, 5, 6}, {8, 9, 10}, {11, 12, 13},        /* asml, adsl, asmi, adsi   */
    {4, 5, 7}, {11, 12, 14}};                              /* forl, fori               */

static void radval_sums(double *const *rv, const double *ptaucmc, int nlay, int cloudLM, int cloudMH, int iw, int icol,
                        double wgt, double *z, int pncol) {
    double S[4][15];   /* 0 whole subcolumn, 1 high, 2 mid, 3 low: the {t,h,m,l} order of the outputs */
    const int first[4] = {0, cloudMH + 1, cloudLM + 1, 1}, last[4] = {0, nlay, cloudMH, cloudLM};
    for (int L = 3; L >= 1; --L) {
        double *s = S[L];
        for (int q = 0; q < 15; ++q) s[q] = 0.;
        for (int lay = first[L]; lay <= last[L]; ++lay) s[0] = s[0] + ptaucmc[I3(lay, iw, icol)];
        for (int ph = 0; ph < 2; ++ph) {   /* liquid, ice */
            const double *tao = rv[ph ? RV_ITAOR : RV_LTAOR], *omo = rv[ph ? RV_IOMOR : RV_LOMOR],
                         *aso = rv[ph ? RV_IASOR : RV_LASOR], *tau = rv[ph ? RV_ITAUC : RV_LTAUC],
                         *omg = rv[ph ? RV_IOMGC : RV_LOMGC], *asy = rv[ph ? RV_IASYC : RV_LASYC],
                         *forw = rv[ph ? RV_FORWI : RV_FORWL];
            double *o = s + 1 + 7 * ph;
            for (int lay = first[L]; lay <= last[L]; ++lay) o[0] = o[0] + tao[I3(lay, iw, icol)];
            if (o[0] > 0.)
                for (int lay = first[L]; lay <= last[L]; ++lay) {
                    const size_t k = I3(lay, iw, icol);
                    o[1] = o[1] + tao[k] * omo[k];
                    o[2] = o[2] + tao[k] * omo[k] * aso[k];
                }
            for (int lay = first[L]; lay <= last[L]; ++lay) o[3] = o[3] + tau[I3(lay, iw, icol)];
            if (o[3] > 0.)
                for (int lay = first[L]; lay <= last[L]; ++lay) {
                    const size_t k = I3(lay, iw, icol);
                    o[4] = o[4] + tau[k] * omg[k];
                    o[5] = o[5] + tau[k] * omg[k] * asy[k];
                    o[6] = o[6] + tau[k] * omg[k] * forw[k];
                }
        }
    }
    for (int q = 0; q < 15; ++q) S[0][q] = S[3][q] + S[2][q] + S[1][q];   /* lp + mp + hp, :1048-1090 */
    for (int L = 0; L < 4; ++L)
        for (int f = 0; f < 15; ++f) {
            if (!(S[L][rv_family[f].test] > 0.)) continue;
            double *zd = z + (size_t)(f * 8 + L) * pncol, *zn = z + (size_t)(f * 8 + 4 + L) * pncol;
            *zd = *zd + (rv_family[f].d < 0 ? wgt : wgt * S[L][rv_family[f].d]);
            *zn = *zn + wgt * S[L][rv_family[f].n];
        }
}


static void spcvmc_sw(int cc, int ncol, int nlay, const SwCoef *sc, const Solar *S,
                      const double *palbd, const double *palbp, const unsigned char *pcldymc,
                      const double *ptaucmc, const double *pasycmc, const double *pomgcmc,
                      const double *ptaormc, const double *ptaua, const double *pasya,
                      const double *pomga, const double *prmu0, int cloudLM, int cloudMH,
                      SpcWork *W, double *pbbfd, double *pbbfu, double *pbbcd, double *pbbcu,
                      double *znirr, double *znirf, double *zparr, double *zparf, double *zuvrr,
                      double *zuvrf, double *fndsbnd /* (pncol,14) */, int pncol, double *zcot /* [8][pncol] */,
                      int do_drfband, double *zdrband, double *zdfband, OracleTaps *taps, const int *gcols, int gncol,
                      double *const *rv /* SOLAR_RADVAL: NULL or the RV_CELL_COUNT arrays of cldprmc_sw */,
                      double *zrv /* [RV_NOUT][pncol] */) {
    const SwTables *T = &g_sw;
    const size_t np = (size_t)(nlay + 1) * pncol;
    memset(pbbcd, 0, np * 8); memset(pbbcu, 0, np * 8); memset(pbbfd, 0, np * 8); memset(pbbfu, 0, np * 8);
    for (int i = 0; i < pncol; ++i) { znirr[i] = znirf[i] = zparr[i] = zparf[i] = zuvrr[i] = zuvrf[i] = 0.; }
    for (int i = 0; i < pncol * NB; ++i) fndsbnd[i] = 0.;
    if (do_drfband) for (int i = 0; i < pncol * NB; ++i) { zdrband[i] = 0.; zdfband[i] = 0.; }

    taumol_sw(sc, S, ncol, nlay, W->ssi, W->zsflxzen, W->ztaug, W->ztaur);
    if (taps)
        for (int icol = 1; icol <= ncol; ++icol) {
            const size_t o3 = (size_t)nlay * NG * gcols[icol - 1], s3 = (size_t)nlay * NG * (icol - 1);
            if (taps->taug) memcpy(taps->taug + o3, W->ztaug + s3, (size_t)nlay * NG * 8);
            if (taps->pfracs) memcpy(taps->pfracs + o3, W->ztaur + s3, (size_t)nlay * NG * 8);
            if (taps->ssi)
                memcpy(taps->ssi + (size_t)NG * gcols[icol - 1],
                       (S->isolvar < 0 ? W->zsflxzen : W->ssi) + (size_t)NG * (icol - 1), NG * 8);
        }
    (void)gncol;

    for (int icol = 1; icol <= ncol; ++icol)
        for (int iw = 1; iw <= NG; ++iw) {
            const int ibm = T->ngb[iw - 1] - 15;
            W->zref[I3P(nlay + 1, iw, icol)] = palbp[(ibm - 1) + NB * (icol - 1)];
            W->zrefd[I3P(nlay + 1, iw, icol)] = palbd[(ibm - 1) + NB * (icol - 1)];
            W->ztra[I3P(nlay + 1, iw, icol)] = 0.;
            W->ztrad[I3P(nlay + 1, iw, icol)] = 0.;
            W->ztdbt[I3P(1, iw, icol)] = 1.;
        }
    /* clear-sky optical properties with delta scaling (:413-437) */
    for (int icol = 1; icol <= ncol; ++icol)
        for (int iw = 1; iw <= NG; ++iw) {
            const int ibm = T->ngb[iw - 1] - 15;
            for (int jk = 1; jk <= nlay; ++jk) {
                const int ikl = nlay + 1 - jk;
                const size_t k = I3(jk, iw, icol), kl = I3(ikl, iw, icol), ka = IB(ikl, ibm, icol);
                W->ztauo[k] = W->ztaur[kl] + W->ztaug[kl] + ptaua[ka];
                W->zomco[k] = W->ztaur[kl] + ptaua[ka] * pomga[ka];
                W->zgco[k] = (pasya[ka] * pomga[ka] * ptaua[ka]) / W->zomco[k];
                W->zomco[k] = W->zomco[k] / W->ztauo[k];
                const double zf = W->zgco[k] * W->zgco[k];
                const double zwf = W->zomco[k] * zf;
                W->ztauo[k] = (1. - zwf) * W->ztauo[k];
                W->zomco[k] = (W->zomco[k] - zwf) / (1. - zwf);
                W->zgco[k] = (W->zgco[k] - zf) / (1. - zf);
            }
        }
    reftra_sw(ncol, nlay, pcldymc, W->zgco, prmu0, W->ztauo, W->zomco, W->zref, W->zrefd, W->ztra, W->ztrad, 0);
    for (int icol = 1; icol <= ncol; ++icol)
        for (int iw = 1; iw <= NG; ++iw)
            for (int jk = 1; jk <= nlay; ++jk) {
                W->zdbt[I3(jk, iw, icol)] = exp(-W->ztauo[I3(jk, iw, icol)] / prmu0[icol - 1]);
                W->ztdbt[I3P(jk + 1, iw, icol)] = W->zdbt[I3(jk, iw, icol)] * W->ztdbt[I3P(jk, iw, icol)];
            }
    vrtqdr_sw(ncol, nlay, W->zref, W->zrefd, W->ztra, W->ztrad, W->zdbt, W->ztdbt, W->zfd, W->zfu, W->w1, W->w2,
              W->w3, W->w4);
#define ZINCFLX(withmu)                                                                         \
    (S->isolvar < 0 ? S->adjflux[jb - 16] * W->zsflxzen[IGC(iw, icol)] * (withmu)               \
                    : S->adjflux[jb - 16] * W->ssi[IGC(iw, icol)] * (withmu))
    for (int icol = 1; icol <= ncol; ++icol)
        for (int iw = 1; iw <= NG; ++iw) {
            const int jb = T->ngb[iw - 1];
            const double zincflx = ZINCFLX(prmu0[icol - 1]);
            for (int ikl = 1; ikl <= nlay + 1; ++ikl) {
                const int jk = nlay + 2 - ikl;
                pbbcu[I2P(ikl, icol)] = pbbcu[I2P(ikl, icol)] + zincflx * W->zfu[I3P(jk, iw, icol)];
                pbbcd[I2P(ikl, icol)] = pbbcd[I2P(ikl, icol)] + zincflx * W->zfd[I3P(jk, iw, icol)];
            }
        }
    if (cc == 2) {
        /* add cloud to the cloudy cells (:512-536) */
        for (int icol = 1; icol <= ncol; ++icol)
            for (int iw = 1; iw <= NG; ++iw)
                for (int jk = 1; jk <= nlay; ++jk) {
                    const int ikl = nlay + 1 - jk;
                    const size_t k = I3(jk, iw, icol), kl = I3(ikl, iw, icol);
                    if (pcldymc[kl]) {
                        W->zgco[k] = W->ztauo[k] * W->zomco[k] * W->zgco[k] + ptaucmc[kl] * pomgcmc[kl] * pasycmc[kl];
                        W->zomco[k] = W->ztauo[k] * W->zomco[k] + ptaucmc[kl] * pomgcmc[kl];
                        W->ztauo[k] = W->ztauo[k] + ptaucmc[kl];
                        W->zgco[k] = W->zgco[k] / W->zomco[k];
                        W->zomco[k] = W->zomco[k] / W->ztauo[k];
                    }
                }
        reftra_sw(ncol, nlay, pcldymc, W->zgco, prmu0, W->ztauo, W->zomco, W->zref, W->zrefd, W->ztra, W->ztrad, 1);
        for (int icol = 1; icol <= ncol; ++icol)
            for (int iw = 1; iw <= NG; ++iw)
                for (int jk = 1; jk <= nlay; ++jk) {
                    const int ikl = nlay + 1 - jk;
                    if (pcldymc[I3(ikl, iw, icol)])
                        W->zdbt[I3(jk, iw, icol)] = exp(-W->ztauo[I3(jk, iw, icol)] / prmu0[icol - 1]);
                    W->ztdbt[I3P(jk + 1, iw, icol)] = W->zdbt[I3(jk, iw, icol)] * W->ztdbt[I3P(jk, iw, icol)];
                }
        vrtqdr_sw(ncol, nlay, W->zref, W->zrefd, W->ztra, W->ztrad, W->zdbt, W->ztdbt, W->zfd, W->zfu, W->w1,
                  W->w2, W->w3, W->w4);
        for (int icol = 1; icol <= ncol; ++icol)
            for (int iw = 1; iw <= NG; ++iw) {
                const int jb = T->ngb[iw - 1];
                const double zincflx = ZINCFLX(prmu0[icol - 1]);
                for (int ikl = 1; ikl <= nlay + 1; ++ikl) {
                    const int jk = nlay + 2 - ikl;
                    pbbfu[I2P(ikl, icol)] = pbbfu[I2P(ikl, icol)] + zincflx * W->zfu[I3P(jk, iw, icol)];
                    pbbfd[I2P(ikl, icol)] = pbbfd[I2P(ikl, icol)] + zincflx * W->zfd[I3P(jk, iw, icol)];
                }
            }
    } else {
        memcpy(pbbfu, pbbcu, np * 8);
        memcpy(pbbfd, pbbcd, np * 8);
    }
    /* surface band fluxes (:624-668) */
    for (int icol = 1; icol <= ncol; ++icol)
        for (int iw = 1; iw <= NG; ++iw) {
            const int jb = T->ngb[iw - 1], ibm = jb - 15;
            const double zincflx = ZINCFLX(prmu0[icol - 1]);
            const double tdb = W->ztdbt[I3P(nlay + 1, iw, icol)], fd = W->zfd[I3P(nlay + 1, iw, icol)],
                         fu = W->zfu[I3P(nlay + 1, iw, icol)];
            const int i = icol - 1;
            if (ibm == 14 || ibm <= 8) {
                znirr[i] = znirr[i] + zincflx * tdb;
                znirf[i] = znirf[i] + zincflx * fd;
            } else if (ibm >= 10 && ibm <= 11) {
                zparr[i] = zparr[i] + zincflx * tdb;
                zparf[i] = zparf[i] + zincflx * fd;
            } else if (ibm >= 12 && ibm <= 13) {
                zuvrr[i] = zuvrr[i] + zincflx * tdb;
                zuvrf[i] = zuvrf[i] + zincflx * fd;
            } else if (ibm == 9) {
                zparr[i] = zparr[i] + 0.5 * zincflx * tdb;
                zparf[i] = zparf[i] + 0.5 * zincflx * fd;
                znirr[i] = znirr[i] + 0.5 * zincflx * tdb;
                znirf[i] = znirf[i] + 0.5 * zincflx * fd;
            }
            fndsbnd[i + pncol * (ibm - 1)] = fndsbnd[i + pncol * (ibm - 1)] + zincflx * (fd - fu);
            if (do_drfband) {
                zdrband[i + pncol * (ibm - 1)] = zdrband[i + pncol * (ibm - 1)] + zincflx * tdb;
                zdfband[i + pncol * (ibm - 1)] = zdfband[i + pncol * (ibm - 1)] + zincflx * fd;
            }
        }
    if (do_drfband)
        for (int i = 0; i < pncol * NB; ++i) zdfband[i] = zdfband[i] - zdrband[i];

    /* PAR-weighted in-cloud optical thickness per pressure super-layer (:748-1108); zcot rows:
     * 0 cotdtp 1 cotdhp 2 cotdmp 3 cotdlp 4 cotntp 5 cotnhp 6 cotnmp 7 cotnlp */
    for (int i = 0; i < 8 * pncol; ++i) zcot[i] = 0.;
    if (rv) for (int i = 0; i < RV_NOUT * pncol; ++i) zrv[i] = 0.;   /* :681-745 */
    if (cc == 2)
        for (int icol = 1; icol <= ncol; ++icol)
            for (int iw = 1; iw <= NG; ++iw) {
                const int jb = T->ngb[iw - 1], ibm = jb - 15;
                double wgt;
                if (ibm >= 10 && ibm <= 11) wgt = 1.0;
                else if (ibm == 9) wgt = 0.5;
                else continue;
                const double zincflx = S->isolvar < 0 ? S->adjflux[jb - 16] * W->zsflxzen[IGC(iw, icol)]
                                                      : S->adjflux[jb - 16] * W->ssi[IGC(iw, icol)];
                wgt = wgt * zincflx;
                double staolp = 0., staomp = 0., staohp = 0.;
                for (int lay = 1; lay <= cloudLM; ++lay) staolp = staolp + ptaormc[I3(lay, iw, icol)];
                for (int lay = cloudLM + 1; lay <= cloudMH; ++lay) staomp = staomp + ptaormc[I3(lay, iw, icol)];
                for (int lay = cloudMH + 1; lay <= nlay; ++lay) staohp = staohp + ptaormc[I3(lay, iw, icol)];
                const int i = icol - 1;
                if (staolp > 0.) { zcot[3 * pncol + i] += wgt; zcot[7 * pncol + i] += wgt * staolp; }
                if (staomp > 0.) { zcot[2 * pncol + i] += wgt; zcot[6 * pncol + i] += wgt * staomp; }
                if (staohp > 0.) { zcot[1 * pncol + i] += wgt; zcot[5 * pncol + i] += wgt * staohp; }
                const double staotp = staolp + staomp + staohp;
                if (staotp > 0.) { zcot[0 * pncol + i] += wgt; zcot[4 * pncol + i] += wgt * staotp; }
                if (rv) radval_sums(rv, ptaucmc, nlay, cloudLM, cloudMH, iw, icol, wgt, zrv + i, pncol);
            }
#undef ZINCFLX
}

/* ------------------------------------------------------------------------------------------
 * rrtmg_sw_sub partitions, SW/src/rrtmg_sw_rad.F90:1164-1759
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    const double *coszen, *play, *plev, *tlay, *h2ovmr, *o3vmr, *co2vmr, *ch4vmr, *o2vmr, *cld,
        *ciwp, *clwp, *rei, *rel, *zm, *alat, *tauaer, *ssaaer, *asmaer, *asdir, *asdif, *aldir,
        *aldif;
} SwIn;

typedef struct {
    int *clearCounts;
    double *swuflx, *swdflx, *swuflxc, *swdflxc, *nirr, *nirf, *parr, *parf, *uvrr, *uvrf,
        *fswband, *cot[8], *drband, *dfband;
    double *radval;   /* SOLAR_RADVAL build: (gncol, RV_NOUT) or NULL */
} SwOut;

static int sw_partition(int cc, const int *gcols /* 0-based global columns */, int ncol, int gncol,
                        int nlay, const SwIn *in, const Solar *S, int iceflgsw, int liqflgsw,
                        int dyofyr, int iaer, int cloudLM, int cloudMH, int do_drfband,
                        const SwOut *out, OracleTaps *taps) {
    const int pncol = ncol;
    const double amd = 28.9660, amw = 18.0160, zepzen = 1.e-10;
    const double avogad = g_sw.avogad, grav = g_sw.grav;
    const size_t n2 = (size_t)nlay * pncol, n3 = n2 * NG, n3p = (size_t)(nlay + 1) * NG * pncol,
                 nb3 = n2 * NB, np = (size_t)(nlay + 1) * pncol;
    const size_t total = n2 * 24 + np * 5 + (size_t)pncol * 12 + nb3 * 3 + (size_t)NB * pncol * 5 +
                         n3 * 12 + n3p * 11 + (size_t)NG * pncol * 2 + (size_t)8 * pncol;
    double *buf = (double *)zalloc(sizeof(double) * total), *p = buf;
    if (!buf) return -5;
#define TAKE(name, cnt) double *name = p; p += (cnt)
    TAKE(play, n2); TAKE(tlay, n2); TAKE(cld, n2); TAKE(ciwp, n2); TAKE(clwp, n2); TAKE(rei, n2);
    TAKE(rel, n2); TAKE(zm, n2); TAKE(coldry, n2); TAKE(colh2o, n2); TAKE(colco2, n2); TAKE(colo3, n2);
    TAKE(colch4, n2); TAKE(colo2, n2); TAKE(colmol, n2); TAKE(selffac, n2); TAKE(selffrac, n2);
    TAKE(forfac, n2); TAKE(forfrac, n2); TAKE(fac00, n2); TAKE(fac01, n2); TAKE(fac10, n2); TAKE(fac11, n2);
    TAKE(spare, n2);
    TAKE(plev, np); TAKE(zbbfd, np); TAKE(zbbfu, np); TAKE(zbbcd, np); TAKE(zbbcu, np);
    TAKE(alat, pncol); TAKE(coszen, pncol); TAKE(cossza, pncol); TAKE(znirr, pncol); TAKE(znirf, pncol);
    TAKE(zparr, pncol); TAKE(zparf, pncol); TAKE(zuvrr, pncol); TAKE(zuvrf, pncol); TAKE(sp1, pncol);
    TAKE(sp2, pncol); TAKE(sp3, pncol);
    TAKE(taua, nb3); TAKE(asya, nb3); TAKE(omga, nb3);
    TAKE(albdir, (size_t)NB * pncol); TAKE(albdif, (size_t)NB * pncol); TAKE(fndsbnd, (size_t)NB * pncol);
    TAKE(zdrband, (size_t)NB * pncol); TAKE(zdfband, (size_t)NB * pncol);
    TAKE(ciwpmcl, n3); TAKE(clwpmcl, n3); TAKE(taormc, n3); TAKE(taucmc, n3); TAKE(ssacmc, n3); TAKE(asmcmc, n3);
    SpcWork W;
    W.zgco = p; p += n3; W.zomco = p; p += n3; W.ztauo = p; p += n3; W.zdbt = p; p += n3;
    W.ztaur = p; p += n3; W.ztaug = p; p += n3;
    W.ztdbt = p; p += n3p; W.zfd = p; p += n3p; W.zfu = p; p += n3p; W.zref = p; p += n3p;
    W.zrefd = p; p += n3p; W.ztra = p; p += n3p; W.ztrad = p; p += n3p; W.w1 = p; p += n3p;
    W.w2 = p; p += n3p; W.w3 = p; p += n3p; W.w4 = p; p += n3p;
    W.zsflxzen = p; p += (size_t)NG * pncol; W.ssi = p; p += (size_t)NG * pncol;
    TAKE(zcot, (size_t)8 * pncol);
#undef TAKE
    (void)spare; (void)sp1; (void)sp2; (void)sp3;
    unsigned char *cldymcl = (unsigned char *)zalloc(n3);
    double *rvbuf = out->radval ? (double *)zalloc(sizeof(double) * (n3 * RV_CELL_COUNT + (size_t)RV_NOUT * pncol)) : NULL;
    double *rvcell[RV_CELL_COUNT], *zrv = rvbuf ? rvbuf + n3 * RV_CELL_COUNT : NULL;
    for (int q = 0; q < RV_CELL_COUNT; ++q) rvcell[q] = rvbuf ? rvbuf + n3 * q : NULL;
    int *ibuf = (int *)zalloc(sizeof(int) * (n2 * 5 + (size_t)pncol * 5));
    SwCoef sc;
    sc.jp = ibuf; sc.jt = ibuf + n2; sc.jt1 = ibuf + 2 * n2; sc.indself = ibuf + 3 * n2; sc.indfor = ibuf + 4 * n2;
    sc.laytrop = ibuf + 5 * n2;
    int *p_clearCounts = ibuf + 5 * n2 + pncol;
    sc.colh2o = colh2o; sc.colco2 = colco2; sc.colo3 = colo3; sc.colch4 = colch4; sc.colo2 = colo2;
    sc.colmol = colmol; sc.coldry = coldry; sc.selffac = selffac; sc.selffrac = selffrac; sc.forfac = forfac;
    sc.forfrac = forfrac; sc.fac00 = fac00; sc.fac01 = fac01; sc.fac10 = fac10; sc.fac11 = fac11;
    int rc = 0;

    /* gather + albedo band map (:1217-1359) */
    for (int j = 0; j < ncol; ++j) {
        const size_t g = (size_t)gcols[j];
        for (int ib = 1; ib <= 8; ++ib) { albdir[(ib - 1) + NB * j] = in->aldir[g]; albdif[(ib - 1) + NB * j] = in->aldif[g]; }
        albdir[(NB - 1) + NB * j] = in->aldir[g];
        albdif[(NB - 1) + NB * j] = in->aldif[g];
        for (int ib = 10; ib <= 13; ++ib) { albdir[(ib - 1) + NB * j] = in->asdir[g]; albdif[(ib - 1) + NB * j] = in->asdif[g]; }
        albdir[8 + NB * j] = (in->asdir[g] + in->aldir[g]) / 2.;
        albdif[8 + NB * j] = (in->asdif[g] + in->aldif[g]) / 2.;
        coszen[j] = in->coszen[g];
        alat[j] = in->alat[g];
        for (int l = 0; l < nlay; ++l) {
            const size_t d = (size_t)l + (size_t)nlay * j, s = g + (size_t)gncol * l;
            play[d] = in->play[s]; tlay[d] = in->tlay[s];
            colh2o[d] = in->h2ovmr[s]; colco2[d] = in->co2vmr[s]; colo3[d] = in->o3vmr[s];
            colch4[d] = in->ch4vmr[s]; colo2[d] = in->o2vmr[s];
            if (cc == 2) {
                cld[d] = in->cld[s]; ciwp[d] = in->ciwp[s]; clwp[d] = in->clwp[s]; rei[d] = in->rei[s];
                rel[d] = in->rel[s]; zm[d] = in->zm[s];
            }
            for (int ib = 0; ib < NB; ++ib) {
                const size_t da = (size_t)l + (size_t)nlay * (ib + (size_t)NB * j);
                if (iaer == 10) {
                    const size_t sa = g + (size_t)gncol * (l + (size_t)nlay * ib);
                    taua[da] = in->tauaer[sa]; asya[da] = in->asmaer[sa]; omga[da] = in->ssaaer[sa];
                } else {
                    taua[da] = 0.; asya[da] = 0.; omga[da] = 1.;
                }
            }
        }
        for (int l = 0; l <= nlay; ++l) plev[(size_t)l + (size_t)(nlay + 1) * j] = in->plev[g + (size_t)gncol * l];
    }
    for (int j = 0; j < ncol; ++j) cossza[j] = f_max(zepzen, coszen[j]);
    for (int j = 0; j < ncol; ++j)
        for (int l = 0; l < nlay; ++l) {
            const size_t d = (size_t)l + (size_t)nlay * j;
            coldry[d] = (plev[(size_t)l + (size_t)(nlay + 1) * j] - plev[(size_t)l + 1 + (size_t)(nlay + 1) * j]) *
                        1.e3 * avogad / (1.e2 * grav * ((1. - colh2o[d]) * amd + colh2o[d] * amw) * (1. + colh2o[d]));
        }
    for (size_t d = 0; d < (size_t)nlay * ncol; ++d) {
        colh2o[d] = coldry[d] * colh2o[d];
        colco2[d] = coldry[d] * colco2[d];
        colo3[d] = coldry[d] * colo3[d];
        colch4[d] = coldry[d] * colch4[d];
        colo2[d] = coldry[d] * colo2[d];
    }
    if (cc == 2) {
        static const int seed_order[4] = {4, 3, 2, 1};
        rc = oracle_generate_stochastic_clouds(pncol, ncol, NG, nlay, zm, alat, dyofyr, play, cld, ciwp, clwp,
                                               1.e-20, cldymcl, ciwpmcl, clwpmcl, seed_order);
        if (!rc) rc = oracle_clearCounts_threeBand(pncol, ncol, NG, nlay, cloudLM, cloudMH, cldymcl, p_clearCounts);
        if (!rc)
            rc = cldprmc_sw(ncol, nlay, iceflgsw, liqflgsw, cldymcl, ciwpmcl, clwpmcl, rei, rel, taormc, taucmc,
                            ssacmc, asmcmc, rvbuf ? rvcell : NULL);
    }
    if (!rc) {
        setcoef_sw(&sc, ncol, nlay, play, tlay);
        spcvmc_sw(cc, ncol, nlay, &sc, S, albdif, albdir, cldymcl, taucmc, asmcmc, ssacmc, taormc, taua, asya,
                  omga, cossza, cloudLM, cloudMH, &W, zbbfd, zbbfu, zbbcd, zbbcu, znirr, znirf, zparr, zparf,
                  zuvrr, zuvrf, fndsbnd, pncol, zcot, do_drfband, zdrband, zdfband, taps, gcols, gncol,
                  rvbuf ? rvcell : NULL, zrv);
        for (int j = 0; j < ncol; ++j) {
            const size_t g = (size_t)gcols[j];
            for (int n = 0; n < 4; ++n)
                out->clearCounts[g + (size_t)gncol * n] = cc == 1 ? NG : p_clearCounts[n + 4 * j];
            for (int l = 0; l <= nlay; ++l) {
                const size_t d = g + (size_t)gncol * l, s = (size_t)l + (size_t)(nlay + 1) * j;
                out->swuflxc[d] = zbbcu[s]; out->swdflxc[d] = zbbcd[s];
                out->swuflx[d] = zbbfu[s]; out->swdflx[d] = zbbfd[s];
            }
            for (int q = 0; q < 8; ++q) out->cot[q][g] = cc == 1 ? 0. : zcot[(size_t)q * pncol + j];
            if (out->radval)   /* SW/src/rrtmg_sw_rad.F90:1539-1603 (zeros), :1657-1726 */
                for (int q = 0; q < RV_NOUT; ++q)
                    out->radval[g + (size_t)gncol * q] = cc == 1 ? 0. : zrv[(size_t)q * pncol + j];
            out->nirr[g] = znirr[j]; out->nirf[g] = znirf[j] - znirr[j];
            out->parr[g] = zparr[j]; out->parf[g] = zparf[j] - zparr[j];
            out->uvrr[g] = zuvrr[j]; out->uvrf[g] = zuvrf[j] - zuvrr[j];
            for (int ib = 0; ib < NB; ++ib) {
                out->fswband[g + (size_t)gncol * ib] = fndsbnd[j + pncol * ib];
                if (do_drfband) {
                    out->drband[g + (size_t)gncol * ib] = zdrband[j + pncol * ib];
                    out->dfband[g + (size_t)gncol * ib] = zdfband[j + pncol * ib];
                }
            }
        }
        if (taps)
            for (int j = 0; j < ncol; ++j) {
                const size_t g = (size_t)gcols[j];
                if (taps->laytrop) taps->laytrop[g] = sc.laytrop[j];
                for (int l = 0; l < nlay; ++l) {
                    const size_t d = g + (size_t)gncol * l, s = (size_t)l + (size_t)nlay * j;
                    if (taps->jp) taps->jp[d] = sc.jp[s];
                    if (taps->jt) taps->jt[d] = sc.jt[s];
                    if (taps->jt1) taps->jt1[d] = sc.jt1[s];
                    if (taps->indfor) taps->indfor[d] = sc.indfor[s];
                    if (taps->indself) taps->indself[d] = sc.indself[s];
                    if (taps->fac00) taps->fac00[d] = sc.fac00[s];
                    if (taps->fac01) taps->fac01[d] = sc.fac01[s];
                    if (taps->fac10) taps->fac10[d] = sc.fac10[s];
                    if (taps->fac11) taps->fac11[d] = sc.fac11[s];
                }
                const size_t o3 = (size_t)nlay * NG * g, s3 = (size_t)nlay * NG * j, cnt = (size_t)nlay * NG;
                if (cc == 2) {
                    if (taps->cldymc) memcpy(taps->cldymc + o3, cldymcl + s3, cnt);
                    if (taps->ciwpmc) memcpy(taps->ciwpmc + o3, ciwpmcl + s3, cnt * 8);
                    if (taps->clwpmc) memcpy(taps->clwpmc + o3, clwpmcl + s3, cnt * 8);
                    if (taps->taucmc) memcpy(taps->taucmc + o3, taucmc + s3, cnt * 8);
                }
            }
    }
    free(buf); free(cldymcl); free(ibuf); free(rvbuf);
    return rc;
}

/* _ASSERT(all(x >= 0.)) (SW/src/rrtmg_sw_rad.F90:365-383): unlike the LW driver's any(x < 0.) this also fires on NaN */
static int any_negative(const double *x, size_t n) {
    for (size_t i = 0; i < n; ++i)
        if (!(x[i] >= 0.)) return 1;
    return 0;
}

/* SW/src/rrtmg_sw_rad.F90:68-452 + rrtmg_sw_sub */
int oracle_rrtmg_sw(
    int rpart, int ncol, int nlay, double scon, double adjes, const double *coszen, int isolvar,
    const double *play, const double *plev, const double *tlay, const double *h2ovmr,
    const double *o3vmr, const double *co2vmr, const double *ch4vmr, const double *o2vmr,
    int iceflgsw, int liqflgsw, const double *cld, const double *ciwp, const double *clwp,
    const double *rei, const double *rel, int dyofyr, const double *zm, const double *alat,
    int iaer, const double *tauaer, const double *ssaaer, const double *asmaer,
    const double *asdir, const double *asdif, const double *aldir, const double *aldif,
    int cloudLM, int cloudMH, int normFlx, int *clearCounts, double *swuflx, double *swdflx,
    double *swuflxc, double *swdflxc, double *nirr, double *nirf, double *parr, double *parf,
    double *uvrr, double *uvrf, double *fswband, double *cotdtp, double *cotdhp, double *cotdmp,
    double *cotdlp, double *cotntp, double *cotnhp, double *cotnmp, double *cotnlp,
    int do_drfband, double *drband, double *dfband, const double *bndscl, const double *indsolvar,
    const double *solcycfrac, OracleTaps *taps) {
    const size_t n2 = (size_t)ncol * nlay, n2p = (size_t)ncol * (nlay + 1);
    /* _ASSERTs :365-383, in the reference's order; code = -(100 + position) */
    struct { const double *x; size_t n; } chk[] = {
        {play, n2}, {plev, n2p}, {tlay, n2}, {h2ovmr, n2}, {o3vmr, n2}, {co2vmr, n2}, {ch4vmr, n2},
        {o2vmr, n2}, {asdir, (size_t)ncol}, {aldir, (size_t)ncol}, {asdif, (size_t)ncol},
        {aldif, (size_t)ncol}, {cld, n2}, {ciwp, n2}, {clwp, n2}, {rei, n2}, {rel, n2},
        {tauaer, n2 * NB}, {ssaaer, n2 * NB}};
    for (size_t i = 0; i < sizeof chk / sizeof chk[0]; ++i)
        if (any_negative(chk[i].x, chk[i].n)) return -(101 + (int)i);
    const int pncol = rpart > 0 ? rpart : 2;

    Solar S;
    int rc = solar_setup(&S, isolvar, scon, adjes, bndscl, indsolvar, solcycfrac);
    if (rc) return rc;

    /* clear / cloudy split (:1138-1148) */
    int *gicol_clr = (int *)zalloc(sizeof(int) * ncol), *gicol_cld = (int *)zalloc(sizeof(int) * ncol);
    int ncol_clr = 0, ncol_cld = 0;
    for (int g = 0; g < ncol; ++g) {
        int any = 0;
        for (int l = 0; l < nlay && !any; ++l) any = cld[(size_t)g + (size_t)ncol * l] > 0;
        if (any) gicol_cld[ncol_cld++] = g; else gicol_clr[ncol_clr++] = g;
    }
    SwIn in = {coszen, play, plev, tlay, h2ovmr, o3vmr, co2vmr, ch4vmr, o2vmr, cld, ciwp, clwp, rei, rel,
               zm, alat, tauaer, ssaaer, asmaer, asdir, asdif, aldir, aldif};
    SwOut out = {clearCounts, swuflx, swdflx, swuflxc, swdflxc, nirr, nirf, parr, parf, uvrr, uvrf, fswband,
                 {cotdtp, cotdhp, cotdmp, cotdlp, cotntp, cotnhp, cotnmp, cotnlp}, drband, dfband,
                 taps ? taps->radval : NULL};
    int rc_all = 0;
    for (int cc = 1; cc <= 2; ++cc) {
        const int *list = cc == 1 ? gicol_clr : gicol_cld;
        const int col_last = cc == 1 ? ncol_clr : ncol_cld;
        const int npart = (col_last + pncol - 1) / pncol;
#pragma omp parallel for schedule(dynamic)
        for (int ipart = 0; ipart < npart; ++ipart) {
            const int cols = ipart * pncol;
            int cole = (ipart + 1) * pncol;
            if (cole > col_last) cole = col_last;
            int r = sw_partition(cc, list + cols, cole - cols, ncol, nlay, &in, &S, iceflgsw, liqflgsw, dyofyr,
                                 iaer, cloudLM, cloudMH, do_drfband, &out, taps);
            if (r) {
#pragma omp critical
                if (!rc_all) rc_all = r;
            }
        }
    }
    free(gicol_clr); free(gicol_cld);
    if (rc_all) return rc_all;

    if (normFlx == 1) { /* :1769-1798 */
        for (int g = 0; g < ncol; ++g) {
            const double top = f_max(swdflx[(size_t)g + (size_t)ncol * nlay], 1e-7);
            for (int l = 0; l <= nlay; ++l) {
                const size_t d = (size_t)g + (size_t)ncol * l;
                swuflxc[d] = swuflxc[d] / top; swdflxc[d] = swdflxc[d] / top;
                swuflx[d] = swuflx[d] / top; swdflx[d] = swdflx[d] / top;
            }
            nirr[g] = nirr[g] / top; nirf[g] = nirf[g] / top; parr[g] = parr[g] / top;
            parf[g] = parf[g] / top; uvrr[g] = uvrr[g] / top; uvrf[g] = uvrf[g] / top;
            for (int ib = 0; ib < NB; ++ib) {
                fswband[(size_t)g + (size_t)ncol * ib] = fswband[(size_t)g + (size_t)ncol * ib] / top;
                if (do_drfband) {
                    drband[(size_t)g + (size_t)ncol * ib] = drband[(size_t)g + (size_t)ncol * ib] / top;
                    dfband[(size_t)g + (size_t)ncol * ib] = dfband[(size_t)g + (size_t)ncol * ib] / top;
                }
            }
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Test hooks: the two-stream and the adding routine on their own, so that tests/ can hold them
 * against an independent numpy restatement of SW/src/rrtmg_sw_spcvmc.F90:1115-1588.
 * Arrays as in the routines: inputs (nlay,ngpt,ncol), outputs and level arrays (nlay+1,ngpt,ncol).
 * ---------------------------------------------------------------------------------------- */
int oracle_reftra_sw(int ncol, int nlay, const double *pgg, const double *prmuz, const double *ptau,
                     const double *pw, double *pref, double *prefd, double *ptra, double *ptrad) {
    reftra_sw(ncol, nlay, NULL, pgg, prmuz, ptau, pw, pref, prefd, ptra, ptrad, 0);
    return 0;
}

int oracle_vrtqdr_sw(int ncol, int nlay, const double *pref, const double *prefd, const double *ptra,
                     const double *ptrad, const double *pdbt, const double *ptdbt, double *pfd, double *pfu) {
    const size_t n = (size_t)(nlay + 1) * NG * (size_t)ncol;
    double *w = (double *)malloc(4 * n * sizeof(double));
    if (!w) return -1;
    vrtqdr_sw(ncol, nlay, pref, prefd, ptra, ptrad, pdbt, ptdbt, pfd, pfu, w, w + n, w + 2 * n, w + 3 * n);
    free(w);
    return 0;
}
