/*
 * oracle/oracle.h -- CPU restatement of the reference RRTMG LW + SW + McICA column path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (geosradiation_gridcomp_b200/) may
 * include, link or call this.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, as the checker and the timed CPU baseline.
 *
 * PARITY PINNED BY THE REFERENCE'S OWN SOURCE TEXT: the reference tree ships no golden vectors,
 * known-answer tests or fixtures for this path and no Fortran compiler
 * exists in the build container, so the reference is executed by translation (oracle/refexec/:
 * the 91 Fortran files of the path, read where they lie, turned into Python in memory) and its
 * output is committed as tests/golden/rrtmg_refexec_golden.npz.  tests/test_refexec_pin_cpu.py
 * holds this restatement to those numbers: integers bit for bit, reals <= 1e-12 (seen: 3.9e-13).
 * The drivers' Run-phase glue (glue.c) is held bit for bit to the reference's own line ranges,
 * read from the driver files and executed (oracle/refexec/glue.py).  Not pinned: output of a
 * COMPILED reference (oracle/build_ref.sh is the recipe; it needs a Fortran compiler).  This file is a routine-by-routine restatement of the Fortran, with `real`
 * promoted to 8 bytes (the north_star fp64 contract), same loop nests, same expression order,
 * compiled with -ffp-contract=off.
 *
 * Array layouts are those of the reference driver interfaces (column index fastest):
 *   x(ncol,nlay) -> x[icol + ncol*ilay].
 */
#ifndef RRTMG_ORACLE_H
#define RRTMG_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

enum { NBNDLW = 16, NGPTLW = 140, NBNDSW = 14, NGPTSW = 112 };

/* Optional taps on intermediates for parity tests; any pointer may be NULL. */
typedef struct {
    int *jp, *jt, *jt1;                /* (ncol,nlay)  1-based reference indices          */
    int *indfor, *indself, *indminor;  /* (ncol,nlay)  (indself/indminor: LW lower only)  */
    int *laytrop;                      /* (ncol)                                          */
    double *fac00, *fac01, *fac10, *fac11; /* (ncol,nlay)                                 */
    unsigned char *cldymc;             /* [icol][ig][ilay] binary McICA cloud mask        */
    double *ciwpmc, *clwpmc;           /* [icol][ig][ilay]                                */
    double *taug, *pfracs;             /* LW: [icol][ig][ilay]; SW: taug, taur            */
    double *taucmc;                    /* [icol][ig][ilay] (LW: absorption od; SW: delta-scaled) */
    double *pwvcm;                     /* (ncol) LW only                                  */
    double *ssi;                       /* SW: [icol][ig] solar source per g-point          */
    double *radval;                    /* SW, not a tap but a switch: non-NULL runs the SOLAR_RADVAL build of
                                        * rrtmg_sw and receives its 120 extra dummies (SW/src/rrtmg_sw_rad.F90:85-122),
                                        * (ncol,120) column fastest, in the order of the dummy list               */
} OracleTaps;

int oracle_init(const char *blob_path);
void oracle_finalize(void);
/* ih: 0 homogeneous, 1 beta, 2 gamma (SH/cloud_condensate_inhomogeneity.F90:45-73);
 * corr = {adl_am1, adl_am2, adl_am30, adl_am4, rdl_am1, rdl_am2, rdl_am30, rdl_am4}
 * (SH/cloud_subcol_gen.F90:108-129); NULL keeps the Oreopoulos 2012 defaults. */
int oracle_set_mcica(int ih, const double *corr);
int oracle_num_threads(void);

/* LW/src/rrtmg_lw_rad.F90:15-344.  Returns 0 or a negative code mirroring an error stop. */
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
    const int *band_output, double *olrb, double *dolrb_dTs,
    OracleTaps *taps);

/* SW/src/rrtmg_sw_rad.F90:68-1801 (the default build; the SOLAR_RADVAL build when taps->radval is set).  bndscl/indsolvar/solcycfrac may
 * be NULL (absent optional arguments). drband/dfband only touched when do_drfband. */
int oracle_rrtmg_sw(
    int rpart, int ncol, int nlay,
    double scon, double adjes, const double *coszen, int isolvar,
    const double *play, const double *plev, const double *tlay,
    const double *h2ovmr, const double *o3vmr, const double *co2vmr, const double *ch4vmr,
    const double *o2vmr, int iceflgsw, int liqflgsw,
    const double *cld, const double *ciwp, const double *clwp, const double *rei,
    const double *rel, int dyofyr, const double *zm, const double *alat,
    int iaer, const double *tauaer, const double *ssaaer, const double *asmaer,
    const double *asdir, const double *asdif, const double *aldir, const double *aldif,
    int cloudLM, int cloudMH, int normFlx,
    int *clearCounts, double *swuflx, double *swdflx, double *swuflxc, double *swdflxc,
    double *nirr, double *nirf, double *parr, double *parf, double *uvrr, double *uvrf,
    double *fswband,
    double *cotdtp, double *cotdhp, double *cotdmp, double *cotdlp,
    double *cotntp, double *cotnhp, double *cotnmp, double *cotnlp,
    int do_drfband, double *drband, double *dfband,
    const double *bndscl, const double *indsolvar, const double *solcycfrac,
    OracleTaps *taps);

/* stand-alone McICA generator + clear counts for unit tests
 * (SH/cloud_subcol_gen.F90:132-487, 611-769); arrays in the (nlay,dncol) partition layout. */
int oracle_generate_stochastic_clouds(
    int dncol, int ncol, int nsubcol, int nlay,
    const double *zmid, const double *alat, int doy,
    const double *play, const double *cldfrac, const double *ciwp, const double *clwp,
    double cwp_tiny, unsigned char *cldy_stoch, double *ciwp_stoch, double *clwp_stoch,
    const int *seed_order);
int oracle_clearCounts_threeBand(int dncol, int ncol, int nsubcol, int nlay, int cloudLM,
                                 int cloudMH, const unsigned char *cldy_stoch, int *clearCnts);
void oracle_rng_kiss(int *s1, int *s2, int *s3, int *s4, double *ran);

/* ---- Run-phase glue around the RRTMG calls (oracle/glue.c) ----
 * GEOS-native state: (ncol,LM) arrays with level 1 at the model top, PLE (ncol,LM+1). */
typedef struct {
    int ncol, lm, iceflg, liqflg, lcldmh, lcldlm;   /* lcld*: GEOS top-down interface levels */
    double co2_fixed, o2, ccl4;
    double airmw, h2omw, o3mw, rgas, grav;          /* MAPL_AIRMW, _H2OMW, _O3MW, _RGAS, _GRAV */
    const double *ple, *pl, *t, *q, *o3, *ch4, *n2o, *co2 /* (ncol,LM) or NULL */, *cfc11, *cfc12, *hcfc22, *fcld;
    const double *qliq, *qice, *rliq, *rice;        /* CWC / REFF, KLIQUID and KICE */
    const double *ts, *t2m, *emis, *lats;           /* (ncol) */
    const double *taua, *ssaa;                      /* (ncol,LM,16) or NULL */
} OracleIrradState;
typedef struct {                                     /* the arguments of rrtmg_lw, caller-allocated */
    int cloudLM, cloudMH;
    double *play, *plev, *tlay, *tlev, *tsfc, *emis, *h2ovmr, *o3vmr, *co2vmr, *ch4vmr, *n2ovmr, *o2vmr,
        *cfc11vmr, *cfc12vmr, *cfc22vmr, *ccl4vmr, *cldf, *ciwp, *clwp, *rei, *rel, *tauaer, *zm, *alat;
} OracleLwInputs;
typedef struct {                                     /* GEOS convention: (ncol,0:LM) top-down, upward negative */
    double *flxu, *flxd, *flcu, *flcd, *dfdts, *dfdtsc, *sfcem;
    double *cldtt, *cldhi, *cldmd, *cldlo;           /* any may be NULL */
} OracleIrradFluxes;
int oracle_irrad_prepare(const OracleIrradState *s, OracleLwInputs *o);   /* IRR:3237-3371 */
int oracle_irrad_finish(int ncol, int lm, const double *emis, const int *clearCounts, const double *uflx,
                        const double *dflx, const double *uflxc, const double *dflxc, const double *duflx_dTs,
                        const double *duflxc_dTs, OracleIrradFluxes *f);  /* IRR:3486-3533 */

typedef struct {                                     /* exports of the between-refresh Update, any may be NULL */
    double *flx, *flc, *flxu, *flcu, *flxd, *flcd;   /* (ncol,0:LM) */
    double *olr, *olc, *sfcem, *lws, *lcs, *flns, *flnsc;   /* (ncol) */
} OracleIrradExports;
int oracle_irrad_update(int ncol, int lm, const double *flxu_int, const double *flxd_int, const double *flcu_int,
                        const double *flcd_int, const double *dfdts, const double *dfdtsc, const double *sfcem_int,
                        const double *ts_int, const double *tsinst, OracleIrradExports *e);   /* IRR:3861, 3929-3990 */

typedef struct {
    int ncol, lm, iceflg, liqflg, lcldmh, lcldlm;
    double co2, o2;
    double airmw, h2omw, o3mw, rgas, grav;
    const double *ple, *pl, *t, *q, *o3, *ch4, *cl;
    const double *qliq, *qice, *rliq, *rice;        /* QQ3(:,:,2), QQ3(:,:,1), RR3(:,:,2), RR3(:,:,1) */
    const double *ts;                               /* (ncol) */
    const double *taua, *ssaa, *asya;               /* (ncol,LM,14) un-normalised, or NULL */
} OracleSolarState;
typedef struct {
    int cloudLM, cloudMH;
    double *play, *plev, *tlay, *h2ovmr, *o3vmr, *co2vmr, *ch4vmr, *o2vmr, *cld, *ciwp, *clwp, *rei, *rel, *zm,
        *tauaer, *ssaaer, *asmaer;
} OracleSwInputs;
typedef struct {
    double *fsw, *fsc, *fswu, *fscu;                 /* (ncol,LM+1) top-down */
    double *cldts, *cldhs, *cldms, *cldls, *cottp, *cothp, *cotmp, *cotlp;   /* (ncol), any may be NULL */
} OracleSolarFluxes;
int oracle_solar_prepare(const OracleSolarState *s, OracleSwInputs *o);   /* SOL:6113-6223 */
int oracle_solar_finish(int ncol, int lm, double undef, const int *clearCounts, const double *swuflx,
                        const double *swdflx, const double *swuflxc, const double *swdflxc, const double *cotd[4],
                        const double *cotn[4], OracleSolarFluxes *f);     /* SOL:6395-6444 */

/* test hooks: reftra_sw / vrtqdr_sw alone (SW/src/rrtmg_sw_spcvmc.F90:1115-1588), arrays (nlay,ngpt,ncol) in,
 * (nlay+1,ngpt,ncol) out, for the independent numpy restatement in tests/test_oracle_cpu.py */
int oracle_reftra_sw(int ncol, int nlay, const double *pgg, const double *prmuz, const double *ptau,
                     const double *pw, double *pref, double *prefd, double *ptra, double *ptrad);
int oracle_vrtqdr_sw(int ncol, int nlay, const double *pref, const double *prefd, const double *ptra,
                     const double *ptrad, const double *pdbt, const double *ptdbt, double *pfd, double *pfu);

/* reduced (post-cmbgb) tables and lookup tables, for tests of the init restatement */
const double *oracle_lw_table(const char *name, int band, int *n);
const double *oracle_sw_table(const char *name, int band, int *n);

#ifdef __cplusplus
}
#endif
#endif
