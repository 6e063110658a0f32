/*
 * rrtmgx.h -- C ABI of the B200-native RRTMG LW + SW + McICA column path.
 *
 * Drop-in boundary: these entry points are what a Fortran ISO_C_BINDING shim (see
 * geosradiation_gridcomp_b200/fortran/ and INTEGRATION.md) binds to replace the bodies of
 *   rrtmg_lw            GEOSirrad_GridComp/RRTMG/rrtmg_lw/gcm_model/src/rrtmg_lw_rad.F90:15-23
 *   rrtmg_sw            GEOSsolar_GridComp/RRTMG/rrtmg_sw/gcm_model/src/rrtmg_sw_rad.F90:68-124
 *   rrtmg_lw_ini        .../rrtmg_lw/gcm_model/src/rrtmg_lw_init.F90:22
 *   rrtmg_sw_ini        .../rrtmg_sw/gcm_model/src/rrtmg_sw_init.F90:49
 *   set_inhomogeneity   GEOS_RadiationShared/cloud_condensate_inhomogeneity.F90:45
 *   initialize_cloud_subcol_gen  GEOS_RadiationShared/cloud_subcol_gen.F90:108
 * called from GEOS_IrradGridComp.F90:3381,3471, GEOS_SolarGridComp.F90:6225,6331 and
 * GEOS_RadiationGridComp.F90:565,578.
 *
 * All arrays are the caller's, in the reference layout: column-major with the COLUMN index
 * fastest, x(ncol,nlay) -> x[icol + ncol*ilay]; layer 1 at the surface; fp64 ("real" promoted
 * to 8 bytes, the precision contract of this build).  Pointers are HOST pointers unless
 * RRTMGX_DEVICE_PTRS is set in `flags`, in which case every array pointer is a device pointer
 * on the current CUDA device and no copies are made.  logical arguments are int32 (0/1).
 *
 * Every function returns 0 on success or a negative status that mirrors the reference trap
 * (`error stop` in LW, _ASSERT/_FAIL -> RC in SW); rrtmgx_strerror() gives the message.
 * There is no CPU fallback: without a CUDA device rrtmgx_init fails with RRTMGX_ENODEVICE.
 */
#ifndef RRTMGX_H
#define RRTMGX_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum {
    RRTMGX_NBNDLW = 16, RRTMGX_NGPTLW = 140,   /* LW/modules/parrrtm.F90:39 */
    RRTMGX_NBNDSW = 14, RRTMGX_NGPTSW = 112    /* SW/modules/parrrsw.F90:32 */
};

/* flags */
enum {
    RRTMGX_DEVICE_PTRS = 1,   /* array arguments are device pointers                         */
    RRTMGX_NO_SYNC     = 2,   /* device pointers only: return without synchronising; status  */
                              /* traps are then reported by rrtmgx_{lw,sw}_status()           */
    RRTMGX_SKIP_CHECKS = 4,   /* skip the negative-input scans (LW :209-318, SW :365-383)    */
    RRTMGX_KEEP_STATUS = 8,   /* device pointers only: do not clear the path's status word    */
                              /* first, so one status read covers a sequence of NO_SYNC runs  */
    RRTMGX_F32_ARRAYS = 32,   /* every real array argument is real*4 (the production kind of    */
                              /* GEOS); widened exactly on the device while staged, arithmetic  */
                              /* stays fp64, outputs rounded once to real*4.  Halves the bytes  */
                              /* that cross PCIe.  With RRTMGX_DEVICE_PTRS the real*4 arrays are */
                              /* staged device to device (the call is synchronous; NO_SYNC is   */
                              /* refused).  (bndscl, indsolvar and solcycfrac stay double;      */
                              /* clearCounts stays int32.)                                      */
    RRTMGX_REUSE_CLOUDS = 16, /* the cloud inputs (cldf, ciwp, clwp, rei, rel, zm, play, alat,  */
                              /* dyofyr, ice/liq flags, cloudLM/MH) are those of the previous   */
                              /* call on this path: keep its McICA subcolumns, cloud optics and */
                              /* clear counts instead of regenerating them (GEOS calls rrtmg_lw */
                              /* once more per removed gas, IRR:3405-3468, and rrtmg_sw with    */
                              /* and without aerosols, SOL:3249-3287, on one cloud state).      */
                              /* Honoured when the previous run of the path was this very chunk */
                              /* of columns of a call of the same extent (a call that crosses   */
                              /* in several device or host staging chunks leaves only its last  */
                              /* chunk's clouds behind); otherwise the clouds are regenerated.  */
    RRTMGX_LIT_ONLY = 64      /* rrtmgx_solar_refresh only: run the daytime columns alone.  The  */
                              /* Solar driver packs the soundings with ZTH > 0 before SORADCORE  */
                              /* and unpacks afterwards (GEOS_SolarGridComp.F90:3686-3687,       */
                              /* PackIt / UnPackIt :7753-7799); here the glue kernels gather     */
                              /* the native state of the lit columns (ascending order, as PackIt */
                              /* does) into the rrtmg_sw arguments and scatter the results back, */
                              /* so night columns cost nothing.  Night columns receive UnPackIt's */
                              /* DEFAULT: 0 in the fluxes and surface components (a dark sun),   */
                              /* `undef` in the cloud fractions and optical thicknesses (the     */
                              /* DEFAULT = MAPL_UNDEF of their internal specs, :873-989).  The   */
                              /* call reads the zenith cosines back first (one stream            */
                              /* synchronisation at its start); NO_SYNC is refused.              */
};

/* status codes (negative) */
enum {
    RRTMGX_OK = 0,
    RRTMGX_ENODEVICE = -1,      /* no CUDA device / CUDA error                               */
    RRTMGX_ENOTINIT = -2,       /* rrtmgx_init not called                                    */
    RRTMGX_EBLOB = -3,          /* table blob missing or malformed                           */
    RRTMGX_EARG = -4,           /* bad scalar argument (ncol, nlay, flags ...)               */
    RRTMGX_ECUDA = -5,          /* CUDA runtime failure during a run                         */
    RRTMGX_EINHOMO = -6,        /* set_inhomogeneity: unknown ih                             */
    RRTMGX_ESEEDORDER = -11,    /* cloud_subcol_gen.F90:273-300                              */
    RRTMGX_ESUPERLAYER = -21,   /* cloud_subcol_gen.F90:762-766 'invalid pressure super-layers' */
    RRTMGX_EPRESSURE = -31,     /* rrtmg_lw_setcoef.F90:445-453 'RRTMG LW pressure misordering' */
    RRTMGX_EICEFLAG = -41,      /* cldprmc: invalid iceflag                                  */
    RRTMGX_ERADIUS_ICE = -42,   /* cldprmc: ice radius extrapolation forbidden               */
    RRTMGX_ELIQFLAG = -51,      /* cldprmc: invalid liqflag                                  */
    RRTMGX_ERADIUS_LIQ = -52,   /* cldprmc: liquid radius extrapolation forbidden            */
    RRTMGX_ESOLVAR = -61,       /* rrtmg_sw_rad.F90:910,1033,1117 bad isolvar / missing optional */
    RRTMGX_ENEGATIVE = -100     /* -(100+k): k-th checked input array has a negative value   */
};

typedef struct {
    const char *table_blob;   /* path of rrtmg_tables.bin; NULL -> $RRTMGX_TABLES or the     */
                              /* file next to the library                                     */
    int device;               /* CUDA device ordinal, -1 = current                            */
    int inhomogeneity;        /* ih: 0 homogeneous, 1 beta (GEOS default), 2 gamma; -1 = default */
    const double *corr;       /* 8 correlation-length parameters or NULL for the defaults     */
} RrtmgxConfig;

/* rrtmg_lw_ini + rrtmg_sw_ini (+ the initial set_inhomogeneity / initialize_cloud_subcol_gen state).
 * Idempotent: GEOS calls the _ini routines on every refresh (IRR:3381, SOL:6225) and the reference's
 * _ini routines never touch the McICA module state, so a call on an initialised library is a pure
 * no-op whatever cfg holds -- inhomogeneity and corr are applied by the FIRST call only; afterwards
 * rrtmgx_set_mcica is the one way to change them.  After rrtmgx_finalize the next call starts from
 * the built-in defaults and the environment again (nothing of the previous life is inherited). */
int rrtmgx_init(const RrtmgxConfig *cfg);
int rrtmgx_set_mcica(int ih, const double corr[8]);
/* The tuning knobs in force and the McICA inhomogeneity type, for tests and diagnostics:
 * knobs[0] = RRTMGX_CHUNK (0 = automatic), [1] = RRTMGX_HOST_CHUNK (default 8192 columns; real*4 arrays take twice
 * the default: the same bytes per staging chunk; a value from the environment is taken literally), [2] = RRTMGX_STAGES
 * (default 2), [3] = ih.  Returns RRTMGX_ENOTINIT before rrtmgx_init. */
int rrtmgx_get_knobs(long long knobs[4]);
int rrtmgx_finalize(void);
const char *rrtmgx_strerror(int status);
/* number of kernels launched by this library since rrtmgx_init (bench.py gpu_launches) */
long long rrtmgx_launch_count(void);
/* Per-kernel device timing (the reference brackets its stages with MAPL timers,
 * SW/src/rrtmg_sw_rad.F90:1181-1200): while enabled every kernel launch is bracketed by CUDA
 * events on its stream and waited for, so a profiled step is serialised and slower.
 * rrtmgx_profile(1) clears the totals and starts, rrtmgx_profile(0) stops;
 * rrtmgx_profile_report writes "name<TAB>launches<TAB>total_ms" lines into buf (NUL terminated,
 * truncated to cap) and returns the size needed. */
void rrtmgx_profile(int enable);
size_t rrtmgx_profile_report(char *buf, size_t cap);

typedef struct {
    int ncol, nlay;
    int psize;                /* accepted and ignored (cache blocking of the CPU code)        */
    int dudTs;                /* logical                                                      */
    int iceflglw, liqflglw;
    int dyofyr, cloudLM, cloudMH;
    int flags;
    void *stream;             /* cudaStream_t or NULL for the library's LW stream             */
    /* inputs */
    const double *play, *plev, *tlay, *tlev;          /* (ncol,nlay) (ncol,0:nlay) ...        */
    const double *tsfc, *emis;                        /* (ncol) (ncol,16)                     */
    const double *h2ovmr, *o3vmr, *co2vmr, *ch4vmr, *n2ovmr, *o2vmr;
    const double *cfc11vmr, *cfc12vmr, *cfc22vmr, *ccl4vmr;
    const double *cldf, *ciwp, *clwp, *rei, *rel;     /* (ncol,nlay)                          */
    const double *tauaer;                             /* (ncol,nlay,16)                       */
    const double *zm, *alat;                          /* (ncol,nlay) (ncol)                   */
    const int32_t *band_output;                       /* (16) logical                         */
    /* outputs */
    int32_t *clearCounts;                             /* (ncol,4)                             */
    double *uflx, *dflx, *uflxc, *dflxc;              /* (ncol,nlay+1)                        */
    double *duflx_dTs, *duflxc_dTs;                   /* (ncol,nlay+1), written iff dudTs     */
    double *olrb, *dolrb_dTs;                         /* (16,ncol), bands with band_output    */
} RrtmgxLwArgs;

typedef struct {
    int ncol, nlay;
    int rpart;                /* accepted and ignored                                         */
    int isolvar;              /* -1,0,1,2,3 (rrtmg_sw_rad.F90:889-1127)                       */
    int iceflgsw, liqflgsw;
    int dyofyr, cloudLM, cloudMH;
    int iaer;                 /* 0 or 10                                                      */
    int normFlx;              /* logical                                                      */
    int do_drfband;           /* logical                                                      */
    int flags;
    void *stream;             /* cudaStream_t or NULL for the library's SW stream             */
    double scon, adjes;
    const double *bndscl;     /* (14) or NULL (absent optional)                               */
    const double *indsolvar;  /* (2) or NULL                                                  */
    const double *solcycfrac; /* scalar or NULL                                               */
    /* inputs */
    const double *coszen;                             /* (ncol)                               */
    const double *play, *plev, *tlay;                 /* (ncol,nlay) (ncol,nlay+1) (ncol,nlay)*/
    const double *h2ovmr, *o3vmr, *co2vmr, *ch4vmr, *o2vmr;
    const double *cld, *ciwp, *clwp, *rei, *rel;
    const double *zm, *alat;
    const double *tauaer, *ssaaer, *asmaer;           /* (ncol,nlay,14)                       */
    const double *asdir, *asdif, *aldir, *aldif;      /* (ncol)                               */
    /* outputs */
    int32_t *clearCounts;                             /* (ncol,4)                             */
    double *swuflx, *swdflx, *swuflxc, *swdflxc;      /* (ncol,nlay+1)                        */
    double *nirr, *nirf, *parr, *parf, *uvrr, *uvrf;  /* (ncol)                               */
    double *fswband;                                  /* (ncol,14)                            */
    double *cotdtp, *cotdhp, *cotdmp, *cotdlp;        /* (ncol)                               */
    double *cotntp, *cotnhp, *cotnmp, *cotnlp;        /* (ncol)                               */
    double *drband, *dfband;                          /* (ncol,14), touched iff do_drfband    */
    double *radval;           /* NULL: the default build.  Otherwise the SOLAR_RADVAL build of rrtmg_sw      */
                              /* (GEOSsolar_GridComp/CMakeLists.txt:18-20): (ncol,RRTMGX_NRADVAL) receives   */
                              /* its 120 extra dummies, column fastest, in the order of the dummy list       */
                              /* (rrtmg_sw_rad.F90:85-122), see RRTMGX_RADVAL_FAMILIES below                 */
} RrtmgxSwArgs;

/* The SOLAR_RADVAL diagnostics (rrtmg_sw_rad.F90:85-122, 306-345; formed in rrtmg_sw_spcvmc.F90:681-1105 from the
 * phase-split cloud optics of rrtmg_sw_cldprmc.F90:38-47): PAR-weighted sums over the McICA subcolumns of bands
 * 24-26 for the Tot|High|Mid|Low pressure super-layers, fifteen families of eight,
 *   <family>{d,n}{t,h,m,l}p  ->  radval(:, 8*family + 4*(d=0,n=1) + (t=0,h=1,m=2,l=3)),
 * families in this order: cds (delta-scaled in-cloud optical thickness); cotl, cdsl, coti, cdsi (liquid / ice,
 * original / delta-scaled thickness); ssal, sdsl, ssai, sdsi (single-scattering albedo, tau weighted); asml, adsl,
 * asmi, adsi (asymmetry, tau*ssa weighted); forl, fori (forward-scattering fraction, tau*ssa weighted).  A ratio
 * n/d is the diagnosed mean; cloud-free columns hold zeros. */
enum { RRTMGX_NRADVAL = 120 };
#define RRTMGX_RADVAL_FAMILIES "cds cotl cdsl coti cdsi ssal sdsl ssai sdsi asml adsl asmi adsi forl fori"

int rrtmgx_lw_run(const RrtmgxLwArgs *a);

/* The removed-gas diagnostic loop of the LW driver in one call (GEOS_IrradGridComp.F90:3405-3468 followed by
 * the main call :3471-3478): for n = 1..nvar rrtmg_lw is run with the gas gas[n] zeroed and its uflx, dflx and
 * duflx_dTs written to slab n of the (ncol,nlay+1,nvar) arrays below, then once with every gas, into the arrays
 * of `a`.  (Like the loop it replaces, the variants leave their uflxc/dflxc/duflxc_dTs/olrb/clearCounts in the
 * arrays of `a`, where the main call overwrites them.)  The inputs cross PCIe once and every run of a chunk
 * shares its McICA subcolumns and cloud optics; results are those of the separate calls, bit for bit. */
enum { RRTMGX_GAS_H2O = 1, RRTMGX_GAS_O3, RRTMGX_GAS_CO2, RRTMGX_GAS_CH4, RRTMGX_GAS_N2O, RRTMGX_GAS_CFC11,
       RRTMGX_GAS_CFC12, RRTMGX_GAS_HCFC22 };
typedef struct {
    int nvar;                                 /* number of removed-gas runs (0: same as rrtmgx_lw_run)        */
    const int32_t *gas;                       /* (nvar) RRTMGX_GAS_*, host                                     */
    double *uflx, *dflx, *duflx_dTs;          /* (ncol,nlay+1,nvar): UFLXRAT, DFLXRAT, DUFLX_DT_RAT            */
} RrtmgxLwVariants;
int rrtmgx_lw_run_variants(const RrtmgxLwArgs *a, const RrtmgxLwVariants *v);
int rrtmgx_sw_run(const RrtmgxSwArgs *a);

/* The two SORADCORE passes of one solar refresh in one call (GEOS_SolarGridComp.F90:3249-3287): rrtmg_sw without
 * aerosols (aerosol optical depth zero), its fluxes and band fluxes into `na` (FSWNA, FSWUNA, FSCNA, FSCUNA,
 * FSWBANDNA come from these), then the regular run into the arrays of `a`.  The inputs cross PCIe once and both
 * runs of a chunk share the McICA subcolumns, cloud optics and setcoef state; bits of the separate calls. */
typedef struct {
    double *swuflx, *swdflx, *swuflxc, *swdflxc;      /* (ncol,nlay+1)                                */
    double *fswband;                                  /* (ncol,14)                                    */
} RrtmgxSwNoAerosol;
int rrtmgx_sw_run_with_clean(const RrtmgxSwArgs *a, const RrtmgxSwNoAerosol *na);
/* status of the last RRTMGX_NO_SYNC run on that path (synchronises its stream) */
int rrtmgx_lw_status(void);
int rrtmgx_sw_status(void);

/* Optional taps on device intermediates for parity tests; any pointer may be NULL.  Host
 * pointers, filled (with a synchronisation) by the next rrtmgx_{lw,sw}_run. */
typedef struct {
    int32_t *jp, *jt, *jt1, *indfor, *indself, *indminor;   /* (ncol,nlay) 1-based           */
    int32_t *laytrop;                                        /* (ncol)                        */
    double *fac00, *fac01, *fac10, *fac11;                   /* (ncol,nlay)                   */
    uint8_t *cldymc;          /* [ilay][ig][icol] optical cloud mask (taucmc > 0)             */
    double *taucmc;           /* [ilay][ig][icol]; 0 where clear                              */
    double *pwvcm;            /* (ncol) LW                                                    */
    double *taug, *pfracs;    /* [ilay][ig][icol] LW gas optical depth incl. aerosol / Planck fraction; SW taug, taur */
    double *ssi;              /* [ig][icol] SW                                                */
} RrtmgxTaps;
void rrtmgx_set_taps(const RrtmgxTaps *lw_taps, const RrtmgxTaps *sw_taps);

/* GEOS_RadiationGridComp.F90:798-819 heating-rate epilogue (RADLW/RADSW):
 * hr(l) = (F(l-1) - F(l)) * grav / (cp * (plev(l-1) - plev(l)))  with F = up - down flux at
 * levels (ncol,nlay+1), level 0 at the surface, plev in hPa; result (ncol,nlay) in K/day.
 * grav and cp are the caller's MAPL_GRAV and MAPL_CP.  Device or host pointers per flags. */
int rrtmgx_heating_rate(int ncol, int nlay, const double *fnet_up_minus_down, const double *plev,
                        double *hr_K_per_day, double grav, double cp, int flags, void *stream);

/* ---- Run-phase glue fused on the device --------------------------------
 * The two GEOS drivers reshape the model state on the host before every RRTMG call (vertical flip,
 * Pa -> hPa, q -> vmr, content -> path, radius limits, TLEV interpolation, layer heights, negative
 * clean-up) and reshape the fluxes afterwards (unflip, sign convention, SFCEM, FSW = down - up,
 * clear counts -> cloud fractions, COT ratios).  These entry points take the NATIVE GEOS arrays -
 * columns flattened (IM*JM), (ncol,LM) with level 1 at the model top, PLE (ncol,LM+1), SI units -
 * do both reshapes in kernels around the device-resident RRTMG path, and hand back native outputs,
 * so only the native state crosses PCIe, once.
 *   rrtmgx_irrad_refresh   replaces GEOS_IrradGridComp.F90 LW_Driver :3237-3371, :3471-3478, :3486-3547
 *   rrtmgx_solar_refresh   replaces GEOS_SolarGridComp.F90 SORADCORE :6113-6223, :6331-6387, :6395-6447
 *   rrtmgx_*_prepare       only the first reshape: fills the caller-allocated arrays of an
 *                          RrtmgxLwArgs / RrtmgxSwArgs (and its cloudLM/cloudMH), for staged use and tests
 * Pointers are host or device pointers per `flags` as for rrtmgx_lw_run. */
typedef struct {
    int ncol, lm;
    int iceflg, liqflg;       /* RRTMG_ICEFLG / RRTMG_LIQFLG resources (GEOS defaults 3, 1)           */
    int doy;
    int lcldmh, lcldlm;       /* GEOS (top-down) super-layer interface levels                        */
    int flags;
    void *stream;
    double co2_fixed, o2, ccl4;                       /* CO2 where `co2` is NULL; uniform O2, CCl4   */
    double airmw, h2omw, o3mw, rgas, grav;            /* MAPL_AIRMW, _H2OMW, _O3MW, _RGAS, _GRAV      */
    const double *ple;                                /* (ncol,0:LM) Pa                              */
    const double *pl, *t, *q, *o3, *ch4, *n2o;        /* (ncol,LM): Pa, K, kg/kg, mmr, vmr, vmr      */
    const double *co2;                                /* (ncol,LM) vmr or NULL                       */
    const double *cfc11, *cfc12, *hcfc22, *fcld;      /* (ncol,LM)                                   */
    const double *qliq, *qice, *rliq, *rice;          /* CWC [kg/kg] and REFF [um], KLIQUID / KICE   */
    const double *ts, *t2m, *emis, *lats;             /* (ncol)                                      */
    const double *taua, *ssaa;                        /* (ncol,LM,16) extinction, scattering; or NULL */
    const int32_t *band_output;                       /* (16) logical (host)                         */
    /* outputs, GEOS convention: (ncol,0:LM) top-down, upward negative */
    double *flxu, *flxd, *flcu, *flcd, *dfdts, *dfdtsc;
    double *sfcem;                                    /* (ncol)                                      */
    double *cldtt, *cldhi, *cldmd, *cldlo;            /* (ncol), any may be NULL                     */
    double *olrb, *dolrb_dts;                         /* (16,ncol), bands with band_output; or NULL  */
} RrtmgxIrradArgs;

typedef struct {
    int ncol, lm;
    int iceflg, liqflg;
    int doy, isolvar;
    int lcldmh, lcldlm;
    int flags;
    void *stream;
    double sc, dist;                                  /* solar constant, Earth-Sun adjustment (ADJES) */
    double co2, o2;
    double airmw, h2omw, o3mw, rgas, grav, undef;     /* ..., MAPL_UNDEF                              */
    const double *solcycfrac;                         /* scalar or NULL                               */
    const double *ple;                                /* (ncol,LM+1) Pa                               */
    const double *pl, *t, *q, *o3, *ch4, *cl;         /* (ncol,LM)                                    */
    const double *qliq, *qice, *rliq, *rice;          /* QQ3(:,:,2), QQ3(:,:,1), RR3(:,:,2), RR3(:,:,1) */
    const double *ts, *zt, *lats;                     /* (ncol): TS, cos(zenith), latitude            */
    const double *albvr, *albvf, *albnr, *albnf;      /* (ncol)                                       */
    const double *taua, *ssaa, *asya;                 /* (ncol,LM,14) un-normalised, or NULL          */
    /* outputs */
    double *fsw, *fsc, *fswu, *fscu;                  /* (ncol,LM+1) top-down                         */
    double *nirr, *nirf, *parr, *parf, *uvrr, *uvrf;  /* (ncol), any may be NULL                      */
    double *fswband;                                  /* (ncol,14) or NULL                            */
    double *cldts, *cldhs, *cldms, *cldls;            /* (ncol), any may be NULL                      */
    double *cottp, *cothp, *cotmp, *cotlp;            /* (ncol), any may be NULL                      */
} RrtmgxSolarArgs;

/* Between refreshes GEOS updates the LW exports every model step from the fluxes held since the last
 * refresh, linearised in the surface temperature (GEOS_IrradGridComp.F90 Update, :3861 and the
 * USE_RRTMG branch :3929-3990 with FLX_INT = FLXD_INT + FLXU_INT, :3604-3606):
 *   DELT = TSINST - TS_INT;  FLX = FLX_INT + DFDTS*DELT;  FLXU = FLXU_INT + DFDTS*DELT;  FLXD = FLXD_INT;
 *   OLR = -(FLX_INT(0) + DFDTS(0)*DELT);  SFCEM = SFCEM_INT - DFDTS(LM)*DELT;  LWS = FLX_INT(LM) + SFCEM_INT;
 *   FLNS = FLX_INT(LM) + DFDTS(LM)*DELT;  and the clear-sky twins.  Inputs are the outputs of
 * rrtmgx_irrad_refresh (they can stay on the device); any output pointer may be NULL. */
typedef struct {
    int ncol, lm;
    int flags;
    void *stream;
    const double *flxu_int, *flxd_int, *flcu_int, *flcd_int, *dfdts, *dfdtsc;   /* (ncol,0:LM)          */
    const double *sfcem_int, *ts_int, *tsinst;                                    /* (ncol)               */
    double *flx, *flc, *flxu, *flcu, *flxd, *flcd;                                /* (ncol,0:LM)          */
    double *olr, *olc, *sfcem, *lws, *lcs, *flns, *flnsc;                         /* (ncol)               */
} RrtmgxIrradUpdateArgs;
int rrtmgx_irrad_update(const RrtmgxIrradUpdateArgs *u);

int rrtmgx_irrad_prepare(const RrtmgxIrradArgs *g, RrtmgxLwArgs *lw);
int rrtmgx_irrad_refresh(const RrtmgxIrradArgs *g);
int rrtmgx_solar_prepare(const RrtmgxSolarArgs *g, RrtmgxSwArgs *sw);
int rrtmgx_solar_refresh(const RrtmgxSolarArgs *g);

/* Test hook: the band kernels replace the compiler's IEEE fp64 division by its own fast-path
 * instruction sequence without the range test (csrc/common.cuh ddiv/drcp).  Evaluates both on the
 * device for n host pairs (a, b): q_* = a/b, r_* = 1/b, so a test can pin them bit for bit. */
int rrtmgx_debug_divide(size_t n, const double *a, const double *b, double *q_fast, double *q_ieee,
                        double *r_fast, double *r_ieee);

/* Test hook: the device KISS generator on its own (rng_kiss, GEOS_RadiationShared/cloud_subcol_gen.F90:546-607) and
 * the O(1) jump-ahead the McICA kernel uses instead of replaying the sequence.  All arrays are host arrays.
 *   seeds (4,nstream): seed1..seed4 of every stream.
 *   ndraw > 0: kiss / ran8 / ran4 (ndraw,nstream) receive the first ndraw integer draws `kiss` of each stream and
 *     `ran_num = kiss * 2.328306e-10 + 0.5` (:575) evaluated in real*8 (the promoted-real contract the library computes
 *     in) and in real*4 (the production kind, whose range the reference records: [8.9406967E-08, 0.9999999], :597-604).
 *   nsub > 0: for the jump table of (nsub subcolumns, nlay layers, inhomo) - entry 2i jumps i*stride draws, entry 2i+1
 *     i*stride + 2*nlay, stride = 2*nlay or 4*nlay - jumped / replayed (4, 2*nsub, nstream) receive the generator
 *     state after Kiss::jump and after replaying the same number of draws one by one.
 *   nvalue > 0: val8 / val4 (nvalue) = ran_num of the given integers `values` in real*8 and real*4. */
int rrtmgx_debug_kiss(int nstream, const int32_t *seeds, int ndraw, int32_t *kiss, double *ran8, float *ran4, int nsub,
                      int nlay, int inhomo, uint32_t *jumped, uint32_t *replayed, int nvalue, const int32_t *values,
                      double *val8, float *val4);

/* reduced (post-cmbgb) host copies of tables for tests: kind "lw"/"sw", band as in the
 * reference (1..16 / 16..29; 0 for band-independent), g-point fastest layout [lead][ng]. */
const double *rrtmgx_table(const char *kind, const char *name, int band, int *n);

#ifdef __cplusplus
}
#endif
#endif
