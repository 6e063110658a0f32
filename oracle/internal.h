/* oracle/internal.h -- shared internals of the CPU oracle (test infrastructure only). */
#ifndef RRTMG_ORACLE_INTERNAL_H
#define RRTMG_ORACLE_INTERNAL_H
#include <stddef.h>
#include "oracle.h"

/* ---- raw table blob (tools/extract_tables.py) ---- */
typedef struct {
    char name[48];
    int dtype, ndim, dims[6];
    long long offset, nbytes;
} BlobEntry;
int blob_load(const char *path);
void blob_free(void);
const double *blob_f64(const char *name, int *n);  /* NULL when absent */
const int *blob_i32(const char *name, int *n);

/* Fortran int(): truncate toward zero (real -> integer conversion) */
static inline int f_int(double x) { return (int)x; }
static inline double f_min(double a, double b) { return a < b ? a : b; }
static inline double f_max(double a, double b) { return a > b ? a : b; }

/* ---- McICA (SH/) ---- */
typedef struct {
    int inhm;               /* 0 homogeneous, 1 beta, 2 gamma */
    const double *xcw;      /* (1000,140) column-major, NULL when homogeneous */
    double aam1, aam2, aam30, aam4, ram1, ram2, ram30, ram4;
} McicaState;
extern McicaState g_mcica;
double zcw_lookup(double cdf, double sigma_qcw);

/* ---- reduced LW tables (LW/src/rrtmg_lw_init.F90) ---- */
typedef struct {
    int ng;        /* reduced g-points in band */
    int nspa, nspb;
    /* key-species tables flattened as absa(65*nspa, ng), absb(235*nspb, ng) */
    double *absa, *absb;
    double *selfref;  /* (10,ng) */
    double *forref;   /* (4,ng)  */
    double *fracrefa; /* (ng) or (ng,9) */
    double *fracrefb; /* (ng) or (ng,5) */
    /* minor species, (19,ng) or (n,19,ng) */
    double *ka_mn2, *kb_mn2, *ka_mn2o, *kb_mn2o, *ka_mo3, *kb_mo3, *ka_mco2, *kb_mco2,
        *ka_mco, *ka_mo2, *kb_mo2;
    double *ccl4, *cfc11adj, *cfc12, *cfc22adj; /* (ng) */
} LwBand;

typedef struct {
    LwBand b[NBNDLW];
    int ngc[NBNDLW], ngs[NBNDLW], ngb[NGPTLW];
    double rwgt[NBNDLW * 16];
    double tau_tbl[10001], exp_tbl[10001], tfn_tbl[10001];
    double bpade;
    const double *pref, *preflog, *tref, *chi_mls; /* chi_mls(7,59) */
    const double *totplnk, *totplk16, *totplnkderiv, *totplk16deriv;
    const double *absice0, *absice1, *absice2, *absice3, *absice4, *absliq1;
    double delwave[NBNDLW];
    double grav, avogad, oneminus, fluxfac;
} LwTables;
extern LwTables g_lw;
int lw_init(void);
void lw_free(void);

/* ---- reduced SW tables (SW/src/rrtmg_sw_init.F90) ---- */
typedef struct {
    int ng, nspa, nspb, nfor, nsrc; /* nfor: forref rows (3 or 4); nsrc: 1, 5 or 9 */
    double *absa, *absb, *selfref, *forref;
    double *sfluxref, *irradnce, *facbrght, *snsptdrk; /* (ng) or (ng,nsrc) */
    double rayl;            /* scalar Rayleigh coefficient (bands with scalar rayl) */
    double *raylv;          /* (ng) per-g Rayleigh (bands 23,25,26,27) */
    double *rayla, *raylb;  /* band 24: (ng,9), (ng) */
    double *abso3a, *abso3b, *absch4, *absco2, *absh2o;
} SwBand;

typedef struct {
    SwBand b[NBNDSW];
    int ngc[NBNDSW], ngs[NBNDSW], ngb[NGPTSW], icxa[NBNDSW];
    double rwgt[NBNDSW * 16];
    const double *pref, *preflog, *tref;
    const double *extliq1, *ssaliq1, *asyliq1, *extice2, *ssaice2, *asyice2, *extice3,
        *ssaice3, *asyice3, *fdlice3, *extice4, *ssaice4, *asyice4, *abari, *bbari, *cbari,
        *dbari, *ebari, *fbari;
    const double *mgavgcyc, *sbavgcyc;
    double grav, avogad, oneminus;
} SwTables;
extern SwTables g_sw;
int sw_init(void);
void sw_free(void);

#endif
