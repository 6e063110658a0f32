// RRTMG longwave on the device (sm_100a).
//
// Restates, as a different program, what these reference routines compute
// (LW/ = GEOSirrad_GridComp/RRTMG/rrtmg_lw/gcm_model/):
//   LW/src/rrtmg_lw_rad.F90      rrtmg_lw_part :348-610 (orchestration; no transposes here)
//   LW/src/rrtmg_lw_setcoef.F90  setcoef :52-584                -> lw_setcoef_kernel
//   LW/src/rrtmg_lw_cldprmc.F90  cldprmc :24-385                -> LwOptics (inside McICA)
//   LW/src/rrtmg_lw_taumol.F90   taugb1..16 :191-3126, addAerosols :3130-3146
//   LW/src/rrtmg_lw_rtrnmc.F90   rtrnmc :27-390                 -> lw_band_kernel (fused)
//
// Kernel structure.  One thread owns one column.  lw_setcoef_kernel writes the per-(layer,
// column) interpolation state (packed indices + 23 factors) and the band Planck functions
// once; the McICA kernel writes the optical cloud mask and cloud optical depth; then one
// lw_band_kernel instantiation per (band, g-point sub-range) fuses gas optics and the
// radiance sweeps: optical depths and Planck fractions live only in registers, the downward
// sweep stores the two 14-bit transmittance-table indices per cell (4 bytes) and the upward
// sweep re-reads them.  Every unit writes its partial flux profiles; lw_reduce_kernel adds
// the units in a fixed order (deterministic, no floating-point atomics).
//
// Tables are g-point fastest ([lead][ng]): a thread reads the consecutive g-points of its
// sub-range from one table row.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "engine.h"
#include "mcica.cuh"

namespace rrtmgx {

// ---------------------------------------------------------------------------------------------
// device-resident tables
// ---------------------------------------------------------------------------------------------
struct LwBandTab {
    const double *absa, *absb, *selfref, *forref, *fracrefa, *fracrefb;
    const double *ka_mn2, *kb_mn2, *ka_mn2o, *kb_mn2o, *ka_mo3, *kb_mo3, *ka_mco2, *kb_mco2, *ka_mco,
        *ka_mo2, *kb_mo2;
    const double *ccl4, *cfc11adj, *cfc12, *cfc22adj;
};

struct LwDev {
    LwBandTab b[16];
    const double *preflog, *tref, *chi_mls, *rat;   // rat: [6][59]
    const double *totplnk, *totplnkderiv;           // (181,16)
    const double *exptfn;                           // [10001][2] {exp_tbl, tfn_tbl}
    const double *tau_tbl;                          // [10001]
    const double *absice0, *absice1, *absice2, *absice3, *absice4, *absliq1;
    double delwave[16];
    int ngb[140];     // band (1-based) of each g-point
    int ngs[16];      // cumulative g-points
    double bpade, oneminus, fluxfac, grav, avogad;
    // reference-atmosphere ratios used as refrat_* constants by the binary-species bands
    double chi_1_9_over_2_9, chi_1_3_over_2_3, chi_1_13_over_2_13;   // band 3
    double chi_1_11_over_2_11, chi_3_13_over_2_13;                   // band 4
    double chi_1_5_over_2_5, chi_1_7_over_2_7, chi_3_43_over_2_43;   // band 5
    double chi_1_3_over_3_3;                                         // band 7
    double chi_1_9_over_6_9, chi_1_3_over_6_3;                       // band 9
    double chi_1_10_over_2_10;                                       // band 12
    double chi_1_5_over_4_5, chi_1_1_over_4_1, chi_1_3_over_4_3;     // band 13
    double chi_4_1_over_2_1;                                         // band 15
    double chi_1_6_over_6_6;                                         // band 16
};

__constant__ LwDev c_lw;

static int g_lw_ngs[16], g_lw_ngb[140];   // host copies for the debug taps

int lw_upload_tables(const HostTables &ht, const double *d_arena) {
    LwDev h;
    std::memset(&h, 0, sizeof h);
    auto dev = [&](const std::string &name) -> const double * {
        TableRef r = ht.find(name);
        return r.ok() ? d_arena + r.off : nullptr;
    };
    for (int ib = 0; ib < 16; ++ib) {
        char pre[16];
        std::snprintf(pre, sizeof pre, "lw.%02d.", ib + 1);
        const std::string p(pre);
        LwBandTab &B = h.b[ib];
        B.absa = dev(p + "absa"); B.absb = dev(p + "absb");
        B.selfref = dev(p + "selfref"); B.forref = dev(p + "forref");
        B.fracrefa = dev(p + "fracrefa"); B.fracrefb = dev(p + "fracrefb");
        B.ka_mn2 = dev(p + "ka_mn2"); B.kb_mn2 = dev(p + "kb_mn2");
        B.ka_mn2o = dev(p + "ka_mn2o"); B.kb_mn2o = dev(p + "kb_mn2o");
        B.ka_mo3 = dev(p + "ka_mo3"); B.kb_mo3 = dev(p + "kb_mo3");
        B.ka_mco2 = dev(p + "ka_mco2"); B.kb_mco2 = dev(p + "kb_mco2");
        B.ka_mco = dev(p + "ka_mco"); B.ka_mo2 = dev(p + "ka_mo2"); B.kb_mo2 = dev(p + "kb_mo2");
        B.ccl4 = dev(p + "ccl4"); B.cfc11adj = dev(p + "cfc11adj");
        B.cfc12 = dev(p + "cfc12"); B.cfc22adj = dev(p + "cfc22adj");
    }
    h.preflog = dev("lw.ref.preflog"); h.tref = dev("lw.ref.tref");
    h.chi_mls = dev("lw.ref.chi_mls"); h.rat = dev("lw.ref.rat");
    h.totplnk = dev("lw.wvn.totplnk"); h.totplnkderiv = dev("lw.wvn.totplnkderiv");
    h.exptfn = dev("lw.exptfn"); h.tau_tbl = dev("lw.tau_tbl");
    h.absice0 = dev("lw.cld.absice0"); h.absice1 = dev("lw.cld.absice1");
    h.absice2 = dev("lw.cld.absice2"); h.absice3 = dev("lw.cld.absice3");
    h.absice4 = dev("lw.cld.absice4"); h.absliq1 = dev("lw.cld.absliq1");
    if (!h.preflog || !h.tref || !h.chi_mls || !h.rat || !h.totplnk || !h.totplnkderiv || !h.exptfn ||
        !h.tau_tbl || !h.absice3 || !h.absliq1 || !h.b[0].absa || !h.b[15].absb)
        return RRTMGX_EBLOB;
    for (int i = 0; i < 16; ++i) { h.delwave[i] = ht.lw_delwave[i]; h.ngs[i] = ht.lw_ngs[i]; }
    for (int i = 0; i < 140; ++i) h.ngb[i] = ht.lw_ngb[i];
    for (int i = 0; i < 16; ++i) g_lw_ngs[i] = ht.lw_ngs[i];
    for (int i = 0; i < 140; ++i) g_lw_ngb[i] = ht.lw_ngb[i];
    // lwdatinit (LW/src/rrtmg_lw_init.F90:214,222), rrlw_con.F90:37-38, rrlw_tbl.F90:32
    h.bpade = 1.0 / 0.278;
    h.oneminus = 1. - 1.e-6;
    h.fluxfac = 3.14159265358979323846 * 2.e4;
    h.grav = 9.8066;
    h.avogad = 6.02214199e+23;
    const double *chi = ht.ptr(ht.find("lw.ref.chi_mls"));
    auto CHI = [&](int m, int j) { return chi[(m - 1) + 7 * (j - 1)]; };
    h.chi_1_9_over_2_9 = CHI(1, 9) / CHI(2, 9);
    h.chi_1_3_over_2_3 = CHI(1, 3) / CHI(2, 3);
    h.chi_1_13_over_2_13 = CHI(1, 13) / CHI(2, 13);
    h.chi_1_11_over_2_11 = CHI(1, 11) / CHI(2, 11);
    h.chi_3_13_over_2_13 = CHI(3, 13) / CHI(2, 13);
    h.chi_1_5_over_2_5 = CHI(1, 5) / CHI(2, 5);
    h.chi_1_7_over_2_7 = CHI(1, 7) / CHI(2, 7);
    h.chi_3_43_over_2_43 = CHI(3, 43) / CHI(2, 43);
    h.chi_1_3_over_3_3 = CHI(1, 3) / CHI(3, 3);
    h.chi_1_9_over_6_9 = CHI(1, 9) / CHI(6, 9);
    h.chi_1_3_over_6_3 = CHI(1, 3) / CHI(6, 3);
    h.chi_1_10_over_2_10 = CHI(1, 10) / CHI(2, 10);
    h.chi_1_5_over_4_5 = CHI(1, 5) / CHI(4, 5);
    h.chi_1_1_over_4_1 = CHI(1, 1) / CHI(4, 1);
    h.chi_1_3_over_4_3 = CHI(1, 3) / CHI(4, 3);
    h.chi_4_1_over_2_1 = CHI(4, 1) / CHI(2, 1);
    h.chi_1_6_over_6_6 = CHI(1, 6) / CHI(6, 6);
    if (cudaMemcpyToSymbol(c_lw, &h, sizeof h) != cudaSuccess) return RRTMGX_ECUDA;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// per-(layer,column) interpolation state written by setcoef, [lay][c] with c fastest
// ---------------------------------------------------------------------------------------------
enum LwF {
    F_FAC00, F_FAC01, F_FAC10, F_FAC11, F_COLH2O, F_COLCO2, F_COLO3, F_COLN2O, F_COLCH4, F_COLO2,
    F_COLBRD, F_COLCFC11, F_COLCFC12, F_COLCFC22, F_COLCCL4, F_COLDRY, F_FORFAC, F_FORFRAC,
    F_SELFFAC, F_SELFFRAC, F_SCALEMINOR, F_SCALEMINORN2, F_MINORFRAC,
    // minor-gas column amounts adjusted for abundances above the reference (taumol :460-467 etc.),
    // formed once per (layer, column) here instead of once per g-point thread
    F_ADJN2O, F_ADJCO2_6, F_ADJCO2_7, F_ADJCO2_8, F_ADJCO2_13, F_COUNT
};

struct LwWork {
    int nc, nlay;
    const int *trap;          // position of the first refused input of the call (>= 2^30: none), see RRTMGX_TRAPPED
    int *idx;                 // [tile][nlay][32] packed jp|jt|jt1|indfor|indself|indminor
    double *fbase;            // [tile][nlay][F_COUNT][32]: the setcoef state of a 32-column tile is contiguous
    size_t n2;                // nlay*nc
    size_t n2p;               // nlay * (nc padded to 32)
    int ncp;                  // nc padded to 32
    // plane k of the factors, to be indexed with tf(lay, c); packed indices are indexed with ti(lay, c)
    __host__ __device__ __forceinline__ double *f(int k) const { return fbase + (size_t)k * 32; }
    __host__ __device__ __forceinline__ size_t ti(int lay, int c) const { return ((size_t)(c >> 5) * nlay + lay) * 32 + (c & 31); }
    __host__ __device__ __forceinline__ size_t tf(int lay, int c) const {
        return ((size_t)(c >> 5) * nlay + lay) * (F_COUNT * 32) + (c & 31);
    }
    double *planklay;         // [16][nlay][nc]
    double *planklev;         // [16][nlay+1][nc]
    double *plankbnd, *dplankbnd;   // [16][nc]
    double *pwvcm;            // [nc]
    int *laytrop;             // [nc]
    uint32_t *seeds;          // [4][nc]
    double *alpha, *rcorr;    // [nlay][nc]
    long long *thr;           // [tile][nlay][3][32] integer thresholds of the McICA comparisons (alpha, rcorr, cld)
    double *abscoice, *abscoliq;   // [16][nlay][nc] cloud absorption coefficients per band (cloudy layers)
    unsigned char *cldtrap;   // [nlay][nc] bit0: ice radius out of range, bit1: liquid radius out of range
    int *perm;                // [nc] cloudy columns first (build_cloud_partition)
    unsigned char *pflags;    // [nc]
    int *ktop;                // [nc] last layer with cldf > 0 per chunk-local column (caller's order)
    int32_t *clear_save;      // [4][nc] clear counts of the chunk, kept for RRTMGX_REUSE_CLOUDS
    char *ptmp; size_t ptmp_bytes;
    uint32_t *mask;           // [band][nw][nc][ng] optical cloud mask (lw_cell with nw for nlay)
    uint32_t *cloudy_any;     // [nw][nc]
    double *taucmc;           // per-cell layout (lw_cell), valid where the mask bit is set
    uint32_t *it;             // per-cell layout (lw_cell): itgas | ittot << 16 (0xffff: clear cell)
    double *pfs;              // per-cell layout (lw_cell): Planck fraction of the cell (down sweep -> up sweep)
    double *part;             // [16][LP_COUNT][nlay+1][nc]
};

__device__ __forceinline__ int pack_idx(int jp, int jt, int jt1, int indfor, int indself, int indminor) {
    return jp | (jt << 6) | (jt1 << 9) | (indfor << 12) | (indself << 14) | (indminor << 18);
}

// adjusted minor column amount when the gas exceeds its reference abundance (e.g. taumol :460-467)
__device__ __forceinline__ double adjcol(double col, double coldry, double chi, double thresh, double base,
                                         double expo) {
    const double chi_x = col / coldry;
    const double rat = 1.e20 * chi_x / chi;
    if (rat > thresh) {
        const double adjfac = base + pow(rat - base, expo);
        return adjfac * chi * coldry * 1.e-20;
    }
    return col;
}
__device__ __forceinline__ double chi_mls(int m, int j) { return c_lw.chi_mls[(m - 1) + 7 * (j - 1)]; }

// LW/src/rrtmg_lw_setcoef.F90:52-584.  One thread per column, layers bottom-up.
__global__ void __launch_bounds__(128)
lw_setcoef_kernel(int ld, int col0, const int *__restrict__ perm, LwWork W, int dudTs,
                  const double *__restrict__ pavel, const double *__restrict__ tavel,
                  const double *__restrict__ pz, const double *__restrict__ tz,
                  const double *__restrict__ tbound, const double *__restrict__ semiss,
                  const double *__restrict__ h2ovmr, const double *__restrict__ o3vmr,
                  const double *__restrict__ co2vmr, const double *__restrict__ ch4vmr,
                  const double *__restrict__ n2ovmr, const double *__restrict__ o2vmr,
                  const double *__restrict__ cfc11vmr, const double *__restrict__ cfc12vmr,
                  const double *__restrict__ cfc22vmr, const double *__restrict__ ccl4vmr, int *err) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int nc = W.nc, nlay = W.nlay;
    if (c >= nc || RRTMGX_TRAPPED(W.trap)) return;
    const size_t col = gcol(col0, perm, c);
    const double amd = 28.9660, amw = 18.0160;
    const double stpfac = 296. / 1013.;
    const double grav = c_lw.grav, avogad = c_lw.avogad;

    // column sums for precipitable water (:224-270); coldry is kept for the second pass
    double amttl = 0., wvttl = 0.;
    for (int lay = 0; lay < nlay; ++lay) {
        const size_t i = (size_t)lay * ld + col;
        const double h2o = h2ovmr[i];
        const double amm = (1. - h2o) * amd + h2o * amw;
        const double coldry = (pz[i] - pz[i + ld]) * 1.e3 * avogad / (1.e2 * grav * amm * (1. + h2o));
        W.f(F_COLDRY)[W.tf(lay, c)] = coldry;
        const double btemp = h2o * coldry;
        amttl = amttl + coldry + btemp;
        wvttl = wvttl + btemp;
    }
    const double wvsh = (amw * wvttl) / (amd * amttl);
    W.pwvcm[c] = wvsh * (1.e3 * pz[col]) / (1.e2 * grav);

    // surface / lowest-level Planck functions (:278-330)
    {
        const double tb = tbound[col];
        const int indbound = clampi(f_int(tb - 159.), 1, 180);
        const double tbndfrac = tb - 159. - (double)indbound;
        const double t0 = tz[col];
        const int indlev0 = clampi(f_int(t0 - 159.), 1, 180);
        const double t0frac = t0 - 159. - (double)indlev0;
#pragma unroll 4
        for (int ib = 0; ib < 16; ++ib) {
            const double *tp = c_lw.totplnk + 181 * ib;
            const double em = semiss[(size_t)ib * ld + col];
            double dbdtlev = tp[indbound] - tp[indbound - 1];
            W.plankbnd[(size_t)ib * nc + c] = em * (tp[indbound - 1] + tbndfrac * dbdtlev);
            dbdtlev = tp[indlev0] - tp[indlev0 - 1];
            W.planklev[((size_t)ib * (nlay + 1)) * nc + c] = tp[indlev0 - 1] + t0frac * dbdtlev;
            if (dudTs) {
                const double *tpd = c_lw.totplnkderiv + 181 * ib;
                dbdtlev = tpd[indbound] - tpd[indbound - 1];
                W.dplankbnd[(size_t)ib * nc + c] = em * (tpd[indbound - 1] + tbndfrac * dbdtlev);
            }
        }
    }

    int laytrop = 0;
    bool upper_found = false;
    for (int lay = 0; lay < nlay; ++lay) {
        const size_t i = (size_t)lay * ld + col;
        const size_t j = W.tf(lay, c);
        const double coldry = W.f(F_COLDRY)[j];
        const double h2o = h2ovmr[i], co2 = co2vmr[i], o3 = o3vmr[i], n2o = n2ovmr[i], ch4 = ch4vmr[i],
                     o2 = o2vmr[i];
        const double p = pavel[i], t = tavel[i];
        const double summol = co2 + o3 + n2o + ch4 + o2;
        const double wbroad = coldry * (1. - summol);
        const double wv = h2o * coldry;

        // Planck functions at the layer and at its upper level (:340-394)
        {
            const int indlay = clampi(f_int(t - 159.), 1, 180);
            const double tlayfrac = t - 159. - (double)indlay;
            const double tl = tz[i + ld];
            const int indlev = clampi(f_int(tl - 159.), 1, 180);
            const double tlevfrac = tl - 159. - (double)indlev;
#pragma unroll 4
            for (int ib = 0; ib < 16; ++ib) {
                const double *tp = c_lw.totplnk + 181 * ib;
                const double dbdtlev = tp[indlev] - tp[indlev - 1];
                W.planklev[((size_t)ib * (nlay + 1) + lay + 1) * nc + c] = tp[indlev - 1] + tlevfrac * dbdtlev;
                const double dbdtlay = tp[indlay] - tp[indlay - 1];
                W.planklay[((size_t)ib * nlay + lay) * nc + c] = tp[indlay - 1] + tlayfrac * dbdtlay;
            }
        }

        const double plog = log(p);
        const int jp = clampi(f_int(36. - 5. * (plog + 0.04)), 1, 58);
        const double fp = 5. * (c_lw.preflog[jp - 1] - plog);
        const double tref0 = c_lw.tref[jp - 1], tref1 = c_lw.tref[jp];
        const int jt = clampi(f_int(3. + (t - tref0) / 15.), 1, 4);
        const double ft = ((t - tref0) / 15.) - (double)(jt - 3);
        const int jt1 = clampi(f_int(3. + (t - tref1) / 15.), 1, 4);
        const double ft1 = ((t - tref1) / 15.) - (double)(jt1 - 3);

        const double water = wv / coldry;
        const double scalefac = p * stpfac / t;
        double forfac, forfrac, selffac, selffrac = 0.;
        int indfor, indself = 1;
        if (plog > 4.56) {
            if (upper_found) raise(err, RRTMGX_EPRESSURE);
            laytrop += 1;
            forfac = scalefac / (1. + water);
            double factor = (332. - t) / 36.;
            indfor = clampi(f_int(factor), 1, 2);
            forfrac = factor - (double)indfor;
            selffac = water * forfac;
            factor = (t - 188.) / 7.2;
            indself = clampi(f_int(factor) - 7, 1, 9);
            selffrac = factor - (double)(indself + 7);
        } else {
            upper_found = true;
            forfac = scalefac / (1. + water);
            const double factor = (t - 188.) / 36.;
            indfor = 3;
            forfrac = factor - 1.;
            selffac = 0.;
        }
        const double scaleminor = p / t;
        const double scaleminorn2 = (p / t) * (wbroad / (coldry + wv));
        const double factor = (t - 180.8) / 7.2;
        const int indminor = clampi(f_int(factor), 1, 18);
        const double minorfrac = factor - (double)indminor;

        const double colh2o = 1.e-20 * h2o * coldry;
        double colco2 = 1.e-20 * co2 * coldry;
        double colo3 = 1.e-20 * o3 * coldry;
        double coln2o = 1.e-20 * n2o * coldry;
        double colch4 = 1.e-20 * ch4 * coldry;
        if (colco2 == 0.) colco2 = 1.e-32 * coldry;
        if (colo3 == 0.) colo3 = 1.e-32 * coldry;
        if (coln2o == 0.) coln2o = 1.e-32 * coldry;
        if (colch4 == 0.) colch4 = 1.e-32 * coldry;

        const double compfp = 1. - fp;
        W.idx[W.ti(lay, c)] = pack_idx(jp, jt, jt1, indfor, indself, indminor);
        W.f(F_FAC10)[j] = compfp * ft;
        W.f(F_FAC00)[j] = compfp * (1. - ft);
        W.f(F_FAC11)[j] = fp * ft1;
        W.f(F_FAC01)[j] = fp * (1. - ft1);
        W.f(F_COLH2O)[j] = colh2o;
        W.f(F_COLCO2)[j] = colco2;
        W.f(F_COLO3)[j] = colo3;
        W.f(F_COLN2O)[j] = coln2o;
        W.f(F_COLCH4)[j] = colch4;
        W.f(F_COLO2)[j] = 1.e-20 * o2 * coldry;
        W.f(F_COLBRD)[j] = 1.e-20 * wbroad;
        W.f(F_COLCFC11)[j] = 1.e-20 * cfc11vmr[i] * coldry;
        W.f(F_COLCFC12)[j] = 1.e-20 * cfc12vmr[i] * coldry;
        W.f(F_COLCFC22)[j] = 1.e-20 * cfc22vmr[i] * coldry;
        W.f(F_COLCCL4)[j] = 1.e-20 * ccl4vmr[i] * coldry;
        W.f(F_FORFAC)[j] = colh2o * forfac;
        W.f(F_FORFRAC)[j] = forfrac;
        W.f(F_SELFFAC)[j] = colh2o * selffac;
        W.f(F_SELFFRAC)[j] = selffrac;
        W.f(F_SCALEMINOR)[j] = scaleminor;
        W.f(F_SCALEMINORN2)[j] = scaleminorn2;
        W.f(F_MINORFRAC)[j] = minorfrac;
        {
            const bool lower = plog > 4.56;
            const double chi_n2o = chi_mls(4, jp + 1), chi_co2 = chi_mls(2, jp + 1);
            W.f(F_ADJN2O)[j] = adjcol(coln2o, coldry, chi_n2o, 1.5, 0.5, 0.65);          // bands 3, 9
            W.f(F_ADJCO2_6)[j] = adjcol(colco2, coldry, chi_co2, 3.0, 2.0, 0.77);        // band 6
            W.f(F_ADJCO2_7)[j] = adjcol(colco2, coldry, chi_co2, 3.0, lower ? 3.0 : 2.0, 0.79);   // band 7
            W.f(F_ADJCO2_8)[j] = adjcol(colco2, coldry, chi_co2, 3.0, 2.0, 0.65);        // band 8
            W.f(F_ADJCO2_13)[j] = adjcol(colco2, coldry, 3.55e-4, 3.0, 2.0, 0.68);       // band 13
        }
    }
    W.laytrop[c] = laytrop;
}

// ---------------------------------------------------------------------------------------------
// cloud optics inside the McICA sweep: LW/src/rrtmg_lw_cldprmc.F90:24-385
// ---------------------------------------------------------------------------------------------
// Absorption coefficients per (band, layer, column), formed once instead of once per subcolumn:
// ice by iceflag 0-4 (:227-268), liquid by liqflag 1 (:318-360).  Radii outside the table range
// are flagged; the trap fires only if a McICA-cloudy cell uses the layer, as in the reference.
__global__ void lw_cldcoef_kernel(int ld, int col0, const int *__restrict__ perm, int nc, int nlay, int iceflag,
                                  const double *__restrict__ cldf, const double *__restrict__ reice,
                                  const double *__restrict__ reliq, double *__restrict__ abscoice,
                                  double *__restrict__ abscoliq, unsigned char *__restrict__ cldtrap) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int lay = blockIdx.y;
    if (c >= nc) return;
    const size_t i2 = (size_t)lay * ld + gcol(col0, perm, c);
    const size_t j = (size_t)lay * nc + c, n2 = (size_t)nlay * nc;
    if (!(cldf[i2] > 0.)) return;   // no subcolumn of this layer can be cloudy
    // index clamp / extrapolation traps shared by iceflag 2,3,4 and liqflag 1
    auto lookup_index = [](double factor, int hi, int &index) {
        int idx = f_int(factor);
        if (idx >= hi) {
            if (idx == hi) idx = hi - 1; else return false;
        } else if (idx <= 0) {
            if (idx == 0) idx = 1; else return false;
        }
        index = idx;
        return true;
    };
    unsigned char trap = 0;
    const double re = reice[i2];
    const double *itab = nullptr;
    int ilead = 0, iindex = 1;
    double ifint = 0.;
    if (iceflag >= 2) {
        double factor;
        if (iceflag == 2) { factor = (re - 2.) / 3.; ilead = 43; itab = c_lw.absice2; }
        else if (iceflag == 3) { factor = (re - 2.) / 3.; ilead = 46; itab = c_lw.absice3; }
        else { factor = re; ilead = 200; itab = c_lw.absice4; }
        if (!lookup_index(factor, ilead, iindex)) { trap |= 1; iindex = 1; }
        ifint = factor - (double)iindex;
    }
    const double lfactor = reliq[i2] - 1.5;
    int lindex = 1;
    if (!lookup_index(lfactor, 58, lindex)) { trap |= 2; lindex = 1; }
    const double lfint = lfactor - (double)lindex;
    cldtrap[j] = trap;
    for (int ib = 1; ib <= 16; ++ib) {
        double ai;
        if (iceflag == 0) {
            ai = c_lw.absice0[0] + c_lw.absice0[1] / re;
        } else if (iceflag == 1) {
            const int k = ib <= 2 ? ib : (ib <= 5 ? 3 : (ib <= 8 ? 4 : 5));   // rrlw_cld.F90 ice1b map
            ai = c_lw.absice1[2 * (k - 1)] + c_lw.absice1[1 + 2 * (k - 1)] / re;
        } else {
            const double *cb = itab + (size_t)ilead * (ib - 1);
            ai = cb[iindex - 1] + ifint * (cb[iindex] - (cb[iindex - 1]));
        }
        const double *cl = c_lw.absliq1 + (size_t)58 * (ib - 1);
        abscoice[(size_t)(ib - 1) * n2 + j] = ai;
        abscoliq[(size_t)(ib - 1) * n2 + j] = cl[lindex - 1] + lfint * (cl[lindex] - (cl[lindex - 1]));
    }
}

// Per-cell scratch of the LW path is g-point fastest inside a band and tiled by 32 columns:
// [band][tile][row][32 columns][ng_band], i.e. the g-points of one (band, layer, column) are adjacent, the
// columns of a tile follow, then the layers of the tile: what a block streams is one contiguous region.  The
// McICA kernel (lanes = subcolumns of a column) and the band kernels (lanes = g-point groups of a few columns)
// read and write contiguous runs.  `first` = first g-point of the band, `ng` its count, rows = layers (or
// 32-layer mask words) per column, ncp = columns padded to 32.
__host__ __device__ __forceinline__ size_t lw_cell(int first, int ng, int rows, int ncp, int row, int c, int gi) {
    return (size_t)first * rows * ncp + ((((size_t)(c >> 5) * rows + row) * 32) + (c & 31)) * ng + gi;
}

struct LwOptics {
    int nc, nlay, ncp;
    const double *abscoice, *abscoliq;   // [16][nlay][nc]
    const unsigned char *cldtrap;        // [nlay][nc]
    double *taucmc;                      // lw_cell layout
    struct State {};
    __device__ __forceinline__ void finish(int, int, State &) const {}
    __device__ __forceinline__ size_t cell_index(int rows, int row, int ig, int c) const {
        const int ib = c_lw.ngb[ig] - 1;
        const int first = ib ? c_lw.ngs[ib - 1] : 0;
        return lw_cell(first, c_lw.ngs[ib] - first, rows, ncp, row, c, ig - first);
    }
    __device__ __forceinline__ size_t mask_index(int w, int nw, int ig, int c) const { return cell_index(nw, w, ig, c); }

    // Called for every McICA-cloudy cell: taucmc = ciwp*abscoice + clwp*abscoliq (:362-383).  The
    // reference derives both radius indices (and traps) for every layer holding such a cell,
    // whichever phase holds water (:193-205).
    __device__ __forceinline__ bool cell(int lay, int ig, int c, double ciw, double clw, int *err, State &) const {
        const size_t j = (size_t)lay * nc + c, n2 = (size_t)nlay * nc;
        const int ib = c_lw.ngb[ig];   // 1-based band
        const unsigned char trap = cldtrap[j];
        if (trap) raise(err, (trap & 1) ? RRTMGX_ERADIUS_ICE : RRTMGX_ERADIUS_LIQ);
        double tau = 0.;
        if (ciw > 0.) tau = ciw * abscoice[(size_t)(ib - 1) * n2 + j];
        if (clw > 0.) tau = tau + clw * abscoliq[(size_t)(ib - 1) * n2 + j];
        const bool optical = tau > 0.;
        if (optical) __stcs(&taucmc[cell_index(nlay, lay, ig, c)], tau);
        return optical;
    }
};

// ---------------------------------------------------------------------------------------------
// gas optics: taugb1..16 restated per (band, g sub-range)
// ---------------------------------------------------------------------------------------------
#define FORG _Pragma("unroll") for (int ig = 0; ig < GN; ++ig)

struct Spec { double speccomb, specparm, fs; int js; };
// binary-species parameter (e.g. LW/src/rrtmg_lw_taumol.F90:436-441)
__device__ __forceinline__ Spec spec(double cola, double rat, double colb, double mult) {
    Spec r;
    r.speccomb = cola + rat * colb;
    r.specparm = ddiv(cola, r.speccomb);
    if (r.specparm >= c_lw.oneminus) r.specparm = c_lw.oneminus;
    const double specmult = mult * r.specparm;
    const int k = f_int(specmult);
    r.js = 1 + k;
    r.fs = specmult - (double)k;   // mod(specmult, 1.) for specmult >= 0
    return r;
}

// everything a band needs about one (layer, column)
struct Lay {
    int jp, jt, jt1, indfor, indself, indminor;
    const double *fj;  // factor base + lay*nc + c
    __device__ __forceinline__ double f(int k) const { return fj[k * 32]; }   // planes of a tile row are 32 apart
};

// GN consecutive g-points of one table row.  The tables are g-point fastest and 16-byte aligned in
// the arena, ng and the thread's first g-point are even when GN is, so a row slice is read with
// 16-byte loads: half the load instructions - and half the L1 wavefronts, which is what these
// kernels are bound by (profiles/r2_d_*) - of one 8-byte gather per g-point.
template <int GN> struct GRow {
    double v[GN];
    __device__ __forceinline__ double operator[](int i) const { return v[i]; }
};
template <int GN>
__device__ __forceinline__ GRow<GN> ldrow(const double *__restrict__ p) {
    GRow<GN> r;
    if constexpr (GN % 2 == 0) {
#pragma unroll
        for (int i = 0; i < GN; i += 2) {
            const double2 t = __ldg(reinterpret_cast<const double2 *>(p + i));
            r.v[i] = t.x; r.v[i + 1] = t.y;
        }
    } else {
#pragma unroll
        for (int i = 0; i < GN; ++i) r.v[i] = __ldg(p + i);
    }
    return r;
}

// lower-atmosphere key-species sum for a binary band at one reference pressure
// (e.g. taugb3 :482-511, :554-576): three-point stencils near specparm 0 and 1, else bilinear
template <int GN>
__device__ __forceinline__ void stencil_lower(const double *__restrict__ row /* absa row ind (1-based ind -> row ind-1), at G0 */,
                                              int ng, double specparm, double fs, double fa, double fb,
                                              double speccomb, double (&out)[GN]) {
    if (specparm < 0.125) {
        const double p = fs - 1.;
        const double p2 = p * p, p4 = p2 * p2;
        const double fk0 = p4, fk1 = 1. - p - 2.0 * p4, fk2 = p + p4;
        const double w0 = fk0 * fa, w1 = fk1 * fa, w2 = fk2 * fa, w3 = fk0 * fb, w4 = fk1 * fb, w5 = fk2 * fb;
        const GRow<GN> r0 = ldrow<GN>(row), r1 = ldrow<GN>(row + ng), r2 = ldrow<GN>(row + 2 * ng),
                       r9 = ldrow<GN>(row + 9 * ng), r10 = ldrow<GN>(row + 10 * ng), r11 = ldrow<GN>(row + 11 * ng);
        FORG out[ig] = speccomb * (w0 * r0[ig] + w1 * r1[ig] + w2 * r2[ig] + w3 * r9[ig] + w4 * r10[ig] + w5 * r11[ig]);
    } else if (specparm > 0.875) {
        const double p = -fs;
        const double p2 = p * p, p4 = p2 * p2;
        const double fk0 = p4, fk1 = 1. - p - 2.0 * p4, fk2 = p + p4;
        const double w0 = fk2 * fa, w1 = fk1 * fa, w2 = fk0 * fa, w3 = fk2 * fb, w4 = fk1 * fb, w5 = fk0 * fb;
        const GRow<GN> rm = ldrow<GN>(row - ng), r0 = ldrow<GN>(row), r1 = ldrow<GN>(row + ng),
                       r8 = ldrow<GN>(row + 8 * ng), r9 = ldrow<GN>(row + 9 * ng), r10 = ldrow<GN>(row + 10 * ng);
        FORG out[ig] = speccomb * (w0 * rm[ig] + w1 * r0[ig] + w2 * r1[ig] + w3 * r8[ig] + w4 * r9[ig] + w5 * r10[ig]);
    } else {
        const double w0 = (1. - fs) * fa, w1 = fs * fa, w2 = (1. - fs) * fb, w3 = fs * fb;
        const GRow<GN> r0 = ldrow<GN>(row), r1 = ldrow<GN>(row + ng), r9 = ldrow<GN>(row + 9 * ng),
                       r10 = ldrow<GN>(row + 10 * ng);
        FORG out[ig] = speccomb * (w0 * r0[ig] + w1 * r1[ig] + w2 * r9[ig] + w3 * r10[ig]);
    }
}

// lerp in the first index of a [lead][ng] table: t(i) + f*(t(i+1) - t(i)), i 1-based
template <int GN>
__device__ __forceinline__ void lerp_rows(const double *__restrict__ t, int ng, int g0, int i, double f,
                                          double (&out)[GN]) {
    const double *r = t + ((i - 1) * ng + g0);
    const GRow<GN> a = ldrow<GN>(r), b = ldrow<GN>(r + ng);
    FORG out[ig] = a[ig] + f * (b[ig] - a[ig]);
}

// binary minor species k(jm,indm,g) of shape (nj,19,ng) (e.g. taugb3 :548-552)
template <int GN>
__device__ __forceinline__ void minor2(const double *__restrict__ k, int ng, int g0, int nj, int jm, int indm,
                                       double fm, double minorfrac, double (&out)[GN]) {
    const double *a = k + (((jm - 1) + nj * (indm - 1)) * ng + g0);   // K(jm,indm)
    const double *b = a + nj * ng;                                    // K(jm,indm+1)
    const GRow<GN> a0 = ldrow<GN>(a), a1 = ldrow<GN>(a + ng), b0 = ldrow<GN>(b), b1 = ldrow<GN>(b + ng);
    FORG {
        const double m1 = a0[ig] + fm * (a1[ig] - a0[ig]);
        const double m2 = b0[ig] + fm * (b1[ig] - b0[ig]);
        out[ig] = m1 + minorfrac * (m2 - m1);
    }
}

// four-point (p,T) interpolation of a single-species table: rows ind0, ind0+1, ind1, ind1+1
template <int GN>
__device__ __forceinline__ void key4(const double *__restrict__ tab, int ng, int g0, int ind0, int ind1,
                                     const Lay &L, double (&out)[GN]) {
    const double fac00 = L.f(F_FAC00), fac10 = L.f(F_FAC10), fac01 = L.f(F_FAC01), fac11 = L.f(F_FAC11);
    const double *r0 = tab + ((ind0 - 1) * ng + g0);
    const double *r1 = tab + ((ind1 - 1) * ng + g0);
    const GRow<GN> a0 = ldrow<GN>(r0), a1 = ldrow<GN>(r0 + ng), b0 = ldrow<GN>(r1), b1 = ldrow<GN>(r1 + ng);
    FORG out[ig] = fac00 * a0[ig] + fac10 * a1[ig] + fac01 * b0[ig] + fac11 * b1[ig];
}

// upper-atmosphere binary key species with nspb = 5 (bands 3, 4, 5; e.g. :675-685)
template <int GN>
__device__ __forceinline__ void key_upper5(const double *__restrict__ tab, int ng, int g0, int ind0, int ind1,
                                           const Spec &s0, const Spec &s1, const Lay &L, double (&out)[GN]) {
    const double fac00 = L.f(F_FAC00), fac10 = L.f(F_FAC10), fac01 = L.f(F_FAC01), fac11 = L.f(F_FAC11);
    const double fac000 = (1. - s0.fs) * fac00, fac010 = (1. - s0.fs) * fac10;
    const double fac100 = s0.fs * fac00, fac110 = s0.fs * fac10;
    const double fac001 = (1. - s1.fs) * fac01, fac011 = (1. - s1.fs) * fac11;
    const double fac101 = s1.fs * fac01, fac111 = s1.fs * fac11;
    const double *r0 = tab + ((ind0 - 1) * ng + g0);
    const double *r1 = tab + ((ind1 - 1) * ng + g0);
    const GRow<GN> a0 = ldrow<GN>(r0), a1 = ldrow<GN>(r0 + ng), a5 = ldrow<GN>(r0 + 5 * ng), a6 = ldrow<GN>(r0 + 6 * ng);
    const GRow<GN> b0 = ldrow<GN>(r1), b1 = ldrow<GN>(r1 + ng), b5 = ldrow<GN>(r1 + 5 * ng), b6 = ldrow<GN>(r1 + 6 * ng);
    FORG out[ig] = s0.speccomb * (fac000 * a0[ig] + fac100 * a1[ig] + fac010 * a5[ig] + fac110 * a6[ig]) +
                   s1.speccomb * (fac001 * b0[ig] + fac101 * b1[ig] + fac011 * b5[ig] + fac111 * b6[ig]);
}

__device__ __forceinline__ double rat_tab(int which, int jp1 /* 1-based */) { return c_lw.rat[which * 59 + jp1 - 1]; }
enum { R_H2OCO2 = 0, R_H2OO3 = 1, R_H2ON2O = 2, R_H2OCH4 = 3, R_N2OCO2 = 4, R_O3CO2 = 5 };
// Planck fraction interpolated in the binary-species parameter (e.g. :604-605)
template <int GN>
__device__ __forceinline__ void pfrac2(const double *__restrict__ fr, int ng, int g0, const Spec &sp,
                                       double (&pf)[GN]) {
    const double *r = fr + ((sp.js - 1) * ng + g0);
    const GRow<GN> a = ldrow<GN>(r), b = ldrow<GN>(r + ng);
    FORG pf[ig] = a[ig] + sp.fs * (b[ig] - a[ig]);
}
template <int GN>
__device__ __forceinline__ void pfrac1(const double *__restrict__ fr, int g0, double (&pf)[GN]) {
    const GRow<GN> a = ldrow<GN>(fr + g0);
    FORG pf[ig] = a[ig];
}

// Gas optical depth (TAU = true) and Planck fraction of one layer for g-points [G0, G0+GN) of
// BAND (G0 is the thread's first g-point within the band).  `lower` selects the lower/upper-atmosphere branch (lay <= laytrop).  The aerosol term
// of addAerosols is added by the caller.
template <int BAND, int GN, bool TAU>
__device__ __forceinline__ void lw_band_layer(const Lay &L, bool lower, double pavel, const int G0,
                                              double (&taug)[GN], double (&pf)[GN]) {
    const LwBandTab &B = c_lw.b[BAND - 1];
    constexpr int ng = BAND == 1 ? 10 : BAND == 2 ? 12 : BAND == 3 ? 16 : BAND == 4 ? 14 : BAND == 5 ? 16
                     : BAND == 6 ? 8 : BAND == 7 ? 12 : BAND == 8 ? 8 : BAND == 9 ? 12 : BAND == 10 ? 6
                     : BAND == 11 ? 8 : BAND == 12 ? 8 : BAND == 13 ? 4 : 2;
    // nspa = 1 1 9 9 9 1 9 1 9 1 1 9 9 1 9 9 ; nspb = 1 1 5 5 5 0 1 1 1 1 1 0 0 1 0 0 (rrtmg_lw_init.F90:194-195)
    constexpr int nspa = (BAND == 3 || BAND == 4 || BAND == 5 || BAND == 7 || BAND == 9 || BAND == 12 ||
                          BAND == 13 || BAND == 15 || BAND == 16) ? 9 : 1;
    // band 16 multiplies its upper-atmosphere index by nspb(16) = 0, i.e. always reads rows 1,2
    constexpr int nspb = (BAND == 3 || BAND == 4 || BAND == 5) ? 5 : (BAND == 16 ? 0 : 1);
    const int ind0lo = ((L.jp - 1) * 5 + (L.jt - 1)) * nspa;
    const int ind1lo = (L.jp * 5 + (L.jt1 - 1)) * nspa;
    const int ind0up = ((L.jp - 13) * 5 + (L.jt - 1)) * nspb;
    const int ind1up = ((L.jp - 12) * 5 + (L.jt1 - 1)) * nspb;
    double t1[GN], t2[GN], t3[GN];

    // self and foreign continuum shared by most lower-atmosphere branches
    auto self_for = [&](double (&acc)[GN]) {   // acc += tauself + taufor, in that order
        const double selffac = L.f(F_SELFFAC), selffrac = L.f(F_SELFFRAC);
        const double forfac = L.f(F_FORFAC), forfrac = L.f(F_FORFRAC);
        lerp_rows<GN>(B.selfref, ng, G0, L.indself, selffrac, t2);
        lerp_rows<GN>(B.forref, ng, G0, L.indfor, forfrac, t3);
        FORG acc[ig] = acc[ig] + selffac * t2[ig] + forfac * t3[ig];
    };
    auto for_only = [&](double (&acc)[GN]) {
        const double forfac = L.f(F_FORFAC), forfrac = L.f(F_FORFRAC);
        lerp_rows<GN>(B.forref, ng, G0, L.indfor, forfrac, t3);
        FORG acc[ig] = acc[ig] + forfac * t3[ig];
    };
    // lower-atmosphere binary key species: tau_major + tau_major1
    auto binary_lower = [&](const Spec &s0, const Spec &s1) {
        const double *r0 = B.absa + ((ind0lo + s0.js - 1) * ng + G0);
        const double *r1 = B.absa + ((ind1lo + s1.js - 1) * ng + G0);
        stencil_lower<GN>(r0, ng, s0.specparm, s0.fs, L.f(F_FAC00), L.f(F_FAC10), s0.speccomb, t1);
        stencil_lower<GN>(r1, ng, s1.specparm, s1.fs, L.f(F_FAC01), L.f(F_FAC11), s1.speccomb, t2);
        FORG taug[ig] = t1[ig] + t2[ig];
    };

    if constexpr (BAND == 1) {   // :191-285  H2O (+N2 continuum)
        if (lower) {
            if constexpr (TAU) {
                double corradj = 1.;
                if (pavel < 250.) corradj = 1. - 0.15 * (250. - pavel) / 154.4;
                const double scalen2 = L.f(F_COLBRD) * L.f(F_SCALEMINORN2), colh2o = L.f(F_COLH2O);
                key4<GN>(B.absa, ng, G0, ind0lo + 1, ind1lo + 1, L, t1);
                FORG taug[ig] = colh2o * t1[ig];
                self_for(taug);
                lerp_rows<GN>(B.ka_mn2, ng, G0, L.indminor, L.f(F_MINORFRAC), t1);
                FORG taug[ig] = corradj * (taug[ig] + scalen2 * t1[ig]);
            }
            pfrac1<GN>(B.fracrefa, G0, pf);
        } else {
            if constexpr (TAU) {
                const double corradj = 1. - 0.15 * (pavel / 95.6);
                const double scalen2 = L.f(F_COLBRD) * L.f(F_SCALEMINORN2), colh2o = L.f(F_COLH2O);
                key4<GN>(B.absb, ng, G0, ind0up + 1, ind1up + 1, L, t1);
                FORG taug[ig] = colh2o * t1[ig];
                for_only(taug);
                lerp_rows<GN>(B.kb_mn2, ng, G0, L.indminor, L.f(F_MINORFRAC), t1);
                FORG taug[ig] = corradj * (taug[ig] + scalen2 * t1[ig]);
            }
            pfrac1<GN>(B.fracrefb, G0, pf);
        }
    } else if constexpr (BAND == 2) {   // :289-363  H2O
        if (lower) {
            if constexpr (TAU) {
                const double corradj = 1. - .05 * (pavel - 100.) / 900.;
                const double colh2o = L.f(F_COLH2O);
                key4<GN>(B.absa, ng, G0, ind0lo + 1, ind1lo + 1, L, t1);
                FORG taug[ig] = colh2o * t1[ig];
                self_for(taug);
                FORG taug[ig] = corradj * taug[ig];
            }
            pfrac1<GN>(B.fracrefa, G0, pf);
        } else {
            if constexpr (TAU) {
                const double colh2o = L.f(F_COLH2O);
                key4<GN>(B.absb, ng, G0, ind0up + 1, ind1up + 1, L, t1);
                FORG taug[ig] = colh2o * t1[ig];
                for_only(taug);
            }
            pfrac1<GN>(B.fracrefb, G0, pf);
        }
    } else if constexpr (BAND == 3) {   // :367-695  H2O,CO2 (+N2O)
        const double colh2o = L.f(F_COLH2O), colco2 = L.f(F_COLCO2);
        if (lower) {
            if constexpr (TAU) {
                const Spec s0 = spec(colh2o, rat_tab(R_H2OCO2, L.jp), colco2, 8.);
                const Spec s1 = spec(colh2o, rat_tab(R_H2OCO2, L.jp + 1), colco2, 8.);
                const Spec sm = spec(colh2o, c_lw.chi_1_3_over_2_3, colco2, 8.);
                const double adjcoln2o = L.f(F_ADJN2O);
                binary_lower(s0, s1);
                self_for(taug);
                minor2<GN>(B.ka_mn2o, ng, G0, 9, sm.js, L.indminor, sm.fs, L.f(F_MINORFRAC), t1);
                FORG taug[ig] = taug[ig] + adjcoln2o * t1[ig];
            }
            pfrac2<GN>(B.fracrefa, ng, G0, spec(colh2o, c_lw.chi_1_9_over_2_9, colco2, 8.), pf);
        } else {
            if constexpr (TAU) {
                const Spec s0 = spec(colh2o, rat_tab(R_H2OCO2, L.jp), colco2, 4.);
                const Spec s1 = spec(colh2o, rat_tab(R_H2OCO2, L.jp + 1), colco2, 4.);
                const Spec sm = spec(colh2o, c_lw.chi_1_13_over_2_13, colco2, 4.);
                const double adjcoln2o = L.f(F_ADJN2O);
                key_upper5<GN>(B.absb, ng, G0, ind0up + s0.js, ind1up + s1.js, s0, s1, L, taug);
                for_only(taug);
                minor2<GN>(B.kb_mn2o, ng, G0, 5, sm.js, L.indminor, sm.fs, L.f(F_MINORFRAC), t1);
                FORG taug[ig] = taug[ig] + adjcoln2o * t1[ig];
            }
            pfrac2<GN>(B.fracrefb, ng, G0, spec(colh2o, c_lw.chi_1_13_over_2_13, colco2, 4.), pf);
        }
    } else if constexpr (BAND == 4) {   // :699-960  H2O,CO2 / O3,CO2
        const double colco2 = L.f(F_COLCO2);
        if (lower) {
            const double colh2o = L.f(F_COLH2O);
            if constexpr (TAU) {
                const Spec s0 = spec(colh2o, rat_tab(R_H2OCO2, L.jp), colco2, 8.);
                const Spec s1 = spec(colh2o, rat_tab(R_H2OCO2, L.jp + 1), colco2, 8.);
                binary_lower(s0, s1);
                self_for(taug);
            }
            pfrac2<GN>(B.fracrefa, ng, G0, spec(colh2o, c_lw.chi_1_11_over_2_11, colco2, 8.), pf);
        } else {
            const double colo3 = L.f(F_COLO3);
            if constexpr (TAU) {
                const Spec s0 = spec(colo3, rat_tab(R_O3CO2, L.jp), colco2, 4.);
                const Spec s1 = spec(colo3, rat_tab(R_O3CO2, L.jp + 1), colco2, 4.);
                key_upper5<GN>(B.absb, ng, G0, ind0up + s0.js, ind1up + s1.js, s0, s1, L, taug);
                // empirical stratospheric CO2 cooling-rate fix, :948-954 (g-points 8..14 of the band)
                constexpr double fix[14] = {1., 1., 1., 1., 1., 1., 1., 0.92, 0.88, 1.07, 1.1, 0.99, 0.88, 0.943};
                FORG if (G0 + ig >= 7) taug[ig] = taug[ig] * fix[G0 + ig];
            }
            pfrac2<GN>(B.fracrefb, ng, G0, spec(colo3, c_lw.chi_3_13_over_2_13, colco2, 4.), pf);
        }
    } else if constexpr (BAND == 5) {   // :964-1239  H2O,CO2 / O3,CO2 (+O3, CCl4)
        const double colco2 = L.f(F_COLCO2);
        if (lower) {
            const double colh2o = L.f(F_COLH2O);
            if constexpr (TAU) {
                const Spec s0 = spec(colh2o, rat_tab(R_H2OCO2, L.jp), colco2, 8.);
                const Spec s1 = spec(colh2o, rat_tab(R_H2OCO2, L.jp + 1), colco2, 8.);
                const Spec sm = spec(colh2o, c_lw.chi_1_7_over_2_7, colco2, 8.);
                const double colo3 = L.f(F_COLO3), colccl4 = L.f(F_COLCCL4);
                binary_lower(s0, s1);
                self_for(taug);
                minor2<GN>(B.ka_mo3, ng, G0, 9, sm.js, L.indminor, sm.fs, L.f(F_MINORFRAC), t1);
                FORG taug[ig] = taug[ig] + t1[ig] * colo3 + colccl4 * B.ccl4[G0 + ig];
            }
            pfrac2<GN>(B.fracrefa, ng, G0, spec(colh2o, c_lw.chi_1_5_over_2_5, colco2, 8.), pf);
        } else {
            const double colo3 = L.f(F_COLO3);
            if constexpr (TAU) {
                const Spec s0 = spec(colo3, rat_tab(R_O3CO2, L.jp), colco2, 4.);
                const Spec s1 = spec(colo3, rat_tab(R_O3CO2, L.jp + 1), colco2, 4.);
                const double colccl4 = L.f(F_COLCCL4);
                key_upper5<GN>(B.absb, ng, G0, ind0up + s0.js, ind1up + s1.js, s0, s1, L, taug);
                FORG taug[ig] = taug[ig] + colccl4 * B.ccl4[G0 + ig];
            }
            pfrac2<GN>(B.fracrefb, ng, G0, spec(colo3, c_lw.chi_3_43_over_2_43, colco2, 4.), pf);
        }
    } else if constexpr (BAND == 6) {   // :1243-1327  H2O (+CO2, CFC11, CFC12)
        if constexpr (TAU) {
            const double colcfc11 = L.f(F_COLCFC11), colcfc12 = L.f(F_COLCFC12);
            if (lower) {
                const double adjcolco2 = L.f(F_ADJCO2_6);
                const double colh2o = L.f(F_COLH2O);
                key4<GN>(B.absa, ng, G0, ind0lo + 1, ind1lo + 1, L, t1);
                FORG taug[ig] = colh2o * t1[ig];
                self_for(taug);
                lerp_rows<GN>(B.ka_mco2, ng, G0, L.indminor, L.f(F_MINORFRAC), t1);
                FORG taug[ig] = taug[ig] + adjcolco2 * t1[ig] + colcfc11 * B.cfc11adj[G0 + ig] +
                                colcfc12 * B.cfc12[G0 + ig];
            } else {
                FORG taug[ig] = 0.0 + colcfc11 * B.cfc11adj[G0 + ig] + colcfc12 * B.cfc12[G0 + ig];
            }
        }
        pfrac1<GN>(B.fracrefa, G0, pf);
    } else if constexpr (BAND == 7) {   // :1331-1603  H2O,O3 / O3 (+CO2)
        if (lower) {
            const double colh2o = L.f(F_COLH2O), colo3 = L.f(F_COLO3);
            if constexpr (TAU) {
                const Spec s0 = spec(colh2o, rat_tab(R_H2OO3, L.jp), colo3, 8.);
                const Spec s1 = spec(colh2o, rat_tab(R_H2OO3, L.jp + 1), colo3, 8.);
                const Spec sm = spec(colh2o, c_lw.chi_1_3_over_3_3, colo3, 8.);
                const double adjcolco2 = L.f(F_ADJCO2_7);
                binary_lower(s0, s1);
                self_for(taug);
                minor2<GN>(B.ka_mco2, ng, G0, 9, sm.js, L.indminor, sm.fs, L.f(F_MINORFRAC), t1);
                FORG taug[ig] = taug[ig] + adjcolco2 * t1[ig];
            }
            pfrac2<GN>(B.fracrefa, ng, G0, spec(colh2o, c_lw.chi_1_3_over_3_3, colo3, 8.), pf);
        } else {
            if constexpr (TAU) {
                const double adjcolco2 = L.f(F_ADJCO2_7);
                const double colo3 = L.f(F_COLO3);
                key4<GN>(B.absb, ng, G0, ind0up + 1, ind1up + 1, L, t1);
                lerp_rows<GN>(B.kb_mco2, ng, G0, L.indminor, L.f(F_MINORFRAC), t2);
                FORG taug[ig] = colo3 * t1[ig] + adjcolco2 * t2[ig];
                // empirical stratospheric O3 cooling-rate fix, :1592-1597 (g-points 6..11)
                constexpr double fix[12] = {1., 1., 1., 1., 1., 0.92, 0.88, 1.07, 1.1, 0.99, 0.855, 1.};
                FORG if (G0 + ig >= 5 && G0 + ig <= 10) taug[ig] = taug[ig] * fix[G0 + ig];
            }
            pfrac1<GN>(B.fracrefb, G0, pf);
        }
    } else if constexpr (BAND == 8) {   // :1607-1728  H2O / O3 (+CO2, O3, N2O, CFC12, CFC22)
        if constexpr (TAU) {
            const double adjcolco2 = L.f(F_ADJCO2_8);
            const double colo3 = L.f(F_COLO3), coln2o = L.f(F_COLN2O), colcfc12 = L.f(F_COLCFC12),
                         colcfc22 = L.f(F_COLCFC22), minorfrac = L.f(F_MINORFRAC);
            if (lower) {
                const double colh2o = L.f(F_COLH2O);
                key4<GN>(B.absa, ng, G0, ind0lo + 1, ind1lo + 1, L, t1);
                FORG taug[ig] = colh2o * t1[ig];
                self_for(taug);
                lerp_rows<GN>(B.ka_mco2, ng, G0, L.indminor, minorfrac, t1);
                lerp_rows<GN>(B.ka_mo3, ng, G0, L.indminor, minorfrac, t2);
                lerp_rows<GN>(B.ka_mn2o, ng, G0, L.indminor, minorfrac, t3);
                FORG taug[ig] = taug[ig] + adjcolco2 * t1[ig] + colo3 * t2[ig] + coln2o * t3[ig] +
                                colcfc12 * B.cfc12[G0 + ig] + colcfc22 * B.cfc22adj[G0 + ig];
            } else {
                key4<GN>(B.absb, ng, G0, ind0up + 1, ind1up + 1, L, t1);
                lerp_rows<GN>(B.kb_mco2, ng, G0, L.indminor, minorfrac, t2);
                lerp_rows<GN>(B.kb_mn2o, ng, G0, L.indminor, minorfrac, t3);
                FORG taug[ig] = colo3 * t1[ig] + adjcolco2 * t2[ig] + coln2o * t3[ig] +
                                colcfc12 * B.cfc12[G0 + ig] + colcfc22 * B.cfc22adj[G0 + ig];
            }
        }
        if (lower) pfrac1<GN>(B.fracrefa, G0, pf); else pfrac1<GN>(B.fracrefb, G0, pf);
    } else if constexpr (BAND == 9) {   // :1732-1994  H2O,CH4 / CH4 (+N2O)
        if (lower) {
            const double colh2o = L.f(F_COLH2O), colch4 = L.f(F_COLCH4);
            if constexpr (TAU) {
                const Spec s0 = spec(colh2o, rat_tab(R_H2OCH4, L.jp), colch4, 8.);
                const Spec s1 = spec(colh2o, rat_tab(R_H2OCH4, L.jp + 1), colch4, 8.);
                const Spec sm = spec(colh2o, c_lw.chi_1_3_over_6_3, colch4, 8.);
                const double adjcoln2o = L.f(F_ADJN2O);
                binary_lower(s0, s1);
                self_for(taug);
                minor2<GN>(B.ka_mn2o, ng, G0, 9, sm.js, L.indminor, sm.fs, L.f(F_MINORFRAC), t1);
                FORG taug[ig] = taug[ig] + adjcoln2o * t1[ig];
            }
            pfrac2<GN>(B.fracrefa, ng, G0, spec(colh2o, c_lw.chi_1_9_over_6_9, colch4, 8.), pf);
        } else {
            if constexpr (TAU) {
                const double adjcoln2o = L.f(F_ADJN2O);
                const double colch4 = L.f(F_COLCH4);
                key4<GN>(B.absb, ng, G0, ind0up + 1, ind1up + 1, L, t1);
                lerp_rows<GN>(B.kb_mn2o, ng, G0, L.indminor, L.f(F_MINORFRAC), t2);
                FORG taug[ig] = colch4 * t1[ig] + adjcoln2o * t2[ig];
            }
            pfrac1<GN>(B.fracrefb, G0, pf);
        }
    } else if constexpr (BAND == 10 || BAND == 11) {   // :1998-2066 H2O ; :2070-2149 H2O (+O2)
        if constexpr (TAU) {
            const double colh2o = L.f(F_COLH2O);
            if (lower) {
                key4<GN>(B.absa, ng, G0, ind0lo + 1, ind1lo + 1, L, t1);
                FORG taug[ig] = colh2o * t1[ig];
                self_for(taug);
            } else {
                key4<GN>(B.absb, ng, G0, ind0up + 1, ind1up + 1, L, t1);
                FORG taug[ig] = colh2o * t1[ig];
                for_only(taug);
            }
            if constexpr (BAND == 11) {
                const double scaleo2 = L.f(F_COLO2) * L.f(F_SCALEMINOR);
                lerp_rows<GN>(lower ? B.ka_mo2 : B.kb_mo2, ng, G0, L.indminor, L.f(F_MINORFRAC), t1);
                FORG taug[ig] = taug[ig] + scaleo2 * t1[ig];
            }
        }
        if (lower) pfrac1<GN>(B.fracrefa, G0, pf); else pfrac1<GN>(B.fracrefb, G0, pf);
    } else if constexpr (BAND == 12) {   // :2153-2356  H2O,CO2 / nothing
        if (lower) {
            const double colh2o = L.f(F_COLH2O), colco2 = L.f(F_COLCO2);
            if constexpr (TAU) {
                const Spec s0 = spec(colh2o, rat_tab(R_H2OCO2, L.jp), colco2, 8.);
                const Spec s1 = spec(colh2o, rat_tab(R_H2OCO2, L.jp + 1), colco2, 8.);
                binary_lower(s0, s1);
                self_for(taug);
            }
            pfrac2<GN>(B.fracrefa, ng, G0, spec(colh2o, c_lw.chi_1_10_over_2_10, colco2, 8.), pf);
        } else {
            FORG { if constexpr (TAU) taug[ig] = 0.0; pf[ig] = 0.0; }
        }
    } else if constexpr (BAND == 13) {   // :2360-2620  H2O,N2O (+CO2, CO) / O3 minor
        if (lower) {
            const double colh2o = L.f(F_COLH2O), coln2o = L.f(F_COLN2O);
            if constexpr (TAU) {
                const Spec s0 = spec(colh2o, rat_tab(R_H2ON2O, L.jp), coln2o, 8.);
                const Spec s1 = spec(colh2o, rat_tab(R_H2ON2O, L.jp + 1), coln2o, 8.);
                const Spec smco2 = spec(colh2o, c_lw.chi_1_1_over_4_1, coln2o, 8.);
                const double coldry = L.f(F_COLDRY);
                const double adjcolco2 = L.f(F_ADJCO2_13);
                const Spec smco = spec(colh2o, c_lw.chi_1_3_over_4_3, coln2o, 8.);
                // covmr = 0 in GEOS (LW/src/rrtmg_lw_rad.F90:520) -> colco takes its 1e-32 floor
                const double colco = 1.e-32 * coldry;
                binary_lower(s0, s1);
                self_for(taug);
                minor2<GN>(B.ka_mco2, ng, G0, 9, smco2.js, L.indminor, smco2.fs, L.f(F_MINORFRAC), t1);
                minor2<GN>(B.ka_mco, ng, G0, 9, smco.js, L.indminor, smco.fs, L.f(F_MINORFRAC), t2);
                FORG taug[ig] = taug[ig] + adjcolco2 * t1[ig] + colco * t2[ig];
            }
            pfrac2<GN>(B.fracrefa, ng, G0, spec(colh2o, c_lw.chi_1_5_over_4_5, coln2o, 8.), pf);
        } else {
            if constexpr (TAU) {
                const double colo3 = L.f(F_COLO3);
                lerp_rows<GN>(B.kb_mo3, ng, G0, L.indminor, L.f(F_MINORFRAC), t1);
                FORG taug[ig] = colo3 * t1[ig];
            }
            pfrac1<GN>(B.fracrefb, G0, pf);
        }
    } else if constexpr (BAND == 14) {   // :2624-2686  CO2
        if constexpr (TAU) {
            const double colco2 = L.f(F_COLCO2);
            if (lower) {
                key4<GN>(B.absa, ng, G0, ind0lo + 1, ind1lo + 1, L, t1);
                FORG taug[ig] = colco2 * t1[ig];
                self_for(taug);
            } else {
                key4<GN>(B.absb, ng, G0, ind0up + 1, ind1up + 1, L, t1);
                FORG taug[ig] = colco2 * t1[ig];
            }
        }
        if (lower) pfrac1<GN>(B.fracrefa, G0, pf); else pfrac1<GN>(B.fracrefb, G0, pf);
    } else if constexpr (BAND == 15) {   // :2690-2913  N2O,CO2 (+N2) / nothing
        if (lower) {
            const double coln2o = L.f(F_COLN2O), colco2 = L.f(F_COLCO2);
            const Spec sp = spec(coln2o, c_lw.chi_4_1_over_2_1, colco2, 8.);
            if constexpr (TAU) {
                const Spec s0 = spec(coln2o, rat_tab(R_N2OCO2, L.jp), colco2, 8.);
                const Spec s1 = spec(coln2o, rat_tab(R_N2OCO2, L.jp + 1), colco2, 8.);
                const double scalen2 = L.f(F_COLBRD) * L.f(F_SCALEMINOR);
                binary_lower(s0, s1);
                self_for(taug);
                minor2<GN>(B.ka_mn2, ng, G0, 9, sp.js, L.indminor, sp.fs, L.f(F_MINORFRAC), t1);
                FORG taug[ig] = taug[ig] + scalen2 * t1[ig];
            }
            pfrac2<GN>(B.fracrefa, ng, G0, sp, pf);
        } else {
            FORG { if constexpr (TAU) taug[ig] = 0.0; pf[ig] = 0.0; }
        }
    } else {   // BAND == 16, :2917-3126  H2O,CH4 / CH4
        if (lower) {
            const double colh2o = L.f(F_COLH2O), colch4 = L.f(F_COLCH4);
            if constexpr (TAU) {
                const Spec s0 = spec(colh2o, rat_tab(R_H2OCH4, L.jp), colch4, 8.);
                const Spec s1 = spec(colh2o, rat_tab(R_H2OCH4, L.jp + 1), colch4, 8.);
                binary_lower(s0, s1);
                self_for(taug);
            }
            pfrac2<GN>(B.fracrefa, ng, G0, spec(colh2o, c_lw.chi_1_6_over_6_6, colch4, 8.), pf);
        } else {
            if constexpr (TAU) {
                const double colch4 = L.f(F_COLCH4);
                key4<GN>(B.absb, ng, G0, ind0up + 1, ind1up + 1, L, t1);
                FORG taug[ig] = colch4 * t1[ig];
            }
            pfrac1<GN>(B.fracrefb, G0, pf);
        }
    }
    (void)t1; (void)t2; (void)t3; (void)ind0lo; (void)ind1lo; (void)ind0up; (void)ind1up; (void)pavel;
}

// ---------------------------------------------------------------------------------------------
// fused gas optics + radiance sweeps for one (band, g sub-range): LW/src/rrtmg_lw_rtrnmc.F90:27-390
// ---------------------------------------------------------------------------------------------
struct LwBandArgs {
    int ld, col0;
    const int *perm;         // optional column grouping of the chunk
    LwWork W;
    int dudTs;
    const double *pavel;     // caller (ld,nlay)
    const double *semiss;    // caller (ld,16)
    const double *taua;      // caller (ld,nlay,16)
    double *dbg_taug, *dbg_pfracs;   // optional [nlay][140][nc]
};

// Sum v[q] over the threads of a block that share a column (threadIdx.x runs over the g-point groups
// of the band, threadIdx.y over the block's columns) in ascending g order and store the Q totals at
// dst + q*qstride.  `red` holds Q*CB*NY doubles; callers alternate two buffers so one barrier per call
// suffices.
// the g-point groups of a column are NY adjacent lanes of one warp when NY is a power of two and the block is
// made of whole warps: the sums are then an xor butterfly of warp shuffles (no shared memory, no block barrier)
template <int NY, int CB> __host__ __device__ constexpr bool lw_shuffle_sums() {
    return NY > 1 && NY <= 32 && (NY & (NY - 1)) == 0 && (NY * CB) % 32 == 0;
}
// The xor butterfly over the NY lanes of a column, as a reduce-scatter: while a lane still holds several of the Q
// values, a round halves them (the lane keeps the lower or the upper half by its bit M and sends the other half to
// its partner) instead of exchanging all of them; the last rounds are the plain butterfly on the one value left.
// Every total is the same tree of pairwise sums as in the full butterfly (IEEE addition commutes), so the bits do
// not change; Q = 4 over 8 lanes takes 4 exchanges of a double instead of 12.  The totals end up spread over the
// column's lanes: the lane whose remaining bits are zero stores its value(s).
template <int Q, int M>
__device__ __forceinline__ void lw_shuffle_reduce_scatter(double (&s)[Q], const int ty, const int q0, const bool store,
                                                          double *__restrict__ dst, const size_t qstride) {
    if constexpr (M == 0) {
        if (store) {
#pragma unroll
            for (int q = 0; q < Q; ++q) dst[(size_t)(q0 + q) * qstride] = s[q];
        }
    } else if constexpr (Q > 1 && Q % 2 == 0) {
        constexpr int H = Q / 2;
        const bool upper = (ty & M) != 0;
        double k[H];
#pragma unroll
        for (int q = 0; q < H; ++q) {
            const double keep = upper ? s[H + q] : s[q];
            const double send = upper ? s[q] : s[H + q];
            k[q] = keep + __shfl_xor_sync(0xffffffffu, send, M);
        }
        lw_shuffle_reduce_scatter<H, M / 2>(k, ty, q0 + (upper ? H : 0), store, dst, qstride);
    } else {
#pragma unroll
        for (int q = 0; q < Q; ++q) s[q] = s[q] + __shfl_xor_sync(0xffffffffu, s[q], M);
        lw_shuffle_reduce_scatter<Q, M / 2>(s, ty, q0, store && (ty & M) == 0, dst, qstride);
    }
}

template <int Q, int NY, int CB>
__device__ __forceinline__ void block_sum_store(const double (&v)[Q], double *__restrict__ red,
                                                double *__restrict__ dst, size_t qstride, bool active) {
    const int ty = threadIdx.x, lane = threadIdx.y;
    if (NY == 1) {
        if (active) {
#pragma unroll
            for (int q = 0; q < Q; ++q) dst[q * qstride] = v[q];
        }
        return;
    }
    if constexpr (lw_shuffle_sums<NY, CB>()) {
        double s[Q];
#pragma unroll
        for (int q = 0; q < Q; ++q) s[q] = v[q];
        lw_shuffle_reduce_scatter<Q, NY / 2>(s, ty, 0, active, dst, qstride);
        (void)red; (void)lane;
        return;
    }
#pragma unroll
    for (int q = 0; q < Q; ++q) red[(q * CB + lane) * NY + ty] = v[q];
    __syncthreads();
    for (int q = ty; q < Q; q += NY) {
        const double *r = red + (q * CB + lane) * NY;
        double s = r[0];
#pragma unroll
        for (int y = 1; y < NY; ++y) s = s + r[y];
        if (active) dst[q * qstride] = s;
    }
}

template <int BAND> struct LwBandInfo {
    static constexpr int ng = BAND == 1 ? 10 : BAND == 2 ? 12 : BAND == 3 ? 16 : BAND == 4 ? 14 : BAND == 5 ? 16
                            : BAND == 6 ? 8 : BAND == 7 ? 12 : BAND == 8 ? 8 : BAND == 9 ? 12 : BAND == 10 ? 6
                            : BAND == 11 ? 8 : BAND == 12 ? 8 : BAND == 13 ? 4 : 2;
};

// factor arrays (LwF) a band's gas optics read, for the next-layer prefetch
__host__ __device__ constexpr unsigned lw_fbit(int k) { return 1u << k; }
constexpr unsigned LW_FAC4 = lw_fbit(F_FAC00) | lw_fbit(F_FAC01) | lw_fbit(F_FAC10) | lw_fbit(F_FAC11);
constexpr unsigned LW_SELFFOR = lw_fbit(F_SELFFAC) | lw_fbit(F_SELFFRAC) | lw_fbit(F_FORFAC) | lw_fbit(F_FORFRAC);
template <int BAND> __host__ __device__ constexpr unsigned lw_band_fmask() {
    constexpr unsigned base = LW_FAC4 | LW_SELFFOR;
    return BAND == 1 ? base | lw_fbit(F_COLH2O) | lw_fbit(F_COLBRD) | lw_fbit(F_SCALEMINORN2) | lw_fbit(F_MINORFRAC)
         : BAND == 2 ? base | lw_fbit(F_COLH2O)
         : BAND == 3 ? base | lw_fbit(F_COLH2O) | lw_fbit(F_COLCO2) | lw_fbit(F_ADJN2O) | lw_fbit(F_MINORFRAC)
         : BAND == 4 ? base | lw_fbit(F_COLH2O) | lw_fbit(F_COLCO2) | lw_fbit(F_COLO3)
         : BAND == 5 ? base | lw_fbit(F_COLH2O) | lw_fbit(F_COLCO2) | lw_fbit(F_COLO3) | lw_fbit(F_COLCCL4) | lw_fbit(F_MINORFRAC)
         : BAND == 6 ? base | lw_fbit(F_COLH2O) | lw_fbit(F_COLCFC11) | lw_fbit(F_COLCFC12) | lw_fbit(F_ADJCO2_6) | lw_fbit(F_MINORFRAC)
         : BAND == 7 ? base | lw_fbit(F_COLH2O) | lw_fbit(F_COLO3) | lw_fbit(F_ADJCO2_7) | lw_fbit(F_MINORFRAC)
         : BAND == 8 ? base | lw_fbit(F_COLH2O) | lw_fbit(F_COLO3) | lw_fbit(F_COLN2O) | lw_fbit(F_COLCFC12) | lw_fbit(F_COLCFC22) | lw_fbit(F_ADJCO2_8) | lw_fbit(F_MINORFRAC)
         : BAND == 9 ? base | lw_fbit(F_COLH2O) | lw_fbit(F_COLCH4) | lw_fbit(F_ADJN2O) | lw_fbit(F_MINORFRAC)
         : BAND == 10 ? base | lw_fbit(F_COLH2O)
         : BAND == 11 ? base | lw_fbit(F_COLH2O) | lw_fbit(F_COLO2) | lw_fbit(F_SCALEMINOR) | lw_fbit(F_MINORFRAC)
         : BAND == 12 ? base | lw_fbit(F_COLH2O) | lw_fbit(F_COLCO2)
         : BAND == 13 ? base | lw_fbit(F_COLH2O) | lw_fbit(F_COLN2O) | lw_fbit(F_COLO3) | lw_fbit(F_COLDRY) | lw_fbit(F_ADJCO2_13) | lw_fbit(F_MINORFRAC)
         : BAND == 14 ? base | lw_fbit(F_COLCO2)
         : BAND == 15 ? base | lw_fbit(F_COLN2O) | lw_fbit(F_COLCO2) | lw_fbit(F_COLBRD) | lw_fbit(F_SCALEMINOR) | lw_fbit(F_MINORFRAC)
         : base | lw_fbit(F_COLH2O) | lw_fbit(F_COLCH4);
}
// Planck fractions of the binary-species bands need the two key columns in the upward sweep too
template <int BAND> __host__ __device__ constexpr unsigned lw_band_fmask_up() {
    return lw_band_fmask<BAND>() & (lw_fbit(F_COLH2O) | lw_fbit(F_COLCO2) | lw_fbit(F_COLO3) | lw_fbit(F_COLN2O) |
                                    lw_fbit(F_COLCH4));
}

// partial flux profiles of a band: part[band][LP_*][lev][c]
enum LwPart { LP_U, LP_UC, LP_DU, LP_DUC, LP_D, LP_DC, LP_COUNT };

// Block = (ng/GN) g-point groups (x, fastest) x CB columns of BAND: a warp is the g-point groups of
// 32/(ng/GN) neighbouring columns.  The lanes of a column share jp/jt/js, so every k-table gather of a
// warp falls on a few rows read in full (the tables are g-point fastest), the per-(layer, column) setcoef
// state is one broadcast sector, and the per-cell scratch (lw_cell: g-point fastest inside a band) is
// one contiguous run.  (With 32 columns per warp every gather touched 32 scattered rows: the L1 data
// pipe was 88 % busy, profiles/r2_d_*.)  The g-point sums of every level are formed in the block
// (block_sum_store) in ascending g order, like the reference's sequential accumulation.
// Both sweeps of one (column, GN g-points of BAND) thread: rtrnmc :198-385 on the band's gas optics.  `ty` = the
// thread's g-point group, `sum` forms the g-point sums of a level over the column's threads (a block-level or a
// warp-level policy, below), `etab` = the {exp, tfn} transmittance table (global memory or the block's shared copy).
template <int BAND, int GN, class Sum>
__device__ __forceinline__ void lw_column_sweeps(const LwBandArgs &A, const int c, const bool active, const int ty,
                                                 const double2 *__restrict__ etab, Sum &sum) {
    constexpr int NY = LwBandInfo<BAND>::ng / GN;
    constexpr int NG = LwBandInfo<BAND>::ng;
    const LwWork &W = A.W;
    const int nc = W.nc, nlay = W.nlay;
    const size_t col = gcol(A.col0, A.perm, c);
    constexpr int ib = BAND - 1;
    const int gs = BAND == 1 ? 0 : c_lw.ngs[ib - 1];   // first g-point (0-based) of the band
    const int G0 = ty * GN;
    const int g_first = gs + G0;
    const int laytrop = W.laytrop[c];

    // diffusivity angle, :177-186
    double secdiff;
    if (BAND == 1 || BAND == 4 || BAND >= 10) {
        secdiff = 1.66;
    } else {
        constexpr double a0[9] = {1.66, 1.55, 1.58, 1.66, 1.54, 1.454, 1.89, 1.33, 1.668};
        constexpr double a1[9] = {0.00, 0.25, 0.22, 0.00, 0.13, 0.446, -0.10, 0.40, -0.006};
        constexpr double a2[9] = {0.00, -12.0, -11.7, 0.00, -0.72, -0.243, 0.19, -0.062, 0.414};
        secdiff = a0[ib < 9 ? ib : 0] + a1[ib < 9 ? ib : 0] * exp(a2[ib < 9 ? ib : 0] * W.pwvcm[c]);
        if (secdiff > 1.80) secdiff = 1.80;
        else if (secdiff < 1.50) secdiff = 1.50;
    }
    const double sumfac = 0.5 * c_lw.delwave[ib] * c_lw.fluxfac;
    const double bpade = c_lw.bpade;
    const double tblint = 10000.0;
    const size_t fstride = (size_t)(nlay + 1) * nc;
    double *part = W.part + (size_t)ib * LP_COUNT * fstride + c;   // [f][lev][c]
    const double *planklay = W.planklay + (size_t)ib * nlay * nc + c;
    const double *planklev = W.planklev + (size_t)ib * (nlay + 1) * nc + c;

    double radld[GN], radclrd[GN], taug[GN], pf[GN];
    FORG { radld[ig] = 0.; radclrd[ig] = 0.; }
    bool diverge = false;
    if (active && ty == 0) {
        part[LP_D * fstride + (size_t)nlay * nc] = 0.;
        part[LP_DC * fstride + (size_t)nlay * nc] = 0.;
    }
    // running addresses: per-(layer, column) planes by 32-bit offsets from the column's pointer, the
    // thread's cell of the lw_cell scratch by one 64-bit offset (GN adjacent elements per thread)
    const int *pidx = W.idx + W.ti(0, c);          // + 32 per layer
    const double *pfac = W.fbase + W.tf(0, c);     // + F_COUNT*32 per layer
    constexpr int lay_cell = 32 * NG;              // layer stride of the tiled per-cell scratch
    size_t koff = lw_cell(gs, NG, nlay, W.ncp, nlay - 1, c, G0);              // cell (nlay-1, c, G0)
    size_t aoff = ((size_t)ib * nlay + nlay - 1) * A.ld + col;                // aerosol at layer nlay-1
    const int nw = (nlay + 31) >> 5;
    uint32_t any_word = 0u, mword[GN];
    FORG mword[ig] = 0u;

    // ---- downward sweep, :198-309 ----
    for (int lay = nlay - 1; lay >= 0; --lay) {
        const int jl = lay * nc;
        if (lay > 0 && ty == 0) {   // next layer's per-(layer, column) state -> L1 while this one computes
            prefetch_l1(pidx + (lay - 1) * 32);
            constexpr unsigned fm = lw_band_fmask<BAND>();
#pragma unroll
            for (int k = 0; k < F_COUNT; ++k)
                if ((fm >> k) & 1u) prefetch_l1(pfac + ((lay - 1) * F_COUNT + k) * 32);
            prefetch_l1(planklay + jl - nc);
            prefetch_l1(planklev + jl - nc);
            prefetch_l1(A.taua + aoff - A.ld);
        }
        if ((lay & 31) == 31 || lay == nlay - 1) {   // cloud words of the next (up to) 32 layers
            any_word = W.cloudy_any[(lay >> 5) * nc + c];
            const uint32_t *pm = W.mask + lw_cell(gs, NG, nw, W.ncp, lay >> 5, c, G0);
            FORG mword[ig] = any_word ? pm[ig] : 0u;
        }
        Lay L;
        L.fj = pfac + lay * (F_COUNT * 32);
        const int pk = pidx[lay * 32];
        L.jp = pk & 63; L.jt = (pk >> 6) & 7; L.jt1 = (pk >> 9) & 7;
        L.indfor = (pk >> 12) & 3; L.indself = (pk >> 14) & 15; L.indminor = (pk >> 18) & 31;
        const double pavel = (BAND <= 2) ? A.pavel[(size_t)lay * A.ld + col] : 0.;
        lw_band_layer<BAND, GN, true>(L, lay < laytrop, pavel, G0, taug, pf);
        const double taer = A.taua[aoff];
        aoff -= A.ld;
        FORG taug[ig] = taug[ig] + taer;
        if (A.dbg_taug && active) FORG A.dbg_taug[((size_t)lay * 140 + g_first + ig) * nc + c] = taug[ig];
        if (A.dbg_pfracs && active) FORG A.dbg_pfracs[((size_t)lay * 140 + g_first + ig) * nc + c] = pf[ig];

        const double blay = planklay[jl];
        const double dplankdn = planklev[jl] - blay;
        const bool layer_cloudy = (any_word >> (lay & 31)) & 1u;
        if (!diverge && layer_cloudy) diverge = true;
        double sums[2] = {0., 0.};
        FORG {
            double odepth = secdiff * taug[ig];
            if (odepth < 0.) odepth = 0.;
            double tblind = ddiv(odepth, bpade + odepth);
            const int itgas = f_int(tblint * tblind + 0.5);
            const double2 et = etab[itgas];
            const double agas = 1. - et.x;
            const double bbdgas = pf[ig] * (blay + et.y * dplankdn);
            uint32_t code = (uint32_t)itgas | 0xffff0000u;
            if (!((mword[ig] >> (lay & 31)) & 1u)) {
                radld[ig] = radld[ig] + (bbdgas - radld[ig]) * agas;
            } else {
                const double odcld = secdiff * __ldcs(W.taucmc + koff + ig);
                const double odtot = c_lw.tau_tbl[itgas] + odcld;
                tblind = ddiv(odtot, bpade + odtot);
                const int ittot = f_int(tblint * tblind + 0.5);
                const double2 ett = etab[ittot];
                const double atot = 1. - ett.x;
                const double bbdtot = pf[ig] * (blay + ett.y * dplankdn);
                radld[ig] = radld[ig] + (bbdtot - radld[ig]) * atot;
                code = (uint32_t)itgas | ((uint32_t)ittot << 16);
            }
            if (active) __stcs(W.it + koff + ig, code);
            sums[0] = sums[0] + sumfac * radld[ig];
            if (diverge) radclrd[ig] = radclrd[ig] + (bbdgas - radclrd[ig]) * agas;
            else radclrd[ig] = radld[ig];
            sums[1] = sums[1] + sumfac * radclrd[ig];
        }
        if (active) {   // the upward sweep reads the fractions back instead of redoing the band's spectral interpolation
            if constexpr (GN == 2) __stcs(reinterpret_cast<double2 *>(W.pfs + koff), make_double2(pf[0], pf[1]));
            else FORG __stcs(W.pfs + koff + ig, pf[ig]);
        }
        sum.template store<2>(sums, part + LP_D * fstride + jl, fstride, active);
        koff -= lay_cell;
    }
    koff += lay_cell;   // back on layer 0

    // ---- surface, :319-333 (pf now holds the Planck fractions of layer 1) ----
    const double plankbnd = W.plankbnd[(size_t)ib * nc + c];
    const double reflect = 1. - A.semiss[(size_t)ib * A.ld + col];
    double radlu[GN], radclru[GN], drad[GN], dradc[GN];
    {
        const double dpb = A.dudTs ? W.dplankbnd[(size_t)ib * nc + c] : 0.;
        double sums[4] = {0., 0., 0., 0.};
        FORG {
            const double rad0 = pf[ig] * plankbnd;
            radlu[ig] = rad0 + reflect * radld[ig];
            radclru[ig] = rad0 + reflect * radclrd[ig];
            sums[0] = sums[0] + sumfac * radlu[ig];
            sums[1] = sums[1] + sumfac * radclru[ig];
            drad[ig] = pf[ig] * dpb;
            dradc[ig] = drad[ig];
            sums[2] = sums[2] + sumfac * drad[ig];
        }
        sums[3] = sums[2];
        sum.template store<4>(sums, part, fstride, active);
    }

    // ---- upward sweep, :336-379 ----
    // The cell's table indices (it) and Planck fractions (pfs) come back from the downward sweep.  The gas look-up
    // is software-pipelined: the indices of layer lay+2 are loaded while the table entries of layer lay+1 are
    // fetched and layer lay is computed, so neither the index load nor the data-dependent table read is waited
    // for (they were 16 % of the kernel's stall samples, profiles/s2_a_*).
    auto ld_code = [&](size_t off, uint32_t (&code)[GN]) {
        if constexpr (GN == 2) {
            const uint2 t = __ldcs(reinterpret_cast<const uint2 *>(W.it + off));
            code[0] = t.x; code[1] = t.y;
        } else {
            FORG code[ig] = __ldcs(W.it + off + ig);
        }
    };
    auto ld_tab = [&](const uint32_t (&code)[GN], double2 (&eg)[GN]) {
        FORG eg[ig] = etab[code[ig] & 0xffffu];
    };
    uint32_t cur[GN], code1[GN], code2[GN];   // indices of layers lay, lay+1, lay+2
    double2 eg0[GN], eg1[GN];                 // gas entries of layers lay, lay+1 (cloudy cells read theirs on demand)
    ld_code(koff, cur);
    ld_tab(cur, eg0);                  // layer 0
    FORG { code1[ig] = 0xffff0000u; code2[ig] = 0xffff0000u; }
    if (nlay > 1) ld_code(koff + lay_cell, code1);
    if (nlay > 2) ld_code(koff + 2 * (size_t)lay_cell, code2);
    ld_tab(code1, eg1);                // layer 1
    for (int lay = 0; lay < nlay; ++lay) {
        const int jl = lay * nc;
        double2 ett[GN];
        FORG {
            const uint32_t ittot = cur[ig] >> 16;
            if (ittot != 0xffffu) ett[ig] = etab[ittot];
        }
        if (lay + 3 < nlay) prefetch_l1(W.it + koff + 3 * (size_t)lay_cell);
        if (lay + 1 < nlay) {
            prefetch_l1(W.pfs + koff + lay_cell);
            if (ty == 0) {
                prefetch_l1(planklay + jl + nc);
                prefetch_l1(planklev + jl + 2 * nc);
            }
        }
        if constexpr (GN == 2) {
            const double2 t = __ldcs(reinterpret_cast<const double2 *>(W.pfs + koff));
            pf[0] = t.x; pf[1] = t.y;
        } else {
            FORG pf[ig] = __ldcs(W.pfs + koff + ig);
        }
        const double blay = planklay[jl];
        const double dplankup = planklev[jl + nc] - blay;
        double sums[4] = {0., 0., 0., 0.};
        FORG {
            const double2 et = eg0[ig];
            const double agas = 1. - et.x;
            const double bbugas = pf[ig] * (blay + et.y * dplankup);
            if ((cur[ig] >> 16) == 0xffffu) {
                radlu[ig] = radlu[ig] + (bbugas - radlu[ig]) * agas;
                if (A.dudTs) drad[ig] = drad[ig] - drad[ig] * agas;
            } else {
                const double atot = 1. - ett[ig].x;
                const double bbutot = pf[ig] * (blay + ett[ig].y * dplankup);
                radlu[ig] = radlu[ig] + (bbutot - radlu[ig]) * atot;
                if (A.dudTs) drad[ig] = drad[ig] - drad[ig] * atot;
            }
            sums[0] = sums[0] + sumfac * radlu[ig];
            if (diverge) radclru[ig] = radclru[ig] + (bbugas - radclru[ig]) * agas;
            else radclru[ig] = radlu[ig];
            sums[1] = sums[1] + sumfac * radclru[ig];
            if (A.dudTs) {
                if (diverge) dradc[ig] = dradc[ig] - dradc[ig] * agas;
                else dradc[ig] = drad[ig];
                sums[2] = sums[2] + sumfac * drad[ig];
                sums[3] = sums[3] + sumfac * dradc[ig];
            }
        }
        sum.template store<4>(sums, part + jl + nc, fstride, active);
        // rotate the pipeline: layer lay+1 becomes current, the entries of lay+2 are fetched, the indices of lay+3 loaded
        FORG { cur[ig] = code1[ig]; eg0[ig] = eg1[ig]; code1[ig] = code2[ig]; }
        if (lay + 2 < nlay) ld_tab(code1, eg1);
        if (lay + 3 < nlay) ld_code(koff + 3 * (size_t)lay_cell, code2);
        koff += lay_cell;
    }
}

// g-point sums of a level over the threads of a column, block-level: shared memory + barrier, or warp shuffles when
// the column's threads are a power-of-two run of lanes (block_sum_store).
// (A persistent variant - one 768-thread block per SM, the 160 KB {exp, tfn} table copied into shared memory by one
// cp.async.bulk, warp-level sums - was built and measured in commit 4e691fe: it is 20-40 % SLOWER on the 14- and
// 16-g bands, whose k-tables (75-225 KB) then fight over the 67 KB of L1 that are left, and 2-16 % faster only on
// bands 1 and 2; with every look-up pinned to entry 0 the kernels gain 7 %: the look-ups are not what binds them.
// profiles/s2_band_kernel_experiments.txt.)
template <int NY, int CB> struct LwBlockSum {
    double *red_buf;
    int flip = 0;
    template <int Q>
    __device__ __forceinline__ void store(const double (&v)[Q], double *__restrict__ dst, size_t qstride, bool active) {
        flip ^= 1;
        block_sum_store<Q, NY, CB>(v, red_buf + ((NY > 1 && !lw_shuffle_sums<NY, CB>()) ? flip * 4 * NY * CB : 0), dst,
                                   qstride, active);
    }
};
template <int BAND, int GN, int REGS, int CB>
__global__ void __launch_bounds__(CB * (LwBandInfo<BAND>::ng / GN), min_blocks(CB * (LwBandInfo<BAND>::ng / GN), REGS))
lw_band_kernel(const LwBandArgs A) {
    constexpr int NY = LwBandInfo<BAND>::ng / GN;
    static_assert(NY * GN == LwBandInfo<BAND>::ng, "GN must divide the band's g-points");
    __shared__ double red_buf[(NY > 1 && !lw_shuffle_sums<NY, CB>()) ? 2 * 4 * NY * CB : 1];
    const LwWork &W = A.W;
    if (RRTMGX_TRAPPED(W.trap)) return;   // block-uniform
    const int c0 = blockIdx.x * CB + threadIdx.y;
    const bool active = c0 < W.nc;
    const int c = active ? c0 : W.nc - 1;   // idle lanes shadow the last column and never store
    LwBlockSum<NY, CB> sum{red_buf};
    lw_column_sweeps<BAND, GN>(A, c, active, (int)threadIdx.x, reinterpret_cast<const double2 *>(c_lw.exptfn), sum);
}

// Compiled variants per band: the (g-points per thread, register budget) pair tuned for the band
// (profiles/r2_cb_tuning.txt, run r3c) at CB = 32, 16, 8, 4 columns per block.  The one used is picked
// per band from lw_variant[] (tuned on B200; RRTMGX_LW_GN="vvv..." overrides).
typedef void (*LwBandLauncher)(int, cudaStream_t, const LwBandArgs &);
template <int BAND, int GN, int REGS, int CB>
static void lw_launch_band(int nc, cudaStream_t st, const LwBandArgs &A) {
    static char tag[48] = "";
    if (!tag[0]) {
        std::snprintf(tag, sizeof tag, "lw_band_kernel<%d,gn%d,r%d,c%d>", BAND, GN, REGS, CB);
        // experiment (profiles/t1_g_*): preferred shared-memory share of the SM in percent, unset = the driver's choice
        if (const char *e = std::getenv("RRTMGX_CARVEOUT"))
            cudaFuncSetAttribute(lw_band_kernel<BAND, GN, REGS, CB>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                 std::atoi(e));
    }
    RRTMGX_LAUNCH_TAG(tag, (lw_band_kernel<BAND, GN, REGS, CB>), dim3((nc + CB - 1) / CB),
                      dim3(LwBandInfo<BAND>::ng / GN, CB), 0, st, A);
}
#define X(BAND, G, R) \
    {lw_launch_band<BAND, G, R, 32>, lw_launch_band<BAND, G, R, 16>, lw_launch_band<BAND, G, R, 8>, lw_launch_band<BAND, G, R, 4>},
static const LwBandLauncher lw_launchers[16][4] = {
    X(1, 2, 56) X(2, 2, 56) X(3, 2, 80) X(4, 2, 56) X(5, 2, 80) X(6, 2, 80) X(7, 2, 48) X(8, 2, 80)
    X(9, 2, 56) X(10, 2, 64) X(11, 2, 80) X(12, 2, 80) X(13, 1, 80) X(14, 1, 64) X(15, 1, 80) X(16, 1, 80)};
#undef X
static int lw_variant[16] = {0, 1, 3, 3, 3, 2, 3, 2, 3, 2, 2, 2, 2, 1, 1, 1};   // columns per block, profiles/r2_cb_tuning.txt (run r2h)
static const int lw_variant_default[16] = {0, 1, 3, 3, 3, 2, 3, 2, 3, 2, 2, 2, 2, 1, 1, 1};
void lw_read_env() {   // once per rrtmgx_init, under the library lock
    for (int b = 0; b < 16; ++b) lw_variant[b] = lw_variant_default[b];
    const char *e = std::getenv("RRTMGX_LW_GN");
    if (!e) return;
    int b = 0;
    for (const char *q = e; *q && b < 16; ++q)
        if (*q >= '0' && *q <= '3') lw_variant[b++] = *q - '0';
    for (; b > 0 && b < 16; ++b) lw_variant[b] = lw_variant[b - 1];
}

// fixed-order sum of the band partials -> caller arrays; band OLR (:382-385, rad.F90:586-605)
__global__ void lw_reduce_kernel(int ld, int col0, const int *__restrict__ perm, int nc, int nlay, int dudTs, const double *__restrict__ part,
                                 double *__restrict__ uflx, double *__restrict__ dflx,
                                 double *__restrict__ uflxc, double *__restrict__ dflxc,
                                 double *__restrict__ duflx, double *__restrict__ duflxc, int band_mask,
                                 double *__restrict__ olrb, double *__restrict__ dolrb) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int lev = blockIdx.y;
    if (c >= nc) return;
    const size_t fstride = (size_t)(nlay + 1) * nc;
    const size_t o = (size_t)lev * nc + c;
    double s[LP_COUNT] = {0., 0., 0., 0., 0., 0.};
    const size_t col = gcol(col0, perm, c);
    const bool top = lev == nlay;
    for (int b = 0; b < 16; ++b) {
        const double *p = part + (size_t)b * LP_COUNT * fstride + o;
        const double vu = p[LP_U * fstride];
        s[LP_U] = s[LP_U] + vu;
        s[LP_UC] = s[LP_UC] + p[LP_UC * fstride];
        s[LP_D] = s[LP_D] + p[LP_D * fstride];
        s[LP_DC] = s[LP_DC] + p[LP_DC * fstride];
        double vd = 0.;
        if (dudTs) {
            vd = p[LP_DU * fstride];
            s[LP_DU] = s[LP_DU] + vd;
            s[LP_DUC] = s[LP_DUC] + p[LP_DUC * fstride];
        }
        if (top && ((band_mask >> b) & 1)) {
            olrb[b + 16 * col] = vu;
            if (dudTs) dolrb[b + 16 * col] = vd;
        }
    }
    const size_t oo = (size_t)lev * ld + col;
    uflx[oo] = s[LP_U]; dflx[oo] = s[LP_D]; uflxc[oo] = s[LP_UC]; dflxc[oo] = s[LP_DC];
    if (dudTs) { duflx[oo] = s[LP_DU]; duflxc[oo] = s[LP_DUC]; }
}

// ---------------------------------------------------------------------------------------------
// host orchestration of one chunk of columns
// ---------------------------------------------------------------------------------------------
static LwWork lw_carve(Slab &slab, int nc, int nlay) {
    LwWork W;
    W.nc = nc; W.nlay = nlay; W.trap = nullptr;
    const size_t n2 = (size_t)nlay * nc, nw = (size_t)((nlay + 31) / 32);
    W.ncp = (nc + 31) & ~31;
    W.n2p = (size_t)nlay * W.ncp;
    W.idx = slab.take<int>(W.n2p);
    W.fbase = slab.take<double>((size_t)F_COUNT * W.n2p);
    W.n2 = n2;
    W.planklay = slab.take<double>(16 * n2);
    W.planklev = slab.take<double>((size_t)16 * (nlay + 1) * nc);
    W.plankbnd = slab.take<double>((size_t)16 * nc);
    W.dplankbnd = slab.take<double>((size_t)16 * nc);
    W.pwvcm = slab.take<double>(nc);
    W.laytrop = slab.take<int>(nc);
    W.seeds = slab.take<uint32_t>((size_t)4 * nc);
    W.alpha = slab.take<double>(n2);
    W.rcorr = slab.take<double>(n2);
    W.thr = slab.take<long long>(3 * W.n2p);
    W.abscoice = slab.take<double>(16 * n2);
    W.abscoliq = slab.take<double>(16 * n2);
    W.cldtrap = slab.take<unsigned char>(n2);
    W.perm = slab.take<int>(nc);
    W.pflags = slab.take<unsigned char>(nc);
    W.ktop = slab.take<int>(nc);
    W.clear_save = slab.take<int32_t>((size_t)4 * nc);
    W.ptmp_bytes = cloud_partition_tmp_bytes(nc);
    W.ptmp = slab.take<char>(W.ptmp_bytes);
    W.mask = slab.take<uint32_t>(nw * 140 * W.ncp);
    W.cloudy_any = slab.take<uint32_t>(nw * nc);
    W.taucmc = slab.take<double>(W.n2p * 140);
    W.it = slab.take<uint32_t>(W.n2p * 140);
    W.pfs = slab.take<double>(W.n2p * 140);
    W.part = slab.take<double>((size_t)16 * LP_COUNT * (nlay + 1) * nc);
    return W;
}

size_t lw_scratch_bytes(int nc, int nlay, bool debug) {
    Slab s;   // dry run of the carve
    lw_carve(s, nc, nlay);
    size_t b = s.used;
    if (debug) b += 2 * (((size_t)nlay * 140 * nc * 8 + 255) & ~(size_t)255);
    return b + 4096;
}

// what RRTMGX_REUSE_CLOUDS may keep from the previous call of this path (see lw_run_chunk)
static CloudCache g_lw_cloud_cache;
void lw_forget_clouds() { g_lw_cloud_cache = CloudCache(); }

int lw_run_chunk(const RrtmgxLwArgs *a, int col0, int nc, const ChunkId &id, const McicaParams &mp_in, const KissJump *d_jumps,
                 Slab &slab, int *d_err, cudaStream_t stream, cudaStream_t *side, int nside, cudaEvent_t *ev,
                 const RrtmgxTaps *taps, int *d_negpos) {
    const int ld = a->ncol, nlay = a->nlay;
    slab.used = 0;
    LwWork W = lw_carve(slab, nc, nlay);
    W.trap = d_negpos;
    McicaParams mp = mp_in;
    mp.trap = d_negpos;
    const bool want_dbg = taps && (taps->taug || taps->pfracs);
    double *dbg_taug = nullptr, *dbg_pfracs = nullptr;
    if (want_dbg) {
        dbg_taug = slab.take<double>((size_t)nlay * 140 * nc);
        dbg_pfracs = slab.take<double>((size_t)nlay * 140 * nc);
    }
    const int nw = (nlay + 31) / 32;
    const dim3 blk(128), grd((nc + 127) / 128);

    // RRTMGX_REUSE_CLOUDS: the previous call left this chunk's column grouping, McICA mask, cloud optical
    // depths and clear counts in the slab (same carve: same shape, same slab, whole call in one chunk)
    CloudCache &cache = g_lw_cloud_cache;
    // the slab holds the clouds of ONE chunk: the previous run of this path must have been this very chunk
    const bool keep = !taps;
    const bool reuse = (a->flags & RRTMGX_REUSE_CLOUDS) && keep && cache.matches(slab.base, id, nc, nlay);
    const int *perm = nullptr;
    if (reuse) {
        perm = cache.perm ? W.perm : nullptr;
        for (int k = 0; k < 4; ++k)
            cudaMemcpyAsync(a->clearCounts + (size_t)k * ld + col0, W.clear_save + (size_t)k * nc,
                            sizeof(int32_t) * (size_t)nc, cudaMemcpyDeviceToDevice, stream);
    } else {
        cache.valid = false;
        cudaMemsetAsync(W.cloudy_any, 0, sizeof(uint32_t) * (size_t)nw * nc, stream);
        // clearCounts of this chunk: (ld,4) -> four strided segments
        for (int k = 0; k < 4; ++k)
            cudaMemsetAsync(a->clearCounts + (size_t)k * ld + col0, 0, sizeof(int32_t) * (size_t)nc, stream);
        // group cloudy and cloud-free columns (not under debug taps, whose layouts assume identity order)
        if (!taps) {
            if (int rc = build_cloud_partition(ld, col0, nc, nlay, a->cldf, W.perm, W.pflags, W.ktop, W.ptmp, W.ptmp_bytes, stream))
                return rc;
            perm = W.perm;
        }
    }
    RRTMGX_LAUNCH(lw_setcoef_kernel, grd, blk, 0, stream, ld, col0, perm, W, a->dudTs, a->play, a->tlay, a->plev,
                  a->tlev, a->tsfc, a->emis, a->h2ovmr, a->o3vmr, a->co2vmr, a->ch4vmr, a->n2ovmr, a->o2vmr,
                  a->cfc11vmr, a->cfc12vmr, a->cfc22vmr, a->ccl4vmr, d_err);
    if (!reuse) {
        RRTMGX_LAUNCH(mcica_prep_kernel, grd, blk, 0, stream, ld, col0, perm, nc, nlay, mp, a->zm, a->play, a->alat,
                      perm ? W.ktop : nullptr, W.seeds, W.alpha, W.rcorr);
        RRTMGX_LAUNCH(mcica_threshold_kernel, dim3(grd.x, nlay), blk, 0, stream, ld, col0, perm, nc, nlay, mp.inhomo,
                      W.alpha, W.rcorr, a->cldf, perm ? W.ktop : nullptr, W.thr);
        RRTMGX_LAUNCH(lw_cldcoef_kernel, dim3(grd.x, nlay), blk, 0, stream, ld, col0, perm, nc, nlay, a->iceflglw, a->cldf,
                      a->rei, a->rel, W.abscoice, W.abscoliq, W.cldtrap);
        LwOptics opt{nc, nlay, W.ncp, W.abscoice, W.abscoliq, W.cldtrap, W.taucmc};
        RRTMGX_LAUNCH(mcica_kernel<LwOptics>, dim3(140 / MCICA_XS, (nc + MCICA_YC - 1) / MCICA_YC),
                      dim3(MCICA_XS, MCICA_YC), 0, stream, ld, col0, perm, nc, nlay, 140, mp, d_jumps, W.seeds, W.thr, a->cldf, a->ciwp, a->clwp, 1.e-20, a->cloudLM, a->cloudMH,
                      perm ? (const int *)W.ptmp : nullptr, perm ? W.ktop : nullptr, a->clearCounts, W.cloudy_any, W.mask,
                      opt, d_err);
        if (keep) {
            for (int k = 0; k < 4; ++k)
                cudaMemcpyAsync(W.clear_save + (size_t)k * nc, a->clearCounts + (size_t)k * ld + col0,
                                sizeof(int32_t) * (size_t)nc, cudaMemcpyDeviceToDevice, stream);
            cache = {slab.base, id.first, id.total, nc, nlay, perm != nullptr, true};
        }
    }

    LwBandArgs A{ld, col0, perm, W, a->dudTs, a->play, a->emis, a->tauaer, dbg_taug, dbg_pfracs};
    // fan the independent band units out over the side streams
    cudaEventRecord(ev[0], stream);
    for (int s = 0; s < nside; ++s) cudaStreamWaitEvent(side[s], ev[0], 0);
    for (int b = 0; b < 16; ++b) lw_launchers[b][lw_variant[b]](nc, nside ? side[b % nside] : stream, A);
    for (int s = 0; s < nside; ++s) {
        cudaEventRecord(ev[1 + s], side[s]);
        cudaStreamWaitEvent(stream, ev[1 + s], 0);
    }
    int band_mask = 0;
    for (int b = 0; b < 16; ++b)
        if (a->band_output && a->band_output[b]) band_mask |= 1 << b;
    RRTMGX_LAUNCH(lw_reduce_kernel, dim3(grd.x, nlay + 1), blk, 0, stream, ld, col0, perm, nc, nlay, a->dudTs, W.part,
                  a->uflx, a->dflx, a->uflxc, a->dflxc, a->duflx_dTs, a->duflxc_dTs, band_mask, a->olrb,
                  a->dolrb_dTs);

    if (taps) {   // debug / parity taps: synchronous strided copies into the host arrays
        if (cudaStreamSynchronize(stream) != cudaSuccess) return RRTMGX_ECUDA;
        std::vector<int> hidx(W.n2p);
        auto copy2d = [&](void *dst_host, const void *src_dev, size_t elem, size_t rows) {
            // chunk-local [rows][nc] -> host [rows][ld] at column col0
            cudaMemcpy2D((char *)dst_host + (size_t)col0 * elem, (size_t)ld * elem, src_dev, (size_t)nc * elem,
                         (size_t)nc * elem, rows, cudaMemcpyDeviceToHost);
        };
        if (taps->jp || taps->jt || taps->jt1 || taps->indfor || taps->indself || taps->indminor) {
            cudaMemcpy(hidx.data(), W.idx, W.n2p * sizeof(int), cudaMemcpyDeviceToHost);
            std::vector<int> hl(nc);
            cudaMemcpy(hl.data(), W.laytrop, nc * sizeof(int), cudaMemcpyDeviceToHost);
            for (int lay = 0; lay < nlay; ++lay)
                for (int c = 0; c < nc; ++c) {
                    const int pk = hidx[W.ti(lay, c)];
                    const size_t o = (size_t)lay * ld + col0 + c;
                    const bool lower = lay < hl[c];
                    if (taps->jp) taps->jp[o] = pk & 63;
                    if (taps->jt) taps->jt[o] = (pk >> 6) & 7;
                    if (taps->jt1) taps->jt1[o] = (pk >> 9) & 7;
                    if (taps->indfor) taps->indfor[o] = (pk >> 12) & 3;
                    // the reference leaves indself unset above the tropopause (calloc'd 0 in the oracle)
                    if (taps->indself) taps->indself[o] = lower ? (pk >> 14) & 15 : 0;
                    if (taps->indminor) taps->indminor[o] = (pk >> 18) & 31;
                }
        }
        if (taps->laytrop) cudaMemcpy(taps->laytrop + col0, W.laytrop, nc * sizeof(int), cudaMemcpyDeviceToHost);
        if (taps->pwvcm) cudaMemcpy(taps->pwvcm + col0, W.pwvcm, nc * sizeof(double), cudaMemcpyDeviceToHost);
        if (taps->fac00 || taps->fac01 || taps->fac10 || taps->fac11) {   // de-tile on the host
            std::vector<double> hf((size_t)F_COUNT * W.n2p);
            cudaMemcpy(hf.data(), W.fbase, hf.size() * 8, cudaMemcpyDeviceToHost);
            double *dst[4] = {taps->fac00, taps->fac01, taps->fac10, taps->fac11};
            const int plane[4] = {F_FAC00, F_FAC01, F_FAC10, F_FAC11};
            for (int q = 0; q < 4; ++q)
                if (dst[q])
                    for (int lay = 0; lay < nlay; ++lay)
                        for (int c = 0; c < nc; ++c)
                            dst[q][(size_t)lay * ld + col0 + c] = hf[W.tf(lay, c) + (size_t)plane[q] * 32];
        }
        if (taps->taug) copy2d(taps->taug, dbg_taug, 8, (size_t)nlay * 140);
        if (taps->pfracs) copy2d(taps->pfracs, dbg_pfracs, 8, (size_t)nlay * 140);
        if (taps->cldymc || taps->taucmc) {
            std::vector<uint32_t> hm((size_t)nw * 140 * W.ncp);
            std::vector<double> ht;
            cudaMemcpy(hm.data(), W.mask, hm.size() * 4, cudaMemcpyDeviceToHost);
            if (taps->taucmc) {
                ht.resize(W.n2p * 140);
                cudaMemcpy(ht.data(), W.taucmc, ht.size() * 8, cudaMemcpyDeviceToHost);
            }
            for (int lay = 0; lay < nlay; ++lay)
                for (int g = 0; g < 140; ++g)
                    for (int c = 0; c < nc; ++c) {
                        const int ib = g_lw_ngb[g] - 1, first = ib ? g_lw_ngs[ib - 1] : 0, ngb = g_lw_ngs[ib] - first;
                        auto cell = [&](int rows, int row) { return lw_cell(first, ngb, rows, W.ncp, row, c, g - first); };
                        const bool on = (hm[cell(nw, lay >> 5)] >> (lay & 31)) & 1u;
                        const size_t o = ((size_t)lay * 140 + g) * ld + col0 + c;
                        if (taps->cldymc) taps->cldymc[o] = on;
                        if (taps->taucmc) taps->taucmc[o] = on ? ht[cell(nlay, lay)] : 0.;
                    }
        }
        if (cudaGetLastError() != cudaSuccess) return RRTMGX_ECUDA;
    }
    return cudaGetLastError() == cudaSuccess ? 0 : RRTMGX_ECUDA;
}

// ---- test hook: the device KISS generator on its own (include/rrtmgx.h rrtmgx_debug_kiss) -----------------------
// One thread per stream: the first ndraw draws of SH/cloud_subcol_gen.F90:568-575, `ran_num` in real*8 (the
// promoted-real contract) and in real*4 (the production kind: int -> real*4 conversion, real*4 product and sum).
static __global__ void debug_kiss_draw_kernel(int nstream, const int32_t *__restrict__ seeds, int ndraw,
                                              int32_t *__restrict__ kiss, double *__restrict__ ran8, float *__restrict__ ran4) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nstream) return;
    Kiss k{(uint32_t)seeds[4 * i], (uint32_t)seeds[4 * i + 1], (uint32_t)seeds[4 * i + 2], (uint32_t)seeds[4 * i + 3]};
    for (int d = 0; d < ndraw; ++d) {
        const int32_t v = k.draw_int();
        const size_t o = (size_t)i * ndraw + d;
        kiss[o] = v;
        ran8[o] = kiss_value(v);
        ran4[o] = (float)v * 2.328306e-10f + 0.5f;
    }
}
// One thread per (jump entry, stream): the state Kiss::jump reaches beside the state after replaying J.n draws.
static __global__ void debug_kiss_jump_kernel(int nstream, const int32_t *__restrict__ seeds, int nentry,
                                              const KissJump *__restrict__ J, uint32_t *__restrict__ jumped,
                                              uint32_t *__restrict__ replayed) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nstream * nentry) return;
    const int i = t / nentry, e = t % nentry;
    Kiss a{(uint32_t)seeds[4 * i], (uint32_t)seeds[4 * i + 1], (uint32_t)seeds[4 * i + 2], (uint32_t)seeds[4 * i + 3]};
    Kiss b = a;
    a.jump(J[e]);
    for (uint32_t d = 0; d < J[e].n; ++d) (void)b.draw_int();
    const size_t o = 4 * (size_t)t;
    jumped[o] = a.s1; jumped[o + 1] = a.s2; jumped[o + 2] = a.s3; jumped[o + 3] = a.s4;
    replayed[o] = b.s1; replayed[o + 1] = b.s2; replayed[o + 2] = b.s3; replayed[o + 3] = b.s4;
}
static __global__ void debug_kiss_value_kernel(int n, const int32_t *__restrict__ k, double *__restrict__ r8, float *__restrict__ r4) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    r8[i] = kiss_value(k[i]);
    r4[i] = (float)k[i] * 2.328306e-10f + 0.5f;
}

int debug_kiss(int nstream, const int32_t *seeds, int ndraw, int32_t *kiss, double *ran8, float *ran4, int nsub, int nlay,
               int inhomo, uint32_t *jumped, uint32_t *replayed, int nvalue, const int32_t *values, double *val8, float *val4,
               cudaStream_t st) {
    const size_t nd = (size_t)nstream * ndraw, nentry = 2 * (size_t)nsub, nj = 4 * nentry * nstream;
    char *d = nullptr;
    const size_t bytes = 16 * (size_t)nstream + nd * 16 + nentry * sizeof(KissJump) + nj * 8 + (size_t)nvalue * 16 + 4096;
    if (cudaMalloc((void **)&d, bytes) != cudaSuccess) { cudaGetLastError(); return RRTMGX_ECUDA; }
    Slab s;
    s.base = d; s.cap = bytes;
    int32_t *d_seed = s.take<int32_t>(4 * (size_t)nstream);
    double *d_r8 = s.take<double>(nd + 1);
    int32_t *d_k = s.take<int32_t>(nd + 1);
    float *d_r4 = s.take<float>(nd + 1);
    KissJump *d_J = s.take<KissJump>(nentry + 1);
    uint32_t *d_a = s.take<uint32_t>(nj + 1), *d_b = s.take<uint32_t>(nj + 1);
    double *d_v8 = s.take<double>((size_t)nvalue + 1);
    int32_t *d_v = s.take<int32_t>((size_t)nvalue + 1);
    float *d_v4 = s.take<float>((size_t)nvalue + 1);
    cudaMemcpyAsync(d_seed, seeds, 16 * (size_t)nstream, cudaMemcpyHostToDevice, st);
    if (ndraw > 0) {
        debug_kiss_draw_kernel<<<(nstream + 63) / 64, 64, 0, st>>>(nstream, d_seed, ndraw, d_k, d_r8, d_r4);
        cudaMemcpyAsync(kiss, d_k, nd * 4, cudaMemcpyDeviceToHost, st);
        cudaMemcpyAsync(ran8, d_r8, nd * 8, cudaMemcpyDeviceToHost, st);
        cudaMemcpyAsync(ran4, d_r4, nd * 4, cudaMemcpyDeviceToHost, st);
    }
    if (nsub > 0) {
        std::vector<KissJump> h(nentry);
        kiss_jump_table(nsub, nlay, inhomo != 0, h.data());
        cudaMemcpyAsync(d_J, h.data(), nentry * sizeof(KissJump), cudaMemcpyHostToDevice, st);
        const int nt = nstream * (int)nentry;
        debug_kiss_jump_kernel<<<(nt + 63) / 64, 64, 0, st>>>(nstream, d_seed, (int)nentry, d_J, d_a, d_b);
        cudaMemcpyAsync(jumped, d_a, nj * 4, cudaMemcpyDeviceToHost, st);
        cudaMemcpyAsync(replayed, d_b, nj * 4, cudaMemcpyDeviceToHost, st);
        cudaStreamSynchronize(st);   // `h` is pageable host memory
    }
    if (nvalue > 0) {
        cudaMemcpyAsync(d_v, values, (size_t)nvalue * 4, cudaMemcpyHostToDevice, st);
        debug_kiss_value_kernel<<<(nvalue + 127) / 128, 128, 0, st>>>(nvalue, d_v, d_v8, d_v4);
        cudaMemcpyAsync(val8, d_v8, (size_t)nvalue * 8, cudaMemcpyDeviceToHost, st);
        cudaMemcpyAsync(val4, d_v4, (size_t)nvalue * 4, cudaMemcpyDeviceToHost, st);
    }
    const cudaError_t e = cudaStreamSynchronize(st);
    cudaFree(d);
    return (e == cudaSuccess && cudaGetLastError() == cudaSuccess) ? 0 : RRTMGX_ECUDA;
}


}  // namespace rrtmgx
