// Host-side table initialisation (see tables.h).  Follows
//   LW/src/rrtmg_lw_init.F90  lookup tables :96-113, rwgt :120-144, cmbgb1..16 :329-1978
//   SW/src/rrtmg_sw_init.F90  rwgt :128-152, cmbgb16s..29 :463-1660
//   SW/src/rrtmg_sw_k_g_29.F90:80-81 (irradnce scaling of band 29)
// The sixteen + fourteen cmbgbN routines are instances of one rule: absorption-like data are
// summed over each group of original g-points weighted by rwgt, in ascending original g-point
// order; Planck fractions and solar source terms are summed unweighted.
#include "tables.h"

#include <cmath>
#include <cstdio>
#include <cstring>

#include "../../include/rrtmgx.h"

namespace rrtmgx {

namespace {

struct BlobEntry {
    std::string name;
    int dtype = 0, ndim = 0, dims[6] = {0, 0, 0, 0, 0, 0};
    long long offset = 0, nbytes = 0;
};

struct Blob {
    std::vector<unsigned char> buf;
    std::map<std::string, BlobEntry> ent;
    int load(const std::string &path) {
        FILE *f = std::fopen(path.c_str(), "rb");
        if (!f) return -1;
        std::fseek(f, 0, SEEK_END);
        long sz = std::ftell(f);
        std::fseek(f, 0, SEEK_SET);
        buf.resize((size_t)sz);
        size_t got = std::fread(buf.data(), 1, (size_t)sz, f);
        std::fclose(f);
        if (got != (size_t)sz || sz < 12 || std::memcmp(buf.data(), "RRTMGTB1", 8) != 0) return -2;
        int n = 0;
        std::memcpy(&n, buf.data() + 8, 4);
        size_t pos = 12;
        for (int i = 0; i < n; ++i) {
            if (pos + 96 > buf.size()) return -3;
            BlobEntry e;
            char nm[49];
            std::memcpy(nm, buf.data() + pos, 48);
            nm[48] = 0;
            e.name = nm;
            std::memcpy(&e.dtype, buf.data() + pos + 48, 4);
            std::memcpy(&e.ndim, buf.data() + pos + 52, 4);
            std::memcpy(e.dims, buf.data() + pos + 56, 24);
            std::memcpy(&e.offset, buf.data() + pos + 80, 8);
            std::memcpy(&e.nbytes, buf.data() + pos + 88, 8);
            if (e.offset < 0 || (size_t)(e.offset + e.nbytes) > buf.size()) return -4;
            ent[e.name] = e;
            pos += 96;
        }
        return 0;
    }
    const BlobEntry *get(const std::string &n) const {
        auto it = ent.find(n);
        return it == ent.end() ? nullptr : &it->second;
    }
    const double *f64(const BlobEntry *e) const { return (const double *)(buf.data() + e->offset); }
    const int *i32(const BlobEntry *e) const { return (const int *)(buf.data() + e->offset); }
    bool ints(const std::string &n, int *dst, int cnt) const {
        const BlobEntry *e = get(n);
        if (!e || e->dtype != 1 || e->nbytes != (long long)cnt * 4) return false;
        std::memcpy(dst, buf.data() + e->offset, (size_t)cnt * 4);
        return true;
    }
    std::vector<double> vec(const std::string &n) const {
        const BlobEntry *e = get(n);
        if (!e || e->dtype != 0) return {};
        const double *p = f64(e);
        return std::vector<double>(p, p + e->nbytes / 8);
    }
};

// relative weights of the original 16 g-points inside their reduced group
void relative_weights(int nbnd, const int *ngc, const int *ngn, const int *ngm, const double *wt,
                      std::vector<double> &rwgt) {
    rwgt.assign((size_t)nbnd * 16, 0.0);
    int igcsm = 0;
    for (int ibnd = 0; ibnd < nbnd; ++ibnd) {
        if (ngc[ibnd] < 16) {
            double wtsm[16];
            int iprsm = 0;
            for (int igc = 0; igc < ngc[ibnd]; ++igc) {
                double wtsum = 0.0;
                for (int ipr = 0; ipr < ngn[igcsm]; ++ipr) wtsum = wtsum + wt[iprsm++];
                wtsm[igc] = wtsum;
                ++igcsm;
            }
            for (int ig = 0; ig < 16; ++ig)
                rwgt[ibnd * 16 + ig] = wt[ig] / wtsm[ngm[ibnd * 16 + ig] - 1];
        } else {
            for (int ig = 0; ig < 16; ++ig) rwgt[ibnd * 16 + ig] = 1.0;
            igcsm += 16;
        }
    }
}

// Reduce one original table to `ngc` g-points, output [lead][ngc] (g fastest).
// g_first: source is (16, lead) (Planck fractions / solar source / rayla); else (lead, 16).
std::vector<double> reduce(const double *src, int lead, bool g_first, bool weighted, int ngc,
                           const int *ngn_band, const double *rwgt_band) {
    std::vector<double> dst((size_t)lead * ngc);
    for (int l = 0; l < lead; ++l) {
        int iprsm = 0;
        for (int igc = 0; igc < ngc; ++igc) {
            double sum = 0.0;
            for (int ipr = 0; ipr < ngn_band[igc]; ++ipr, ++iprsm) {
                double v = g_first ? src[iprsm + 16 * l] : src[l + (size_t)lead * iprsm];
                sum = weighted ? sum + v * rwgt_band[iprsm] : sum + v;
            }
            dst[(size_t)l * ngc + igc] = sum;
        }
    }
    return dst;
}

bool unweighted_name(const std::string &n) {
    return n == "fracrefao" || n == "fracrefbo" || n == "sfluxrefo" || n == "irradnceo" ||
           n == "facbrghto" || n == "snsptdrko";
}

std::string reduced_name(const std::string &orig) {
    // kao -> absa, kbo -> absb (the reference's equivalence(ka,absa) flattened view);
    // kao_mX -> ka_mX; every other "<name>o" -> "<name>"
    if (orig == "kao") return "absa";
    if (orig == "kbo") return "absb";
    if (orig.compare(0, 4, "kao_") == 0) return "ka_" + orig.substr(4);
    if (orig.compare(0, 4, "kbo_") == 0) return "kb_" + orig.substr(4);
    if (!orig.empty() && orig.back() == 'o') return orig.substr(0, orig.size() - 1);
    return orig;
}

}  // namespace

TableRef HostTables::add(const std::string &name, const std::vector<double> &v) {
    TableRef r;
    r.off = arena.size();
    r.n = v.size();
    arena.insert(arena.end(), v.begin(), v.end());
    while (arena.size() % 2) arena.push_back(0.0);  // keep every table 16-byte aligned
    index[name] = r;
    return r;
}

TableRef HostTables::find(const std::string &name) const {
    auto it = index.find(name);
    return it == index.end() ? TableRef() : it->second;
}

int HostTables::load(const std::string &blob_path) {
    Blob b;
    if (b.load(blob_path) != 0) return RRTMGX_EBLOB;
    arena.clear();
    index.clear();
    arena.reserve(1 << 20);

    for (int pass = 0; pass < 2; ++pass) {
        const bool lw = pass == 0;
        const int nbnd = lw ? 16 : 14, ngpt = lw ? 140 : 112, band0 = lw ? 1 : 16;
        const std::string p = lw ? "lw" : "sw";
        int *ngc = lw ? lw_ngc : sw_ngc, *ngs = lw ? lw_ngs : sw_ngs, *ngb = lw ? lw_ngb : sw_ngb;
        int *nspa = lw ? lw_nspa : sw_nspa, *nspb = lw ? lw_nspb : sw_nspb;
        std::vector<int> ngn(ngpt), ngm(nbnd * 16);
        if (!b.ints(p + ".wvn.ngc", ngc, nbnd) || !b.ints(p + ".wvn.ngs", ngs, nbnd) ||
            !b.ints(p + ".wvn.ngb", ngb, ngpt) || !b.ints(p + ".wvn.nspa", nspa, nbnd) ||
            !b.ints(p + ".wvn.nspb", nspb, nbnd) || !b.ints(p + ".wvn.ngn", ngn.data(), ngpt) ||
            !b.ints(p + ".wvn.ngm", ngm.data(), nbnd * 16))
            return RRTMGX_EBLOB;
        if (!lw && !b.ints("sw.wvn.icxa", sw_icxa, 14)) return RRTMGX_EBLOB;
        std::vector<double> wt = b.vec(p + ".wvn.wt");
        if (wt.size() != 16) return RRTMGX_EBLOB;
        std::vector<double> rwgt;
        relative_weights(nbnd, ngc, ngn.data(), ngm.data(), wt.data(), rwgt);
        add(p + ".rwgt", rwgt);

        for (int ib = 0; ib < nbnd; ++ib) {
            const int band = band0 + ib;
            char pre[32];
            std::snprintf(pre, sizeof pre, lw ? "lw.kg%02d." : "sw.kg%d.", band);
            char outpre[32];
            std::snprintf(outpre, sizeof outpre, "%s.%02d.", p.c_str(), band);
            const int g0 = ib == 0 ? 0 : ngs[ib - 1];
            if (!lw) { sw_nfor[ib] = 0; sw_nsrc[ib] = 1; sw_rayl_scalar[ib] = 0.0; sw_has_raylv[ib] = false; }
            for (const auto &kv : b.ent) {
                if (kv.first.compare(0, std::strlen(pre), pre) != 0) continue;
                const BlobEntry &e = kv.second;
                const std::string orig = kv.first.substr(std::strlen(pre));
                const double *src = b.f64(&e);
                if (orig == "rayl") {  // scalar Rayleigh coefficient, not reduced
                    sw_rayl_scalar[ib] = src[0];
                    continue;
                }
                const long long total = e.nbytes / 8;
                const bool g_first = e.ndim == 2 && e.dims[0] == 16 && e.dims[1] != 16;
                if (!g_first && e.dims[e.ndim - 1] != 16) return RRTMGX_EBLOB;
                const int lead = (int)(total / 16);
                std::vector<double> tmp;
                if (!lw && band == 29 && orig == "irradnceo") {
                    // SW/src/rrtmg_sw_k_g_29.F90:80-81
                    const double irradscl = 13.221 / (13.221 - 0.455);
                    tmp.assign(src, src + 16);
                    for (double &v : tmp) v = irradscl * v;
                    src = tmp.data();
                }
                std::vector<double> red = reduce(src, lead, g_first, !unweighted_name(orig), ngc[ib],
                                                 ngn.data() + g0, rwgt.data() + 16 * ib);
                add(std::string(outpre) + reduced_name(orig), red);
                if (!lw && orig == "forrefo") sw_nfor[ib] = lead;
                if (!lw && orig == "sfluxrefo") sw_nsrc[ib] = lead;
                if (!lw && orig == "raylo") sw_has_raylv[ib] = true;
            }
        }
    }

    // SW: row pairs {t[row][g], t[row+1][g]} of the tables the band kernels interpolate between neighbouring rows
    // (absa/absb along the binary-species and temperature index, selfref/forref, rayla): a thread then reads both
    // operands of a lerp with one 16-byte load, which halves the gather instructions and the L1 tag look-ups of
    // the SW band kernels (their lanes are 32 columns on 32 different rows).  The last row is paired with itself.
    for (int ib = 0; ib < 14; ++ib) {
        const int ng = sw_ngc[ib];
        for (const char *nm : {"absa", "absb", "selfref", "forref", "rayla"}) {
            char name[48];
            std::snprintf(name, sizeof name, "sw.%02d.%s", ib + 16, nm);
            const TableRef r = find(name);
            if (!r.ok() || r.n % ng) continue;
            const size_t rows = r.n / ng;
            std::vector<double> pr(2 * r.n);
            for (size_t row = 0; row < rows; ++row)
                for (int g = 0; g < ng; ++g) {
                    const size_t nxt = row + 1 < rows ? row + 1 : row;
                    pr[2 * (row * ng + g)] = arena[r.off + row * ng + g];
                    pr[2 * (row * ng + g) + 1] = arena[r.off + nxt * ng + g];
                }
            add(std::string(name) + "2", pr);
        }
    }

    // LW lookup tables, rrtmg_lw_init.F90:96-113
    {
        const int ntbl = 10000;
        const double pade = 0.278, expeps = 1.e-20;
        const double bpade = 1.0 / pade;
        std::vector<double> tau(ntbl + 1), ex(ntbl + 1), tfn(ntbl + 1);
        tau[0] = 0.0; tau[ntbl] = 1.e10;
        ex[0] = 1.0; ex[ntbl] = expeps;
        tfn[0] = 0.0; tfn[ntbl] = 1.0;
        for (int itr = 1; itr <= ntbl - 1; ++itr) {
            double f = (double)itr / (double)ntbl;
            tau[itr] = bpade * f / (1. - f);
            ex[itr] = std::exp(-tau[itr]);
            if (ex[itr] <= expeps) ex[itr] = expeps;
            if (tau[itr] < 0.06)
                tfn[itr] = tau[itr] / 6.;
            else
                tfn[itr] = 1. - 2. * ((1. / tau[itr]) - (ex[itr] / (1. - ex[itr])));
        }
        add("lw.tau_tbl", tau);
        add("lw.exp_tbl", ex);
        add("lw.tfn_tbl", tfn);
        // device view: {exp_tbl, tfn_tbl} interleaved so one 16-byte load serves both lookups
        std::vector<double> et(2 * (ntbl + 1));
        for (int i = 0; i <= ntbl; ++i) { et[2 * i] = ex[i]; et[2 * i + 1] = tfn[i]; }
        add("lw.exptfn", et);
    }
    // band widths, LW/modules/rrlw_wvn.F90 via lwcmbdat: delwave = wavenum2 - wavenum1
    {
        static const double w1[16] = {10., 350., 500., 630., 700., 820., 980., 1080.,
                                      1180., 1390., 1480., 1800., 2080., 2250., 2380., 2600.};
        static const double w2[16] = {350., 500., 630., 700., 820., 980., 1080., 1180.,
                                      1390., 1480., 1800., 2080., 2250., 2380., 2600., 3250.};
        for (int i = 0; i < 16; ++i) lw_delwave[i] = w2[i] - w1[i];
    }
    // data used as is (reference atmosphere, Planck, cloud optics, McICA, NRLSSI2)
    static const char *raw[] = {
        "lw.ref.pref", "lw.ref.preflog", "lw.ref.tref", "lw.ref.chi_mls", "lw.wvn.totplnk",
        "lw.wvn.totplk16", "lw.wvn.totplnkderiv", "lw.wvn.totplk16deriv", "lw.cld.absice0",
        "lw.cld.absice1", "lw.cld.absice2", "lw.cld.absice3", "lw.cld.absice4", "lw.cld.absliq1",
        "sw.ref.pref", "sw.ref.preflog", "sw.ref.tref", "sw.cld.extliq1", "sw.cld.ssaliq1",
        "sw.cld.asyliq1", "sw.cld.extice2", "sw.cld.ssaice2", "sw.cld.asyice2", "sw.cld.extice3",
        "sw.cld.ssaice3", "sw.cld.asyice3", "sw.cld.fdlice3", "sw.cld.extice4", "sw.cld.ssaice4",
        "sw.cld.asyice4", "sw.cld.abari", "sw.cld.bbari", "sw.cld.cbari", "sw.cld.dbari",
        "sw.cld.ebari", "sw.cld.fbari", "sw.nrlssi2.mgavgcyc", "sw.nrlssi2.sbavgcyc",
        "mcica.xcw_beta", "mcica.xcw_gamma"};
    for (const char *nm : raw) {
        std::vector<double> v = b.vec(nm);
        if (v.empty()) return RRTMGX_EBLOB;
        add(nm, v);
    }
    // species ratios of the reference atmosphere used by setcoef (rrtmg_lw_setcoef.F90:487-541):
    // rat(j) = chi_mls(a,j)/chi_mls(b,j), tabulated once instead of per layer
    {
        std::vector<double> chi = b.vec("lw.ref.chi_mls");  // (7,59)
        static const int pairs[6][2] = {{1, 2}, {1, 3}, {1, 4}, {1, 6}, {4, 2}, {3, 2}};
        std::vector<double> rat(6 * 59);
        for (int r = 0; r < 6; ++r)
            for (int j = 0; j < 59; ++j)
                rat[r * 59 + j] = chi[(pairs[r][0] - 1) + 7 * j] / chi[(pairs[r][1] - 1) + 7 * j];
        add("lw.ref.rat", rat);
    }
    return 0;
}

}  // namespace rrtmgx
