// Host side of the McICA KISS jump-ahead (see mcica.cuh): for a given number of draws n, the
// state of SH/cloud_subcol_gen.F90:568-575 after n calls is
//   s1: affine map (a,c)^n over Z/2^32          s2: GF(2)-linear map L^n (32x32 bit matrix)
//   s3, s4: y <- a*y mod (a*2^16-1) iterated, once two explicit steps have made y canonical
// so a (column, subcolumn) thread can start at draw isub*(2 or 4)*nlay without replaying the
// sequence.  Also evaluates the day-of-year dependent correlation-length parameter (:491-516).
#include <cstring>

#include "engine.h"

namespace rrtmgx {

namespace {

uint32_t xs_step(uint32_t s) {
    s ^= s << 13;
    s ^= s >> 17;
    s ^= s << 5;
    return s;
}

struct Gf2 { uint32_t col[32]; };   // col[j] = image of bit j

uint32_t apply(const Gf2 &m, uint32_t v) {
    uint32_t r = 0;
    for (int j = 0; j < 32; ++j)
        if ((v >> j) & 1u) r ^= m.col[j];
    return r;
}

Gf2 mul(const Gf2 &a, const Gf2 &b) {   // a after b
    Gf2 r;
    for (int j = 0; j < 32; ++j) r.col[j] = apply(a, b.col[j]);
    return r;
}

uint64_t powmod(uint64_t base, uint64_t e, uint64_t m) {
    uint64_t r = 1 % m;
    base %= m;
    while (e) {
        if (e & 1) r = (r * base) % m;   // operands < 2^31: no overflow
        base = (base * base) % m;
        e >>= 1;
    }
    return r;
}

KissJump jump_entry(uint64_t n) {
    KissJump J;
    std::memset(&J, 0, sizeof J);
    J.n = (uint32_t)n;
    // LCG: compose (a,c) n times by squaring: (a2,c2)o(a1,c1) = (a2*a1, a2*c1 + c2)
    uint32_t ra = 1, rc = 0, ba = 69069u, bc = 1327217885u;
    for (uint64_t e = n; e; e >>= 1) {
        if (e & 1) { rc = ba * rc + bc; ra = ba * ra; }
        bc = ba * bc + bc;
        ba = ba * ba;
    }
    J.lcg_a = ra;
    J.lcg_c = rc;
    // xorshift
    Gf2 r, b;
    for (int j = 0; j < 32; ++j) { r.col[j] = 1u << j; b.col[j] = xs_step(1u << j); }
    for (uint64_t e = n; e; e >>= 1) {
        if (e & 1) r = mul(b, r);
        b = mul(b, b);
    }
    for (int j = 0; j < 32; ++j) J.xs[j] = r.col[j];
    // multiply-with-carry lanes: a^(n-2) mod (a*2^16 - 1)
    if (n >= 2) {
        J.mwc3 = (uint32_t)powmod(18000ull, n - 2, 18000ull * 65536ull - 1ull);
        J.mwc4 = (uint32_t)powmod(30903ull, n - 2, 30903ull * 65536ull - 1ull);
    }
    return J;
}

}  // namespace

void kiss_jump_table(int nsub, int nlay, bool inhomo, KissJump *out) {
    const uint64_t stride = (uint64_t)(inhomo ? 4 : 2) * (uint64_t)nlay;
    for (int i = 0; i < nsub; ++i) {
        out[2 * i] = jump_entry((uint64_t)i * stride);
        out[2 * i + 1] = jump_entry((uint64_t)i * stride + 2ull * (uint64_t)nlay);
    }
}

McicaParams mcica_params(const McicaConfig &cfg, const double *d_xcw_beta, const double *d_xcw_gamma, int doy,
                         const int seed_order[4]) {
    McicaParams P;
    P.inhomo = cfg.ih > 0;
    P.xcw = cfg.ih == 1 ? d_xcw_beta : (cfg.ih == 2 ? d_xcw_gamma : nullptr);
    auto am3 = [&](double am30) {   // SH/cloud_subcol_gen.F90:505-509
        if (doy > 181) return -(4. * am30 / 365. * (double)(doy - 272));
        return 4. * am30 / 365. * (double)(doy - 91);
    };
    P.adl_am1 = cfg.corr[0]; P.adl_am2 = cfg.corr[1]; P.adl_am3 = am3(cfg.corr[2]); P.adl_am4 = cfg.corr[3];
    P.rdl_am1 = cfg.corr[4]; P.rdl_am2 = cfg.corr[5]; P.rdl_am3 = am3(cfg.corr[6]); P.rdl_am4 = cfg.corr[7];
    for (int i = 0; i < 4; ++i) P.seed_order[i] = seed_order[i];
    P.trap = nullptr;
    return P;
}

}  // namespace rrtmgx
