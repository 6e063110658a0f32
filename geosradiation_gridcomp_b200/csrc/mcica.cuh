// McICA stochastic subcolumn generator on the device.
//
// Restates GEOS_RadiationShared/cloud_subcol_gen.F90 (generate_stochastic_clouds :132-487,
// rng_kiss :546-607, correlation_length :491-542, clearCounts_threeBand :611-769) and
// zcw_lookup (GEOS_RadiationShared/cloud_condensate_inhomogeneity.F90:86-124).
//
// The reference draws one KISS stream per column: for subcolumn i and layer k the draws are
// [cdf1(k),cdf2(k)] interleaved for k=1..nlay, then (inhomogeneous condensate) [cdf2'(k),cdf3(k)],
// i.e. subcolumn i starts at draw i*4*nlay.  Here one thread owns one (column, subcolumn) and
// jumps its four KISS lanes ahead in O(1) with host-precomputed jump entries (affine power for
// the LCG, GF(2) matrix power for the xorshift, modular power for the two multiply-with-carry
// lanes), so the integer sequence - and hence the cloud mask - is bit-identical to the
// sequential generator.  Layers are then swept once, carrying cdf1/cdf3 in registers.
#pragma once
#include "common.cuh"

namespace rrtmgx {

__device__ __forceinline__ uint32_t mwc_step(uint32_t y, uint32_t a) { return a * (y & 65535u) + (y >> 16); }

// n steps of y <- a*(y & 65535) + (y >> 16); M = a^(n-2) mod m, m = a*2^16 - 1.
// After two explicit steps y is in [0, m+1] and y_{k+1} = a*y_k mod m holds as an equality of
// canonical residues except at the fixed points 0 and m and the transient m+1 (handled).
template <uint32_t A>
__device__ __forceinline__ uint32_t mwc_jump(uint32_t y, uint32_t n, uint32_t M) {
    constexpr uint64_t m = (uint64_t)A * 65536ull - 1ull;
    if (n == 0) return y;
    y = mwc_step(y, A);
    if (n == 1) return y;
    y = mwc_step(y, A);
    if (n == 2) return y;
    if (y == 0u || (uint64_t)y == m) return y;
    uint64_t mult = M;
    if ((uint64_t)y == m + 1ull) {          // one more explicit step, one fewer modular one
        y = mwc_step(y, A);
        mult = (mult * 65536ull) % m;       // times a^-1 = 2^16 (mod m)
    }
    return (uint32_t)((mult * (uint64_t)y) % m);
}

struct Kiss {
    uint32_t s1, s2, s3, s4;
    __device__ __forceinline__ void jump(const KissJump &J) {
        s1 = J.lcg_a * s1 + J.lcg_c;
        uint32_t r = 0;
#pragma unroll 8
        for (int j = 0; j < 32; ++j)
            if ((s2 >> j) & 1u) r ^= J.xs[j];
        s2 = r;
        s3 = mwc_jump<18000u>(s3, J.n, J.mwc3);
        s4 = mwc_jump<30903u>(s4, J.n, J.mwc4);
    }
    // SH/cloud_subcol_gen.F90:568-574 (int32 wraparound, logical shifts): the integer `kiss`
    __device__ __forceinline__ int32_t draw_int() {
        s1 = 69069u * s1 + 1327217885u;
        s2 = s2 ^ (s2 << 13);
        s2 = s2 ^ (s2 >> 17);
        s2 = s2 ^ (s2 << 5);
        s3 = 18000u * (s3 & 65535u) + (s3 >> 16);
        s4 = 30903u * (s4 & 65535u) + (s4 >> 16);
        return (int32_t)(s1 + s2 + (s3 << 16) + s4);
    }
};

// ran_num of an integer draw, SH/cloud_subcol_gen.F90:575 (the literal is not 2^-32)
__device__ __forceinline__ double kiss_value(int32_t kiss) { return (double)kiss * 2.328306e-10 + 0.5; }

// kiss_value is monotone non-decreasing in the integer (exact conversion, one rounded multiply by
// a positive constant, one rounded add), so every comparison the generator makes between a random
// number and a real threshold t is equivalent to an integer comparison with
//   K(t) = min{ k in [-2^31, 2^31) : kiss_value(k) >= t }   (2^31 when no k qualifies):
//   ran < t  <=>  kiss < K(t),      ran >= t  <=>  kiss >= K(t).
// K is found with the same floating-point operations, so the cloud masks stay bit-identical.
__device__ __forceinline__ long long kiss_threshold(double t) {
    long long lo = -2147483648LL, hi = 2147483648LL;
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;   // floor
        if (kiss_value((int32_t)mid) >= t) hi = mid; else lo = mid + 1;
    }
    return lo;
}

// Per-(layer, column) scratch that one column's sweep walks layer by layer is tiled by 32 columns,
// [tile][lay][plane][32]: consecutive layers of a column are a few hundred bytes apart instead of nc*8.
__host__ __device__ __forceinline__ size_t tile_index(int nlay, int planes, int lay, int c) {
    return (((size_t)(c >> 5) * nlay + lay) * planes) * 32 + (c & 31);
}

// integer thresholds of the three comparisons of the layer sweep, one thread per (layer, column):
// thr = [tile][lay][alpha, rcorr, cld][32]
static __global__ void mcica_threshold_kernel(int ld, int col0, const int *__restrict__ perm, int nc, int nlay, int inhomo,
                                              const double *__restrict__ alpha, const double *__restrict__ rcorr,
                                              const double *__restrict__ cldf, const int *__restrict__ ktop,
                                              long long *__restrict__ thr) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int k = blockIdx.y;
    if (c >= nc) return;
    if (ktop && k > ktop[perm[c]]) return;   // the sweep of this column stops below this layer (or never starts)
    const size_t j = (size_t)k * nc + c;
    long long *t = thr + tile_index(nlay, 3, k, c);
    if (k > 0) {
        t[0] = kiss_threshold(alpha[j]);                 // cdf2 < alpha(k), :411
        if (inhomo) t[32] = kiss_threshold(rcorr[j]);    // cdf2 < rcorr(k), :424
    }
    t[64] = kiss_threshold(1. - cldf[(size_t)k * ld + gcol(col0, perm, c)]);   // cdf1 >= 1 - cldfrac, :435
}

// SH/cloud_condensate_inhomogeneity.F90:86-124
__device__ __forceinline__ double zcw_lookup(const double *__restrict__ xcw, double cdf, double sigma_qcw) {
    const int n1 = 1000, n2 = 140;
    double rind1 = cdf * (double)(n1 - 1) + 1.;
    int ind1 = clampi(f_int(rind1), 1, n1 - 1);
    rind1 = rind1 - (double)ind1;
    double rind2 = 40. * sigma_qcw - 3.;
    int ind2 = clampi(f_int(rind2), 1, n2 - 1);
    rind2 = rind2 - (double)ind2;
    const double *c0 = xcw + (size_t)1000 * (ind2 - 1) + (ind1 - 1);
    const double *c1 = c0 + 1000;
    return (1.0 - rind1) * (1.0 - rind2) * __ldg(c0) + (1.0 - rind1) * rind2 * __ldg(c1) +
           rind1 * (1.0 - rind2) * __ldg(c0 + 1) + rind1 * rind2 * __ldg(c1 + 1);
}

// Per-column preparation shared by all subcolumns: KISS seeds from the four lowest layer
// pressures (:375-400) and the inter-layer overlap / condensate correlations (:314-321).
// Inputs are the caller's arrays (leading dimension ld, first column col0); outputs are
// chunk-local [..][nc].
static __global__ void mcica_prep_kernel(int ld, int col0, const int *__restrict__ perm, int nc, int nlay, McicaParams P,
                                  const double *__restrict__ zm, const double *__restrict__ play,
                                  const double *__restrict__ alat,
                                  const int *__restrict__ ktop,   // by caller-order column, -1: no cloud (null: unknown)
                                  uint32_t *__restrict__ seeds,   // [4][nc]
                                  double *__restrict__ alpha,     // [nlay][nc], k >= 1
                                  double *__restrict__ rcorr) {   // [nlay][nc], k >= 1
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nc || RRTMGX_TRAPPED(P.trap)) return;
    if (ktop && ktop[perm[c]] < 0) return;   // a cloud-free column draws nothing
    const size_t col = gcol(col0, perm, c);
    const double r2d = 180.0 / 3.14159265358979323846;
    double d = alat[col] * r2d - P.adl_am3;
    double adl = (P.adl_am1 + P.adl_am2 * exp(-((d * d) / (P.adl_am4 * P.adl_am4)))) * 1.e3;
    double rdl = 1.0;
    if (P.inhomo) {
        d = alat[col] * r2d - P.rdl_am3;
        rdl = (P.rdl_am1 + P.rdl_am2 * exp(-((d * d) / (P.rdl_am4 * P.rdl_am4)))) * 1.e3;
    }
    double zprev = zm[col];
    for (int k = 1; k < nlay; ++k) {
        double z = zm[(size_t)k * ld + col];
        double dz = fabs(z - zprev);
        alpha[(size_t)k * nc + c] = exp(-dz / adl);
        if (P.inhomo) rcorr[(size_t)k * nc + c] = exp(-dz / rdl);
        zprev = z;
    }
    const int maximo = 2147483647 - 1;
    // the reference tests play(1,1) > play(nlay,1) on the first column of a partition (:267);
    // every column shares the ordering, so each column tests itself
    const bool surface_at_one = play[col] > play[(size_t)(nlay - 1) * ld + col];
    double pseed[4];
#pragma unroll
    for (int n = 0; n < 4; ++n)
        pseed[n] = play[(size_t)(surface_at_one ? n : nlay - 1 - n) * ld + col] * 100.;
#pragma unroll
    for (int n = 0; n < 4; ++n) {
        double p = pseed[0];
        int so = P.seed_order[n];
        if (so == 2) p = pseed[1];
        if (so == 3) p = pseed[2];
        if (so == 4) p = pseed[3];
        seeds[(size_t)n * nc + c] = (uint32_t)f_int((p - (double)f_int(p)) * (double)maximo + 1.);
    }
}

// One thread per (column, subcolumn).  Optics::cell(lay, isub, c, ciwp, clwp, err, state) turns
// the stochastic water paths of a McICA-cloudy cell into cloud optical properties, stores them,
// and returns whether the cell goes into the cloud mask; Optics::State is per-thread scratch
// carried along the layer sweep and handed to Optics::finish(isub, c, state) at the end.  Outputs: clearCounts (caller layout
// (ncol,4), integer atomics, so deterministic), the optical cloud mask bit-packed over layers
// (laid out by Optics::mask_index) and its OR over subcolumns cloudy_any [nw][nc] (the reference's
// cloudy(lay,col) after cldprmc).
// Block = MCICA_XS subcolumns (x, fastest) x MCICA_YC columns: a warp is (almost) one column at 28 of its
// subcolumns.  The subcolumns of a column share cldfrac, so at a given layer they are cloudy or clear
// TOGETHER far more often than 32 different columns are: the cloud-optics branch runs with most lanes
// active instead of a few, and the per-(layer, column) thresholds and water paths are one broadcast
// load per warp.  (The cloudy-cell stores then fall 8 bytes per sector; the L2 merges them with the
// neighbouring columns' before they reach DRAM.)
constexpr int MCICA_XS = 28;   // divides 140 (LW) and 112 (SW)
constexpr int MCICA_YC = 8;

template <class Optics>
__global__ void __launch_bounds__(MCICA_XS * MCICA_YC, 5)
mcica_kernel(int ld, int col0, const int *__restrict__ perm, int nc, int nlay, int nsub, McicaParams P,
             const KissJump *__restrict__ jumps, const uint32_t *__restrict__ seeds,
             const long long *__restrict__ thr, const double *__restrict__ cldf, const double *__restrict__ ciwp,
             const double *__restrict__ clwp, double cwp_tiny, int cloudLM, int cloudMH,
             const int *__restrict__ ncloudy,    // columns c >= *ncloudy hold no cloud at all (null: unknown)
             const int *__restrict__ ktop,       // by caller-order column: last layer with cldf > 0 (null: unknown)
             int *__restrict__ clearCounts,      // (ld,4)
             uint32_t *__restrict__ cloudy_any,  // [nw][nc]
             uint32_t *__restrict__ mask,        // indexed by Optics::mask_index(w, nw, isub, c)
             Optics opt, int *err) {
    const int isub = blockIdx.x * MCICA_XS + threadIdx.x;
    const int c = blockIdx.y * MCICA_YC + threadIdx.y;
    if (c >= nc || isub >= nsub || RRTMGX_TRAPPED(P.trap)) return;
    const size_t col = gcol(col0, perm, c);
    if (ncloudy && c >= *ncloudy) {
        // cldfrac = 0 in every layer: cdf1 >= 1 - cldfrac never holds (ran_num < 1), so whatever the
        // generator draws the subcolumn is clear everywhere (:435); nothing to draw
        const int nw = (nlay + 31) >> 5;
        for (int w = 0; w < nw; ++w) mask[opt.mask_index(w, nw, isub, c)] = 0u;
        if (isub == 0)
            for (int q = 0; q < 4; ++q) atomicAdd(&clearCounts[(size_t)q * ld + col], nsub);
        return;
    }
    Kiss a, b;
    a.s1 = seeds[c]; a.s2 = seeds[(size_t)nc + c];
    a.s3 = seeds[(size_t)2 * nc + c]; a.s4 = seeds[(size_t)3 * nc + c];
    b = a;
    a.jump(jumps[2 * isub]);
    if (P.inhomo) b.jump(jumps[2 * isub + 1]);

    const bool surf1 = cloudLM < cloudMH;
    bool any_all = false, any_low = false, any_mid = false, any_high = false;
    int32_t k1 = 0, k3 = 0;   // integer draws behind cdf1 / cdf3, carried down the column
    uint32_t word = 0;
    typename Optics::State ost{};
    // above the last layer that holds cloud no draw can make a cell cloudy (cdf1 >= 1 - 0 never holds) and
    // nothing later reads the generator: the sweep stops there
    const int klast = ktop ? ktop[perm[c]] : nlay - 1;
    for (int k = 0; k <= klast; ++k) {
        const int32_t d1 = a.draw_int();
        const int32_t d2 = a.draw_int();
        const long long *t = thr + tile_index(nlay, 3, k, c);
        if (!(k > 0 && (long long)d2 < t[0])) k1 = d1;                 // else cdf1(k) = cdf1(k-1)
        if (P.inhomo) {
            const int32_t e2 = b.draw_int();
            const int32_t e3 = b.draw_int();
            if (!(k > 0 && (long long)e2 < t[32])) k3 = e3;            // else cdf3(k) = cdf3(k-1)
        }
        bool optical = false;
        if ((long long)k1 >= t[64]) {
            const size_t i2 = (size_t)k * ld + col;
            double ciw = ciwp[i2], clw = clwp[i2];
            if (P.inhomo) {
                const double cf = cldf[i2];
                double sigma = cf > 0.99 ? 0.5 : (cf > 0.9 ? 0.71 : 1.0);
                double zcw = zcw_lookup(P.xcw, kiss_value(k3), sigma);
                ciw = ciw * zcw;
                clw = clw * zcw;
            }
            bool ineg = ciw <= cwp_tiny, lneg = clw <= cwp_tiny;
            if (ineg) ciw = 0.;
            if (lneg) clw = 0.;
            if (!(ineg && lneg)) {
                // McICA-cloudy cell (cldy_stoch = .true.)
                const int lay1 = k + 1;
                any_all = true;
                if (surf1) {
                    if (lay1 <= cloudLM) any_low = true;
                    else if (lay1 <= cloudMH) any_mid = true;
                    else any_high = true;
                } else {
                    if (lay1 < cloudMH) any_high = true;
                    else if (lay1 < cloudLM) any_mid = true;
                    else any_low = true;
                }
                optical = opt.cell(k, isub, c, ciw, clw, err, ost);
            }
        }
        if (optical) word |= 1u << (k & 31);
        if ((k & 31) == 31 || k == klast) {
            const int w = k >> 5;
            mask[opt.mask_index(w, (nlay + 31) >> 5, isub, c)] = word;
            if (word) atomicOr(&cloudy_any[(size_t)w * nc + c], word);
            word = 0;
        }
    }
    for (int w = (klast >> 5) + 1; w < ((nlay + 31) >> 5); ++w) mask[opt.mask_index(w, (nlay + 31) >> 5, isub, c)] = 0u;
    opt.finish(isub, c, ost);
    if (!any_all) atomicAdd(&clearCounts[col], 1);
    if (!any_high) atomicAdd(&clearCounts[(size_t)ld + col], 1);
    if (!any_mid) atomicAdd(&clearCounts[(size_t)2 * ld + col], 1);
    if (!any_low) atomicAdd(&clearCounts[(size_t)3 * ld + col], 1);
}

}  // namespace rrtmgx
