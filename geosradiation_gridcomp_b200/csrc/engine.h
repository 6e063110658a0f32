// Internal host-side interface between the C ABI (api.cu) and the LW / SW kernel translation
// units (lw.cu, sw.cu).  Not installed; the public surface is include/rrtmgx.h.
#pragma once
#include <atomic>
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/rrtmgx.h"
#include "common.cuh"
#include "tables.h"

namespace rrtmgx {

// every kernel launch of the library goes through this counter (rrtmgx_launch_count); with
// profiling switched on (rrtmgx_profile) each launch is bracketed by CUDA events on its own
// stream and waited for, which serialises the step and attributes device time per kernel
extern std::atomic<long long> g_launches;   // LW and SW may be driven from two host threads
extern bool g_profile;
void profile_add(const char *name, float ms);
struct ProfScope {
    const char *name;
    cudaStream_t stream;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    ProfScope(const char *n, cudaStream_t s) : name(n), stream(s) {
        if (g_profile) {
            cudaEventCreate(&e0);
            cudaEventCreate(&e1);
            cudaEventRecord(e0, stream);
        }
    }
    ~ProfScope() {
        if (e0) {
            cudaEventRecord(e1, stream);
            cudaEventSynchronize(e1);
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            profile_add(name, ms);
            cudaEventDestroy(e0);
            cudaEventDestroy(e1);
        }
    }
};
#define RRTMGX_LAUNCH_TAG(tag, kernel, grid, block, smem, stream, ...)      \
    do {                                                                    \
        ::rrtmgx::ProfScope prof_scope_(tag, stream);                       \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);          \
        ++::rrtmgx::g_launches;                                             \
    } while (0)
#define RRTMGX_LAUNCH(kernel, grid, block, smem, stream, ...) \
    RRTMGX_LAUNCH_TAG(#kernel, kernel, grid, block, smem, stream, __VA_ARGS__)

// device bump allocator over one cudaMalloc'ed scratch slab (re-grown on demand by api.cu)
struct Slab {
    char *base = nullptr;
    size_t cap = 0, used = 0;
    template <class T> T *take(size_t n) {
        size_t bytes = (n * sizeof(T) + 255) & ~(size_t)255;
        T *p = (T *)(base + used);
        used += bytes;
        return p;
    }
};

// McICA state shared by LW and SW (SH/cloud_subcol_gen.F90:87-95 module variables)
struct McicaConfig {
    int ih = 1;
    double corr[8] = {1.4315, 2.1219, 7., -25.584, 0.72192, 0.78996, 8.5, 40.404};
};
// host: jump entries for subcolumn starts; entry 2*i: i*stride draws, 2*i+1: i*stride + 2*nlay
void kiss_jump_table(int nsub, int nlay, bool inhomo, KissJump *out /* 2*nsub */);
// test hook behind rrtmgx_debug_kiss (lw.cu): host arrays in and out
int debug_kiss(int nstream, const int32_t *seeds, int ndraw, int32_t *kiss, double *ran8, float *ran4, int nsub, int nlay,
               int inhomo, uint32_t *jumped, uint32_t *replayed, int nvalue, const int32_t *values, double *val8, float *val4,
               cudaStream_t st);
McicaParams mcica_params(const McicaConfig &cfg, const double *d_xcw_beta, const double *d_xcw_gamma,
                         int doy, const int seed_order[4]);

// Which columns of the CALLER's call a chunk covers.  The scratch slab of a path holds the McICA clouds of exactly
// one chunk; RRTMGX_REUSE_CLOUDS may keep them only when the previous run of the path was this very chunk of a
// call of the same extent.  A host-array call crosses in several staging chunks, each of which is run with
// chunk-local arrays (col0 = 0, ld = nc): without the call-level identity all of them would look alike.
struct ChunkId {
    long long first = 0;   // first column of the chunk in the caller's arrays
    long long total = 0;   // columns of the caller's call
    int nchunks = 1;       // chunks the call is run in (1: the slab still holds all of its clouds afterwards)
};
struct CloudCache {
    const char *base = nullptr;
    long long first = -1, total = -1;
    int nc = 0, nlay = 0;
    bool perm = false, valid = false;
    bool radval = false;   // SW: the SOLAR_RADVAL layer sums of the chunk were formed with the clouds
    bool matches(const char *b, const ChunkId &id, int nc_, int nlay_) const {
        return valid && base == b && first == id.first && total == id.total && nc == nc_ && nlay == nlay_;
    }
};

// ---- LW ------------------------------------------------------------------------------------
int lw_upload_tables(const HostTables &ht, const double *d_arena);   // fills __constant__ state
size_t lw_scratch_bytes(int nc, int nlay, bool debug);
// runs columns [col0, col0+nc) of the caller's arrays (device pointers, leading dimension
// a->ncol); `err` is a device word receiving negative trap codes
struct LwDebug {                  // optional device taps, chunk-local [..][nc] layouts
    double *taug = nullptr, *pfracs = nullptr;   // [nlay][140][nc]
};
// `id` names the chunk within the caller's whole call (RRTMGX_REUSE_CLOUDS keys on it, see ChunkId)
int lw_run_chunk(const RrtmgxLwArgs *a, int col0, int nc, const ChunkId &id, const McicaParams &mp,
                 const KissJump *d_jumps, Slab &slab, int *d_err, cudaStream_t stream,
                 cudaStream_t *side, int nside, cudaEvent_t *ev, const RrtmgxTaps *taps, int *d_negpos);
void lw_read_env();   // RRTMGX_LW_GN (called once per rrtmgx_init, under the library lock)

// ---- SW ------------------------------------------------------------------------------------
int sw_upload_tables(const HostTables &ht, const double *d_arena);
size_t sw_scratch_bytes(int nc, int nlay, bool debug, bool radval = false);
struct SwSolar {                  // host-evaluated scalars of rrtmg_sw_sub :889-1127
    double adjflux[14];           // adjes (* solvar) per band
    double svar_f, svar_s, svar_i;
    double svar_bnd[14];          // isolvar == 3: one multiplier per band for all three terms
    int isolvar;
};
// rrtmg_sw_sub :889-1127 + NRLSSI2.F90 (host scalars only); 0 or RRTMGX_ESOLVAR
int sw_solar_setup(const RrtmgxSwArgs *a, const HostTables &ht, SwSolar *out);
int sw_run_chunk(const RrtmgxSwArgs *a, const SwSolar &sol, int col0, int nc, const ChunkId &id, const McicaParams &mp,
                 const KissJump *d_jumps, Slab &slab, int *d_err, cudaStream_t stream,
                 cudaStream_t *side, int nside, cudaEvent_t *ev, const RrtmgxTaps *taps, int *d_negpos);
void sw_read_env();   // RRTMGX_SW_GN

void lw_forget_clouds();   // drop what RRTMGX_REUSE_CLOUDS would reuse (slab freed or re-grown)
void sw_forget_clouds();

// generic helpers (api.cu)
// perm[0..nc): the chunk's columns with any cldf > 0 first, the cloud-free ones after (device)
// ktop[0..nc): per chunk-local column (caller's order) the last layer with cldf > 0, -1 if none
int build_cloud_partition(int ld, int col0, int nc, int nlay, const double *cldf, int *perm, unsigned char *flags,
                          int *ktop, void *tmp, size_t tmp_bytes, cudaStream_t stream);
size_t cloud_partition_tmp_bytes(int nc);
void launch_check_negative(const double *const *x, const size_t *cnt, int narr, int *d_negpos, cudaStream_t s,
                           bool trap_nan);

}  // namespace rrtmgx
