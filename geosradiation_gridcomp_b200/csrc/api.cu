// C ABI of the library (include/rrtmgx.h): lifecycle, table upload, column chunking, the
// host<->device staging pipeline for host-pointer callers, and the input traps of the
// reference drivers (LW/src/rrtmg_lw_rad.F90:209-318, SW/src/rrtmg_sw_rad.F90:365-383).
//
// There is no CPU fallback: every entry point fails with RRTMGX_ENODEVICE / RRTMGX_ENOTINIT
// when no CUDA device is usable.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include <cub/device/device_partition.cuh>
#include <thrust/iterator/counting_iterator.h>

#include "engine.h"
#include "glue.cuh"

namespace rrtmgx {

std::atomic<long long> g_launches{0};
bool g_profile = false;
namespace {
struct ProfEntry { long long n = 0; double ms = 0.; };
std::map<std::string, ProfEntry> g_prof;
std::mutex g_prof_mu;
}  // namespace
void profile_add(const char *name, float ms) {
    std::lock_guard<std::mutex> lock(g_prof_mu);
    ProfEntry &e = g_prof[name];
    e.n += 1;
    e.ms += ms;
}

namespace {

constexpr int NSIDE = 4;
constexpr int NSTAGE = 4;   // staging sets of the host-array pipeline that exist (g.stages of them are used)

struct Path {   // per-path (LW or SW) execution resources
    cudaStream_t stream = nullptr, h2d = nullptr, d2h = nullptr;
    cudaStream_t side[NSIDE] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev[1 + NSIDE] = {};
    cudaEvent_t ev_in[NSTAGE] = {}, ev_done[NSTAGE] = {}, ev_free[NSTAGE] = {};
    Slab slab;            // kernel scratch
    Slab stage[NSTAGE];   // host-pointer mode: device copies of one chunk's boundary arrays
    Slab glue;            // fused Run-phase glue: the RRTMG argument arrays of one chunk
    Slab zeros;           // removed-gas runs: the array that stands in for the zeroed gas
    KissJump *d_jumps = nullptr;
    int jumps_nlay = -1, jumps_inhomo = -1;
    int *d_err = nullptr;      // [0] trap code, [1] first negative-input position
    int last_status = 0;
    bool pending = false;
    cudaStream_t run_stream = nullptr;   // stream of the last run (status is read behind it)
    RrtmgxTaps taps;
    bool has_taps = false;
};

constexpr size_t kHostChunkDefault = 8192;
constexpr int kStagesDefault = 2;

struct Ctx {
    bool ready = false;
    int device = 0;
    HostTables ht;
    double *d_arena = nullptr;
    McicaConfig mc;
    Path lw, sw;
    size_t chunk_cols = 0;   // 0: automatic
    size_t host_chunk_cols = kHostChunkDefault;   // staging chunk of the host-array pipeline (RRTMGX_HOST_CHUNK; sweep: profiles/r4_c_*, r4_d_*)
    bool host_chunk_from_env = false;   // RRTMGX_HOST_CHUNK set: taken literally, whatever the element size
    int stages = kStagesDefault;   // staging sets in flight (RRTMGX_STAGES): deeper than double buffering measured slower, profiles/r4_d_*
    std::mutex mu;
};

Ctx g;

const int kErrInit[2] = {0, 1 << 30};

bool ok(cudaError_t e) { return e == cudaSuccess; }

Slab g_sw_lit;   // RRTMGX_LIT_ONLY: the daytime-column index list of the Solar call in flight (solar_lit_index)

int grow(Slab &s, size_t bytes) {
    if (s.cap >= bytes) return 0;
    // a re-grown kernel slab holds nothing a later RRTMGX_REUSE_CLOUDS call of that path could keep (the other
    // path's cache is left alone: LW and SW may be called from two host threads)
    if (&s == &g.lw.slab) lw_forget_clouds();
    if (&s == &g.sw.slab) sw_forget_clouds();
    if (s.base) cudaFree(s.base);
    s.base = nullptr;
    s.cap = 0;
    if (!ok(cudaMalloc((void **)&s.base, bytes))) { cudaGetLastError(); return RRTMGX_ECUDA; }
    s.cap = bytes;
    return 0;
}

int path_init(Path &p) {
    if (!ok(cudaStreamCreateWithFlags(&p.stream, cudaStreamNonBlocking))) return RRTMGX_ECUDA;
    if (!ok(cudaStreamCreateWithFlags(&p.h2d, cudaStreamNonBlocking))) return RRTMGX_ECUDA;
    if (!ok(cudaStreamCreateWithFlags(&p.d2h, cudaStreamNonBlocking))) return RRTMGX_ECUDA;
    for (auto &s : p.side)
        if (!ok(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking))) return RRTMGX_ECUDA;
    for (auto &e : p.ev)
        if (!ok(cudaEventCreateWithFlags(&e, cudaEventDisableTiming))) return RRTMGX_ECUDA;
    for (int i = 0; i < NSTAGE; ++i) {
        if (!ok(cudaEventCreateWithFlags(&p.ev_in[i], cudaEventDisableTiming))) return RRTMGX_ECUDA;
        if (!ok(cudaEventCreateWithFlags(&p.ev_done[i], cudaEventDisableTiming))) return RRTMGX_ECUDA;
        if (!ok(cudaEventCreateWithFlags(&p.ev_free[i], cudaEventDisableTiming))) return RRTMGX_ECUDA;
    }
    if (!ok(cudaMalloc((void **)&p.d_err, 2 * sizeof(int)))) return RRTMGX_ECUDA;
    return 0;
}

void path_free(Path &p) {
    if (p.stream) cudaStreamDestroy(p.stream);
    if (p.h2d) cudaStreamDestroy(p.h2d);
    if (p.d2h) cudaStreamDestroy(p.d2h);
    for (auto &s : p.side) if (s) cudaStreamDestroy(s);
    for (auto &e : p.ev) if (e) cudaEventDestroy(e);
    for (int i = 0; i < NSTAGE; ++i) {
        if (p.ev_in[i]) cudaEventDestroy(p.ev_in[i]);
        if (p.ev_done[i]) cudaEventDestroy(p.ev_done[i]);
        if (p.ev_free[i]) cudaEventDestroy(p.ev_free[i]);
    }
    if (p.slab.base) cudaFree(p.slab.base);
    for (auto &s : p.stage) if (s.base) cudaFree(s.base);
    if (p.glue.base) cudaFree(p.glue.base);
    if (p.zeros.base) cudaFree(p.zeros.base);
    if (p.d_jumps) cudaFree(p.d_jumps);
    if (p.d_err) cudaFree(p.d_err);
    p = Path();
}

std::string default_blob() {
    if (const char *e = std::getenv("RRTMGX_TABLES")) return e;
    Dl_info info;
    if (dladdr((void *)&default_blob, &info) && info.dli_fname) {
        std::string so(info.dli_fname);
        size_t k = so.find_last_of('/');
        std::string dir = k == std::string::npos ? "." : so.substr(0, k);
        return dir + "/data/rrtmg_tables.bin";
    }
    return "rrtmg_tables.bin";
}

int ensure_jumps(Path &p, int nsub, int nlay) {
    const int inhomo = g.mc.ih > 0;
    if (p.d_jumps && p.jumps_nlay == nlay && p.jumps_inhomo == inhomo) return 0;
    std::vector<KissJump> h(2 * (size_t)nsub);
    kiss_jump_table(nsub, nlay, inhomo, h.data());
    if (!p.d_jumps && !ok(cudaMalloc((void **)&p.d_jumps, sizeof(KissJump) * 2 * 140))) return RRTMGX_ECUDA;
    // stream-ordered behind any work still reading the old table
    cudaStreamSynchronize(p.stream);
    if (!ok(cudaMemcpy(p.d_jumps, h.data(), sizeof(KissJump) * h.size(), cudaMemcpyHostToDevice))) return RRTMGX_ECUDA;
    p.jumps_nlay = nlay;
    p.jumps_inhomo = inhomo;
    return 0;
}

// The reference scans every input for negative values before it starts (LW :209-318, SW :365-383) and names
// the first offending array.  One launch scans all arrays of a call: blockIdx.y picks the array, the blocks of
// a row stride over it with 16-byte loads; the smallest offending array position wins.
struct NegScan {
    const double *x[24];
    unsigned long long n[24];
    int trap_nan;   // SW: _ASSERT(all(x >= 0.)) also fires on NaN (SW :365-383); LW: any(x < 0.) (LW :209-318) would not, see the call
};
__device__ __forceinline__ bool neg_bad(double v, int trap_nan) { return trap_nan ? !(v >= 0.) : (v < 0.); }
__global__ void __launch_bounds__(256) check_negative_kernel(NegScan S, int *negpos) {
    const int a = blockIdx.y;
    const double *__restrict__ x = S.x[a];
    const size_t n = S.n[a], n2 = n / 2;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    bool bad = false;
    const double2 *__restrict__ x2 = reinterpret_cast<const double2 *>(x);   // arrays are 16-byte aligned slabs
    if ((reinterpret_cast<uintptr_t>(x) & 15) == 0) {
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
            const double2 v = x2[i];
            bad |= neg_bad(v.x, S.trap_nan) | neg_bad(v.y, S.trap_nan);
        }
        if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) bad |= neg_bad(x[n - 1], S.trap_nan);
    } else {
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) bad |= neg_bad(x[i], S.trap_nan);
    }
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicMin(negpos, a);
}

// the band kernels address the per-(layer, column) planes of a chunk with 32-bit element offsets
// (at most 32 planes of nlay*nc elements)
size_t chunk_cap(int nlay) { return (((size_t)1 << 31) - 1) / ((size_t)32 * std::max(nlay, 1)); }

// Staging chunk of the host-array pipeline in columns.  The default is sized for fp64 arrays; real*4 arrays take
// twice the columns (the same bytes per chunk): with half the bytes the kernels, not the link, are the longer stage
// of the pipeline, and 8192 columns are 256 blocks per band kernel, not one wave of a B200 (measured,
// profiles/t1_b_e2e_chunk_sweep.jsonl: +6 % end to end with real*4 arrays and through the real*4 glue at 16384,
// nothing with fp64 arrays).  RRTMGX_HOST_CHUNK is taken literally.
size_t host_chunk(bool f32) { return g.host_chunk_cols * ((f32 && !g.host_chunk_from_env) ? 2 : 1); }

size_t pick_chunk(int ncol, int nlay, size_t per_col_bytes, bool host_mode, size_t own_bytes, bool f32 = false) {
    if (g.chunk_cols) return std::min<size_t>(std::min<size_t>(g.chunk_cols, chunk_cap(nlay)), (size_t)ncol);
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    // scratch budget per path: 55% of what is free, at most 80 GiB (B200: 180 GB of HBM3e; the SW path
    // holds 1.1 MB per column, so this is what lets it run 65 536-column chunks: fewer, fuller launches)
    // (the path's own slab from earlier calls counts as free: the choice must not shrink once it is allocated)
    size_t budget = std::min<size_t>((free_b + own_bytes) / 100 * 55, (size_t)80 << 30);
    size_t nc = std::max<size_t>(1024, budget / std::max<size_t>(per_col_bytes, 1));
    nc = std::min<size_t>(std::min<size_t>(nc, 65536), chunk_cap(nlay));
    // host arrays: smaller chunks shorten the fill/drain of the H2D -> kernels -> D2H pipeline
    if (host_mode) nc = std::min<size_t>(nc, host_chunk(f32));
    nc &= ~(size_t)127;
    return std::min<size_t>(nc, (size_t)ncol);
}

int status_from(Path &p) {
    int h[2];
    cudaStream_t st = p.run_stream ? p.run_stream : p.stream;
    if (!ok(cudaMemcpyAsync(h, p.d_err, sizeof h, cudaMemcpyDeviceToHost, st))) return RRTMGX_ECUDA;
    if (!ok(cudaStreamSynchronize(st))) { cudaGetLastError(); return RRTMGX_ECUDA; }
    p.pending = false;
    if (h[1] < (1 << 30)) return RRTMGX_ENEGATIVE - 1 - h[1];   // -(101 + position)
    return h[0];
}

// One boundary array as seen by the chunk pipeline: `rows` rows of ncol elements (column
// fastest), or band-fastest (16,ncol) outputs when `colmajor_inner` is set.
struct Arr {
    const void *host;   // caller pointer (host or device)
    void **slot;        // where the per-chunk device pointer goes in the chunk args
    size_t rows, elem;
    bool inner;         // (k,ncol) layout: the chunk is contiguous
    bool in, out;
    bool f32 = false;   // RRTMGX_F32_ARRAYS: the caller's array is real*4; widened / narrowed on the device
};

// real*4 boundary arrays (the production kind of GEOS): exact widening on the way in, one rounding on the way out
__global__ void widen_kernel(const float *__restrict__ x, double *__restrict__ y, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        y[i] = (double)x[i];
}
__global__ void narrow_kernel(const double *__restrict__ x, float *__restrict__ y, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        y[i] = (float)x[i];
}

// Host-pointer mode: run `fn(chunk_args, nc)` over chunks with H2D / compute / D2H overlapped
// on three streams and g.stages staging sets (a set is reused once the D2H of its previous chunk is done).
template <class Args, class Fn>
int run_staged(Path &p, const Args &a, Args &ca, std::vector<Arr> &arrs, int ncol, size_t chunk, Fn fn) {
    size_t stage_bytes = 0;
    for (auto &r : arrs)
        stage_bytes += (((r.rows * chunk * r.elem) + 255) & ~(size_t)255) +
                       (r.f32 ? ((r.rows * chunk * 8 + 255) & ~(size_t)255) : 0);
    const int nstage = g.stages;
    for (int s = 0; s < nstage; ++s)
        if (int rc = grow(p.stage[s], stage_bytes + 4096)) return rc;
    // the first failing copy / event call of the loop decides the status; on any failure the three streams are
    // drained before returning, so no copy into the caller's arrays is still in flight behind the call
    cudaError_t first_err = cudaSuccess;
    auto note = [&](cudaError_t e) { if (e != cudaSuccess && first_err == cudaSuccess) first_err = e; };
    auto drain = [&]() { note(cudaStreamSynchronize(p.h2d)); note(cudaStreamSynchronize(p.stream)); note(cudaStreamSynchronize(p.d2h)); };
    int k = 0;
    for (size_t col0 = 0; col0 < (size_t)ncol; col0 += chunk, ++k) {
        const int s = k % nstage;
        const size_t nc = std::min(chunk, (size_t)ncol - col0);
        Slab &st = p.stage[s];
        st.used = 0;
        std::vector<void *> dev(arrs.size()), raw(arrs.size());
        if (k >= nstage) note(cudaStreamWaitEvent(p.h2d, p.ev_free[s], 0));
        for (size_t i = 0; i < arrs.size(); ++i) {
            Arr &r = arrs[i];
            raw[i] = r.host ? (void *)st.take<char>(r.rows * nc * r.elem) : nullptr;
            dev[i] = (r.host && r.f32) ? (void *)st.take<char>(r.rows * nc * 8) : raw[i];
            if (!r.host || !r.in) continue;
            if (r.inner)
                note(cudaMemcpyAsync(raw[i], (const char *)r.host + col0 * r.rows * r.elem, r.rows * nc * r.elem,
                                     cudaMemcpyDefault, p.h2d));
            else
                note(cudaMemcpy2DAsync(raw[i], nc * r.elem, (const char *)r.host + col0 * r.elem, (size_t)ncol * r.elem,
                                       nc * r.elem, r.rows, cudaMemcpyDefault, p.h2d));
        }
        note(cudaEventRecord(p.ev_in[s], p.h2d));
        note(cudaStreamWaitEvent(p.stream, p.ev_in[s], 0));
        for (size_t i = 0; i < arrs.size(); ++i)
            if (arrs[i].host && arrs[i].f32 && arrs[i].in) {
                const size_t n = arrs[i].rows * nc;
                RRTMGX_LAUNCH(widen_kernel, (unsigned)std::min<size_t>((n + 255) / 256, 148 * 16), 256, 0, p.stream,
                              (const float *)raw[i], (double *)dev[i], n);
            }
        ca = a;
        ca.ncol = (int)nc;
        for (size_t i = 0; i < arrs.size(); ++i) *arrs[i].slot = dev[i];
        if (int rc = fn(ca, (int)nc, col0)) { drain(); return rc; }
        for (size_t i = 0; i < arrs.size(); ++i)
            if (arrs[i].host && arrs[i].f32 && arrs[i].out) {
                const size_t n = arrs[i].rows * nc;
                RRTMGX_LAUNCH(narrow_kernel, (unsigned)std::min<size_t>((n + 255) / 256, 148 * 16), 256, 0, p.stream,
                              (const double *)dev[i], (float *)raw[i], n);
            }
        note(cudaEventRecord(p.ev_done[s], p.stream));
        note(cudaStreamWaitEvent(p.d2h, p.ev_done[s], 0));
        for (size_t i = 0; i < arrs.size(); ++i) {
            Arr &r = arrs[i];
            if (!r.host || !r.out) continue;
            if (r.inner)
                note(cudaMemcpyAsync((char *)r.host + col0 * r.rows * r.elem, raw[i], r.rows * nc * r.elem,
                                     cudaMemcpyDefault, p.d2h));
            else
                note(cudaMemcpy2DAsync((char *)r.host + col0 * r.elem, (size_t)ncol * r.elem, raw[i], nc * r.elem,
                                       nc * r.elem, r.rows, cudaMemcpyDefault, p.d2h));
        }
        note(cudaEventRecord(p.ev_free[s], p.d2h));
        if (first_err != cudaSuccess) break;
    }
    if (first_err != cudaSuccess) { drain(); cudaGetLastError(); return RRTMGX_ECUDA; }
    if (!ok(cudaStreamSynchronize(p.d2h))) { cudaGetLastError(); return RRTMGX_ECUDA; }
    return 0;
}

const double *dev_table(const char *name) {
    TableRef r = g.ht.find(name);
    return r.ok() ? g.d_arena + r.off : nullptr;
}

// RAD:798-819: heating rate of layer l from the net flux divergence across it
__global__ void heating_rate_kernel(int ncol, int nlay, const double *__restrict__ fnet,
                                    const double *__restrict__ plev, double *__restrict__ hr, double gcp) {
    const size_t n = (size_t)ncol * nlay;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // fnet = up - down at levels (ncol,nlay+1), level 0 at the surface; plev in hPa
    hr[i] = (fnet[i] - fnet[i + ncol]) * gcp / ((plev[i] - plev[i + ncol]) * 100.) * 86400.;
}

// the band kernels' branch-free division and reciprocal beside the compiler's IEEE ones
__global__ void debug_divide_kernel(size_t n, const double *__restrict__ a, const double *__restrict__ b,
                                    double *__restrict__ q_fast, double *__restrict__ q_ieee,
                                    double *__restrict__ r_fast, double *__restrict__ r_ieee) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    q_fast[i] = ddiv(a[i], b[i]);
    q_ieee[i] = a[i] / b[i];
    r_fast[i] = drcp(b[i]);
    r_ieee[i] = 1. / b[i];
}

}  // namespace

namespace {
__global__ void cloudy_flag_kernel(int ld, int col0, int nc, int nlay, const double *__restrict__ cldf,
                                   unsigned char *__restrict__ flag, int *__restrict__ ktop) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nc) return;
    int top = -1;   // last layer (array order) that can hold cloud
    for (int k = 0; k < nlay; ++k)
        if (cldf[(size_t)k * ld + col0 + c] > 0.) top = k;
    flag[c] = top >= 0 ? 1 : 0;
    ktop[c] = top;
}

}  // namespace

size_t cloud_partition_tmp_bytes(int nc) {
    size_t bytes = 0;
    thrust::counting_iterator<int> it(0);
    cub::DevicePartition::Flagged(nullptr, bytes, it, (const unsigned char *)nullptr, (int *)nullptr, (int *)nullptr, nc);
    return bytes + 256;
}

// perm[0 .. nc): the chunk's columns, those holding cloud in any layer first (their count is left at
// (int*)tmp); both groups keep the caller's order, so neighbouring columns stay neighbours and the
// boundary arrays are still read in whole sectors.  (Ordering each group by surface pressure as well,
// so that a warp walks the same k-table rows, was measured: no gain on distinct columns and 5 % off
// the host-array path, profiles/r2_cb_tuning.txt.)
int build_cloud_partition(int ld, int col0, int nc, int nlay, const double *cldf, int *perm, unsigned char *flags,
                          int *ktop, void *tmp, size_t tmp_bytes, cudaStream_t stream) {
    RRTMGX_LAUNCH(cloudy_flag_kernel, (nc + 255) / 256, 256, 0, stream, ld, col0, nc, nlay, cldf, flags, ktop);
    thrust::counting_iterator<int> it(0);
    int *d_nsel = (int *)tmp;   // first 256 bytes of tmp hold the selected count
    size_t bytes = tmp_bytes - 256;
    ++g_launches;
    return cub::DevicePartition::Flagged((char *)tmp + 256, bytes, it, flags, perm, d_nsel, nc, stream) == cudaSuccess
               ? 0 : RRTMGX_ECUDA;
}

// arrays x[i] of cnt[i] elements, position i in the reference's order of checks; null or empty arrays are skipped
void launch_check_negative(const double *const *x, const size_t *cnt, int narr, int *d_negpos, cudaStream_t s,
                           bool trap_nan) {
    NegScan S{};
    S.trap_nan = trap_nan ? 1 : 0;
    size_t most = 0;
    for (int i = 0; i < narr && i < 24; ++i) {
        S.x[i] = x[i];
        S.n[i] = x[i] ? cnt[i] : 0;
        most = std::max<size_t>(most, S.n[i]);
    }
    if (!most) return;
    const unsigned bx = (unsigned)std::min<size_t>((most / 2 + 255) / 256, 148 * 2);
    RRTMGX_LAUNCH(check_negative_kernel, dim3(std::max(bx, 1u), narr), 256, 0, s, S, d_negpos);
}

}  // namespace rrtmgx

using namespace rrtmgx;

extern "C" {

int rrtmgx_init(const RrtmgxConfig *cfg) {
    std::lock_guard<std::mutex> lock(g.mu);
    // the reference's *_ini routines never touch the McICA module state (set_inhomogeneity and
    // initialize_cloud_subcol_gen own it, RAD:565,578) and GEOS calls them on every refresh (IRR:3381,
    // SOL:6225): a call on an initialised library is a pure no-op, whatever cfg holds
    if (g.ready) return 0;
    // cfg->inhomogeneity: 0..2 sets ih, -1 leaves the default (1 = beta, RAD:564); validated before anything is allocated
    if (cfg && (cfg->inhomogeneity < -1 || cfg->inhomogeneity > 2)) return RRTMGX_EINHOMO;
    int ndev = 0;
    if (!ok(cudaGetDeviceCount(&ndev)) || ndev == 0) { cudaGetLastError(); return RRTMGX_ENODEVICE; }
    if (cfg && cfg->device >= 0) {
        if (!ok(cudaSetDevice(cfg->device))) { cudaGetLastError(); return RRTMGX_ENODEVICE; }
    }
    if (!ok(cudaGetDevice(&g.device))) return RRTMGX_ENODEVICE;
    const std::string blob = (cfg && cfg->table_blob) ? cfg->table_blob : default_blob();
    if (int rc = g.ht.load(blob)) return rc;
    auto fail = [&](int rc) {   // a failed init leaves nothing allocated
        path_free(g.lw);
        path_free(g.sw);
        if (g.d_arena) cudaFree(g.d_arena);
        g.d_arena = nullptr;
        cudaGetLastError();
        return rc;
    };
    const size_t bytes = g.ht.arena.size() * sizeof(double);
    if (!ok(cudaMalloc((void **)&g.d_arena, bytes))) return fail(RRTMGX_ENODEVICE);
    if (!ok(cudaMemcpy(g.d_arena, g.ht.arena.data(), bytes, cudaMemcpyHostToDevice))) return fail(RRTMGX_ECUDA);
    if (int rc = lw_upload_tables(g.ht, g.d_arena)) return fail(rc);
    if (int rc = sw_upload_tables(g.ht, g.d_arena)) return fail(rc);
    if (int rc = path_init(g.lw)) return fail(rc);
    if (int rc = path_init(g.sw)) return fail(rc);
    g.mc = McicaConfig();
    if (cfg) {
        if (cfg->inhomogeneity >= 0) g.mc.ih = cfg->inhomogeneity;
        if (cfg->corr) std::memcpy(g.mc.corr, cfg->corr, sizeof g.mc.corr);
    }
    // tuning knobs: the defaults first, so that a finalize -> init cycle never inherits the previous environment
    g.chunk_cols = 0;
    g.host_chunk_cols = kHostChunkDefault;
    g.host_chunk_from_env = false;
    g.stages = kStagesDefault;
    if (const char *e = std::getenv("RRTMGX_CHUNK")) g.chunk_cols = (size_t)std::atoll(e);
    if (const char *e = std::getenv("RRTMGX_HOST_CHUNK")) {
        g.host_chunk_cols = std::max<size_t>(1024, (size_t)std::atoll(e));
        g.host_chunk_from_env = true;
    }
    if (const char *e = std::getenv("RRTMGX_STAGES")) g.stages = std::min(NSTAGE, std::max(2, std::atoi(e)));
    lw_read_env();
    sw_read_env();
    g.ready = true;
    return 0;
}

// set_inhomogeneity + initialize_cloud_subcol_gen (SH/cloud_condensate_inhomogeneity.F90:45, SH/cloud_subcol_gen.F90:108):
// the module state both paths read.  Taken under the library lock and stream-ordered behind the work in flight.
int rrtmgx_set_mcica(int ih, const double corr[8]) {
    std::lock_guard<std::mutex> lock(g.mu);
    if (!g.ready) return RRTMGX_ENOTINIT;
    if (ih < 0 || ih > 2) return RRTMGX_EINHOMO;
    McicaConfig mc;
    mc.ih = ih;
    if (corr) std::memcpy(mc.corr, corr, sizeof mc.corr);
    g.mc = mc;
    // the clouds a path still holds were drawn under the previous settings
    lw_forget_clouds();
    sw_forget_clouds();
    return 0;
}

int rrtmgx_get_knobs(long long knobs[4]) {
    std::lock_guard<std::mutex> lock(g.mu);
    if (!g.ready) return RRTMGX_ENOTINIT;
    if (!knobs) return RRTMGX_EARG;
    knobs[0] = (long long)g.chunk_cols;
    knobs[1] = (long long)g.host_chunk_cols;
    knobs[2] = g.stages;
    knobs[3] = g.mc.ih;
    return 0;
}

int rrtmgx_finalize(void) {
    std::lock_guard<std::mutex> lock(g.mu);
    if (!g.ready) return 0;
    cudaDeviceSynchronize();
    path_free(g.lw);
    path_free(g.sw);
    lw_forget_clouds();
    sw_forget_clouds();
    if (g.d_arena) cudaFree(g.d_arena);
    g.d_arena = nullptr;
    if (g_sw_lit.base) cudaFree(g_sw_lit.base);
    g_sw_lit = Slab();
    g.ready = false;
    return 0;
}

const char *rrtmgx_strerror(int status) {
    switch (status) {
        case RRTMGX_OK: return "success";
        case RRTMGX_ENODEVICE: return "no usable CUDA device (this library has no CPU fallback)";
        case RRTMGX_ENOTINIT: return "rrtmgx_init has not been called";
        case RRTMGX_EBLOB: return "table blob missing or malformed";
        case RRTMGX_EARG: return "bad scalar argument";
        case RRTMGX_ECUDA: return "CUDA runtime failure";
        case RRTMGX_EINHOMO: return "set_inhomogeneity: unknown inhomogeneity type";
        case RRTMGX_ESEEDORDER: return "generate_stochastic_clouds: bad seed_order";
        case RRTMGX_ESUPERLAYER: return "clearCounts_threeBand: invalid pressure super-layers!";
        case RRTMGX_EPRESSURE: return "RRTMG LW pressure misordering";
        case RRTMGX_EICEFLAG: return "cldprmc: invalid iceflag";
        case RRTMGX_ERADIUS_ICE: return "cldprmc: ice radius extrapolation forbidden";
        case RRTMGX_ELIQFLAG: return "cldprmc: invalid liqflag";
        case RRTMGX_ERADIUS_LIQ: return "cldprmc: liquid radius extrapolation forbidden";
        case RRTMGX_ESOLVAR: return "rrtmg_sw: invalid isolvar or missing optional argument";
        default: break;
    }
    if (status <= RRTMGX_ENEGATIVE) return "negative values in an input array";
    return "unknown status";
}

long long rrtmgx_launch_count(void) { return g_launches.load(); }

void rrtmgx_profile(int enable) {
    std::lock_guard<std::mutex> lock(g_prof_mu);
    g_profile = enable != 0;
    if (enable) g_prof.clear();
}

size_t rrtmgx_profile_report(char *buf, size_t cap) {
    std::lock_guard<std::mutex> lock(g_prof_mu);
    std::string out;
    char line[256];
    for (const auto &kv : g_prof) {
        std::snprintf(line, sizeof line, "%s\t%lld\t%.6f\n", kv.first.c_str(), kv.second.n, kv.second.ms);
        out += line;
    }
    if (buf && cap) {
        const size_t n = std::min(cap - 1, out.size());
        std::memcpy(buf, out.data(), n);
        buf[n] = 0;
    }
    return out.size() + 1;
}

void rrtmgx_set_taps(const RrtmgxTaps *lw_taps, const RrtmgxTaps *sw_taps) {
    std::lock_guard<std::mutex> lock(g.mu);
    g.lw.has_taps = lw_taps != nullptr;
    if (lw_taps) g.lw.taps = *lw_taps;
    g.sw.has_taps = sw_taps != nullptr;
    if (sw_taps) g.sw.taps = *sw_taps;
}

const double *rrtmgx_table(const char *kind, const char *name, int band, int *n) {
    if (n) *n = 0;
    if (!g.ready) return nullptr;
    char key[96];
    const bool lw = std::strcmp(kind, "lw") == 0;
    if (band > 0)
        std::snprintf(key, sizeof key, lw ? "lw.%02d.%s" : "sw.%02d.%s", band, name);
    else
        std::snprintf(key, sizeof key, "%s.%s", kind, name);
    TableRef r = g.ht.find(key);
    if (!r.ok()) return nullptr;
    if (n) *n = (int)r.n;
    return g.ht.ptr(r);
}

int rrtmgx_lw_status(void) {
    if (!g.ready) return RRTMGX_ENOTINIT;
    cudaSetDevice(g.device);
    if (!g.lw.pending) return g.lw.last_status;
    g.lw.last_status = status_from(g.lw);
    return g.lw.last_status;
}

int rrtmgx_sw_status(void) {
    if (!g.ready) return RRTMGX_ENOTINIT;
    cudaSetDevice(g.device);
    if (!g.sw.pending) return g.sw.last_status;
    g.sw.last_status = status_from(g.sw);
    return g.sw.last_status;
}

int rrtmgx_lw_run(const RrtmgxLwArgs *a) { return rrtmgx_lw_run_variants(a, nullptr); }

int rrtmgx_lw_run_variants(const RrtmgxLwArgs *a, const RrtmgxLwVariants *var) {
    if (!g.ready) return RRTMGX_ENOTINIT;
    // the CUDA current device is per host thread: callers may run LW and SW from different threads
    if (!ok(cudaSetDevice(g.device))) { cudaGetLastError(); return RRTMGX_ENODEVICE; }
    if (!a || a->ncol <= 0 || a->nlay <= 0 || a->nlay > 1000) return RRTMGX_EARG;
    if (a->cloudLM == a->cloudMH) return RRTMGX_ESUPERLAYER;   // cloud_subcol_gen.F90:762-766
    if (a->iceflglw < 0 || a->iceflglw > 4) return RRTMGX_EICEFLAG;
    if (a->liqflglw != 1) return RRTMGX_ELIQFLAG;
    const int nvar = var ? var->nvar : 0;
    if (nvar < 0 || nvar > 64 || (nvar && (!var->gas || !var->uflx || !var->dflx || (a->dudTs && !var->duflx_dTs))))
        return RRTMGX_EARG;
    for (int v = 0; v < nvar; ++v)
        if (var->gas[v] < RRTMGX_GAS_H2O || var->gas[v] > RRTMGX_GAS_HCFC22) return RRTMGX_EARG;
    Path &p = g.lw;
    const bool devptr = a->flags & RRTMGX_DEVICE_PTRS;
    if ((a->flags & RRTMGX_NO_SYNC) && !devptr) return RRTMGX_EARG;
    // real*4 arrays are widened while staged: real*4 DEVICE arrays take the staging path too (device-to-device
    // copies instead of PCIe), so the call is synchronous for them
    if ((a->flags & RRTMGX_F32_ARRAYS) && (a->flags & RRTMGX_NO_SYNC)) return RRTMGX_EARG;
    const bool staged = !devptr || (a->flags & RRTMGX_F32_ARRAYS);
    const int ncol = a->ncol, nlay = a->nlay;
    if (int rc = ensure_jumps(p, 140, nlay)) return rc;
    static const int seed_order[4] = {1, 2, 3, 4};   // LW/src/rrtmg_lw_rad.F90:541-546
    const McicaParams mp = mcica_params(g.mc, dev_table("mcica.xcw_beta"), dev_table("mcica.xcw_gamma"), a->dyofyr,
                                        seed_order);
    const RrtmgxTaps *taps = p.has_taps ? &p.taps : nullptr;
    const bool dbg = taps && (taps->taug || taps->pfracs);
    const size_t per_col = lw_scratch_bytes(1024, nlay, dbg) / 1024;
    size_t chunk = pick_chunk(ncol, nlay, per_col + (!staged ? 0 : 2 * (size_t)(37 * nlay + 60) * 8), staged, p.slab.cap,
                              (a->flags & RRTMGX_F32_ARRAYS) != 0);
    if (taps) {   // taps are laid out for the whole call
        if ((size_t)ncol > chunk_cap(nlay)) return RRTMGX_EARG;
        chunk = (size_t)ncol;
    }
    if (int rc = grow(p.slab, lw_scratch_bytes((int)chunk, nlay, dbg))) return rc;
    cudaStream_t stream = (!staged && a->stream) ? (cudaStream_t)a->stream : p.stream;
    p.run_stream = stream;
    if (!(devptr && (a->flags & RRTMGX_KEEP_STATUS)) &&
        !ok(cudaMemcpyAsync(p.d_err, kErrInit, sizeof kErrInit, cudaMemcpyHostToDevice, stream)))
        return RRTMGX_ECUDA;

    // removed-gas runs: a device array of zeros stands in for the gas (as large as one staged chunk or the whole
    // device-resident call), d_var = the (.., nvar) flux arrays as the device sees them
    const double *d_zero = nullptr;
    double *d_var[3] = {nullptr, nullptr, nullptr};
    if (nvar) {
        const size_t nz = (size_t)(!staged ? ncol : (int)std::min<size_t>(chunk, ncol)) * nlay;
        if (int rc = grow(p.zeros, nz * sizeof(double) + 256)) return rc;
        if (!ok(cudaMemsetAsync(p.zeros.base, 0, nz * sizeof(double), stream))) return RRTMGX_ECUDA;
        d_zero = (const double *)p.zeros.base;
        d_var[0] = var->uflx; d_var[1] = var->dflx; d_var[2] = var->duflx_dTs;
    }

    // `da` holds device pointers for columns [first, first + da.ncol) of the caller's call
    const int nchunks_dev = (int)(((size_t)ncol + chunk - 1) / chunk);
    auto run_chunks_device = [&](const RrtmgxLwArgs &da, size_t first) -> int {
        const int n = da.ncol;
        if (!(da.flags & RRTMGX_SKIP_CHECKS)) {
            const size_t n2 = (size_t)n * nlay, n2p = (size_t)n * (nlay + 1);
            struct { const double *x; size_t cnt; } chk[] = {
                {da.play, n2}, {da.plev, n2p}, {da.tlay, n2}, {da.tlev, n2p}, {da.tsfc, (size_t)n},
                {da.h2ovmr, n2}, {da.o3vmr, n2}, {da.co2vmr, n2}, {da.ch4vmr, n2}, {da.n2ovmr, n2},
                {da.o2vmr, n2}, {da.cfc11vmr, n2}, {da.cfc12vmr, n2}, {da.cfc22vmr, n2}, {da.ccl4vmr, n2},
                {da.emis, (size_t)n * 16}, {da.cldf, n2}, {da.ciwp, n2}, {da.clwp, n2}, {da.rei, n2},
                {da.rel, n2}, {da.tauaer, n2 * 16}};
            constexpr int narr = (int)(sizeof chk / sizeof chk[0]);
            const double *xs[narr];
            size_t cnts[narr];
            for (int i = 0; i < narr; ++i) { xs[i] = chk[i].x; cnts[i] = chk[i].cnt; }
            // NaN trapped too: the reference's any(x < 0.) lets a NaN through to int() conversions of undefined result
            // (out-of-range table reads); here it is refused at the same position as a negative value
            launch_check_negative(xs, cnts, narr, p.d_err + 1, stream, true);
        }
        // chunk by chunk: the removed-gas runs of the chunk (gas array replaced by zeros, fluxes into slab v of the
        // variant arrays), then the run with every gas; all of them on the clouds the first one generated
        const size_t vslab = (size_t)n * (nlay + 1);
        for (size_t col0 = 0; col0 < (size_t)n; col0 += chunk) {
            const int nc = (int)std::min(chunk, (size_t)n - col0);
            for (int v = 0; v <= nvar; ++v) {
                RrtmgxLwArgs dv = da;
                if (v < nvar) {
                    const double **gasp[] = {nullptr, &dv.h2ovmr, &dv.o3vmr, &dv.co2vmr, &dv.ch4vmr, &dv.n2ovmr,
                                             &dv.cfc11vmr, &dv.cfc12vmr, &dv.cfc22vmr};
                    // a zero array with the leading dimension of the call: rows of the gas array are n apart
                    *gasp[var->gas[v]] = d_zero;
                    dv.uflx = d_var[0] + v * vslab; dv.dflx = d_var[1] + v * vslab;
                    if (da.dudTs) dv.duflx_dTs = d_var[2] + v * vslab;
                }
                if (v > 0) dv.flags |= RRTMGX_REUSE_CLOUDS;
                const ChunkId id{(long long)(first + col0), (long long)ncol, nchunks_dev};
                if (int rc = lw_run_chunk(&dv, (int)col0, nc, id, mp, p.d_jumps, p.slab, p.d_err, stream, p.side, NSIDE, p.ev,
                                          taps, p.d_err + 1))
                    return rc;
            }
        }
        return 0;
    };
    if (!staged) {
        if (int rc = run_chunks_device(*a, 0)) return rc;
        p.pending = true;
        if (a->flags & RRTMGX_NO_SYNC) return 0;
        p.last_status = status_from(p);
        return p.last_status;
    }

    // host pointers: stage chunk by chunk
    RrtmgxLwArgs ca = *a;
    std::vector<Arr> arrs;
    const bool f32 = a->flags & RRTMGX_F32_ARRAYS;   // the caller's real arrays are real*4
    const size_t esz = f32 ? 4 : 8;
    const size_t L = nlay, L1 = nlay + 1;
    auto in = [&](const double *const &field, const double **slot, size_t rows) {
        arrs.push_back({field, (void **)slot, rows, esz, false, true, false, f32});
    };
    auto out = [&](double *const &field, double **slot, size_t rows, bool inner = false) {
        arrs.push_back({field, (void **)slot, rows, esz, inner, false, true, f32});
    };
    in(a->play, &ca.play, L); in(a->plev, &ca.plev, L1); in(a->tlay, &ca.tlay, L); in(a->tlev, &ca.tlev, L1);
    in(a->tsfc, &ca.tsfc, 1); in(a->emis, &ca.emis, 16);
    in(a->h2ovmr, &ca.h2ovmr, L); in(a->o3vmr, &ca.o3vmr, L); in(a->co2vmr, &ca.co2vmr, L);
    in(a->ch4vmr, &ca.ch4vmr, L); in(a->n2ovmr, &ca.n2ovmr, L); in(a->o2vmr, &ca.o2vmr, L);
    in(a->cfc11vmr, &ca.cfc11vmr, L); in(a->cfc12vmr, &ca.cfc12vmr, L); in(a->cfc22vmr, &ca.cfc22vmr, L);
    in(a->ccl4vmr, &ca.ccl4vmr, L); in(a->cldf, &ca.cldf, L); in(a->ciwp, &ca.ciwp, L); in(a->clwp, &ca.clwp, L);
    in(a->rei, &ca.rei, L); in(a->rel, &ca.rel, L); in(a->tauaer, &ca.tauaer, L * 16);
    in(a->zm, &ca.zm, L); in(a->alat, &ca.alat, 1);
    arrs.push_back({a->clearCounts, (void **)&ca.clearCounts, 4, 4, false, false, true});
    out(a->uflx, &ca.uflx, L1); out(a->dflx, &ca.dflx, L1); out(a->uflxc, &ca.uflxc, L1); out(a->dflxc, &ca.dflxc, L1);
    if (a->dudTs) { out(a->duflx_dTs, &ca.duflx_dTs, L1); out(a->duflxc_dTs, &ca.duflxc_dTs, L1); }
    bool any_bo = false;
    for (int b = 0; b < 16; ++b) any_bo |= a->band_output && a->band_output[b];
    if (any_bo) {
        // (16,ncol): untouched bands must survive the round trip, so olrb is staged in as well
        arrs.push_back({a->olrb, (void **)&ca.olrb, 16, esz, true, true, true, f32});
        if (a->dudTs) arrs.push_back({a->dolrb_dTs, (void **)&ca.dolrb_dTs, 16, esz, true, true, true, f32});
    }
    for (int k = 0; k < 3 && nvar; ++k)   // (ncol, nlay+1, nvar): nvar*(nlay+1) rows of ncol
        if (d_var[k]) arrs.push_back({d_var[k], (void **)&d_var[k], (size_t)nvar * L1, esz, false, false, true, f32});
    int rc = run_staged(p, *a, ca, arrs, ncol, chunk, [&](RrtmgxLwArgs &c, int nc, size_t first) -> int {
        (void)nc;
        c.flags |= RRTMGX_DEVICE_PTRS;
        return run_chunks_device(c, first);
    });
    if (rc) return rc;
    p.pending = true;
    p.last_status = status_from(p);
    return p.last_status;
}

int rrtmgx_heating_rate(int ncol, int nlay, const double *fnet_up_minus_down, const double *plev,
                        double *hr_K_per_day, double grav, double cp, int flags, void *stream) {
    if (!g.ready) return RRTMGX_ENOTINIT;
    if (ncol <= 0 || nlay <= 0 || !fnet_up_minus_down || !plev || !hr_K_per_day || cp <= 0.) return RRTMGX_EARG;
    cudaSetDevice(g.device);
    const size_t n = (size_t)ncol * nlay, n1 = (size_t)ncol * (nlay + 1);
    const int blocks = (int)((n + 255) / 256);
    if (flags & RRTMGX_DEVICE_PTRS) {
        cudaStream_t st = stream ? (cudaStream_t)stream : g.lw.stream;
        RRTMGX_LAUNCH(heating_rate_kernel, blocks, 256, 0, st, ncol, nlay, fnet_up_minus_down, plev, hr_K_per_day,
                      grav / cp);
        if (flags & RRTMGX_NO_SYNC) return ok(cudaGetLastError()) ? 0 : RRTMGX_ECUDA;
        return ok(cudaStreamSynchronize(st)) ? 0 : RRTMGX_ECUDA;
    }
    double *d = nullptr;
    if (!ok(cudaMalloc((void **)&d, (2 * n1 + n) * sizeof(double)))) { cudaGetLastError(); return RRTMGX_ECUDA; }
    cudaMemcpy(d, fnet_up_minus_down, n1 * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(d + n1, plev, n1 * 8, cudaMemcpyHostToDevice);
    RRTMGX_LAUNCH(heating_rate_kernel, blocks, 256, 0, g.lw.stream, ncol, nlay, d, d + n1, d + 2 * n1, grav / cp);
    cudaStreamSynchronize(g.lw.stream);
    cudaError_t e = cudaMemcpy(hr_K_per_day, d + 2 * n1, n * 8, cudaMemcpyDeviceToHost);
    cudaFree(d);
    return ok(e) ? 0 : RRTMGX_ECUDA;
}

int rrtmgx_debug_divide(size_t n, const double *a, const double *b, double *q_fast, double *q_ieee,
                        double *r_fast, double *r_ieee) {
    if (!g.ready) return RRTMGX_ENOTINIT;
    if (!n || !a || !b || !q_fast || !q_ieee || !r_fast || !r_ieee) return RRTMGX_EARG;
    cudaSetDevice(g.device);
    double *d = nullptr;
    if (!ok(cudaMalloc((void **)&d, 6 * n * sizeof(double)))) { cudaGetLastError(); return RRTMGX_ECUDA; }
    cudaMemcpy(d, a, n * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(d + n, b, n * 8, cudaMemcpyHostToDevice);
    RRTMGX_LAUNCH(debug_divide_kernel, (unsigned)((n + 255) / 256), 256, 0, g.lw.stream, n, d, d + n, d + 2 * n,
                  d + 3 * n, d + 4 * n, d + 5 * n);
    cudaStreamSynchronize(g.lw.stream);
    cudaMemcpy(q_fast, d + 2 * n, n * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(q_ieee, d + 3 * n, n * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(r_fast, d + 4 * n, n * 8, cudaMemcpyDeviceToHost);
    cudaError_t e = cudaMemcpy(r_ieee, d + 5 * n, n * 8, cudaMemcpyDeviceToHost);
    cudaFree(d);
    return ok(e) ? 0 : RRTMGX_ECUDA;
}

int rrtmgx_debug_kiss(int nstream, const int32_t *seeds, int ndraw, int32_t *kiss, double *ran8, float *ran4, int nsub,
                      int nlay, int inhomo, uint32_t *jumped, uint32_t *replayed, int nvalue, const int32_t *values,
                      double *val8, float *val4) {
    if (!g.ready) return RRTMGX_ENOTINIT;
    if (nstream <= 0 || !seeds || ndraw < 0 || nsub < 0 || nsub > 140 || nvalue < 0 || (ndraw && (!kiss || !ran8 || !ran4)) ||
        (nsub && (nlay <= 0 || !jumped || !replayed)) || (nvalue && (!values || !val8 || !val4)))
        return RRTMGX_EARG;
    cudaSetDevice(g.device);
    return debug_kiss(nstream, seeds, ndraw, kiss, ran8, ran4, nsub, nlay, inhomo, jumped, replayed, nvalue, values, val8,
                      val4, g.lw.stream);
}

#ifndef RRTMGX_WITH_SW
int rrtmgx_sw_run(const RrtmgxSwArgs *) { return RRTMGX_EARG; }
int rrtmgx_sw_run_with_clean(const RrtmgxSwArgs *, const RrtmgxSwNoAerosol *) { return RRTMGX_EARG; }
#else
int rrtmgx_sw_run(const RrtmgxSwArgs *a) { return rrtmgx_sw_run_with_clean(a, nullptr); }

int rrtmgx_sw_run_with_clean(const RrtmgxSwArgs *a, const RrtmgxSwNoAerosol *na) {
    if (!g.ready) return RRTMGX_ENOTINIT;
    // the CUDA current device is per host thread: callers may run LW and SW from different threads
    if (!ok(cudaSetDevice(g.device))) { cudaGetLastError(); return RRTMGX_ENODEVICE; }
    if (!a || a->ncol <= 0 || a->nlay <= 1 || a->nlay > 1000) return RRTMGX_EARG;
    if (a->cloudLM == a->cloudMH) return RRTMGX_ESUPERLAYER;   // cloud_subcol_gen.F90:762-766
    if (a->iceflgsw < 1 || a->iceflgsw > 4) return RRTMGX_EICEFLAG;
    if (a->liqflgsw != 1) return RRTMGX_ELIQFLAG;
    Path &p = g.sw;
    const bool devptr = a->flags & RRTMGX_DEVICE_PTRS;
    if ((a->flags & RRTMGX_NO_SYNC) && !devptr) return RRTMGX_EARG;
    // real*4 arrays are widened while staged: real*4 DEVICE arrays take the staging path too (device-to-device
    // copies instead of PCIe), so the call is synchronous for them
    if ((a->flags & RRTMGX_F32_ARRAYS) && (a->flags & RRTMGX_NO_SYNC)) return RRTMGX_EARG;
    const bool staged = !devptr || (a->flags & RRTMGX_F32_ARRAYS);
    if (na && (!na->swuflx || !na->swdflx || !na->swuflxc || !na->swdflxc || !na->fswband)) return RRTMGX_EARG;
    double *d_na[5] = {na ? na->swuflx : nullptr, na ? na->swdflx : nullptr, na ? na->swuflxc : nullptr,
                       na ? na->swdflxc : nullptr, na ? na->fswband : nullptr};
    SwSolar sol;
    if (int rc = sw_solar_setup(a, g.ht, &sol)) return rc;   // rrtmg_sw_sub :889-1127
    const int ncol = a->ncol, nlay = a->nlay;
    if (int rc = ensure_jumps(p, 112, nlay)) return rc;
    static const int seed_order[4] = {4, 3, 2, 1};   // SW/src/rrtmg_sw_rad.F90:1395-1400
    const McicaParams mp = mcica_params(g.mc, dev_table("mcica.xcw_beta"), dev_table("mcica.xcw_gamma"), a->dyofyr,
                                        seed_order);
    const RrtmgxTaps *taps = p.has_taps ? &p.taps : nullptr;
    const bool dbg = taps && (taps->taug || taps->pfracs || taps->ssi);
    const bool radval = a->radval != nullptr;   // the SOLAR_RADVAL build of rrtmg_sw (include/rrtmgx.h)
    const size_t per_col = sw_scratch_bytes(1024, nlay, dbg, radval) / 1024;
    size_t chunk = pick_chunk(ncol, nlay, per_col + (!staged ? 0 : 2 * (size_t)(57 * nlay + 60 + (radval ? RRTMGX_NRADVAL : 0)) * 8),
                              staged, p.slab.cap, (a->flags & RRTMGX_F32_ARRAYS) != 0);
    if (taps) {   // taps are laid out for the whole call
        if ((size_t)ncol > chunk_cap(nlay)) return RRTMGX_EARG;
        chunk = (size_t)ncol;
    }
    if (int rc = grow(p.slab, sw_scratch_bytes((int)chunk, nlay, dbg, radval))) return rc;
    cudaStream_t stream = (!staged && a->stream) ? (cudaStream_t)a->stream : p.stream;
    p.run_stream = stream;
    if (!(devptr && (a->flags & RRTMGX_KEEP_STATUS)) &&
        !ok(cudaMemcpyAsync(p.d_err, kErrInit, sizeof kErrInit, cudaMemcpyHostToDevice, stream)))
        return RRTMGX_ECUDA;

    const int nchunks_dev = (int)(((size_t)ncol + chunk - 1) / chunk);
    auto run_chunks_device = [&](const RrtmgxSwArgs &da, size_t first) -> int {
        const int n = da.ncol;
        if (!(da.flags & RRTMGX_SKIP_CHECKS)) {   // _ASSERTs :365-383 in the reference's order
            const size_t n2 = (size_t)n * nlay, n2p = (size_t)n * (nlay + 1);
            struct { const double *x; size_t cnt; } chk[] = {
                {da.play, n2}, {da.plev, n2p}, {da.tlay, n2}, {da.h2ovmr, n2}, {da.o3vmr, n2}, {da.co2vmr, n2},
                {da.ch4vmr, n2}, {da.o2vmr, n2}, {da.asdir, (size_t)n}, {da.aldir, (size_t)n}, {da.asdif, (size_t)n},
                {da.aldif, (size_t)n}, {da.cld, n2}, {da.ciwp, n2}, {da.clwp, n2}, {da.rei, n2}, {da.rel, n2},
                {da.tauaer, n2 * 14}, {da.ssaaer, n2 * 14}};
            constexpr int narr = (int)(sizeof chk / sizeof chk[0]);
            const double *xs[narr];
            size_t cnts[narr];
            for (int i = 0; i < narr; ++i) { xs[i] = chk[i].x; cnts[i] = chk[i].cnt; }
            launch_check_negative(xs, cnts, narr, p.d_err + 1, stream, true);
        }
        for (size_t col0 = 0; col0 < (size_t)n; col0 += chunk) {
            const int nc = (int)std::min(chunk, (size_t)n - col0);
            const ChunkId id{(long long)(first + col0), (long long)ncol, nchunks_dev};
            if (na) {   // the no-aerosol run of the chunk first (SOL:3249-3258), its fluxes into the NA arrays
                RrtmgxSwArgs dv = da;
                dv.iaer = 0;
                dv.swuflx = d_na[0]; dv.swdflx = d_na[1]; dv.swuflxc = d_na[2]; dv.swdflxc = d_na[3]; dv.fswband = d_na[4];
                if (int rc = sw_run_chunk(&dv, sol, (int)col0, nc, id, mp, p.d_jumps, p.slab, p.d_err, stream, p.side, NSIDE,
                                          p.ev, taps, p.d_err + 1))
                    return rc;
            }
            RrtmgxSwArgs dm = da;
            if (na) dm.flags |= RRTMGX_REUSE_CLOUDS;
            if (int rc = sw_run_chunk(&dm, sol, (int)col0, nc, id, mp, p.d_jumps, p.slab, p.d_err, stream, p.side, NSIDE,
                                      p.ev, taps, p.d_err + 1))
                return rc;
        }
        return 0;
    };

    if (!staged) {
        if (int rc = run_chunks_device(*a, 0)) return rc;
        p.pending = true;
        if (a->flags & RRTMGX_NO_SYNC) return 0;
        p.last_status = status_from(p);
        return p.last_status;
    }

    RrtmgxSwArgs ca = *a;
    std::vector<Arr> arrs;
    const bool f32 = a->flags & RRTMGX_F32_ARRAYS;   // the caller's real arrays are real*4
    const size_t esz = f32 ? 4 : 8;
    const size_t L = nlay, L1 = nlay + 1;
    auto in = [&](const double *const &field, const double **slot, size_t rows) {
        arrs.push_back({field, (void **)slot, rows, esz, false, true, false, f32});
    };
    auto out = [&](double *const &field, double **slot, size_t rows) {
        arrs.push_back({field, (void **)slot, rows, esz, false, false, true, f32});
    };
    in(a->coszen, &ca.coszen, 1); in(a->play, &ca.play, L); in(a->plev, &ca.plev, L1); in(a->tlay, &ca.tlay, L);
    in(a->h2ovmr, &ca.h2ovmr, L); in(a->o3vmr, &ca.o3vmr, L); in(a->co2vmr, &ca.co2vmr, L);
    in(a->ch4vmr, &ca.ch4vmr, L); in(a->o2vmr, &ca.o2vmr, L);
    in(a->cld, &ca.cld, L); in(a->ciwp, &ca.ciwp, L); in(a->clwp, &ca.clwp, L); in(a->rei, &ca.rei, L);
    in(a->rel, &ca.rel, L); in(a->zm, &ca.zm, L); in(a->alat, &ca.alat, 1);
    in(a->tauaer, &ca.tauaer, L * 14); in(a->ssaaer, &ca.ssaaer, L * 14); in(a->asmaer, &ca.asmaer, L * 14);
    in(a->asdir, &ca.asdir, 1); in(a->asdif, &ca.asdif, 1); in(a->aldir, &ca.aldir, 1); in(a->aldif, &ca.aldif, 1);
    arrs.push_back({a->clearCounts, (void **)&ca.clearCounts, 4, 4, false, false, true});
    out(a->swuflx, &ca.swuflx, L1); out(a->swdflx, &ca.swdflx, L1); out(a->swuflxc, &ca.swuflxc, L1);
    out(a->swdflxc, &ca.swdflxc, L1);
    out(a->nirr, &ca.nirr, 1); out(a->nirf, &ca.nirf, 1); out(a->parr, &ca.parr, 1); out(a->parf, &ca.parf, 1);
    out(a->uvrr, &ca.uvrr, 1); out(a->uvrf, &ca.uvrf, 1); out(a->fswband, &ca.fswband, 14);
    out(a->cotdtp, &ca.cotdtp, 1); out(a->cotdhp, &ca.cotdhp, 1); out(a->cotdmp, &ca.cotdmp, 1);
    out(a->cotdlp, &ca.cotdlp, 1); out(a->cotntp, &ca.cotntp, 1); out(a->cotnhp, &ca.cotnhp, 1);
    out(a->cotnmp, &ca.cotnmp, 1); out(a->cotnlp, &ca.cotnlp, 1);
    if (a->do_drfband) { out(a->drband, &ca.drband, 14); out(a->dfband, &ca.dfband, 14); }
    if (radval) out(a->radval, &ca.radval, RRTMGX_NRADVAL);
    if (na) {
        for (int k = 0; k < 4; ++k) arrs.push_back({d_na[k], (void **)&d_na[k], L1, esz, false, false, true, f32});
        arrs.push_back({d_na[4], (void **)&d_na[4], 14, esz, false, false, true, f32});
    }
    int rc = run_staged(p, *a, ca, arrs, ncol, chunk, [&](RrtmgxSwArgs &c, int nc, size_t first) -> int {
        (void)nc;
        c.flags |= RRTMGX_DEVICE_PTRS;
        return run_chunks_device(c, first);
    });
    if (rc) return rc;
    p.pending = true;
    p.last_status = status_from(p);
    return p.last_status;
}
#endif

}  // extern "C"

// ---- fused Run-phase glue (include/rrtmgx.h; kernels in glue.cuh) -------------------------------
namespace {
// the argument arrays of one rrtmg_lw / rrtmg_sw call over nc columns, carved from a slab
void carve_lw_args(Slab &s, int nc, int L, RrtmgxLwArgs &a) {
    const size_t n2 = (size_t)nc * L, n2p = (size_t)nc * (L + 1);
    auto D = [&](size_t n) { return s.take<double>(n); };
    a.play = D(n2); a.plev = D(n2p); a.tlay = D(n2); a.tlev = D(n2p); a.tsfc = D(nc); a.emis = D((size_t)nc * 16);
    a.h2ovmr = D(n2); a.o3vmr = D(n2); a.co2vmr = D(n2); a.ch4vmr = D(n2); a.n2ovmr = D(n2); a.o2vmr = D(n2);
    a.cfc11vmr = D(n2); a.cfc12vmr = D(n2); a.cfc22vmr = D(n2); a.ccl4vmr = D(n2);
    a.cldf = D(n2); a.ciwp = D(n2); a.clwp = D(n2); a.rei = D(n2); a.rel = D(n2);
    a.tauaer = D(n2 * 16); a.zm = D(n2); a.alat = D(nc);
    a.clearCounts = s.take<int32_t>((size_t)nc * 4);
    a.uflx = D(n2p); a.dflx = D(n2p); a.uflxc = D(n2p); a.dflxc = D(n2p); a.duflx_dTs = D(n2p); a.duflxc_dTs = D(n2p);
    a.olrb = D((size_t)nc * 16); a.dolrb_dTs = D((size_t)nc * 16);
}
void carve_sw_args(Slab &s, int nc, int L, RrtmgxSwArgs &a) {
    const size_t n2 = (size_t)nc * L, n2p = (size_t)nc * (L + 1);
    auto D = [&](size_t n) { return s.take<double>(n); };
    a.coszen = D(nc); a.play = D(n2); a.plev = D(n2p); a.tlay = D(n2);
    a.h2ovmr = D(n2); a.o3vmr = D(n2); a.co2vmr = D(n2); a.ch4vmr = D(n2); a.o2vmr = D(n2);
    a.cld = D(n2); a.ciwp = D(n2); a.clwp = D(n2); a.rei = D(n2); a.rel = D(n2); a.zm = D(n2); a.alat = D(nc);
    a.tauaer = D(n2 * 14); a.ssaaer = D(n2 * 14); a.asmaer = D(n2 * 14);
    a.asdir = D(nc); a.asdif = D(nc); a.aldir = D(nc); a.aldif = D(nc);
    a.clearCounts = s.take<int32_t>((size_t)nc * 4);
    a.swuflx = D(n2p); a.swdflx = D(n2p); a.swuflxc = D(n2p); a.swdflxc = D(n2p);
    a.nirr = D(nc); a.nirf = D(nc); a.parr = D(nc); a.parf = D(nc); a.uvrr = D(nc); a.uvrf = D(nc);
    a.fswband = D((size_t)nc * 14);
    a.cotdtp = D(nc); a.cotdhp = D(nc); a.cotdmp = D(nc); a.cotdlp = D(nc);
    a.cotntp = D(nc); a.cotnhp = D(nc); a.cotnmp = D(nc); a.cotnlp = D(nc);
    a.drband = nullptr; a.dfband = nullptr;
    a.radval = nullptr;
}
template <class A, class Carve> size_t carve_bytes(int nc, int L, Carve carve) {
    Slab s;
    A a{};
    carve(s, nc, L, a);
    return s.used + 4096;
}

void lw_scalars(const RrtmgxIrradArgs &S, int nc, RrtmgxLwArgs &L) {
    L.ncol = nc; L.nlay = S.lm; L.psize = 0; L.dudTs = 1;            // Ts_derivs = .true., IRR:3232
    L.iceflglw = S.iceflg; L.liqflglw = S.liqflg; L.dyofyr = S.doy;
    L.cloudMH = S.lm - S.lcldmh + 1; L.cloudLM = S.lm - S.lcldlm + 1;   // IRR:3239-3240
    L.band_output = S.band_output;
}
void sw_scalars(const RrtmgxSolarArgs &S, int nc, RrtmgxSwArgs &L) {
    L.ncol = nc; L.nlay = S.lm; L.rpart = 0; L.isolvar = S.isolvar;
    L.iceflgsw = S.iceflg; L.liqflgsw = S.liqflg; L.dyofyr = S.doy;
    L.cloudLM = S.lm - S.lcldlm + 1; L.cloudMH = S.lm - S.lcldmh + 1;   // SOL:6347
    L.iaer = 10; L.normFlx = 1; L.do_drfband = 0;                       // SOL:6234-6240
    L.scon = S.sc; L.adjes = S.dist; L.bndscl = nullptr; L.indsolvar = nullptr; L.solcycfrac = S.solcycfrac;
}
int band_mask_of(const int32_t *bo) {
    int m = 0;
    for (int b = 0; b < 16; ++b) m |= (bo && bo[b]) ? 1 << b : 0;
    return m;
}

// one chunk, everything on the device: native arrays `S` (leading dimension lds, first column col0)
// -> prepare -> rrtmg_lw -> finish -> native outputs
int irrad_chunk(const RrtmgxIrradArgs &S, int lds, int col0, int nc, cudaStream_t st) {
    Path &p = g.lw;
    if (int rc = grow(p.glue, carve_bytes<RrtmgxLwArgs>(nc, S.lm, carve_lw_args))) return rc;
    p.glue.used = 0;
    RrtmgxLwArgs L{};
    lw_scalars(S, nc, L);
    carve_lw_args(p.glue, nc, S.lm, L);
    L.flags = RRTMGX_DEVICE_PTRS | RRTMGX_NO_SYNC | RRTMGX_KEEP_STATUS | (S.flags & RRTMGX_SKIP_CHECKS);
    L.stream = st;
    RRTMGX_LAUNCH(irrad_prepare_kernel, (nc + 127) / 128, 128, 0, st, nc, lds, col0, S, L);
    if (int rc = rrtmgx_lw_run(&L)) return rc;
    RRTMGX_LAUNCH(irrad_finish_kernel, (nc + 127) / 128, 128, 0, st, nc, lds, col0, S, L, band_mask_of(S.band_output));
    lw_forget_clouds();   // the glue workspace is rewritten per chunk: nothing a later RRTMGX_REUSE_CLOUDS call may keep
    return ok(cudaGetLastError()) ? 0 : RRTMGX_ECUDA;
}
#ifdef RRTMGX_WITH_SW
int solar_chunk(const RrtmgxSolarArgs &S, int lds, int col0, int nc, cudaStream_t st, const int *lit = nullptr) {
    Path &p = g.sw;
    if (int rc = grow(p.glue, carve_bytes<RrtmgxSwArgs>(nc, S.lm, carve_sw_args))) return rc;
    p.glue.used = 0;
    RrtmgxSwArgs L{};
    sw_scalars(S, nc, L);
    carve_sw_args(p.glue, nc, S.lm, L);
    L.flags = RRTMGX_DEVICE_PTRS | RRTMGX_NO_SYNC | RRTMGX_KEEP_STATUS | (S.flags & RRTMGX_SKIP_CHECKS);
    L.stream = st;
    RRTMGX_LAUNCH(solar_prepare_kernel, (nc + 127) / 128, 128, 0, st, nc, lds, col0, S, L, lit);
    if (int rc = rrtmgx_sw_run(&L)) return rc;
    RRTMGX_LAUNCH(solar_finish_kernel, (nc + 127) / 128, 128, 0, st, nc, lds, col0, S, L, lit);
    sw_forget_clouds();
    return ok(cudaGetLastError()) ? 0 : RRTMGX_ECUDA;
}
#endif
// RRTMGX_LIT_ONLY: the daytime columns of native columns [first, first + n) of a Solar call, ascending, as offsets
// into the arrays the kernels will see (`base` + local index), uploaded to g.sw.lit on `st`.  `zt` is the caller's
// array (host or device, real*8 or real*4).  Returns the number of lit columns, or a negative status.
int solar_lit_index(const void *zt, bool on_device, bool f32, size_t first, int n, int base, cudaStream_t st,
                    const int **d_lit) {
    std::vector<double> z8;
    std::vector<float> z4;
    const void *h = nullptr;
    const size_t esz = f32 ? 4 : 8;
    if (on_device) {
        if (f32) z4.resize(n); else z8.resize(n);
        void *dst = f32 ? (void *)z4.data() : (void *)z8.data();
        if (!ok(cudaMemcpyAsync(dst, (const char *)zt + first * esz, (size_t)n * esz, cudaMemcpyDeviceToHost, st)) ||
            !ok(cudaStreamSynchronize(st))) { cudaGetLastError(); return -RRTMGX_ECUDA; }
        h = dst;
    } else {
        h = (const char *)zt + first * esz;
    }
    std::vector<int> lit;
    lit.reserve(n);
    for (int c = 0; c < n; ++c) {
        const double z = f32 ? (double)((const float *)h)[c] : ((const double *)h)[c];
        if (z > 0.) lit.push_back(base + c);   // daytime = ZTH > 0., SOL:3686
    }
    if (int rc = grow(g_sw_lit, (size_t)n * sizeof(int) + 4096)) return -rc;
    if (!lit.empty() &&
        !ok(cudaMemcpyAsync(g_sw_lit.base, lit.data(), lit.size() * sizeof(int), cudaMemcpyHostToDevice, st))) {
        cudaGetLastError();
        return -RRTMGX_ECUDA;
    }
    // the host vector goes out of scope: the copy from pageable memory has been staged when the call returns
    *d_lit = (const int *)g_sw_lit.base;
    return (int)lit.size();
}

constexpr size_t kGlueChunk = 131072;   // columns per pass of the device-pointer refresh (workspace ~25-35 KB each)

template <class A> bool glue_args_ok(const A *a) {
    return a && a->ncol > 0 && a->lm >= 2 && a->lm <= 1000 && a->ple && a->pl && a->t && a->q && a->o3 && a->ch4 &&
           a->qliq && a->qice && a->rliq && a->rice && a->ts && a->lats;
}
}  // namespace

extern "C" {

int rrtmgx_irrad_refresh(const RrtmgxIrradArgs *a) {
    if (!g.ready) return RRTMGX_ENOTINIT;
    if (!ok(cudaSetDevice(g.device))) { cudaGetLastError(); return RRTMGX_ENODEVICE; }
    if (!glue_args_ok(a) || !a->n2o || !a->cfc11 || !a->cfc12 || !a->hcfc22 || !a->fcld || !a->t2m || !a->emis ||
        !a->flxu || !a->flxd || !a->flcu || !a->flcd || !a->dfdts || !a->dfdtsc || !a->sfcem || (!a->taua != !a->ssaa))
        return RRTMGX_EARG;
    Path &p = g.lw;
    const bool devptr = a->flags & RRTMGX_DEVICE_PTRS;
    if ((a->flags & RRTMGX_NO_SYNC) && !devptr) return RRTMGX_EARG;
    // real*4 arrays are widened while staged: real*4 DEVICE arrays take the staging path too (device-to-device
    // copies instead of PCIe), so the call is synchronous for them
    if ((a->flags & RRTMGX_F32_ARRAYS) && (a->flags & RRTMGX_NO_SYNC)) return RRTMGX_EARG;
    const bool staged = !devptr || (a->flags & RRTMGX_F32_ARRAYS);
    const int ncol = a->ncol;
    cudaStream_t st = (!staged && a->stream) ? (cudaStream_t)a->stream : p.stream;
    if (!ok(cudaMemcpyAsync(p.d_err, kErrInit, sizeof kErrInit, cudaMemcpyHostToDevice, st))) return RRTMGX_ECUDA;
    if (!staged) {
        for (size_t col0 = 0; col0 < (size_t)ncol; col0 += kGlueChunk)
            if (int rc = irrad_chunk(*a, ncol, (int)col0, (int)std::min(kGlueChunk, (size_t)ncol - col0), st)) return rc;
        p.pending = true;
        if (a->flags & RRTMGX_NO_SYNC) return 0;
        return p.last_status = status_from(p);
    }
    RrtmgxIrradArgs ca = *a;
    std::vector<Arr> arrs;
    const bool f32 = a->flags & RRTMGX_F32_ARRAYS;   // the caller's real arrays are real*4
    const size_t esz = f32 ? 4 : 8;
    const size_t L = a->lm, L1 = a->lm + 1;
    auto in = [&](const double *const &f, const double **slot, size_t rows) {
        arrs.push_back({f, (void **)slot, rows, esz, false, true, false, f32});
    };
    auto out = [&](double *const &f, double **slot, size_t rows) {
        arrs.push_back({f, (void **)slot, rows, esz, false, false, true, f32});
    };
    in(a->ple, &ca.ple, L1); in(a->pl, &ca.pl, L); in(a->t, &ca.t, L); in(a->q, &ca.q, L); in(a->o3, &ca.o3, L);
    in(a->ch4, &ca.ch4, L); in(a->n2o, &ca.n2o, L); in(a->co2, &ca.co2, L); in(a->cfc11, &ca.cfc11, L);
    in(a->cfc12, &ca.cfc12, L); in(a->hcfc22, &ca.hcfc22, L); in(a->fcld, &ca.fcld, L);
    in(a->qliq, &ca.qliq, L); in(a->qice, &ca.qice, L); in(a->rliq, &ca.rliq, L); in(a->rice, &ca.rice, L);
    in(a->ts, &ca.ts, 1); in(a->t2m, &ca.t2m, 1); in(a->emis, &ca.emis, 1); in(a->lats, &ca.lats, 1);
    in(a->taua, &ca.taua, L * 16); in(a->ssaa, &ca.ssaa, L * 16);
    out(a->flxu, &ca.flxu, L1); out(a->flxd, &ca.flxd, L1); out(a->flcu, &ca.flcu, L1); out(a->flcd, &ca.flcd, L1);
    out(a->dfdts, &ca.dfdts, L1); out(a->dfdtsc, &ca.dfdtsc, L1); out(a->sfcem, &ca.sfcem, 1);
    out(a->cldtt, &ca.cldtt, 1); out(a->cldhi, &ca.cldhi, 1); out(a->cldmd, &ca.cldmd, 1); out(a->cldlo, &ca.cldlo, 1);
    if (band_mask_of(a->band_output)) {   // (16,ncol): untouched bands survive the round trip
        arrs.push_back({a->olrb, (void **)&ca.olrb, 16, esz, true, true, true, f32});
        arrs.push_back({a->dolrb_dts, (void **)&ca.dolrb_dts, 16, esz, true, true, true, f32});
    } else {
        ca.olrb = nullptr; ca.dolrb_dts = nullptr;
    }
    const size_t chunk = std::min<size_t>(host_chunk(f32), (size_t)ncol);
    int rc = run_staged(p, *a, ca, arrs, ncol, chunk, [&](RrtmgxIrradArgs &c, int nc, size_t) -> int {
        return irrad_chunk(c, nc, 0, nc, p.stream);
    });
    if (rc) return rc;
    p.pending = true;
    return p.last_status = status_from(p);
}

int rrtmgx_irrad_prepare(const RrtmgxIrradArgs *a, RrtmgxLwArgs *lw) {
    if (!g.ready) return RRTMGX_ENOTINIT;
    if (!ok(cudaSetDevice(g.device))) { cudaGetLastError(); return RRTMGX_ENODEVICE; }
    if (!glue_args_ok(a) || !lw || !a->n2o || !a->cfc11 || !a->cfc12 || !a->hcfc22 || !a->fcld || !a->t2m || !a->emis)
        return RRTMGX_EARG;
    Path &p = g.lw;
    const int nc = a->ncol, L = a->lm;
    lw_scalars(*a, nc, *lw);
    if (a->flags & RRTMGX_DEVICE_PTRS) {
        cudaStream_t st = a->stream ? (cudaStream_t)a->stream : p.stream;
        RRTMGX_LAUNCH(irrad_prepare_kernel, (nc + 127) / 128, 128, 0, st, nc, nc, 0, *a, *lw);
        if (a->flags & RRTMGX_NO_SYNC) return ok(cudaGetLastError()) ? 0 : RRTMGX_ECUDA;
        return ok(cudaStreamSynchronize(st)) ? 0 : RRTMGX_ECUDA;
    }
    // host arrays: native state up, prepared arguments down (whole call at once: a test / staging utility)
    const size_t n2 = (size_t)nc * L, n2p = (size_t)nc * (L + 1);
    Slab nat;
    RrtmgxIrradArgs d = *a;
    struct In { const double **f; size_t n; } ins[] = {
        {&d.ple, n2p}, {&d.pl, n2}, {&d.t, n2}, {&d.q, n2}, {&d.o3, n2}, {&d.ch4, n2}, {&d.n2o, n2}, {&d.co2, n2},
        {&d.cfc11, n2}, {&d.cfc12, n2}, {&d.hcfc22, n2}, {&d.fcld, n2}, {&d.qliq, n2}, {&d.qice, n2}, {&d.rliq, n2},
        {&d.rice, n2}, {&d.ts, (size_t)nc}, {&d.t2m, (size_t)nc}, {&d.emis, (size_t)nc}, {&d.lats, (size_t)nc},
        {&d.taua, n2 * 16}, {&d.ssaa, n2 * 16}};
    size_t bytes = 4096;
    for (auto &i : ins) bytes += *i.f ? ((i.n * 8 + 255) & ~(size_t)255) : 0;
    if (int rc = grow(nat, bytes)) return rc;
    for (auto &i : ins)
        if (*i.f) {
            double *dev = nat.take<double>(i.n);
            cudaMemcpyAsync(dev, *i.f, i.n * 8, cudaMemcpyHostToDevice, p.stream);
            *i.f = dev;
        }
    int rc = grow(p.glue, carve_bytes<RrtmgxLwArgs>(nc, L, carve_lw_args));
    if (!rc) {
        p.glue.used = 0;
        RrtmgxLwArgs D = *lw;
        carve_lw_args(p.glue, nc, L, D);
        RRTMGX_LAUNCH(irrad_prepare_kernel, (nc + 127) / 128, 128, 0, p.stream, nc, nc, 0, d, D);
        struct Out { const double *src; const double *dst; size_t n; } outs[] = {
            {D.play, lw->play, n2}, {D.plev, lw->plev, n2p}, {D.tlay, lw->tlay, n2}, {D.tlev, lw->tlev, n2p},
            {D.tsfc, lw->tsfc, (size_t)nc}, {D.emis, lw->emis, (size_t)nc * 16}, {D.h2ovmr, lw->h2ovmr, n2},
            {D.o3vmr, lw->o3vmr, n2}, {D.co2vmr, lw->co2vmr, n2}, {D.ch4vmr, lw->ch4vmr, n2}, {D.n2ovmr, lw->n2ovmr, n2},
            {D.o2vmr, lw->o2vmr, n2}, {D.cfc11vmr, lw->cfc11vmr, n2}, {D.cfc12vmr, lw->cfc12vmr, n2},
            {D.cfc22vmr, lw->cfc22vmr, n2}, {D.ccl4vmr, lw->ccl4vmr, n2}, {D.cldf, lw->cldf, n2}, {D.ciwp, lw->ciwp, n2},
            {D.clwp, lw->clwp, n2}, {D.rei, lw->rei, n2}, {D.rel, lw->rel, n2}, {D.tauaer, lw->tauaer, n2 * 16},
            {D.zm, lw->zm, n2}, {D.alat, lw->alat, (size_t)nc}};
        for (auto &o : outs)
            if (o.dst) cudaMemcpyAsync(const_cast<double *>(o.dst), o.src, o.n * 8, cudaMemcpyDeviceToHost, p.stream);
        if (!ok(cudaStreamSynchronize(p.stream))) rc = RRTMGX_ECUDA;
    }
    cudaFree(nat.base);
    return rc;
}

int rrtmgx_irrad_update(const RrtmgxIrradUpdateArgs *u) {
    if (!g.ready) return RRTMGX_ENOTINIT;
    if (!ok(cudaSetDevice(g.device))) { cudaGetLastError(); return RRTMGX_ENODEVICE; }
    if (!u || u->ncol <= 0 || u->lm < 1 || !u->flxu_int || !u->flxd_int || !u->flcu_int || !u->flcd_int || !u->dfdts ||
        !u->dfdtsc || !u->sfcem_int || !u->ts_int || !u->tsinst)
        return RRTMGX_EARG;
    Path &p = g.lw;
    const int nc = u->ncol, L1 = u->lm + 1;
    const dim3 grid((nc + 255) / 256, L1);
    if (u->flags & RRTMGX_DEVICE_PTRS) {
        cudaStream_t st = u->stream ? (cudaStream_t)u->stream : p.stream;
        RRTMGX_LAUNCH(irrad_update_kernel, grid, 256, 0, st, *u);
        if (u->flags & RRTMGX_NO_SYNC) return ok(cudaGetLastError()) ? 0 : RRTMGX_ECUDA;
        return ok(cudaStreamSynchronize(st)) ? 0 : RRTMGX_ECUDA;
    }
    // host arrays: a handful of (ncol, LM+1) fields up, the requested exports down
    RrtmgxIrradUpdateArgs d = *u;
    const size_t n3 = (size_t)nc * L1, n1 = (size_t)nc;
    struct F { const double **in; double **out; size_t n; };
    F fields[] = {{&d.flxu_int, nullptr, n3}, {&d.flxd_int, nullptr, n3}, {&d.flcu_int, nullptr, n3},
                  {&d.flcd_int, nullptr, n3}, {&d.dfdts, nullptr, n3}, {&d.dfdtsc, nullptr, n3},
                  {&d.sfcem_int, nullptr, n1}, {&d.ts_int, nullptr, n1}, {&d.tsinst, nullptr, n1},
                  {nullptr, &d.flx, n3}, {nullptr, &d.flc, n3}, {nullptr, &d.flxu, n3}, {nullptr, &d.flcu, n3},
                  {nullptr, &d.flxd, n3}, {nullptr, &d.flcd, n3}, {nullptr, &d.olr, n1}, {nullptr, &d.olc, n1},
                  {nullptr, &d.sfcem, n1}, {nullptr, &d.lws, n1}, {nullptr, &d.lcs, n1}, {nullptr, &d.flns, n1},
                  {nullptr, &d.flnsc, n1}};
    size_t bytes = 4096;
    for (auto &f : fields) bytes += (f.n * 8 + 255) & ~(size_t)255;
    Slab tmp;
    if (int rc = grow(tmp, bytes)) return rc;
    double *host_out[22] = {};
    int k = 0;
    for (auto &f : fields) {
        double *dev = tmp.take<double>(f.n);
        if (f.in) {
            cudaMemcpyAsync(dev, *f.in, f.n * 8, cudaMemcpyHostToDevice, p.stream);
            *f.in = dev;
        } else if (*f.out) {
            host_out[k] = *f.out;
            *f.out = dev;
        }
        ++k;
    }
    RRTMGX_LAUNCH(irrad_update_kernel, grid, 256, 0, p.stream, d);
    k = 0;
    for (auto &f : fields) {
        if (f.out && host_out[k]) cudaMemcpyAsync(host_out[k], *f.out, f.n * 8, cudaMemcpyDeviceToHost, p.stream);
        ++k;
    }
    const bool good = ok(cudaStreamSynchronize(p.stream));
    cudaFree(tmp.base);
    return good ? 0 : RRTMGX_ECUDA;
}

#ifndef RRTMGX_WITH_SW
int rrtmgx_solar_refresh(const RrtmgxSolarArgs *) { return RRTMGX_EARG; }
int rrtmgx_solar_prepare(const RrtmgxSolarArgs *, RrtmgxSwArgs *) { return RRTMGX_EARG; }
#else
int rrtmgx_solar_refresh(const RrtmgxSolarArgs *a) {
    if (!g.ready) return RRTMGX_ENOTINIT;
    if (!ok(cudaSetDevice(g.device))) { cudaGetLastError(); return RRTMGX_ENODEVICE; }
    if (!glue_args_ok(a) || !a->cl || !a->zt || !a->albvr || !a->albvf || !a->albnr || !a->albnf || !a->fsw || !a->fsc ||
        !a->fswu || !a->fscu || (!a->taua != !a->ssaa) || (!a->taua != !a->asya))
        return RRTMGX_EARG;
    Path &p = g.sw;
    const bool devptr = a->flags & RRTMGX_DEVICE_PTRS;
    if ((a->flags & RRTMGX_NO_SYNC) && !devptr) return RRTMGX_EARG;
    // real*4 arrays are widened while staged: real*4 DEVICE arrays take the staging path too (device-to-device
    // copies instead of PCIe), so the call is synchronous for them
    if ((a->flags & RRTMGX_F32_ARRAYS) && (a->flags & RRTMGX_NO_SYNC)) return RRTMGX_EARG;
    const bool lit_only = a->flags & RRTMGX_LIT_ONLY;
    if (lit_only && (a->flags & RRTMGX_NO_SYNC)) return RRTMGX_EARG;
    const bool staged = !devptr || (a->flags & RRTMGX_F32_ARRAYS);
    const int ncol = a->ncol;
    cudaStream_t st = (!staged && a->stream) ? (cudaStream_t)a->stream : p.stream;
    if (!ok(cudaMemcpyAsync(p.d_err, kErrInit, sizeof kErrInit, cudaMemcpyHostToDevice, st))) return RRTMGX_ECUDA;
    if (!staged && lit_only) {
        // pack the daytime columns (SOL:3686-3687, PackIt :7753-7773): the glue kernels address the native arrays
        // through the index list, the rrtmg_sw arguments in between hold lit columns only
        const int *d_lit = nullptr;
        const int nlit = solar_lit_index(a->zt, true, false, 0, ncol, 0, st, &d_lit);
        if (nlit < 0) return -nlit;
        RRTMGX_LAUNCH(solar_night_kernel, (ncol + 127) / 128, 128, 0, st, ncol, ncol, *a);
        for (size_t off = 0; off < (size_t)nlit; off += kGlueChunk)
            if (int rc = solar_chunk(*a, ncol, 0, (int)std::min(kGlueChunk, (size_t)nlit - off), st, d_lit + off)) return rc;
        p.pending = true;
        return p.last_status = status_from(p);
    }
    if (!staged) {
        for (size_t col0 = 0; col0 < (size_t)ncol; col0 += kGlueChunk)
            if (int rc = solar_chunk(*a, ncol, (int)col0, (int)std::min(kGlueChunk, (size_t)ncol - col0), st)) return rc;
        p.pending = true;
        if (a->flags & RRTMGX_NO_SYNC) return 0;
        return p.last_status = status_from(p);
    }
    RrtmgxSolarArgs ca = *a;
    std::vector<Arr> arrs;
    const bool f32 = a->flags & RRTMGX_F32_ARRAYS;   // the caller's real arrays are real*4
    const size_t esz = f32 ? 4 : 8;
    const size_t L = a->lm, L1 = a->lm + 1;
    auto in = [&](const double *const &f, const double **slot, size_t rows) {
        arrs.push_back({f, (void **)slot, rows, esz, false, true, false, f32});
    };
    auto out = [&](double *const &f, double **slot, size_t rows) {
        arrs.push_back({f, (void **)slot, rows, esz, false, false, true, f32});
    };
    in(a->ple, &ca.ple, L1); in(a->pl, &ca.pl, L); in(a->t, &ca.t, L); in(a->q, &ca.q, L); in(a->o3, &ca.o3, L);
    in(a->ch4, &ca.ch4, L); in(a->cl, &ca.cl, L); in(a->qliq, &ca.qliq, L); in(a->qice, &ca.qice, L);
    in(a->rliq, &ca.rliq, L); in(a->rice, &ca.rice, L); in(a->ts, &ca.ts, 1); in(a->zt, &ca.zt, 1);
    in(a->lats, &ca.lats, 1); in(a->albvr, &ca.albvr, 1); in(a->albvf, &ca.albvf, 1); in(a->albnr, &ca.albnr, 1);
    in(a->albnf, &ca.albnf, 1); in(a->taua, &ca.taua, L * 14); in(a->ssaa, &ca.ssaa, L * 14); in(a->asya, &ca.asya, L * 14);
    out(a->fsw, &ca.fsw, L1); out(a->fsc, &ca.fsc, L1); out(a->fswu, &ca.fswu, L1); out(a->fscu, &ca.fscu, L1);
    out(a->nirr, &ca.nirr, 1); out(a->nirf, &ca.nirf, 1); out(a->parr, &ca.parr, 1); out(a->parf, &ca.parf, 1);
    out(a->uvrr, &ca.uvrr, 1); out(a->uvrf, &ca.uvrf, 1); out(a->fswband, &ca.fswband, 14);
    out(a->cldts, &ca.cldts, 1); out(a->cldhs, &ca.cldhs, 1); out(a->cldms, &ca.cldms, 1); out(a->cldls, &ca.cldls, 1);
    out(a->cottp, &ca.cottp, 1); out(a->cothp, &ca.cothp, 1); out(a->cotmp, &ca.cotmp, 1); out(a->cotlp, &ca.cotlp, 1);
    const size_t chunk = std::min<size_t>(host_chunk(f32), (size_t)ncol);
    int rc = run_staged(p, *a, ca, arrs, ncol, chunk, [&](RrtmgxSolarArgs &c, int nc, size_t first) -> int {
        if (!lit_only) return solar_chunk(c, nc, 0, nc, p.stream);
        // the staged chunk holds native columns [first, first + nc): its daytime columns, from the caller's own ZTH
        const int *d_lit = nullptr;
        const int nlit = solar_lit_index(a->zt, devptr, f32, first, nc, 0, p.stream, &d_lit);
        if (nlit < 0) return -nlit;
        RRTMGX_LAUNCH(solar_night_kernel, (nc + 127) / 128, 128, 0, p.stream, nc, nc, c);
        return nlit ? solar_chunk(c, nc, 0, nlit, p.stream, d_lit) : 0;
    });
    if (rc) return rc;
    p.pending = true;
    return p.last_status = status_from(p);
}

int rrtmgx_solar_prepare(const RrtmgxSolarArgs *a, RrtmgxSwArgs *sw) {
    if (!g.ready) return RRTMGX_ENOTINIT;
    if (!ok(cudaSetDevice(g.device))) { cudaGetLastError(); return RRTMGX_ENODEVICE; }
    if (!glue_args_ok(a) || !sw || !a->cl || !a->zt || !a->albvr || !a->albvf || !a->albnr || !a->albnf) return RRTMGX_EARG;
    Path &p = g.sw;
    const int nc = a->ncol, L = a->lm;
    sw_scalars(*a, nc, *sw);
    if (a->flags & RRTMGX_DEVICE_PTRS) {
        cudaStream_t st = a->stream ? (cudaStream_t)a->stream : p.stream;
        RRTMGX_LAUNCH(solar_prepare_kernel, (nc + 127) / 128, 128, 0, st, nc, nc, 0, *a, *sw, (const int *)nullptr);
        if (a->flags & RRTMGX_NO_SYNC) return ok(cudaGetLastError()) ? 0 : RRTMGX_ECUDA;
        return ok(cudaStreamSynchronize(st)) ? 0 : RRTMGX_ECUDA;
    }
    const size_t n2 = (size_t)nc * L, n2p = (size_t)nc * (L + 1);
    Slab nat;
    RrtmgxSolarArgs d = *a;
    struct In { const double **f; size_t n; } ins[] = {
        {&d.ple, n2p}, {&d.pl, n2}, {&d.t, n2}, {&d.q, n2}, {&d.o3, n2}, {&d.ch4, n2}, {&d.cl, n2}, {&d.qliq, n2},
        {&d.qice, n2}, {&d.rliq, n2}, {&d.rice, n2}, {&d.ts, (size_t)nc}, {&d.zt, (size_t)nc}, {&d.lats, (size_t)nc},
        {&d.albvr, (size_t)nc}, {&d.albvf, (size_t)nc}, {&d.albnr, (size_t)nc}, {&d.albnf, (size_t)nc},
        {&d.taua, n2 * 14}, {&d.ssaa, n2 * 14}, {&d.asya, n2 * 14}};
    size_t bytes = 4096;
    for (auto &i : ins) bytes += *i.f ? ((i.n * 8 + 255) & ~(size_t)255) : 0;
    if (int rc = grow(nat, bytes)) return rc;
    for (auto &i : ins)
        if (*i.f) {
            double *dev = nat.take<double>(i.n);
            cudaMemcpyAsync(dev, *i.f, i.n * 8, cudaMemcpyHostToDevice, p.stream);
            *i.f = dev;
        }
    int rc = grow(p.glue, carve_bytes<RrtmgxSwArgs>(nc, L, carve_sw_args));
    if (!rc) {
        p.glue.used = 0;
        RrtmgxSwArgs D = *sw;
        carve_sw_args(p.glue, nc, L, D);
        RRTMGX_LAUNCH(solar_prepare_kernel, (nc + 127) / 128, 128, 0, p.stream, nc, nc, 0, d, D, (const int *)nullptr);
        struct Out { const double *src; const double *dst; size_t n; } outs[] = {
            {D.coszen, sw->coszen, (size_t)nc}, {D.play, sw->play, n2}, {D.plev, sw->plev, n2p}, {D.tlay, sw->tlay, n2},
            {D.h2ovmr, sw->h2ovmr, n2}, {D.o3vmr, sw->o3vmr, n2}, {D.co2vmr, sw->co2vmr, n2}, {D.ch4vmr, sw->ch4vmr, n2},
            {D.o2vmr, sw->o2vmr, n2}, {D.cld, sw->cld, n2}, {D.ciwp, sw->ciwp, n2}, {D.clwp, sw->clwp, n2},
            {D.rei, sw->rei, n2}, {D.rel, sw->rel, n2}, {D.zm, sw->zm, n2}, {D.alat, sw->alat, (size_t)nc},
            {D.tauaer, sw->tauaer, n2 * 14}, {D.ssaaer, sw->ssaaer, n2 * 14}, {D.asmaer, sw->asmaer, n2 * 14},
            {D.asdir, sw->asdir, (size_t)nc}, {D.asdif, sw->asdif, (size_t)nc}, {D.aldir, sw->aldir, (size_t)nc},
            {D.aldif, sw->aldif, (size_t)nc}};
        for (auto &o : outs)
            if (o.dst) cudaMemcpyAsync(const_cast<double *>(o.dst), o.src, o.n * 8, cudaMemcpyDeviceToHost, p.stream);
        if (!ok(cudaStreamSynchronize(p.stream))) rc = RRTMGX_ECUDA;
    }
    cudaFree(nat.base);
    return rc;
}
#endif

}  // extern "C"


#ifndef RRTMGX_WITH_SW
namespace rrtmgx {
int sw_upload_tables(const HostTables &, const double *) { return 0; }
void sw_forget_clouds() {}
}
#endif
