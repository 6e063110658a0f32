// FP64 issue-rate probe for the roofline's second roof ("measure with a
// DFMA loop").  Every thread runs ILP independent dependency chains of one FP64 instruction kind; the figure reported
// is thread-instructions per second over the whole device, i.e. the rate `sm__inst_executed_pipe_fp64 * 32` would
// reach at 100 % pipe activity.  Kinds: dfma (a*b+c, one instruction), dadd, dmul, and "mix" = alternating dmul/dadd,
// which is what the product kernels issue (they are compiled with --fmad=false).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a --fmad=false -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

template <int KIND, int ILP>
__global__ void __launch_bounds__(256) fp64_loop(double *out, int iters, double a, double b) {
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = a + (double)(threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int i = 0; i < ILP; ++i) {
                if (KIND == 0) x[i] = __fma_rn(x[i], a, b);
                else if (KIND == 1) x[i] = __dadd_rn(x[i], b);
                else if (KIND == 2) x[i] = __dmul_rn(x[i], a);
                else x[i] = (u & 1) ? __dadd_rn(x[i], b) : __dmul_rn(x[i], a);
            }
        }
    }
    double s = 0.;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    if (s == 12345.678) out[0] = s;   // never true: keeps the chains alive
}

// The same with every source operand in its own REGISTER pair (the kernels' case: operands are per-thread values, not
// kernel parameters in the constant bank).  KIND 0: x = fma(x, y, z) (three register pairs read), 1: x = x + y, 2: x = x * y,
// 3: alternating mul/add.  y and z are per-thread and change slowly so that they stay in registers.
template <int KIND, int ILP>
__global__ void __launch_bounds__(256) fp64_loop_rrr(double *out, int iters, double a, double b) {
    double x[ILP], y[ILP], z[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
        x[i] = a + (double)(threadIdx.x + i);
        y[i] = 1.0 + 1e-9 * (double)(threadIdx.x + 3 * i);
        z[i] = b * (double)(1 + threadIdx.x + 7 * i);
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int i = 0; i < ILP; ++i) {
                if (KIND == 0) x[i] = __fma_rn(x[i], y[i], z[i]);
                else if (KIND == 1) x[i] = __dadd_rn(x[i], z[i]);
                else if (KIND == 2) x[i] = __dmul_rn(x[i], y[i]);
                else x[i] = (u & 1) ? __dadd_rn(x[i], z[i]) : __dmul_rn(x[i], y[i]);
            }
        }
        if (it == iters + 5) {   // never true: y and z are not loop invariants the compiler may fold
#pragma unroll
            for (int i = 0; i < ILP; ++i) { y[i] += x[i]; z[i] -= x[i]; }
        }
    }
    double s = 0.;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i] + y[i] + z[i];
    if (s == 12345.678) out[0] = s;
}

template <int KIND, int ILP>
static double run_rrr(int sms, int blocks_per_sm, int iters, double *d) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = sms * blocks_per_sm;
    fp64_loop_rrr<KIND, ILP><<<grid, 256>>>(d, iters / 8, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    double best = 0.;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        fp64_loop_rrr<KIND, ILP><<<grid, 256>>>(d, iters, 1.0000001, 1e-9);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double rate = (double)grid * 256. * (double)iters * 8. * ILP / (ms * 1e-3);
        if (rate > best) best = rate;
    }
    return best;
}

template <int KIND, int ILP>
static double run(int sms, int blocks_per_sm, int iters, double *d) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = sms * blocks_per_sm;
    fp64_loop<KIND, ILP><<<grid, 256>>>(d, iters / 8, 1.0000001, 1e-9);   // warm-up
    cudaDeviceSynchronize();
    double best = 0.;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        fp64_loop<KIND, ILP><<<grid, 256>>>(d, iters, 1.0000001, 1e-9);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double instr = (double)grid * 256. * (double)iters * 8. * ILP;
        const double rate = instr / (ms * 1e-3);
        if (rate > best) best = rate;
    }
    return best;
}

int main(int argc, char **argv) {
    int dev = 0;
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) { std::fprintf(stderr, "no CUDA device\n"); return 2; }
    double *d;
    cudaMalloc(&d, 8);
    const int iters = argc > 1 ? std::atoi(argv[1]) : 20000;
    const int sms = p.multiProcessorCount;
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, dev);
    const char *names[4] = {"dfma", "dadd", "dmul", "mix_dmul_dadd"};
    double r[4];
    r[0] = run<0, 8>(sms, 8, iters, d);
    r[1] = run<1, 8>(sms, 8, iters, d);
    r[2] = run<2, 8>(sms, 8, iters, d);
    r[3] = run<3, 8>(sms, 8, iters, d);
    const double rr[4] = {run_rrr<0, 8>(sms, 8, iters, d), run_rrr<1, 8>(sms, 8, iters, d), run_rrr<2, 8>(sms, 8, iters, d),
                          run_rrr<3, 8>(sms, 8, iters, d)};
    const double dep = run<0, 1>(sms, 1, iters, d);   // one chain, 8 warps per SM: exposes the dependent-issue latency
    std::printf("{\"gpu\": \"%s\", \"sms\": %d, \"sm_max_mhz\": %.0f, \"unit\": \"fp64 thread-instructions/s\"", p.name, sms,
                clk / 1e3);
    for (int k = 0; k < 4; ++k) std::printf(", \"%s\": %.4e", names[k], r[k]);
    for (int k = 0; k < 4; ++k) std::printf(", \"%s_register_operands\": %.4e", names[k], rr[k]);
    std::printf(", \"dfma_flops\": %.4e", 2. * r[0]);
    std::printf(", \"per_sm_per_clk_at_max_clock\": %.2f", r[3] / sms / (clk * 1e3));
    // 8 warps/SM = 2 per SMSP, 1 chain each: rate = 2 warps * 32 lanes / latency per SMSP
    std::printf(", \"dependent_dfma_latency_cycles_at_max_clock\": %.1f", 8. * 32. * sms * (clk * 1e3) / dep);
    std::printf(", \"how\": \"256-thread blocks, 8 per SM, 8 independent chains per thread, %d x 8 instructions per chain, best of 5, CUDA events\"}\n", iters);
    return 0;
}
