// Issue-slot microbenchmark for the FP64-heavy per-sample kernels ("what bounds these kernels").
//
// Question it answers: when a warp scheduler (SM sub-partition) issues an FP64 instruction -- the FP64 pipe accepts one
// warp instruction every 2 cycles -- can it issue an instruction for ANOTHER pipe (integer ALU / FMA-pipe IMAD / FP32) in
// the cycle in between, or does the FP64 instruction hold the dispatch port for both cycles?
//
//   model A ("port held"):     cycles = 2*N_fp64 + N_other
//   model B ("pipes overlap"): cycles = max(2*N_fp64, N_other_per_pipe..., N_fp64 + N_other)
//
// Every kernel runs a fixed instruction pattern (period P of D = DFMA, L = LOP3, I = IMAD, W = IMAD.WIDE, F = FFMA,
// A = IADD3, M = MUFU.RCP64H, S = LDS.32 (pointer chase)) on independent accumulator chains (8 per type) so that dependent-issue latency
// does not limit it once a few warps are resident, and reports SM-sub-partition cycles per pattern period from clock64()
// (and from CUDA events x the SM clock), for 1..16 resident warps per sub-partition.
//
// Build:  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o issue_microbench tools/issue_microbench.cu
// Run:    ./issue_microbench [sm_mhz]        (prints a table and JSON lines; tools/run_issue_microbench.sh wraps it)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

constexpr int kChains = 8;

struct Regs {
    double d[kChains];
    uint32_t l[kChains];
    uint32_t i[kChains];
    uint64_t w[kChains];
    float f[kChains];
    uint32_t a[kChains];
    double m[kChains];
    uint32_t s[kChains];
};

template <char OP>
__device__ __forceinline__ void emit(Regs& r, int k, double da, double db, uint32_t ua, uint32_t ub, float fa, float fb,
                                     uint32_t smem_addr) {
    const int c = k % kChains;
    if (OP == 'D') asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(r.d[c]) : "d"(da), "d"(db));
    if (OP == 'L') asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r.l[c]) : "r"(ua), "r"(ub));
    if (OP == 'I') asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r.i[c]) : "r"(ua), "r"(ub));
    if (OP == 'W') {                                        // multiplicand from the chain (as Philox: counter word x constant)
        const uint32_t in = (k / kChains) & 1 ? (uint32_t)(r.w[c] >> 32) : (uint32_t)r.w[c];    // both halves are consumed
        asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(r.w[c]) : "r"(in), "r"(ua));
    }
    if (OP == 'F') asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(r.f[c]) : "f"(fa), "f"(fb));
    if (OP == 'A') asm volatile("add.u32 %0, %0, %1;" : "+r"(r.a[c]) : "r"(ua));
    if (OP == 'M') asm volatile("rcp.approx.ftz.f64 %0, %0;" : "+d"(r.m[c]));
    if (OP == 'S') asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r.s[c]) : "r"(smem_addr + r.s[c]) : "memory");   // pointer chase
}

// The pattern is a template parameter pack of chars; one loop trip = REP periods.
template <char... OPS>
struct Pattern {
    static constexpr int P = sizeof...(OPS);
    static __host__ __device__ constexpr bool has(char c) { return ((OPS == c) || ...); }
    template <int REP>
    static __device__ __forceinline__ void run(Regs& r, double da, double db, uint32_t ua, uint32_t ub, float fa, float fb,
                                               uint32_t sa) {
#pragma unroll
        for (int rep = 0; rep < REP; ++rep) {
            int k = rep * P;
            ((emit<OPS>(r, k++, da, db, ua, ub, fa, fb, sa)), ...);
        }
    }
};

template <class PAT, int REP>
__global__ void __launch_bounds__(128, 12) bench_kernel(int iters, double da, double db, uint32_t ua, uint32_t ub, float fa, float fb,
                                                        double* out, long long* cyc) {
    __shared__ uint32_t sm[128];
    sm[threadIdx.x] = 4u * ((threadIdx.x + 33u) & 127u);        // byte offset of the next element of the chase
    __syncthreads();
    const uint32_t sa = (uint32_t)__cvta_generic_to_shared(sm);
    Regs r;
#pragma unroll
    for (int c = 0; c < kChains; ++c) {                    // only the accumulators the pattern uses stay live
        if (PAT::has('D')) r.d[c] = 1.0 + c + threadIdx.x;
        if (PAT::has('L')) r.l[c] = c + threadIdx.x;
        if (PAT::has('I')) r.i[c] = 3 * c + threadIdx.x;
        if (PAT::has('W')) r.w[c] = 5 * c + threadIdx.x;
        if (PAT::has('F')) r.f[c] = 1.0f + c;
        if (PAT::has('A')) r.a[c] = 7 * c + threadIdx.x;
        if (PAT::has('M')) r.m[c] = 1.5 + c + threadIdx.x;
        if (PAT::has('S')) r.s[c] = 4u * ((threadIdx.x + 5 * c) & 127u);
    }
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) PAT::template run<REP>(r, da, db, ua, ub, fa, fb, sa);
    const long long t1 = clock64();
    double acc = 0.0;
#pragma unroll
    for (int c = 0; c < kChains; ++c) {
        if (PAT::has('D')) acc += r.d[c];
        if (PAT::has('L')) acc += (double)r.l[c];
        if (PAT::has('I')) acc += (double)r.i[c];
        if (PAT::has('W')) acc += (double)r.w[c];
        if (PAT::has('F')) acc += (double)r.f[c];
        if (PAT::has('A')) acc += (double)r.a[c];
        if (PAT::has('M')) acc += r.m[c];
        if (PAT::has('S')) acc += (double)r.s[c];
    }
    if (acc == 12345.678) out[0] = acc;                     // keep everything live
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

static double g_sm_mhz = 1965.0;
static int g_sms = 148;

template <class PAT>
void run_pattern(const char* name, const char* desc) {
    constexpr int REP = (64 + PAT::P - 1) / PAT::P;          // ~64 instructions per loop trip
    double* out; long long* cyc;
    CK(cudaMalloc(&out, 8));
    CK(cudaMalloc(&cyc, sizeof(long long) * g_sms * 16));
    long long* h = (long long*)malloc(sizeof(long long) * g_sms * 16);
    const int iters = 20000;
    printf("%-10s %-44s", name, desc);
    for (int wps : {1, 2, 4, 6, 8, 12}) {                 // resident warps per SM sub-partition (one 128-thread CTA = 1 each)
        const int blocks = g_sms * wps;
        bench_kernel<PAT, REP><<<blocks, 128>>>(200, 1.0000001, 1e-9, 0x9E3779B9u, 0x85EBCA6Bu, 1.0000001f, 1e-9f, out, cyc);
        CK(cudaDeviceSynchronize());
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        CK(cudaEventRecord(e0));
        bench_kernel<PAT, REP><<<blocks, 128>>>(iters, 1.0000001, 1e-9, 0x9E3779B9u, 0x85EBCA6Bu, 1.0000001f, 1e-9f, out, cyc);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        CK(cudaMemcpy(h, cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost));
        double mean = 0;
        for (int b = 0; b < blocks; ++b) mean += (double)h[b];
        mean /= blocks;
        // every sub-partition issues wps warps x iters x REP periods in `mean` cycles
        const double per_period = mean / ((double)iters * REP * wps);
        const double per_period_ev = ms * 1e-3 * g_sm_mhz * 1e6 / ((double)iters * REP * wps);
        printf(" | w%-2d %6.2f (%6.2f)", wps, per_period, per_period_ev);
        printf("\n#JSON {\"pattern\": \"%s\", \"period\": %d, \"warps_per_smsp\": %d, \"cycles_per_period_clock64\": %.4f, "
               "\"cycles_per_period_events\": %.4f}\n%-55s", name, PAT::P, wps, per_period, per_period_ev, "");
        CK(cudaEventDestroy(e0)); CK(cudaEventDestroy(e1));
    }
    printf("\n");
    free(h);
    CK(cudaFree(out)); CK(cudaFree(cyc));
}

int main(int argc, char** argv) {
    if (argc > 1) g_sm_mhz = atof(argv[1]);
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    g_sms = p.multiProcessorCount;
    printf("# %s, %d SMs, assumed SM clock for the event-based column %.0f MHz\n", p.name, g_sms, g_sm_mhz);
    printf("# cycles per pattern period per SM sub-partition: clock64 (events); model A = 2*nD + others, model B = max(2*nD, total)\n");
    run_pattern<Pattern<'D'>>("D", "DFMA only                       A=2  B=2");
    run_pattern<Pattern<'L'>>("L", "LOP3 only");
    run_pattern<Pattern<'I'>>("I", "IMAD only");
    run_pattern<Pattern<'W'>>("W", "IMAD.WIDE only");
    run_pattern<Pattern<'F'>>("F", "FFMA only");
    run_pattern<Pattern<'A'>>("A", "IADD only");
    run_pattern<Pattern<'M'>>("M", "MUFU.RCP64H only");
    run_pattern<Pattern<'S'>>("S", "LDS.32 only");
    run_pattern<Pattern<'D', 'L'>>("DL", "1 DFMA : 1 LOP3                 A=3  B=2");
    run_pattern<Pattern<'D', 'I'>>("DI", "1 DFMA : 1 IMAD                 A=3  B=2");
    run_pattern<Pattern<'D', 'F'>>("DF", "1 DFMA : 1 FFMA                 A=3  B=2");
    run_pattern<Pattern<'D', 'L', 'I'>>("DLI", "1 DFMA : 1 LOP3 : 1 IMAD        A=4  B=3");
    run_pattern<Pattern<'D', 'L', 'D', 'I', 'L'>>("DLDIL", "2 DFMA : 3 int (chain kernel mix) A=7  B=5");
    run_pattern<Pattern<'D', 'L', 'I', 'L'>>("DLIL", "1 DFMA : 3 int                  A=5  B=4");
    run_pattern<Pattern<'D', 'D', 'L'>>("DDL", "2 DFMA : 1 LOP3                 A=5  B=4");
    run_pattern<Pattern<'D', 'D', 'D', 'L'>>("DDDL", "3 DFMA : 1 LOP3                 A=7  B=6");
    run_pattern<Pattern<'L', 'I'>>("LI", "1 LOP3 : 1 IMAD (alu + fma pipes)");
    run_pattern<Pattern<'W', 'W', 'L', 'L'>>("WWLL", "Philox round: 2 IMAD.WIDE + 2 LOP3");
    run_pattern<Pattern<'D', 'W', 'D', 'W', 'D', 'L', 'D', 'L'>>("DWDWDLDL", "Philox round interleaved with 4 DFMA  A=12 B=8");
    run_pattern<Pattern<'D', 'S'>>("DS", "1 DFMA : 1 LDS.32");
    run_pattern<Pattern<'D', 'M', 'D', 'D', 'D'>>("DMDDD", "4 DFMA : 1 MUFU.RCP64H");
    return 0;
}
