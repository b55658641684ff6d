// pipe_probe2.cu -- cost of DFMA by number of distinct register operands, and of
// MUFU.RSQ64H / I2F.F64 / LDS.128 (random rows) per warp instruction.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) probe(long iters, double seed, double *sink) {
    __shared__ double2 tab[1024];
    for (int i = threadIdx.x; i < 1024; i += 256) tab[i] = make_double2(i * seed, 1.0);
    __syncthreads();
    double d[8], e[8], f[8];
    unsigned u[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { d[k] = seed + k + threadIdx.x; e[k] = 0.999999 - 1e-9 * (k + threadIdx.x); f[k] = 1e-7 * (k + 1 + threadIdx.x); u[k] = threadIdx.x * 2654435761u + k; }
    const double ub = 0.999999 + seed * 1e-9, uc = 1e-7 * seed;
    for (long i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (MODE == 0) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[k]) : "d"(ub), "d"(uc));           // 1 reg
            if (MODE == 1) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[k]) : "d"(e[k]), "d"(uc));        // 2 regs
            if (MODE == 2) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[k]) : "d"(e[k]), "d"(f[k]));      // 3 regs
            if (MODE == 3) asm volatile("fma.rn.f64 %0, %1, %1, %0;" : "+d"(d[k]) : "d"(e[k]));                 // e*e + d
            if (MODE == 4) { double y; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d[k])); d[k] = y; }
            if (MODE == 5) { asm volatile("cvt.rn.f64.u32 %0, %1;" : "=d"(d[k]) : "r"(u[k])); u[k] += (unsigned)i; }
            if (MODE == 6) { double2 t = tab[u[k] & 1023]; u[k] = u[k] * 1664525u + 1013904223u + (unsigned)__double2loint(t.x); }
            if (MODE == 7) { double2 t = tab[(u[k] & 1023 & ~31) | (threadIdx.x & 31)]; u[k] = u[k] * 1664525u + 1013904223u + (unsigned)__double2loint(t.x); }
            if (MODE == 8) asm volatile("mul.f64 %0, %0, %1;" : "+d"(d[k]) : "d"(e[k]));
            if (MODE == 9) asm volatile("add.f64 %0, %0, %1;" : "+d"(d[k]) : "d"(e[k]));
        }
    }
    double s = 0; unsigned t = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) { s += d[k] + e[k] + f[k]; t ^= u[k]; }
    if (s == 12345.678 || t == 0x12345678u) sink[0] = s + t;
}

template <int MODE>
void run(const char *name, int sms, double ghz) {
    double *sink; cudaMalloc(&sink, 8);
    long iters = 1 << 15;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    int grid = sms * 8;
    probe<MODE><<<grid, 256>>>(iters / 8, 1.0, sink);
    cudaEventRecord(a);
    probe<MODE><<<grid, 256>>>(iters, 1.0, sink);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double cyc = ms * 1e-3 * ghz * 1e9 / (double)iters / 16.0 / 8.0;
    printf("%-44s %8.3f ms  %6.2f cycles per warp-instruction per SMSP\n", name, ms, cyc);
    cudaFree(sink);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount; double ghz = p.clockRate * 1e-6;
    printf("%s, %d SMs, %.3f GHz nominal; 16 warps per SMSP\n", p.name, sms, ghz);
    run<0>("DFMA 1 register operand + 2 uniform", sms, ghz);
    run<1>("DFMA 2 register operands + 1 uniform", sms, ghz);
    run<2>("DFMA 3 distinct register operands", sms, ghz);
    run<3>("DFMA e*e+d (2 distinct registers)", sms, ghz);
    run<8>("DMUL 2 registers", sms, ghz);
    run<9>("DADD 2 registers", sms, ghz);
    run<4>("MUFU.RSQ64H", sms, ghz);
    run<5>("I2F.F64.U32 (+1 IADD)", sms, ghz);
    run<6>("LDS.128 random rows of a 16 KB table (+2 int)", sms, ghz);
    run<7>("LDS.128 conflict-free rows (+4 int)", sms, ghz);
    return 0;
}
