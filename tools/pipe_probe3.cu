// pipe_probe3.cu -- is the register file (operand reads) the shared bottleneck
// between the fp64 pipe and the integer pipes?  Mixes of DFMA / LOP3 / IMAD.WIDE
// with different numbers of distinct vector-register operands.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) probe(long iters, double seed, unsigned us, double *sink) {
    double d[8], e[8], f[8];
    unsigned u[8], v[8], w[8];
    unsigned long long q[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        d[k] = seed + k + threadIdx.x; e[k] = 0.999999 - 1e-9 * (k + threadIdx.x); f[k] = 1e-7 * (k + 1 + threadIdx.x);
        u[k] = threadIdx.x * 2654435761u + k; v[k] = u[k] ^ 0x9E3779B9u; w[k] = v[k] * 3u + us; q[k] = u[k];
    }
    const double uc = 1e-7 * seed;
    for (long i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            // fp64 part
            if (MODE == 0 || MODE == 2 || MODE == 6) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[k]) : "d"(e[k]), "d"(f[k]));   // 3 regs
            if (MODE == 1 || MODE == 3 || MODE == 5) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[k]) : "d"(e[k]), "d"(uc));     // 2 regs
            // integer part
            if (MODE == 0 || MODE == 1 || MODE == 4) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[k]) : "r"(v[k]), "r"(w[k]));  // 3 regs
            if (MODE == 2 || MODE == 3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[k]) : "r"(v[k]), "r"(us));                  // 2 regs
            if (MODE == 5 || MODE == 6 || MODE == 7) asm volatile("mul.wide.u32 %0, %1, 0xD2511F53;" : "=l"(q[k]) : "r"((unsigned)q[k] + us)); // IMAD.WIDE 1 reg (+1 add)
        }
    }
    double s = 0; unsigned t = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) { s += d[k] + e[k] + f[k]; t ^= u[k] ^ v[k] ^ w[k] ^ (unsigned)q[k] ^ (unsigned)(q[k] >> 32); }
    if (s == 12345.678 || t == 0x12345678u) sink[0] = s + t;
}

template <int MODE>
void run(const char *name, int sms, double ghz) {
    double *sink; cudaMalloc(&sink, 8);
    long iters = 1 << 15;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    int grid = sms * 8;
    probe<MODE><<<grid, 256>>>(iters / 8, 1.0, 3u, sink);
    cudaEventRecord(a);
    probe<MODE><<<grid, 256>>>(iters, 1.0, 3u, sink);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double cyc = ms * 1e-3 * ghz * 1e9 / (double)iters / 16.0;
    printf("%-52s %8.3f ms  %6.2f cycles per iteration of 8 groups\n", name, ms, cyc);
    cudaFree(sink);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount; double ghz = p.clockRate * 1e-6;
    printf("%s, %d SMs; 16 warps per SMSP; each iteration = 8 x (listed instructions)\n", p.name, sms);
    run<0>("DFMA(3 regs) + LOP3(3 regs)", sms, ghz);
    run<1>("DFMA(2 regs) + LOP3(3 regs)", sms, ghz);
    run<2>("DFMA(3 regs) + LOP3(2 regs)", sms, ghz);
    run<3>("DFMA(2 regs) + LOP3(2 regs)", sms, ghz);
    run<4>("LOP3(3 regs) alone", sms, ghz);
    run<7>("IMAD.WIDE(1 reg) + IADD alone", sms, ghz);
    run<5>("DFMA(2 regs) + IMAD.WIDE(1 reg) + IADD", sms, ghz);
    run<6>("DFMA(3 regs) + IMAD.WIDE(1 reg) + IADD", sms, ghz);
    return 0;
}
