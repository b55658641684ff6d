// pipe_probe.cu -- B200 dispatch micro-benchmark: do fp64-pipe and integer-ALU /
// FMA-pipe instructions overlap within an SM sub-partition?  Each kernel runs a
// register-resident loop of independent chains; cycles per loop iteration per
// warp are derived from the elapsed time at full occupancy.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/pipe_probe tools/pipe_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ND, int NL, int NI>
__global__ void __launch_bounds__(256) probe(long iters, double seed, unsigned useed, double *sink) {
    double d[8];
    unsigned l[8], m[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { d[k] = seed + k + threadIdx.x; l[k] = useed * (k + 1) + threadIdx.x; m[k] = l[k] ^ 0x9E3779B9u; }
    const double bb = 0.999999 + seed * 1e-9, cc = 1e-7 * seed;
    for (long i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (k < ND) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[k]) : "d"(bb), "d"(cc));
            if (k < NL) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(l[k]) : "r"(m[k]), "r"(useed));
            if (k < NI) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(m[k]) : "r"(useed), "r"(l[(k + 1) & 7]));
        }
    }
    double s = 0; unsigned t = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) { s += d[k]; t ^= l[k] ^ m[k]; }
    if (s == 12345.678 || t == 0x12345678u) sink[0] = s + t;
}

template <int ND, int NL, int NI>
void run(const char *name, int sms, double clock_ghz) {
    double *sink; cudaMalloc(&sink, 8);
    long iters = 1 << 16;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    int grid = sms * 8;   // 8 blocks x 256 threads = 2048 threads = 16 warps / SMSP
    probe<ND, NL, NI><<<grid, 256>>>(iters / 8, 1.0, 3u, sink);
    cudaEventRecord(a);
    probe<ND, NL, NI><<<grid, 256>>>(iters, 1.0, 3u, sink);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    // warps per SMSP = 16; cycles per iteration per SMSP-warp-slot:
    double cyc = ms * 1e-3 * clock_ghz * 1e9 / (double)iters / 16.0;
    printf("%-28s DFMA=%d LOP3=%d IMAD=%d : %8.3f ms  %6.2f cycles per warp-iteration (%d instr)\n",
           name, ND, NL, NI, ms, cyc, ND + NL + NI);
    cudaFree(sink);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount; double ghz = p.clockRate * 1e-6;
    printf("%s, %d SMs, %.3f GHz nominal\n", p.name, sms, ghz);
    run<8, 0, 0>("fp64 only", sms, ghz);
    run<0, 8, 0>("alu only", sms, ghz);
    run<0, 0, 8>("imad only", sms, ghz);
    run<8, 8, 0>("fp64 + alu", sms, ghz);
    run<8, 0, 8>("fp64 + imad", sms, ghz);
    run<0, 8, 8>("alu + imad", sms, ghz);
    run<8, 8, 8>("fp64 + alu + imad", sms, ghz);
    run<4, 8, 8>("fp64(4) + alu + imad", sms, ghz);
    run<8, 4, 4>("fp64 + alu(4) + imad(4)", sms, ghz);
    return 0;
}
