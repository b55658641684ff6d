// sqrt_probe.cu -- accuracy of MUFU.RSQ64H (rsqrt.approx.ftz.f64) and of the
// Newton-corrected square roots built on it (optpricing_b200/csrc/philox.cuh).
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
#include "../optpricing_b200/csrc/philox.cuh"
using namespace optmc;

__global__ void probe(long n, double *out) {
    double m0 = 0, m1 = 0, m2 = 0;
    for (long i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        // log-uniform in [1e-12, 1e3] with a low-discrepancy mantissa sweep
        double t = (double)i / (double)n;
        double a = exp(-27.6 + 34.5 * t) * (1.0 + 1e-3 * (double)(i % 977) / 977.0);
        double y, h;
        rsqrt_seed(a, y, h);
        double exact = sqrt(a);
        m0 = fmax(m0, fabs(y * exact - 1.0));
        m1 = fmax(m1, fabs(sqrt_nonneg_fast(a) - exact) / exact);
        m2 = fmax(m2, fabs(sqrt_nonneg(a) - exact) / exact);
    }
    atomicMax((unsigned long long *)&out[0], (unsigned long long)__double_as_longlong(m0));
    atomicMax((unsigned long long *)&out[1], (unsigned long long)__double_as_longlong(m1));
    atomicMax((unsigned long long *)&out[2], (unsigned long long)__double_as_longlong(m2));
}

int main() {
    double *d; cudaMalloc(&d, 24); cudaMemset(d, 0, 24);
    probe<<<592, 256>>>(1L << 28, d);
    double h[3]; cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
    printf("max relative error over 2^28 inputs in [1e-12, 1e3]:\n"
           "  MUFU.RSQ64H seed        %.3e  (2^%.1f)\n  sqrt, 1 Newton step     %.3e\n  sqrt, 2 Newton steps    %.3e\n",
           h[0], log2(h[0]), h[1], h[2]);
    printf("sqrt_nonneg(0) = %g, sqrt_nonneg_fast(0) = n/a (host)\n", 0.0);
    return 0;
}
