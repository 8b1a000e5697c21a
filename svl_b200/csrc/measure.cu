// measure.cu -- device micro-benchmarks behind bench.py's roofline record (fp64_fraction is quoted
// against a MEASURED FP64-FMA peak, hbm fractions against a measured copy bandwidth).  Not on the step path.
#include <string>
#include "model.h"

namespace svl {

// 16 independent DFMA chains per thread, 512 threads per CTA, 2 CTAs per SM: the FP64 pipe is the only limiter
__global__ void __launch_bounds__(512, 2) k_fp64_fma(double *out, int iters, double a, double b) {
    double x[16];
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = 1e-3 * (threadIdx.x + i);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) x[i] = fma(x[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += x[i];
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;      // never true: keeps the chains alive
}

// plain streaming copy (double2 per thread, grid-stride): read + write bytes / time = what a kernel that reads and
// writes HBM once can reach on this device
__global__ void __launch_bounds__(256) k_copy(const double2 *__restrict__ src, double2 *__restrict__ dst, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

}  // namespace svl
using namespace svl;

extern "C" int svlgpu_measure_peaks(int device, double *fp64_tflops, double *copy_gbs) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) { set_error("measure_peaks: no such CUDA device"); return 1; }
    cudaSetDevice(device);
    cudaDeviceProp pr;
    cudaGetDeviceProperties(&pr, device);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms = 0;
    if (fp64_tflops) {
        double *d = nullptr;
        const int grid = pr.multiProcessorCount * 2, iters = 4096;
        cudaMalloc(&d, sizeof(double) * 512 * grid);
        k_fp64_fma<<<grid, 512>>>(d, 64, 0.999, 1e-9);                   // warm-up
        double best = 0;
        for (int rep = 0; rep < 5; rep++) {
            cudaEventRecord(e0);
            k_fp64_fma<<<grid, 512>>>(d, iters, 0.999, 1e-9);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
            const double tf = 2.0 * 16.0 * iters * 512.0 * grid / (ms * 1e-3) / 1e12;
            if (tf > best) best = tf;
        }
        cudaFree(d);
        *fp64_tflops = best;
    }
    if (copy_gbs) {
        const size_t n = (size_t)1 << 26;                                  // 2 x 1 GiB buffers: far beyond the 126 MB L2
        double2 *a = nullptr, *b = nullptr;
        if (cudaMalloc(&a, n * sizeof(double2)) != cudaSuccess || cudaMalloc(&b, n * sizeof(double2)) != cudaSuccess) {
            cudaFree(a); set_error("measure_peaks: out of device memory"); return 1;
        }
        cudaMemset(a, 0, n * sizeof(double2));
        k_copy<<<pr.multiProcessorCount * 8, 256>>>(a, b, n);
        double best = 0;
        for (int rep = 0; rep < 5; rep++) {
            cudaEventRecord(e0);
            k_copy<<<pr.multiProcessorCount * 8, 256>>>(a, b, n);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
            const double gbs = 2.0 * n * sizeof(double2) / (ms * 1e-3) / 1e9;
            if (gbs > best) best = gbs;
        }
        cudaFree(a); cudaFree(b);
        *copy_gbs = best;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error(std::string("measure_peaks: ") + cudaGetErrorString(e)); return 1; }
    return 0;
}
