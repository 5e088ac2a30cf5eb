// Microbenchmark: peak rate of mma.sync.aligned.m8n8k4.f64 (SASS DMMA) vs a DFMA stream on sm_100a.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/dmma_peak tools/micro/dmma_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void k_dmma(double* out, int iters, double seed) {
    double c[NACC][2];
    for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = 0.0;
    double a = seed + threadIdx.x * 1e-3, b = seed * 0.5 + threadIdx.x * 1e-4;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) dmma(c[i][0], c[i][1], a, b);
    }
    double s = 0;
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void k_dfma(double* out, int iters, double seed) {
    double c[NACC];
    for (int i = 0; i < NACC; ++i) c[i] = i;
    double a = seed + threadIdx.x * 1e-3, b = seed * 0.5;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) c[i] = fma(c[i], a, b);
    }
    double s = 0;
    for (int i = 0; i < NACC; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    double* out;
    cudaMalloc(&out, (size_t)sms * 8 * 1024 * sizeof(double));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int iters = 20000;
    for (int warps = 4; warps <= 32; warps *= 2) {
        for (int ctas = 1; ctas <= 1; ++ctas) {
            float ms;
            k_dmma<16><<<sms * ctas, warps * 32>>>(out, 100, 1.0);
            cudaDeviceSynchronize();
            cudaEventRecord(e0);
            k_dmma<16><<<sms * ctas, warps * 32>>>(out, iters, 1.0);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
            double flop = (double)sms * ctas * warps * iters * 16 * 512.0;
            printf("DMMA  warps/SM %2d acc 16: %.2f TFLOP/s (%.3f ms)\n", warps * ctas, flop / ms / 1e9, ms);
            k_dfma<16><<<sms * ctas, warps * 32>>>(out, 100, 1.0);
            cudaDeviceSynchronize();
            cudaEventRecord(e0);
            k_dfma<16><<<sms * ctas, warps * 32>>>(out, iters, 1.0);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
            flop = (double)sms * ctas * warps * 32 * iters * 16 * 2.0;
            printf("DFMA  warps/SM %2d acc 16: %.2f TFLOP/s (%.3f ms)\n", warps * ctas, flop / ms / 1e9, ms);
        }
    }
    {
        float ms;
        cudaEventRecord(e0);
        k_dmma<4><<<sms, 256>>>(out, iters, 1.0);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("DMMA  8 warps acc 4 (latency-exposed): %.2f TFLOP/s\n", (double)sms * 8 * iters * 4 * 512.0 / ms / 1e9);
        cudaEventRecord(e0);
        k_dmma<1><<<sms, 32>>>(out, iters, 1.0);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("DMMA  dependent chain, 1 warp/SM: %.1f ns per DMMA\n", ms * 1e6 / iters);
    }
    printf("sms %d clock %d kHz\n", sms, p.clockRate);
    return 0;
}
