// Microbenchmark: per-SM throughput of 1-D bulk copies (cp.async.bulk global -> shared, mbarrier completion) as a
// function of the copy size: how small can a tile row be before the TMA unit, not the data, is the limit?
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/bulk_copy_rate tools/micro/bulk_copy_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned s32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__global__ void k(const char* src, size_t src_bytes, int copy_bytes, int copies_per_round, int rounds, int lanes, long long* cycles) {
    extern __shared__ __align__(128) unsigned char sm[];
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(sm);
    unsigned char* dst = sm + 128;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const char* base = src + ((size_t)blockIdx.x * 1048576) % (src_bytes - (size_t)copy_bytes * copies_per_round * 4);
    unsigned par = 0;
    long long t0 = clock64();
    for (int r = 0; r < rounds; ++r) {
        if (threadIdx.x == 0)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(copy_bytes * copies_per_round) : "memory");
        __syncwarp();
        if (threadIdx.x < lanes)
            for (int c = threadIdx.x; c < copies_per_round; c += lanes)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(s32(dst + (size_t)c * copy_bytes)), "l"(base + ((size_t)(r & 3) * copies_per_round + c) * copy_bytes * 1), "r"(copy_bytes), "r"(s32(bar)) : "memory");
        asm volatile("{\n.reg .pred P1;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra D;\nbra W;\nD:\n}" ::"r"(s32(bar)), "r"(par) : "memory");
        par ^= 1;
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    const size_t src_bytes = 512ull << 20;
    char* src;
    cudaMalloc(&src, src_bytes);
    cudaMemset(src, 1, src_bytes);
    long long* cyc;
    cudaMallocManaged(&cyc, sms * sizeof(long long));
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int total = 61440;     // bytes per round (one GEMM stage)
    const int sizes[] = {128, 256, 512, 1024, 2048, 4096, 20480, 61440};
    for (int lanes : {1, 32}) {
        for (int sz : sizes) {
            const int copies = total / sz;
            k<<<sms, 32, 128 + total, 0>>>(src, src_bytes, sz, copies, 200, lanes, cyc);
            cudaDeviceSynchronize();
            k<<<sms, 32, 128 + total, 0>>>(src, src_bytes, sz, copies, 200, lanes, cyc);
            cudaError_t e = cudaDeviceSynchronize();
            double avg = 0;
            for (int i = 0; i < sms; ++i) avg += cyc[i];
            avg /= sms * 200.0;
            printf("lanes %2d copy %6d B x %4d: %8.0f cycles per 60 KB round = %.1f B/clk/SM, %.0f cycles per copy (%s)\n", lanes, sz,
                   copies, avg, total / avg, avg / copies, cudaGetErrorString(e));
        }
    }
    return 0;
}
