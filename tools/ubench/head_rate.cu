// Microbenchmark: rate of IMAD.WIDE.U32 as a carry-chain HEAD (carry-out only) with the carry caught by IADD3.X on the ALU pipe
// (the shape of a product-scanning multiply that keeps saturated 32-bit limbs).  Compare with imad_rate.cu kinds 0-2.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o head_rate head_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
// kind 4: every MAC is a chain HEAD (carry-out, no carry-in) and its carry is caught by an addc on a third word
template <int SHARED>
__global__ void __launch_bounds__(256) k_head(uint32_t* out, int iters, uint32_t seed) {
    uint32_t a[8], b[8];
    uint32_t lo[15], hi[15], top[15];
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = seed * (i + 1) + threadIdx.x; b[i] = seed * (i + 9) + blockIdx.x + threadIdx.x * 7u; }
#pragma unroll
    for (int i = 0; i < 15; i++) { lo[i] = i; hi[i] = i * 3; top[i] = 0; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const uint32_t x = SHARED ? a[0] : a[j], y = SHARED ? b[0] : b[i];
                asm volatile("mad.lo.cc.u32 %0, %3, %4, %0; madc.hi.cc.u32 %1, %3, %4, %1; addc.u32 %2, %2, 0;"
                             : "+r"(lo[i + j]), "+r"(hi[i + j]), "+r"(top[i + j]) : "r"(x), "r"(y));
            }
    }
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < 15; i++) x ^= lo[i] ^ hi[i] ^ top[i];
    if (x == 0x12345678u) out[0] = x;
}
int main() {
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    uint32_t* out; cudaMalloc(&out, 64);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int bps = 2; bps <= 8; bps *= 2)
    for (int kind = 0; kind < 2; kind++) {
        const int blocks = sms * bps, threads = 256;
        const int iters = 1024; const double per_thread = 64.0 * iters;
        float best = 1e30f;
        for (int rep = 0; rep < 4; rep++) {
            cudaEventRecord(e0);
            if (kind == 0) k_head<0><<<blocks, threads>>>(out, iters, 12345u + rep); else k_head<1><<<blocks, threads>>>(out, iters, 12345u + rep);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep > 0 && ms < best) best = ms;
        }
        double rate = per_thread * blocks * threads / (best * 1e-3);
        printf("head+catch shared=%d blocks/SM=%d: %.3f T wide-MAC/s = %.1f lanes/clk/SM, %.3f ms\n", kind, bps, rate / 1e12, rate / sms / (clk * 1e3), best);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
