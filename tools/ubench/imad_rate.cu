// Microbenchmark: issue rate of IMAD.WIDE.U32 in the shapes a field multiplication can use (sm_100a).
//   kind 0: carry chains (mad.lo.cc / madc.hi.cc -> IMAD.WIDE.U32.X), the saturated 8x32 schedule
//   kind 1: plain 64-bit accumulate, product-scanning pattern acc[i+j] += a[j]*b[i]  (carry-free radix-2^26/29 schedule)
//   kind 2: plain 64-bit accumulate, every instruction shares a and b (best case for the operand reuse cache)
//   kind 3: kind 1 plus the explicit carry propagation a radix-2^26 representation needs (shift/mask/add per column)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o imad_rate imad_rate.cu ; run: ./imad_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) k_chain(uint32_t* out, int iters, uint32_t seed) {
    uint32_t a = seed + threadIdx.x, b = seed * 3u + blockIdx.x;
    uint32_t r[16];
#pragma unroll
    for (int i = 0; i < 16; i++) r[i] = a + i;
    for (int it = 0; it < iters; it++) {
        asm volatile(
            "mad.lo.cc.u32 %0, %16, %17, %0; madc.hi.cc.u32 %1, %16, %17, %1; madc.lo.cc.u32 %2, %16, %17, %2; madc.hi.cc.u32 %3, %16, %17, %3;"
            "madc.lo.cc.u32 %4, %16, %17, %4; madc.hi.cc.u32 %5, %16, %17, %5; madc.lo.cc.u32 %6, %16, %17, %6; madc.hi.u32 %7, %16, %17, %7;"
            "mad.lo.cc.u32 %8, %17, %16, %8; madc.hi.cc.u32 %9, %17, %16, %9; madc.lo.cc.u32 %10, %17, %16, %10; madc.hi.cc.u32 %11, %17, %16, %11;"
            "madc.lo.cc.u32 %12, %17, %16, %12; madc.hi.cc.u32 %13, %17, %16, %13; madc.lo.cc.u32 %14, %17, %16, %14; madc.hi.u32 %15, %17, %16, %15;"
            : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
              "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
            : "r"(a), "r"(b));
    }
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) x ^= r[i];
    if (x == 0x12345678u) out[0] = x;
}

template <int KIND>
__global__ void __launch_bounds__(256) k_plain(uint32_t* out, int iters, uint32_t seed) {
    uint32_t a[8], b[8];
    unsigned long long acc[15];
#pragma unroll
    for (int i = 0; i < 8; i++) { a[i] = (seed * (i + 1) + threadIdx.x) & 0x3ffffffu; b[i] = (seed * (i + 9) + blockIdx.x) & 0x3ffffffu; }
#pragma unroll
    for (int i = 0; i < 15; i++) acc[i] = i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const uint32_t x = KIND == 2 ? a[0] : a[j], y = KIND == 2 ? b[0] : b[i];
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i + j]) : "r"(x), "r"(y));
            }
        if (KIND == 3) {
            // carry propagation of a radix-2^26 column set: 15 columns, sequential ripple; result limbs feed the next round
#pragma unroll
            for (int k = 0; k < 14; k++) { acc[k + 1] += acc[k] >> 26; acc[k] &= 0x3ffffffull; }
            acc[0] += (acc[14] >> 26) * 608ull; acc[14] &= 0x3ffffffull;
#pragma unroll
            for (int i = 0; i < 8; i++) { a[i] = (uint32_t)acc[i]; }
        } else {
#pragma unroll
            for (int i = 0; i < 15; i++) acc[i] &= 0x3ffffffffffull;   // keep the accumulators bounded (1 LOP per column per 64 MACs)
        }
    }
    unsigned long long x = 0;
#pragma unroll
    for (int i = 0; i < 15; i++) x ^= acc[i];
    if (x == 0x12345678ull) out[0] = (uint32_t)x;
}

int main() {
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    uint32_t* out; cudaMalloc(&out, 64);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = sms * 8, threads = 256;
    for (int kind = 0; kind < 4; kind++) {
        const int iters = kind == 0 ? 4096 : 1024;
        const double per_thread = kind == 0 ? 8.0 * iters : 64.0 * iters;
        float best = 1e30f;
        for (int rep = 0; rep < 4; rep++) {
            cudaEventRecord(e0);
            if (kind == 0) k_chain<<<blocks, threads>>>(out, iters, 12345u + rep);
            else if (kind == 1) k_plain<1><<<blocks, threads>>>(out, iters, 12345u + rep);
            else if (kind == 2) k_plain<2><<<blocks, threads>>>(out, iters, 12345u + rep);
            else k_plain<3><<<blocks, threads>>>(out, iters, 12345u + rep);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep > 0 && ms < best) best = ms;
        }
        double rate = per_thread * blocks * threads / (best * 1e-3);
        printf("kind %d: %.3f T wide-MAC/s = %.1f lanes/clk/SM (at %d MHz nominal), %.3f ms\n", kind, rate / 1e12,
               rate / sms / (clk * 1e3), clk / 1000, best);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(e));
    return 0;
}
