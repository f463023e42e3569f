// Single-warp latency of the building blocks of the serial tail (window Horner): dependent chains of fe_mul, fe_sqr,
// fe_add, fe_sub, an 8-limb shuffle, and whole quad doublings / additions.  Cycles per operation via clock64().
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o quad_latency quad_latency.cu
#include "../../zkvm_b200/csrc/msm.cu"

__global__ void k_lat(const uint4* in_ext, long long* cyc, uint4* sink) {
    const quad_ctx c = quad_self();
    fe a, b, t;
    quad_ld(a, in_ext, threadIdx.x >> 2, c.q);
    quad_ld(b, in_ext, 8 + (threadIdx.x >> 2), c.q);
    const int N = 64;
    long long t0, t1;
#define TIME(slot, stmt) t0 = clock64(); _Pragma("unroll 1") for (int i = 0; i < N; i++) { stmt; } t1 = clock64(); if (threadIdx.x == 0) cyc[slot] = (t1 - t0) / N;
    TIME(0, fe_mul(a, a, b))
    TIME(1, fe_sqr(a, a))
    TIME(2, fe_add(a, a, b))
    TIME(3, fe_sub(a, a, b))
    TIME(4, fe_shfl_xor(t, a, 1, c.mask); fe_add(a, t, b))          // shuffle + add
    TIME(5, quad_dbl_n(a, 1, c))                                     // one doubling + the lane swap back
    TIME(6, quad_add(a, a, b, c))
    TIME(7, fe_mul(a, a, b); fe_mul(b, b, a))                        // two dependent multiplies
    TIME(8, fe_select(t, a, b, (c.q + i) & 1); fe_add(a, t, b))      // select + add
    TIME(9, quad_dbl_n(a, 2, c))                                     // two chained doublings (role-permuted form)
    quad_st(sink, threadIdx.x >> 2, c.q, a);
}
int main() {
    const int n = 16;
    uint8_t h[n * 64];
    srand(5);
    for (int i = 0; i < n * 64; i++) h[i] = rand();
    uint8_t* d_u; uint4 *d_ext, *d_sink; long long* d_cyc;
    cudaMalloc(&d_u, n * 64); cudaMalloc(&d_ext, n * 128); cudaMalloc(&d_sink, 128 * 128); int* d_bad; cudaMalloc(&d_bad, 4); cudaMemset(d_bad, 0, 4); cudaMalloc(&d_cyc, 128);
    cudaMemcpy(d_u, h, n * 64, cudaMemcpyHostToDevice);
    k_map_uniform<<<1, 128>>>((const uint4*)d_u, n, d_ext);
    for (int rep = 0; rep < 2; rep++) k_lat<<<1, 32>>>(d_ext, d_cyc, d_sink);
    long long cyc[16] = {0};
    cudaMemcpy(cyc, d_cyc, 128, cudaMemcpyDeviceToHost);
    const char* names[] = {"fe_mul", "fe_sqr", "fe_add", "fe_sub", "shfl+add", "dbl_n(1)", "quad_add", "2x fe_mul", "select+add", "dbl_chain"};
    cyc[9] /= 2;
    for (int i = 0; i < 10; i++) printf("%-12s %6lld cycles\n", names[i], cyc[i]);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
