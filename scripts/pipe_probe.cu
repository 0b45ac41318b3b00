// pipe_probe.cu — do packed-16-bit integer min/max (VIMNMX3.U16x2, ALU pipe) and half2 min/max (HMNMX2) share a pipe?
// Times (A) an integer min3/max3 chain, (B) a half2 min/max chain, (C) both interleaved, with 8 independent chains per thread.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/pipe_probe scripts/pipe_probe.cu && /tmp/pipe_probe
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) probe(uint32_t *out, int iters, uint32_t seed)
{
    uint32_t a[8], b[8];
    __half2 h[8], g[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        a[i] = seed * (threadIdx.x + 1) + i * 0x01010101u; b[i] = a[i] ^ 0x5a5a5a5au;
        h[i] = *reinterpret_cast<__half2 *>(&a[i]); g[i] = *reinterpret_cast<__half2 *>(&b[i]);
    }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0 || MODE == 2) { a[i] = __vimin3_u16x2(a[i], b[i], a[(i + 1) & 7]); b[i] = __vimax3_u16x2(b[i], a[i], b[(i + 3) & 7]); }
            if (MODE == 1 || MODE == 2) { h[i] = __hmin2(h[i], g[(i + 1) & 7]); g[i] = __hmax2(g[i], h[(i + 3) & 7]); }
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) r ^= a[i] ^ b[i] ^ *reinterpret_cast<uint32_t *>(&h[i]) ^ *reinterpret_cast<uint32_t *>(&g[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE> static float run(uint32_t *d, int iters)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<MODE><<<148 * 8, 256>>>(d, iters, 12345u);
    cudaEventRecord(e0);
    probe<MODE><<<148 * 8, 256>>>(d, iters, 12345u);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}
int main()
{
    uint32_t *d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    const int iters = 20000;
    const float a = run<0>(d, iters), b = run<1>(d, iters), c = run<2>(d, iters);
    const double ops = 148.0 * 8 * 256 * iters * 16;   // per mode: 16 min/max per iteration and thread
    printf("{\"vimnmx3_u16x2_ms\": %.3f, \"hmnmx2_ms\": %.3f, \"both_ms\": %.3f, \"vimnmx3_per_clk_per_sm\": %.1f, \"hmnmx2_per_clk_per_sm\": %.1f, "
           "\"reading\": \"both ~= sum: same pipe; both ~= max: different pipes\"}\n",
           a, b, c, ops / (a * 1e-3) / 148 / 1.965e9, ops / (b * 1e-3) / 148 / 1.965e9);
    return 0;
}
