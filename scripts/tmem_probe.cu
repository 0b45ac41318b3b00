// tmem_probe.cu — measures the tensor-memory read rate of tcgen05.ld (32x32b.x32) per SM with 4 / 8 / 16 reading warps.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tmem_probe scripts/tmem_probe.cu && /tmp/tmem_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(512, 1) probe(int n_warps, int iters, long long *cycles, unsigned *sink)
{
    __shared__ uint32_t s_tmem;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&s_tmem)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    unsigned acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    if (warp < n_warps) {
        const uint32_t lane_base = 32u * (warp & 3);
        for (int it = 0; it < iters; it++) {
            uint32_t v[32];
            const uint32_t taddr = tmem + (lane_base << 16) + (uint32_t)(((it + warp) * 32) & 511);
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                  "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                  "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                  "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 32; j++) acc ^= v[j];
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

int main()
{
    long long *d_c; unsigned *d_s;
    cudaMalloc(&d_c, 148 * sizeof(long long)); cudaMalloc(&d_s, 148 * 512 * sizeof(unsigned));
    const int iters = 2000;
    for (int nw : {1, 4, 8, 16}) {
        probe<<<148, 512>>>(nw, iters, d_c, d_s);
        cudaDeviceSynchronize();
        probe<<<148, 512>>>(nw, iters, d_c, d_s);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
        long long c[148]; cudaMemcpy(c, d_c, sizeof c, cudaMemcpyDeviceToHost);
        const double bytes = (double)nw * iters * 4096.0;
        printf("{\"reading_warps\": %d, \"cycles\": %lld, \"tmem_read_bytes_per_clk_per_sm\": %.1f, \"clk_per_x32_load_per_warp\": %.1f}\n",
               nw, c[0], bytes / (double)c[0], (double)c[0] / iters);
    }
    return 0;
}
