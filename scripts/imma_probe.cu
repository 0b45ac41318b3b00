// imma_probe.cu — measured int8 tensor-core peak of tcgen05.mma (kind::i8, M128 x N256 x K32, both operands in shared
// memory, accumulators in TMEM), one CTA per SM, one elected thread issuing back-to-back MMAs on resident operands.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/imma_probe scripts/imma_probe.cu && /tmp/imma_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t sbo)   // K-major, no swizzle: LBO 128 B, SBO = 8-row group stride
{
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
extern __shared__ __align__(1024) uint8_t smem[];

__global__ void __launch_bounds__(128, 1) probe(int n_mma, int N, long long *cycles)
{
    __shared__ uint32_t s_tmem;
    __shared__ alignas(8) uint64_t s_bar;
    const int warp = threadIdx.x >> 5;
    uint8_t *sA = smem, *sB = smem + 128 * 32;           // one K step (32 int8) of 128 and N rows: SBO = 2 chunks * 128 B
    for (int i = threadIdx.x; i < (128 + 256) * 32 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x01ff01ffu * (i | 1);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&s_bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(s32(&s_tmem)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    long long t0 = 0, t1 = 0;
    if (warp == 1) {
        uint32_t pred;
        asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
        if (pred) {
            const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            const uint64_t ad = desc(s32(sA), 256), bd = desc(s32(sB), 256);
            t0 = clock64();
            for (int i = 0; i < n_mma; i++) {
                const uint32_t d = tmem + (uint32_t)((i & 1) * 256);     // two accumulators alternate
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                             ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(i > 1 ? 1u : 0u) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&s_bar)) : "memory");
            asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(s32(&s_bar)) : "memory");
            t1 = clock64();
            cycles[blockIdx.x] = t1 - t0;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

int main()
{
    long long *d_c; cudaMalloc(&d_c, 148 * sizeof(long long));
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    const int n_mma = 20000;
    for (int N : {64, 128, 256}) {
        probe<<<148, 128, 64 * 1024>>>(n_mma, N, d_c);
        cudaDeviceSynchronize();
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        probe<<<148, 128, 64 * 1024>>>(n_mma, N, d_c);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        long long c[148]; cudaMemcpy(c, d_c, sizeof c, cudaMemcpyDeviceToHost);
        const double macs = 148.0 * n_mma * 128.0 * N * 32.0;
        printf("{\"N\": %d, \"cycles_per_mma\": %.1f, \"int8_dense_Pop_per_s\": %.3f, \"Tpairs_per_s_at_K256\": %.2f, \"ms\": %.3f}\n",
               N, (double)c[0] / n_mma, 2 * macs / (ms * 1e-3) / 1e15, macs / 256 / (ms * 1e-3) / 1e12, ms);
    }
    return 0;
}
