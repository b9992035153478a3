// Microbenchmark: cycles per tcgen05.mma (kind::f16, bf16, SS operands already resident in smem) for several N and
// cta_group 1 / 2.  Development aid: gives the measured MMA floor that DESIGN.md quotes.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/mma_rate tools/micro/mma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t a) {
    return (uint64_t) ((a & 0x3FFFFu) >> 4) | ((uint64_t) 1 << 16) | ((uint64_t) (1024 >> 4) << 32) | ((uint64_t) 1 << 46) | ((uint64_t) 2 << 61);
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t) (N >> 3) << 17) | ((uint32_t) (M >> 4) << 24);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(c)); }
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}

template <int CG, int N, int KSTEPS_PER_COMMIT>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(long long *out, int iters, int col0) {
    extern __shared__ uint8_t raw[];
    uint8_t *smem = (uint8_t *) (((uintptr_t) raw + 1023) & ~(uintptr_t) 1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) ((uint32_t *) smem)[i] = 0x3c003c00u + i;   // finite bf16 junk
    if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x < 32) {
        if (CG == 1) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (CG == 2) { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); } else __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t d = slot;
    if (threadIdx.x == 0 && rank == 0) {
        const uint64_t adesc = make_sw128_desc(smem_u32(smem)), bdesc = make_sw128_desc(smem_u32(smem + 16384));
        const uint32_t idesc = make_idesc(CG == 2 ? 256 : 128, N);
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int k = 0; k < KSTEPS_PER_COMMIT; ++k) {
                const uint64_t ad = adesc + 2 * (k & 3), bd = bdesc + 2 * (k & 3);
                if (CG == 1) asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d + col0 + (it & 1) * 256), "l"(ad), "l"(bd), "r"(idesc), "r"(1u) : "memory");
                else asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d + col0 + (it & 1) * 256), "l"(ad), "l"(bd), "r"(idesc), "r"(1u) : "memory");
            }
        }
        if (CG == 1) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        else asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "h"((uint16_t) 1) : "memory");
        mbar_wait(&bar, 0);
        long long t1 = clock64();
        out[blockIdx.x] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (CG == 2) { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); } else __syncthreads();
    if (threadIdx.x < 32) {
        if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(d), "r"(512) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(d), "r"(512) : "memory");
    }
}

template <int CG, int N>
void run(const char *name, int grid, int col0 = 0) {
    long long *out;
    cudaMalloc(&out, 1024 * sizeof(long long));
    cudaMemset(out, 0, 1024 * sizeof(long long));
    const int smem = 16384 + 32768 + 1024, iters = 512;
    auto kern = mma_rate_kernel<CG, N, 4>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    for (int rep = 0; rep < 2; rep++) {
        cudaError_t e = cudaLaunchKernelEx(&cfg, kern, out, iters, col0);
        if (e != cudaSuccess) { printf("%s launch failed: %s\n", name, cudaGetErrorString(e)); return; }
        e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s failed: %s\n", name, cudaGetErrorString(e)); return; }
    }
    long long h[1024];
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    double cyc = (double) h[0] / (iters * 4);
    const int M = CG == 2 ? 256 : 128;
    printf("%-28s grid %4d: %7.1f cycles / MMA  -> %6.0f MAC/clk/SM\n", name, grid, cyc, (double) M * N * 16 / cyc / CG);
    cudaFree(out);
}

int main() {
    run<1, 64>("cg1 M128 N64", 1);
    run<1, 128>("cg1 M128 N128", 1);
    run<1, 192>("cg1 M128 N192", 1);
    run<1, 256>("cg1 M128 N256", 1);
    run<2, 64>("cg2 M256 N64", 2);
    run<2, 128>("cg2 M256 N128", 2);
    run<2, 192>("cg2 M256 N192", 2);
    run<2, 256>("cg2 M256 N256", 2);
    run<1, 256>("cg1 M128 N256 (full chip)", 148);
    run<2, 256>("cg2 M256 N256 (full chip)", 148);
    run<2, 192>("cg2 M256 N192 (full chip)", 148);
    run<2, 192>("cg2 M256 N192 col base 64", 2, 64);
    run<2, 192>("cg2 M256 N192 col base 32", 2, 32);
    run<2, 128>("cg2 M256 N128 col base 64", 2, 64);
    run<1, 192>("cg1 M128 N192 col base 64", 1, 64);
    return 0;
}
