// Microbenchmark: TMA tile ingest rate per SM (no consumer): a ring of S stages, one thread issues 2-D tensor loads of
// [rows x 64] bf16 boxes (128B swizzle) from an L2-resident matrix and re-issues each stage as soon as it has landed.
// Development aid for DESIGN.md ("how many bytes per clock can one SM pull from L2 through TMA").
// build+run: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tma_rate tools/micro/tma_rate.cu -lcuda && /tmp/tma_rate
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(c)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap *map, uint64_t *bar, void *dst, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}

template <int ROWS, int STAGES, int ISSUERS>
__global__ void __launch_bounds__(128, 1) tma_rate_kernel(const __grid_constant__ CUtensorMap map, long long *out, int iters, int total_rows, int kcols) {
    extern __shared__ uint8_t raw[];
    uint8_t *smem = (uint8_t *) (((uintptr_t) raw + 1023) & ~(uintptr_t) 1023);
    __shared__ uint64_t bars[ISSUERS * STAGES];
    if (threadIdx.x == 0) {
        for (int s = 0; s < ISSUERS * STAGES; s++) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0 && (threadIdx.x >> 5) < ISSUERS) {
        const int w = threadIdx.x >> 5;
        uint64_t *bar = bars + w * STAGES;
        smem += w * STAGES * ROWS * 128;
        const int row_blocks = total_rows / ROWS, kblocks = kcols / 64;
        int rb = (blockIdx.x * 7 + w * 3) % row_blocks, kb = 0;
        long long t0 = clock64();
        for (int it = 0; it < iters + STAGES; ++it) {
            const int s = it % STAGES, ph = (it / STAGES) & 1;
            if (it >= STAGES) mbar_wait(&bar[s], ph ^ 1);          // previous load of this stage has landed
            if (it < iters) {
                mbar_expect_tx(&bar[s], ROWS * 128);
                tma_load_2d(&map, &bar[s], smem + s * ROWS * 128, kb * 64, rb * ROWS);
                if (++kb == kblocks) { kb = 0; rb = (rb + 1) % row_blocks; }
            }
        }
        long long t1 = clock64();
        if (w == 0) out[blockIdx.x] = t1 - t0;
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int ROWS, int STAGES, int ISSUERS = 1>
void run(EncodeTiledFn fn, void *base, int total_rows, int kcols, int grid) {
    CUtensorMap map;
    const cuuint64_t dims[2] = {(cuuint64_t) kcols, (cuuint64_t) total_rows};
    const cuuint64_t strides[1] = {(cuuint64_t) kcols * 2};
    const cuuint32_t box[2] = {64, ROWS};
    const cuuint32_t estr[2] = {1, 1};
    if (fn(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode failed\n"); return; }
    long long *out;
    cudaMalloc(&out, 1024 * sizeof(long long));
    const int smem = ISSUERS * STAGES * ROWS * 128 + 1024, iters = 2000;
    auto kern = tma_rate_kernel<ROWS, STAGES, ISSUERS>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int rep = 0; rep < 2; rep++) {
        kern<<<grid, 128, smem>>>(map, out, iters, total_rows, kcols);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("failed: %s\n", cudaGetErrorString(e)); return; }
    }
    long long h[1024];
    cudaMemcpy(h, out, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
    double worst = 0, sum = 0;
    for (int i = 0; i < grid; i++) { if ((double) h[i] > worst) worst = (double) h[i]; sum += (double) h[i]; }
    const double bytes = (double) iters * ROWS * 128 * ISSUERS;
    printf("%d issuer(s), box %3d rows x 128 B, %d stages (%3d KB in flight), grid %3d: %6.1f B/clk/SM (mean), %6.1f (slowest SM); %6.0f cycles per load\n", ISSUERS, ROWS, STAGES,
           STAGES * ROWS * 128 / 1024, grid, bytes / (sum / grid), bytes / worst, worst / iters);
    cudaFree(out);
}

int main() {
    void *fnp = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
    EncodeTiledFn fn = (EncodeTiledFn) fnp;
    const int rows = 16384, kcols = 512;            // 16 MB bf16 matrix: L2 resident
    void *base;
    cudaMalloc(&base, (size_t) rows * kcols * 2);
    cudaMemset(base, 0, (size_t) rows * kcols * 2);
    for (int grid : {1, 148}) {
        run<128, 2>(fn, base, rows, kcols, grid);
        run<128, 4>(fn, base, rows, kcols, grid);
        run<128, 7>(fn, base, rows, kcols, grid);
        run<128, 12>(fn, base, rows, kcols, grid);
        run<256, 6>(fn, base, rows, kcols, grid);
        run<64, 14>(fn, base, rows, kcols, grid);
        run<32, 24>(fn, base, rows, kcols, grid);
        run<128, 4, 2>(fn, base, rows, kcols, grid);
        run<128, 3, 4>(fn, base, rows, kcols, grid);
        run<64, 6, 4>(fn, base, rows, kcols, grid);
    }
    // the same 16 MB as a K-blocked matrix: row pitch 128 B, so a [128 x 64] box is one contiguous 16 KB range
    printf("row pitch 128 B (contiguous boxes):\n");
    for (int grid : {1, 148}) {
        run<128, 4>(fn, base, rows * (kcols / 64), 64, grid);
        run<128, 4, 2>(fn, base, rows * (kcols / 64), 64, grid);
        run<128, 3, 4>(fn, base, rows * (kcols / 64), 64, grid);
        run<96, 4, 4>(fn, base, rows * (kcols / 64), 64, grid);
    }
    return 0;
}
