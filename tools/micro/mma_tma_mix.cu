// Microbenchmark: does a concurrent TMA stream into shared memory slow tcgen05.mma down?  One CTA per SM: warp 0 streams
// 2-D tensor tiles into a ring (re-issuing each stage when it lands), warp 1 issues N=192 / N=256 MMAs on a fixed tile.
// Variants: TMA stream on/off, tcgen05.commit after every 4 MMAs on/off.  Development aid for DESIGN.md.
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t a) {
    return (uint64_t) ((a & 0x3FFFFu) >> 4) | ((uint64_t) 1 << 16) | ((uint64_t) (1024 >> 4) << 32) | ((uint64_t) 1 << 46) | ((uint64_t) 2 << 61);
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) { return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t) (N >> 3) << 17) | ((uint32_t) (M >> 4) << 24); }
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

constexpr int STAGES = 4, ROWS = 320;    // 40 KB per stage: A 128 rows + B 192 rows

template <int N, bool TMA_ON, bool COMMIT4, bool WALK>
__global__ void __launch_bounds__(128, 1) mix_kernel(const __grid_constant__ CUtensorMap map, long long *out, int mma_iters, int total_rows) {
    extern __shared__ uint8_t raw[];
    uint8_t *smem = (uint8_t *) (((uintptr_t) raw + 1023) & ~(uintptr_t) 1023);
    uint8_t *ring = smem + 16384 + 32768;
    __shared__ uint64_t bar[STAGES], done_bar, cbar;
    __shared__ uint32_t slot;
    __shared__ volatile int stop;
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) ((uint32_t *) smem)[i] = 0x3c003c00u + i;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) mbar_init(&bar[s], 1);
        mbar_init(&done_bar, 1); mbar_init(&cbar, 1);
        stop = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x >= 32 && threadIdx.x < 64) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t d = slot;
    if (threadIdx.x == 0 && TMA_ON) {
        const int row_blocks = total_rows / ROWS;
        int rb = (blockIdx.x * 5) % row_blocks, kb = 0;
        long long loads = 0;
        for (int it = 0; !stop; ++it) {
            const int s = it % STAGES, ph = (it / STAGES) & 1;
            if (it >= STAGES) mbar_wait(&bar[s], ph ^ 1);
            mbar_expect_tx(&bar[s], ROWS * 128);
            tma_load_2d(&map, &bar[s], ring + s * ROWS * 128, kb * 64, rb * ROWS);
            tma_load_2d(&map, &bar[s], ring + s * ROWS * 128 + (ROWS / 2) * 128, kb * 64, rb * ROWS + ROWS / 2);
            if (++kb == 8) { kb = 0; rb = (rb + 1) % row_blocks; }
            loads++;
        }
        out[1024 + blockIdx.x] = loads;
        // drain: wait for everything in flight (each stage's latest phase)
    }
    if (threadIdx.x == 32) {
        const uint64_t adesc = make_sw128_desc(smem_u32(smem)), bdesc = make_sw128_desc(smem_u32(smem + 16384));
        const uint32_t idesc = make_idesc(128, N);
        long long t0 = clock64();
        for (int it = 0; it < mma_iters; ++it) {
            const uint32_t off = WALK ? (uint32_t) ((it % STAGES) * ROWS * 128) >> 4 : 0u;   // another 28 KB stage each k-block
            const uint64_t ad0 = WALK ? make_sw128_desc(smem_u32(ring)) + off : adesc, bd0 = WALK ? make_sw128_desc(smem_u32(ring) + 16384) + off : bdesc;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d + (it & 1) * 256), "l"(ad0 + 2 * k), "l"(bd0 + 2 * k), "r"(idesc), "r"(1u) : "memory");
            if (COMMIT4) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&cbar)) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&done_bar)) : "memory");
        mbar_wait(&done_bar, 0);
        long long t1 = clock64();
        out[blockIdx.x] = t1 - t0;
        stop = 1;
    }
    __syncthreads();
    // let outstanding TMA loads land before the CTA exits
    if (threadIdx.x == 0 && TMA_ON) { for (volatile int i = 0; i < 20000; i++) {} }
    __syncthreads();
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x >= 32 && threadIdx.x < 64) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(d), "r"(512) : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int N, bool TMA_ON, bool COMMIT4, bool WALK = false>
void run(const CUtensorMap &map, int total_rows, int grid) {
    long long *out;
    cudaMalloc(&out, 2048 * sizeof(long long));
    cudaMemset(out, 0, 2048 * sizeof(long long));
    const int smem = 16384 + 32768 + STAGES * ROWS * 128 + 1024, iters = 1500;
    auto kern = mix_kernel<N, TMA_ON, COMMIT4, WALK>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int rep = 0; rep < 2; rep++) {
        kern<<<grid, 128, smem>>>(map, out, iters, total_rows);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("failed: %s\n", cudaGetErrorString(e)); return; }
    }
    long long h[2048];
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    double sum = 0, loads = 0;
    for (int i = 0; i < grid; i++) { sum += (double) h[i]; loads += (double) h[1024 + i]; }
    const double cyc = sum / grid;
    printf("walk %d N=%3d TMA %-3s commit4 %-3s grid %3d: %6.1f cycles / MMA (%5.0f MAC/clk/SM), TMA ingest %5.1f B/clk/SM\n", (int) WALK, N, TMA_ON ? "on" : "off", COMMIT4 ? "on" : "off",
           grid, cyc / (iters * 4), 128.0 * N * 16 / (cyc / (iters * 4)), loads / grid * ROWS * 128 / cyc);
    cudaFree(out);
}

int main() {
    void *fnp = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
    EncodeTiledFn fn = (EncodeTiledFn) fnp;
    const int rows = 320 * 48, kcols = 512;
    void *base;
    cudaMalloc(&base, (size_t) rows * kcols * 2);
    cudaMemset(base, 0, (size_t) rows * kcols * 2);
    CUtensorMap map;
    const cuuint64_t dims[2] = {(cuuint64_t) kcols, (cuuint64_t) rows};
    const cuuint64_t strides[1] = {(cuuint64_t) kcols * 2};
    const cuuint32_t box[2] = {64, ROWS / 2};
    const cuuint32_t estr[2] = {1, 1};
    if (fn(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode failed\n"); return 1; }
    for (int grid : {148}) {
        run<192, false, false>(map, rows, grid);
        run<192, false, true>(map, rows, grid);
        run<192, true, false>(map, rows, grid);
        run<192, true, true>(map, rows, grid);
        run<256, false, true>(map, rows, grid);
        run<256, true, true>(map, rows, grid);
        run<192, false, true, true>(map, rows, grid);
        run<192, true, true, true>(map, rows, grid);
    }
    return 0;
}
