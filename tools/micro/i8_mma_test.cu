// Known-answer test of tcgen05.mma kind::i8 on sm_100a (development aid for the fixed-point mask path, masknet_i8.cuh):
// one CTA, D[128 x 128] (int32, TMEM) = A[128 x 128] (s8 or u8) * B[128 x 128]^T (s8), operands K-major in 128B-swizzled smem.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o gpurun_i8_mma_test tools/micro/i8_mma_test.cu && ./gpurun_i8_mma_test
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t a) {
    return (uint64_t) ((a & 0x3FFFFu) >> 4) | ((uint64_t) 1 << 16) | ((uint64_t) (1024 >> 4) << 32) | ((uint64_t) 1 << 46) | ((uint64_t) 2 << 61);
}
// kind::i8 instruction descriptor: D s32 (c_format 2), A s8 (1) or u8 (0), B s8, both K-major
__host__ __device__ constexpr uint32_t make_idesc_i8(int M, int N, bool a_signed) {
    return (2u << 4) | ((a_signed ? 1u : 0u) << 7) | (1u << 10) | ((uint32_t) (N >> 3) << 17) | ((uint32_t) (M >> 4) << 24);
}

__global__ void __launch_bounds__(128) test_kernel(const uint8_t *A, const int8_t *B, int32_t *D, int a_signed) {
    extern __shared__ uint8_t raw[];
    uint8_t *smem = (uint8_t *) (((uintptr_t) raw + 1023) & ~(uintptr_t) 1023);
    uint8_t *sa = smem, *sb = smem + 16384;
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    // row `tid` of A and of B: 8 chunks of 16 bytes, chunk c stored at position c ^ (row & 7) (128B swizzle)
    for (int c = 0; c < 8; ++c) {
        *(uint4 *) (sa + tid * 128 + ((c ^ (tid & 7)) << 4)) = *(const uint4 *) (A + tid * 128 + c * 16);
        *(uint4 *) (sb + tid * 128 + ((c ^ (tid & 7)) << 4)) = *(const uint4 *) (B + tid * 128 + c * 16);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(&tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    if (tid == 0) {
        const uint64_t ad = make_sw128_desc(smem_u32(sa)), bd = make_sw128_desc(smem_u32(sb));
        const uint32_t idesc = make_idesc_i8(128, 128, a_signed != 0);
        for (int k = 0; k < 4; ++k) {      // K = 32 int8 = 32 bytes per instruction
            const uint32_t acc = k ? 1u : 0u;
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(ad + 2 * k), "l"(bd + 2 * k), "r"(idesc), "r"(acc)
                : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < 128; c0 += 8) {
        uint32_t r[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                     : "r"(tmem + ((uint32_t) (warp * 32) << 16) + c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 8; ++i) D[tid * 128 + c0 + i] = (int32_t) r[i];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem) : "memory");
}

int main() {
    std::vector<uint8_t> A(128 * 128);
    std::vector<int8_t> B(128 * 128);
    srand(1);
    for (auto &v : A) v = (uint8_t) (rand() & 255);
    for (auto &v : B) v = (int8_t) ((rand() & 255) - 128);
    uint8_t *dA; int8_t *dB; int32_t *dD;
    cudaMalloc(&dA, A.size()); cudaMalloc(&dB, B.size()); cudaMalloc(&dD, 128 * 128 * 4);
    cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 34 * 1024);
    int bad_total = 0;
    for (int a_signed = 0; a_signed < 2; ++a_signed) {
        cudaMemset(dD, 0xff, 128 * 128 * 4);
        test_kernel<<<1, 128, 34 * 1024>>>(dA, dB, dD, a_signed);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("a_signed=%d: CUDA error %s\n", a_signed, cudaGetErrorString(e)); return 1; }
        std::vector<int32_t> D(128 * 128);
        cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < 128; ++n) {
                int32_t ref = 0;
                for (int k = 0; k < 128; ++k) ref += (a_signed ? (int32_t) (int8_t) A[m * 128 + k] : (int32_t) A[m * 128 + k]) * (int32_t) B[n * 128 + k];
                if (ref != D[m * 128 + n] && bad++ < 4) printf("  a_signed=%d D[%d][%d] = %d, expected %d\n", a_signed, m, n, D[m * 128 + n], ref);
            }
        printf("kind::i8, A %s, B s8: %d mismatches of 16384\n", a_signed ? "s8" : "u8", bad);
        bad_total += bad;
    }
    return bad_total != 0;
}
