// koala_b200 -- tcgen05 / TMA / mbarrier building blocks of the tensor-core mask estimator (masknet_fused.cuh): tile
// constants, PTX wrappers, UMMA descriptors, the GRU weight packing and the tensor-map encoder.
//
// The mask estimator is the middle stage of `pv_koala_process` (/root/reference/include/pv_koala.h:65-80); it replaces the
// reference's per-stream int8 x int16 dp2a mat-vec + LUT gate kernels by batched GEMMs over the
// stream dimension: bf16 operands staged by TMA into 128B-swizzled shared memory, tcgen05.mma with cta_group::2 (a CTA
// pair = 256 streams per tile; each SM loads its own 128 activation rows and HALF of the weight rows) accumulating fp32 in
// TMEM, gates / state update / activation fused into the epilogue that reads TMEM back with tcgen05.ld.
//
// GRU tile = 256 streams x 64 hidden units.  TMEM columns per accumulator buffer: [n_x 0..63 | r 64..127 | z 128..191 | n_h 192..255].
//   h part (K = H): one N=192 MMA per k-step, packed W_hh rows ordered r|z|n, D column base 64 -> r, z, n_h
//   x part (K = H): one N=192 MMA per k-step, packed W_ih rows ordered n|r|z, D column base  0 -> n_x, r, z
//   Either part may run first (it overwrites its columns), the other accumulates.  The accumulate flag is per instruction,
//   so the n_x and n_h columns, which only one part touches, are cleared by the epilogue with tcgen05.st before it hands the
//   buffer back.
// Linear tile = 256 streams x 128 outputs, one N=128 MMA per k-step.
#pragma once

#include <cuda.h>
#include <stdlib.h>

#include <string>

#include "koala_common.cuh"

namespace koala {

constexpr int kTcBlockM = 128;                // rows per CTA; a pair covers 256
constexpr int kTcPairM = 256;
constexpr int kTcBlockK = 64;                 // 64 bf16 = 128 bytes = one swizzle row
constexpr int kTcABytes = kTcBlockM * 128;    // 16 KB
constexpr int kTcEpiWarps = 16;               // 4 warps per TMEM lane quarter, each owning a quarter of the tile's columns
constexpr int kTcEpiThreads = kTcEpiWarps * 32;
constexpr int kTcStateWarp = 5 + kTcEpiWarps;    // moves the GRU state tiles (h(t-1) in, h(t) out) by TMA for the epilogue warps
constexpr int kTcAccCols = 256;
constexpr int kGruUnits = 64;                 // hidden units per GRU tile
constexpr int kGruRows = 3 * kGruUnits;       // packed weight rows per tile (both CTAs together)
constexpr int kTcTailBytes = 1024 /*align slack*/ + 256 /*barriers*/ + 2048 /*biases*/;

enum TcMode : int { kTcEnc = 0, kTcGru = 1, kTcDec = 2 };

// ---------------------------------------------------------------------------------------------------------------
// PTX wrappers
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
// one lane of a converged warp; keeps the surrounding code warp-uniform so that descriptors stay in uniform registers
// (issuing from inside `if (lane == 0)` made ptxas wrap every tcgen05.mma / TMA in an R2UR waterfall loop, ~140 cycles each)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `p` (a shared::cta pointer of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(const void *p, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    // relaxed: the only thing this arrival publishes is "my TMEM reads/writes are done", which tcgen05.fence orders;
    // a release at cluster scope compiles to MEMBAR.ALL.GPU and stalls on every outstanding global store
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA tile load whose completion bytes are signalled on a barrier that may live in the peer CTA of the pair
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap *map, uint32_t bar_cluster_addr, void *dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
        : "memory");
}
// same, multicast: the tile lands at the same smem offset in every CTA of `mask`, and each destination's completion bytes
// are signalled on the barrier at this offset in the leader of the destination's pair
__device__ __forceinline__ void tma_load_2d_pair_mc(const CUtensorMap *map, uint32_t bar_cluster_addr, void *dst, int c0, int c1,
                                                    uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, "
        "%4}], [%2], %5;" ::"r"(smem_u32(dst)),
        "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "h"(mask)
        : "memory");
}
// CTA-local tile load (completion on a barrier of this CTA) and tile store (bulk async-group completion)
__device__ __forceinline__ void tma_load_2d_local(const CUtensorMap *map, uint64_t *bar, void *dst, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                     smem_u32(dst)),
                 "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, const void *src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(smem_u32(src)), "r"(c0),
                 "r"(c1)
                 : "memory");
}
// pull a contiguous global range into L2 (no smem, no completion): used to turn the next tile's first-touch HBM misses into L2 hits
__device__ __forceinline__ void prefetch_l2(const void *p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap *map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc_pair(uint32_t *dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// arrives (once the MMAs issued so far have completed) on the barrier at this smem offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_pair(uint64_t *bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"(mask)
                 : "memory");
}
// D[tmem, 256 rows over the pair] (+)= A[smem, 128 rows per CTA] * B[smem, N/2 rows per CTA]^T, bf16 x bf16 -> fp32
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_zero8(uint32_t taddr) {
    const uint32_t z = 0;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr), "r"(z) : "memory");
}
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// K-major, 128B-swizzled operand tile: rows at 128 B pitch, 8-row groups at 1024 B (SBO), version 1 (sm_100), layout 2
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
    return (uint64_t) ((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t) 1 << 16) | ((uint64_t) (1024 >> 4) << 32) |
           ((uint64_t) 1 << 46) | ((uint64_t) 2 << 61);
}
// kind::f16 instruction descriptor: D fp32, A/B bf16, both K-major
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t) (N >> 3) << 17) | ((uint32_t) (M >> 4) << 24);
}

__device__ __forceinline__ uint4 pack_bf16x8(const float *f) {
    uint4 u;
    __nv_bfloat162 p;
    p = __floats2bfloat162_rn(f[0], f[1]); u.x = *reinterpret_cast<uint32_t *>(&p);
    p = __floats2bfloat162_rn(f[2], f[3]); u.y = *reinterpret_cast<uint32_t *>(&p);
    p = __floats2bfloat162_rn(f[4], f[5]); u.z = *reinterpret_cast<uint32_t *>(&p);
    p = __floats2bfloat162_rn(f[6], f[7]); u.w = *reinterpret_cast<uint32_t *>(&p);
    return u;
}

// f[i] -= the bf16 value it was just rounded to (exact in fp32: Sterbenz): what remains is the next operand plane of fp32 mode
__device__ __forceinline__ void sub_bf16x8(float *f, const uint4 &u) {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        f[2 * i] -= __uint_as_float(w[i] << 16);
        f[2 * i + 1] -= __uint_as_float(w[i] & 0xffff0000u);
    }
}

// Packs the three gate rows of each 64-unit tile contiguously: packed[(n * 3 + slot) * 64 + u][k] = W[gate * H + n * 64 + u][k]
// with gate = order[slot] (PyTorch gate numbering r=0, z=1, n=2): W_ih uses n|r|z, W_hh uses r|z|n (see file header).
__global__ void pack_gru_weights_kernel(const __nv_bfloat16 *__restrict__ W, __nv_bfloat16 *__restrict__ packed, int H, int g0,
                                        int g1, int g2) {
    const int prow = blockIdx.x;
    const int n = prow / kGruRows, slot = (prow % kGruRows) / kGruUnits, u = prow % kGruUnits;
    const int gate = slot == 0 ? g0 : slot == 1 ? g1 : g2;
    const uint4 *src = reinterpret_cast<const uint4 *>(W + (size_t) (gate * H + n * kGruUnits + u) * H);
    uint4 *dst = reinterpret_cast<uint4 *>(packed + (size_t) prow * H);
    for (int i = threadIdx.x; i < H / 8; i += blockDim.x) dst[i] = src[i];
}

// ---------------------------------------------------------------------------------------------------------------
// host side
struct TcModel {
    int H = 0, L = 0, Bp = 0;
    int tcap = 1, e_ring = 1;      // step slots of feat / mask, slots of the encoder-output ring e
    int planes = 1;                // bf16 planes per activation operand: 1 (bf16 mode) or 3 (fp32 mode), masknet_fused.cuh
    const __nv_bfloat16 *enc_w = nullptr, *dec_w = nullptr, *wih[kMaxLayers] = {}, *whh[kMaxLayers] = {};
    const float *enc_b = nullptr, *dec_b = nullptr, *bih[kMaxLayers] = {}, *bhh[kMaxLayers] = {};
    __nv_bfloat16 *feat = nullptr, *e = nullptr, *hb[2] = {};
    float *h[2] = {}, *mask = nullptr;
};

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// [rows][cols] row-major matrix, box = 128 bytes x box_rows, 128B swizzle; bf16 (64 elements per box row) or fp32 (32)
static bool encode_2d(EncodeTiledFn fn, CUtensorMap *map, const void *base, uint64_t rows, uint64_t cols, uint32_t box_rows,
                      bool f32 = false, bool plain32 = false) {
    const cuuint64_t dims[2] = {cols, rows};
    const cuuint64_t strides[1] = {cols * (f32 ? 4 : 2)};
    const cuuint32_t box[2] = {(cuuint32_t) ((f32 || plain32) ? 32 : kTcBlockK), box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return fn(map, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, plain32 ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace koala
