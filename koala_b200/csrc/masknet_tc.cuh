// koala_b200 -- mask estimator, tensor-core path (BASELINE.json configs[2..4]: "bf16 tensor-core mask-estimator GEMMs").
//
// Middle stage of `pv_koala_process` (/root/reference/include/pv_koala.h:65-80); replaces the reference's per-stream
// int8 x int16 dp2a mat-vec + LUT gate kernels by batched GEMMs over the stream dimension:
// bf16 operands staged by TMA into 128B-swizzled shared memory, tcgen05.mma with cta_group::2 (a CTA pair = 256 streams
// per tile, each SM loads its own 128 activation rows and HALF of the weight rows) accumulating fp32 in TMEM, gates /
// state update / activation fused into the epilogue that reads TMEM back with tcgen05.ld.
//
// Why CTA pairs and clusters: the first version (cta_group::1, 128 x 192 tiles) needed 40 KB of operands per 64-wide
// k-block against ~400 cycles of MMA; sharing the weight tile between the two SMs of a pair cuts that to 28 KB, and the GRU
// kernel additionally multicasts each activation tile to the two pairs of a 4-CTA cluster that work on neighbouring unit
// tiles (20 KB of L2 reads per CTA per k-block).  Measurements behind every choice: profiles/r01_step_summary.md.
//
// Persistent clusters, 21 warps per CTA: warps 0,1 / 3,4 = TMA producers (activations / weights x even / odd k-blocks),
// warp 2 = TMEM allocator + (leader CTA only) MMA issuer, warps 5-20 = epilogue (four warps per TMEM lane quarter).
// Two accumulator buffers of 256 TMEM columns let the epilogue of tile i overlap the MMAs of tile i+1.
//
// GRU tile = 256 streams x 64 hidden units.  TMEM columns per buffer: [n_x 0..63 | r 64..127 | z 128..191 | n_h 192..255].
//   x-part (K = H): one N=192 MMA per k-step, packed W_ih rows ordered n|r|z, D column base   0 -> n_x, r, z
//   h-part (K = H): one N=192 MMA per k-step, packed W_hh rows ordered r|z|n, D column base  64 -> r, z (accumulated), n_h
//   n_h has no zero-initialising MMA of its own (the accumulate flag is per instruction), so the epilogue clears those
//   64 columns with tcgen05.st before it hands the buffer back.
// Linear tile = 256 streams x 128 outputs, one N=128 MMA per k-step; the output tile is staged in smem and TMA-stored.
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
constexpr int kTcStateWarp = 5 + kTcEpiWarps;    // GRU: moves the state tiles (h(t-1) in, h(t) out) by TMA for the epilogue warps
constexpr int kTcThreads = 160 + kTcEpiThreads + 32; // 2 TMA warps (A) + MMA warp + 2 TMA warps (B) + 16 epilogue warps + state warp
constexpr int kTcAccCols = 256;
constexpr int kGruUnits = 64;                 // hidden units per GRU tile
constexpr int kGruRows = 3 * kGruUnits;       // packed weight rows per tile (both CTAs together)
constexpr int kTcTailBytes = 1024 /*align slack*/ + 256 /*barriers*/ + 2048 /*biases*/;

enum TcMode : int { kTcEnc = 0, kTcGru = 1, kTcDec = 2 };

// A cluster is a PM x PN grid of CTA pairs: pair (qm, qn) works on M-tile cm * PM + qm and N-tile cn * PN + qn of the
// cluster tile (cm, cn).  The PN pairs of a row need the same activation rows and the PM pairs of a column the same
// weight rows, so every CTA fetches only 1/PN of its A block and 1/PM of its B block and TMA-multicasts the piece to
// the CTAs that share it: L2 -> SM traffic per pair-tile drops from (A + B) to (A / PN + B / PM).
template <int MODE> struct TcCfg {
    static constexpr bool kGru = MODE == kTcGru;
    static constexpr int kPM = 1, kPN = kGru ? 2 : 1;   // GRU: 4-CTA clusters, activations multicast across 2 unit tiles
    static constexpr int kPairs = kPM * kPN, kCluster = 2 * kPairs;
    static constexpr int kLinN = 128;                                   // outputs per linear pair-tile (one N=128 MMA per k-step)
    static constexpr int kBRowsHalf = kGru ? kGruRows / 2 : kLinN / 2;  // weight rows each CTA of a pair holds
    static constexpr int kARowsPiece = kTcBlockM / kPN;                  // rows of A this CTA fetches (and multicasts)
    static constexpr int kBRowsPiece = kBRowsHalf / kPM;                 // rows of B this CTA fetches (and multicasts)
    static constexpr int kStageBytes = kTcABytes + kBRowsHalf * 128;     // 28 KB | 24 KB landing per CTA per stage
    static constexpr int kStages = kGru ? 6 : 4;
    // GRU only: the epilogue's h tile travels by TMA too (coalesced, off the LSU): fp32 h(t-1) lands in kEpiF32Bytes, is
    // replaced in place by h(t), and the bf16 copy of h(t) is staged in kEpiBf16Bytes; both leave through TMA stores
    // Linear kernels stage their [128 x 128] output tile the same way (encoder: bf16, decoder: fp32) and TMA-store it.
    // GRU: the state tile is handled in two passes of 32 units, each with its own staging buffers (fp32 box [128][32] that
    // receives h(t-1) by TMA and is overwritten in place by h(t); bf16 box [128][32]) -- 48 KB instead of 80 KB for whole
    // tiles, which pays for a 6th operand stage (the operand pipeline is the bottleneck, the epilogue has slack).
    static constexpr int kEpiF32Bytes = kGru ? kTcBlockM * 32 * 4 : (MODE == kTcDec ? kTcBlockM * kLinN * 4 : 0);    // 128B-swizzled boxes of 32 floats
    static constexpr int kEpiBf16Bytes = kGru ? 2 * kTcBlockM * 32 * 2 : (MODE == kTcEnc ? kTcBlockM * kLinN * 2 : 0);  // GRU: 2 plain boxes of 32 bf16; encoder: swizzled boxes of 64
    static constexpr int kEpiF32Bufs = kGru ? 2 : 1;
    static constexpr int kSmemBytes = kStages * kStageBytes + kEpiF32Bufs * kEpiF32Bytes + kEpiBf16Bytes + kTcTailBytes;
};

struct TcArgs {
    int num_m_tiles, num_n_tiles;   // m tiles of 256 streams
    int kb_per_part;            // K / 64 of one operand part (GRU has two parts: x then h)
    int H;
    const float *bias0;         // enc/dec bias | GRU b_ih
    const float *bias1;         // GRU b_hh
    const float *h_prev;        // GRU fp32 state in  [Bp][H]
    float *h_next;              // GRU fp32 state out [Bp][H]
    __nv_bfloat16 *out_bf16;    // enc: e [Bp][H]; GRU: bf16 copy of h_next
    float *out_f32;             // dec: mask [Bp][256]
    const __nv_bfloat16 *a0, *a1;   // GRU: the two activation operands [Bp][H] (x, h(t-1) in bf16), for L2 prefetch of the next tile
    long long *trace;           // optional clock64() timeline of CTAs 0 and 1 (KOALA_TC_TRACE=1), else nullptr
};

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

// ---------------------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __cluster_dims__(TcCfg<MODE>::kCluster, 1, 1) __launch_bounds__(kTcThreads, 1)
tc_masknet_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
                  const __grid_constant__ CUtensorMap map_b0, const __grid_constant__ CUtensorMap map_b1,
                  const __grid_constant__ CUtensorMap map_hp, const __grid_constant__ CUtensorMap map_hn,
                  const __grid_constant__ CUtensorMap map_hb, const TcArgs args) {
    using Cfg = TcCfg<MODE>;
    constexpr bool kGru = Cfg::kGru;
    constexpr int kStages = Cfg::kStages, kStageBytes = Cfg::kStageBytes, kBRowsHalf = Cfg::kBRowsHalf;
    constexpr int kPM = Cfg::kPM, kPN = Cfg::kPN, kCluster = Cfg::kCluster;
    constexpr uint32_t kPairTx = 2u * kStageBytes;   // bytes landing in both CTAs of a pair per stage, signalled on its leader

    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t) 1023);
    uint8_t *s_hp = smem + kStages * kStageBytes;                 // GRU: fp32 h tile, boxes [128 rows][32 floats] x 2
    uint8_t *s_hb = s_hp + Cfg::kEpiF32Bufs * Cfg::kEpiF32Bytes;  // GRU: bf16 h(t) tile, box [128 rows][64 bf16]
    uint8_t *tail = s_hb + Cfg::kEpiBf16Bytes;
    uint64_t *bars = reinterpret_cast<uint64_t *>(tail);
    uint64_t *full_bar = bars, *empty_bar = bars + kStages;
    uint64_t *tmem_full = bars + 2 * kStages, *tmem_empty = bars + 2 * kStages + 2;
    uint64_t *hp_full = bars + 2 * kStages + 4;                   // [2]: one per fp32 tile buffer
    uint64_t *staged = bars + 2 * kStages + 6;                    // [2]: h(t) of a pass is staged in shared memory
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * kStages + 8);
    float *s_bias = reinterpret_cast<float *>(tail + 256);        // [2 accumulator buffers][256]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t crank = cluster_ctarank();      // rank in the cluster = 2 * pair + position in pair
    const uint32_t rank = crank & 1;               // 0 = pair leader (issues the MMAs), 1 = peer
    const uint32_t leader = crank & ~1u;           // cluster rank of my pair's leader
    const int q = (int) (crank >> 1), qm = q % kPM, qn = q / kPM;
    const uint16_t pair_mask = (uint16_t) (3u << leader), all_mask = (uint16_t) ((1u << kCluster) - 1);
    long long *trace = (args.trace != nullptr && blockIdx.x < 2) ? args.trace + blockIdx.x * 512 : nullptr;
#define KTRACE(slot) do { if (trace && (slot) < 512) trace[(slot)] = clock64(); } while (0)
    if (threadIdx.x == 0) KTRACE(500);
    pdl_launch_dependents();
    const int cluster_id = blockIdx.x / kCluster, num_clusters = gridDim.x / kCluster;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&map_a0);
        prefetch_tmap(&map_b0);
        if (kGru) {
            prefetch_tmap(&map_a1);
            prefetch_tmap(&map_b1);
            prefetch_tmap(&map_hp);
            prefetch_tmap(&map_hn);
            prefetch_tmap(&map_hb);
        } else {
            prefetch_tmap(&map_hn);    // linear kernels: the output tile's store map
        }
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full_bar[s], 1);     // leader's copy is the one in use: 1 arrive.expect_tx + 2 CTAs' TMA bytes
            mbar_init(&empty_bar[s], Cfg::kPairs);   // one multicast tcgen05.commit from every pair leader of the cluster
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&tmem_full[b], 1);    // one multicast tcgen05.commit
            mbar_init(&tmem_empty[b], 2 * kTcEpiThreads); // leader's copy: the epilogue threads of both CTAs
        }
        mbar_init(&hp_full[0], 1);
        mbar_init(&hp_full[1], 1);
        mbar_init(&staged[0], kTcEpiThreads);
        mbar_init(&staged[1], kTcEpiThreads);
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc_pair(tmem_slot, 2 * kTcAccCols);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) KTRACE(501);
    pdl_wait();      // everything above (barriers, TMEM, tensor-map prefetch) overlapped the previous kernel's tail

    // cluster tiles: (num_m_tiles / PM) x (num_n_tiles / PN); every pair of the cluster walks the same sequence
    const int ctiles_n = args.num_n_tiles / kPN;
    const int num_tiles = (args.num_m_tiles / kPM) * ctiles_n;
    const int num_kb = (kGru ? 2 : 1) * args.kb_per_part;

    if (warp == 0 || warp == 1 || warp == 3 || warp == 4) {
        // ===================================================== TMA producers (both CTAs of every pair).  One warp can issue a
        // tensor load only every ~210-380 cycles whatever its size (tools/micro/tma_rate.cu), about one MMA k-block, so
        // the loads are spread over four warps: warps 0/1 fetch the activation (A) tiles of the even/odd k-blocks, warps
        // 3/4 the weight (B) tiles.  Everything stays warp-uniform (elect_one) so descriptors live in uniform registers.
        {
            const bool is_a = warp < 2;
            const int par = is_a ? warp : warp - 3;             // my k-block parity
            int pit = 0;
            long long g = par;                                  // running k-block index across tiles -> stage / phase
            // multicast destinations: A goes to the CTAs with my (qm, position), B to those with my (qn, position)
            uint16_t mask_a = 0, mask_b = 0;
#pragma unroll
            for (int j = 0; j < kPN; ++j) mask_a |= (uint16_t) (1u << (2 * (qm + kPM * j) + rank));
#pragma unroll
            for (int j = 0; j < kPM; ++j) mask_b |= (uint16_t) (1u << (2 * (j + kPM * qn) + rank));
            for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++pit) {
                const int m = (tile / ctiles_n) * kPM + qm, n = (tile % ctiles_n) * kPN + qn;
                if (warp == 0 && lane == 0) KTRACE(pit * 48 + 0);
                if (kGru && is_a) {
                    // my 64 rows of an activation operand are one contiguous 64 KB range: request the next tile's (and, at the
                    // very start, this tile's) into L2 now, a whole mainloop before its loads are issued.  warp 0: x, warp 1: h
                    const __nv_bfloat16 *src = warp == 0 ? args.a0 : args.a1;
                    const uint32_t bytes = (uint32_t) (Cfg::kARowsPiece * args.H * 2);
                    const int next = tile + num_clusters;
                    if (elect_one()) {
                        if (pit == 0) prefetch_l2(src + (size_t) (m * kTcPairM + (int) rank * kTcBlockM + qn * Cfg::kARowsPiece) * args.H, bytes);
                        if (next < num_tiles) {
                            const int m1 = (next / ctiles_n) * kPM + qm;
                            prefetch_l2(src + (size_t) (m1 * kTcPairM + (int) rank * kTcBlockM + qn * Cfg::kARowsPiece) * args.H, bytes);
                        }
                    }
                    __syncwarp();
                }
                for (int kb = par; kb < num_kb; kb += 2, g += 2) {
                    const int stage = (int) (g % kStages), phase = (int) ((g / kStages) & 1);
                    mbar_wait(&empty_bar[stage], phase ^ 1);     // slot free in every CTA of the cluster
                    if (is_a && lane == 0) KTRACE(pit * 48 + 16 + kb);
                    const bool elected = elect_one();
                    if (elected && is_a && rank == 0) mbar_expect_tx(&full_bar[stage], kPairTx);
                    const uint32_t full_leader = map_to_cta(&full_bar[stage], leader);
                    uint8_t *sa = smem + stage * kStageBytes, *sb = sa + kTcABytes;
                    const bool second = kGru && kb >= args.kb_per_part;
                    const int kc = (second ? kb - args.kb_per_part : kb) * kTcBlockK;
                    if (!elected) {
                    } else if (is_a) {
                        const CUtensorMap *ma = second ? &map_a1 : &map_a0;
                        const int arow = m * kTcPairM + (int) rank * kTcBlockM + qn * Cfg::kARowsPiece;
                        if (kPN > 1) tma_load_2d_pair_mc(ma, full_leader, sa + qn * Cfg::kARowsPiece * 128, kc, arow, mask_a);
                        else tma_load_2d_pair(ma, full_leader, sa, kc, arow);
                    } else {
                        const CUtensorMap *mb = second ? &map_b1 : &map_b0;
                        const int brow = n * 2 * kBRowsHalf + (int) rank * kBRowsHalf + qm * Cfg::kBRowsPiece;
                        if (kPM > 1) tma_load_2d_pair_mc(mb, full_leader, sb + qm * Cfg::kBRowsPiece * 128, kc, brow, mask_b);
                        else tma_load_2d_pair(mb, full_leader, sb, kc, brow);
                    }
                    __syncwarp();
                }
                if (warp == 0 && lane == 0) KTRACE(pit * 48 + 1);
            }
        }
    } else if (warp == 2) {
        // ===================================================== MMA issuer (one thread of the leader CTA drives both SMs)
        if (rank == 0) {
            int stage = 0, phase = 0, it = 0;
            for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++it) {
                const int ab = it & 1, aphase = (it >> 1) & 1;
                mbar_wait(&tmem_empty[ab], aphase);      // both CTAs' epilogues have released (and cleared) this buffer
                tc_fence_after();
                if (lane == 0) KTRACE(it * 48 + 2);
                const uint32_t d = tmem_base + ab * kTcAccCols;
                for (int kb = 0; kb < num_kb; ++kb) {
                    if (it == 1 && lane == 0) KTRACE(256 + kb * 4 + 0);
                    mbar_wait(&full_bar[stage], phase);
                    if (it == 1 && lane == 0) KTRACE(256 + kb * 4 + 1);
                    tc_fence_after();
                    if (lane == 0) KTRACE(it * 48 + 32 + kb);
                    const uint32_t sa = smem_u32(smem + stage * kStageBytes), sb = sa + kTcABytes;
                    const uint64_t adesc = make_sw128_desc(sa), bdesc = make_sw128_desc(sb);
                    if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < kTcBlockK / 16; ++k) {
                        const uint64_t ad = adesc + 2 * k, bd = bdesc + 2 * k;   // +32 B per 16-element k-step
                        if (!kGru) umma_bf16_pair(d, ad, bd, make_idesc(256, Cfg::kLinN), (kb | k) != 0);
                        else if (kb < args.kb_per_part) umma_bf16_pair(d, ad, bd, make_idesc(256, kGruRows), (kb | k) != 0);
                        else umma_bf16_pair(d + kGruUnits, ad, bd, make_idesc(256, kGruRows), 1u);
                    }
                    umma_commit_pair(&empty_bar[stage], all_mask);   // partners multicast into my slots, so everyone must know
                    }
                    __syncwarp();
                    if (it == 1 && lane == 0) KTRACE(256 + kb * 4 + 3);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
                if (elect_one()) umma_commit_pair(&tmem_full[ab], pair_mask);
                __syncwarp();
                if (lane == 0) KTRACE(it * 48 + 3);
            }
        }
    } else if (warp == kTcStateWarp) {
        // ===================================================== GRU state traffic: one thread requests the fp32 h(t-1) box of every
        // pass (two passes of 32 units per tile, pass c uses staging buffers c) and stores the fp32 / bf16 h(t) boxes the
        // epilogue warps leave there.  Buffer c is refilled for the next tile as soon as its stores have been read out, a
        // whole pass plus a mainloop before it is needed, and no epilogue thread ever waits on a store.
        if (kGru && elect_one()) {
            const int row0 = (int) rank * kTcBlockM;
            if (cluster_id < num_tiles) {
                const int m0 = (cluster_id / ctiles_n) * kPM + qm, n0 = (cluster_id % ctiles_n) * kPN + qn;
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    mbar_expect_tx(&hp_full[c], Cfg::kEpiF32Bytes);
                    tma_load_2d_local(&map_hp, &hp_full[c], s_hp + c * Cfg::kEpiF32Bytes, n0 * kGruUnits + 32 * c, m0 * kTcPairM + row0);
                }
            }
            int it = 0;
            for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++it) {
                const int m = (tile / ctiles_n) * kPM + qm, n = (tile % ctiles_n) * kPN + qn;
                const int next = tile + num_clusters;
                const int m1 = (next / ctiles_n) * kPM + qm, n1 = (next % ctiles_n) * kPN + qn;
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    uint8_t *f32_box = s_hp + c * Cfg::kEpiF32Bytes, *b16_box = s_hb + c * (Cfg::kEpiBf16Bytes / 2);
                    mbar_wait(&staged[c], it & 1);
                    tma_store_2d(&map_hn, f32_box, n * kGruUnits + c * 32, m * kTcPairM + row0);
                    tma_store_2d(&map_hb, b16_box, n * kGruUnits + c * 32, m * kTcPairM + row0);
                    bulk_commit();
                    KTRACE(it * 48 + 9 + c * 2);
                    if (next < num_tiles) {
                        bulk_wait_read();                        // the stores have read buffers c
                        mbar_expect_tx(&hp_full[c], Cfg::kEpiF32Bytes);
                        tma_load_2d_local(&map_hp, &hp_full[c], f32_box, n1 * kGruUnits + 32 * c, m1 * kTcPairM + row0);
                    }
                }
            }
            bulk_wait_all();                                     // shared memory must outlive the last stores
        }
        __syncwarp();
    } else if (warp >= 5) {
        // ===================================================== epilogue: warps 5..20.  TMEM lane quarter = warp % 4 (a warp can
        // only touch its own 32 lanes); the 4 warps of a quarter split the tile's columns (GRU: 16 of the 64 units each,
        // linear: 32 of the 128 outputs each) and work through them 8 at a time.  16 warps, not 8: the gate math is a long
        // dependent chain (5 MUFU ops per unit) and needs the extra warps per scheduler to hide its latency.
        const int quarter = warp & 3, part = (warp - 5) >> 2, te = threadIdx.x - 160;
        const uint32_t lane_base = tmem_base + ((uint32_t) (quarter * 32) << 16);
        uint32_t empty_leader[2] = {map_to_cta(&tmem_empty[0], leader), map_to_cta(&tmem_empty[1], leader)};
        // hand both buffers to the MMA issuer for the first time (GRU: with the n_h columns cleared)
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            if (kGru) {
                tmem_zero8(lane_base + b * kTcAccCols + 192 + part * 16);
                tmem_zero8(lane_base + b * kTcAccCols + 192 + part * 16 + 8);
            }
        }
        if (kGru) tmem_st_wait();
        tc_fence_before();
        mbar_arrive_cluster(empty_leader[0]);
        mbar_arrive_cluster(empty_leader[1]);

        const int row_in_cta = quarter * 32 + lane;
        const int sw = row_in_cta & 7;
        constexpr float kL2e = 1.4426950408889634f;
        int it = 0;
        for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++it) {
            const int m = (tile / ctiles_n) * kPM + qm, n = (tile % ctiles_n) * kPN + qn;
            const int ab = it & 1, aphase = (it >> 1) & 1;
            float *sb = s_bias + ab * 256;
            if (kGru) {
                // biases of this tile -> smem, ordered like the TMEM columns [n_x | r | z | n_h] x 64; the r and z biases are
                // pre-multiplied by -log2(e) so that the sigmoid argument is one FFMA away from ex2
                if (te < 256) {
                    const int H = args.H, g = te >> 6, u = n * kGruUnits + (te & 63);
                    sb[te] = g == 0 ? __ldg(args.bias0 + 2 * H + u)
                           : g == 1 ? -kL2e * (__ldg(args.bias0 + u) + __ldg(args.bias1 + u))
                           : g == 2 ? -kL2e * (__ldg(args.bias0 + H + u) + __ldg(args.bias1 + H + u))
                                    : __ldg(args.bias1 + 2 * H + u);
                }
            } else {
                if (te < Cfg::kLinN) sb[te] = __ldg(args.bias0 + n * Cfg::kLinN + te);
            }
            asm volatile("bar.sync 1, %0;" ::"n"(kTcEpiThreads) : "memory");   // the epilogue threads only
            if (te == 0) KTRACE(it * 48 + 4);
            mbar_wait(&tmem_full[ab], aphase);
            tc_fence_after();
            if (te == 0) KTRACE(it * 48 + 5);
            const uint32_t t0 = lane_base + ab * kTcAccCols;
            if (kGru) {
                // two passes of 32 units (pass c: units c * 32 .. c * 32 + 31 of the tile, 8 of them mine), each with its own
                // staging buffers, filled and drained by the state warp
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    uint8_t *f32_box = s_hp + c * Cfg::kEpiF32Bytes, *b16_box = s_hb + c * (Cfg::kEpiBf16Bytes / 2);
                    uint8_t *f32_row = f32_box + row_in_cta * 128, *b16_row = b16_box + row_in_cta * 64;
                    const int cu = c * 32 + part * 8;            // first of my 8 units inside the tile
                    float anx[8], ar[8], az[8], anh[8], hp[8], hn[8];
                    tmem_ld8(t0 + 0 + cu, anx);
                    tmem_ld8(t0 + 64 + cu, ar);
                    tmem_ld8(t0 + 128 + cu, az);
                    tmem_ld8(t0 + 192 + cu, anh);
                    mbar_wait(&hp_full[c], it & 1);              // h(t-1) of this pass has landed
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const float4 t = *reinterpret_cast<const float4 *>(f32_row + (((part * 2 + q) ^ sw) << 4));
                        hp[4 * q] = t.x; hp[4 * q + 1] = t.y; hp[4 * q + 2] = t.z; hp[4 * q + 3] = t.w;
                    }
                    tmem_ld_wait();
                    if (te == 0) KTRACE(it * 48 + 8 + c * 2);
                    tmem_zero8(t0 + 192 + cu);                   // n_h columns must be zero when the buffer is reused
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        // r = 1/(1+er), z = 1/(1+ez) share one reciprocal: 5 MUFU ops per unit.  Arguments are clamped so that
                        // (1+er)(1+ez) cannot overflow: sigmoid(-30) is already 9e-14.
                        const float er = ex2_approx(fminf(fmaf(ar[i], -kL2e, sb[64 + cu + i]), 43.0f));
                        const float ez = ex2_approx(fminf(fmaf(az[i], -kL2e, sb[128 + cu + i]), 43.0f));
                        const float pr = 1.0f + er, pz = 1.0f + ez;
                        const float ip = rcp_approx(pr * pz);
                        const float rg = pz * ip, zg = pr * ip;
                        const float t = fmaf(rg, anh[i] + sb[192 + cu + i], anx[i] + sb[cu + i]);
                        const float ng = fmaf(-2.0f, rcp_approx(1.0f + ex2_approx(t * (2.0f * kL2e))), 1.0f);   // tanh(t)
                        hn[i] = fmaf(zg, hp[i] - ng, ng);        // (1 - z) n + z h
                    }
                    if (c == 1) {                                // last TMEM access of this tile: hand the buffer back
                        tmem_st_wait();
                        tc_fence_before();
                        if (te == 0) KTRACE(it * 48 + 6);
                        mbar_arrive_cluster(empty_leader[ab]);
                    }
#pragma unroll
                    for (int q = 0; q < 2; ++q)                  // h(t) replaces h(t-1) in place
                        *reinterpret_cast<float4 *>(f32_row + (((part * 2 + q) ^ sw) << 4)) =
                            make_float4(hn[4 * q], hn[4 * q + 1], hn[4 * q + 2], hn[4 * q + 3]);
                    *reinterpret_cast<uint4 *>(b16_row + part * 16) = pack_bf16x8(hn);
                    fence_proxy_async();                         // my smem writes -> visible to the TMA engine
                    mbar_arrive(&staged[c]);                     // state warp stores the pass once all 512 threads are here
                }
            } else {
                // my 32 of the tile's 128 outputs: bias + activation, staged in 128B-swizzled smem boxes, stored by TMA
                if (te == 0) bulk_wait_read();                   // previous tile's stores have read the staging buffer
                asm volatile("bar.sync 2, %0;" ::"n"(kTcEpiThreads) : "memory");
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int cc = part * 32 + c * 8;
                    float acc[8];
                    tmem_ld8(t0 + cc, acc);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float v = acc[i] + sb[cc + i];
                        acc[i] = MODE == kTcEnc ? fmaxf(v, 0.0f) : sigmoid_f(v);
                    }
                    if (MODE == kTcEnc) {        // bf16 boxes [128 rows][64 bf16]: box part / 2, chunk (part & 1) * 4 + c
                        uint8_t *rowp = s_hb + (part >> 1) * (kTcBlockM * 128) + row_in_cta * 128;
                        *reinterpret_cast<uint4 *>(rowp + ((((part & 1) * 4 + c) ^ sw) << 4)) = pack_bf16x8(acc);
                    } else {                     // fp32 boxes [128 rows][32 floats]: box part, chunks 2c, 2c + 1
                        uint8_t *rowp = s_hp + part * (kTcBlockM * 128) + row_in_cta * 128;
                        *reinterpret_cast<float4 *>(rowp + (((2 * c) ^ sw) << 4)) = make_float4(acc[0], acc[1], acc[2], acc[3]);
                        *reinterpret_cast<float4 *>(rowp + (((2 * c + 1) ^ sw) << 4)) = make_float4(acc[4], acc[5], acc[6], acc[7]);
                    }
                }
                tc_fence_before();
                if (te == 0) KTRACE(it * 48 + 6);
                mbar_arrive_cluster(empty_leader[ab]);
                fence_proxy_async();
                asm volatile("bar.sync 3, %0;" ::"n"(kTcEpiThreads) : "memory");
                if (te == 0) {
                    const int r0 = m * kTcPairM + (int) rank * kTcBlockM, c0 = n * Cfg::kLinN;
                    if (MODE == kTcEnc) {
                        tma_store_2d(&map_hn, s_hb, c0, r0);
                        tma_store_2d(&map_hn, s_hb + kTcBlockM * 128, c0 + 64, r0);
                    } else {
#pragma unroll
                        for (int b4 = 0; b4 < 4; ++b4) tma_store_2d(&map_hn, s_hp + b4 * (kTcBlockM * 128), c0 + 32 * b4, r0);
                    }
                    bulk_commit();
                    if (tile + num_clusters >= num_tiles) bulk_wait_all();
                    KTRACE(it * 48 + 7);
                }
            }
        }
    }
    if (threadIdx.x == 0) KTRACE(502);
    tc_fence_before();
    cluster_sync_all();      // the peer's smem / TMEM are read and written by the leader's MMAs: leave together
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, 2 * kTcAccCols);
    }
    if (threadIdx.x == 0) KTRACE(503);
#undef KTRACE
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
    const __nv_bfloat16 *enc_w = nullptr, *dec_w = nullptr, *wih[kMaxLayers] = {}, *whh[kMaxLayers] = {};
    const float *enc_b = nullptr, *dec_b = nullptr, *bih[kMaxLayers] = {}, *bhh[kMaxLayers] = {};
    __nv_bfloat16 *feat = nullptr, *e = nullptr, *hb[2] = {};
    float *h[2] = {}, *mask = nullptr;
};

struct TcPlan {
    TcModel m;
    int num_sms = 0;
    long long *trace = nullptr;   // 2 x 512 clock64 slots, filled by one kernel: KOALA_TC_TRACE=1 GRU layer 0, 2 decoder, 3 encoder
    int trace_kernel = 0;
    __nv_bfloat16 *wih_p[kMaxLayers] = {}, *whh_p[kMaxLayers] = {};
    CUtensorMap a_feat, a_hb_dec[2];                    // linear kernels: activation operands [Bp][K], box 64 x 128
    CUtensorMap a_e, a_hb[2][kMaxLayers];               // GRU kernel: box 64 x (128 / PN)
    CUtensorMap b_enc, b_dec, b_ih[kMaxLayers], b_hh[kMaxLayers];
    CUtensorMap e128, mask_f32;                         // linear epilogues: encoder output bf16 box 64 x 128, mask fp32 box 32 x 128
    CUtensorMap hf[2][kMaxLayers];                      // GRU epilogue: fp32 state [Bp][H], box 32 floats x 128 rows
    CUtensorMap hb128[2][kMaxLayers];                   // GRU epilogue: bf16 state store, box 32 x 128, no swizzle
    int max_clusters[3] = {0, 0, 0};                    // co-resident clusters per kernel (cudaOccupancyMaxActiveClusters)
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

static void tc_plan_destroy(TcPlan *p) {
    if (!p) return;
    for (int l = 0; l < kMaxLayers; l++) {
        if (p->wih_p[l]) cudaFree(p->wih_p[l]);
        if (p->whh_p[l]) cudaFree(p->whh_p[l]);
    }
    if (p->trace) cudaFree(p->trace);
    delete p;
}

static bool tc_plan_create(const TcModel &m, TcPlan **out, std::string *why) {
    if (m.H % 256 != 0 || m.Bp % (kTcPairM * TcCfg<kTcGru>::kPM) != 0) {
        *why = "hidden size must be a multiple of 256 and the padded stream count a multiple of the cluster tile for the tensor-core path";
        return false;
    }
    void *fnp = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &qres) != cudaSuccess || !fnp ||
        qres != cudaDriverEntryPointSuccess) {
        *why = "cuTensorMapEncodeTiled not available from the driver";
        return false;
    }
    EncodeTiledFn fn = (EncodeTiledFn) fnp;
    TcPlan *p = new TcPlan();
    p->m = m;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&p->num_sms, cudaDevAttrMultiProcessorCount, dev);
    const size_t H = m.H, Bp = m.Bp;
    bool ok = true;
    for (int l = 0; l < m.L && ok; l++) {
        ok = cudaMalloc((void **) &p->wih_p[l], 3 * H * H * 2) == cudaSuccess &&
             cudaMalloc((void **) &p->whh_p[l], 3 * H * H * 2) == cudaSuccess;
        if (!ok) break;
        pack_gru_weights_kernel<<<(unsigned) (3 * H), 64>>>(m.wih[l], p->wih_p[l], (int) H, 2, 0, 1);   // n | r | z
        pack_gru_weights_kernel<<<(unsigned) (3 * H), 64>>>(m.whh[l], p->whh_p[l], (int) H, 0, 1, 2);   // r | z | n
    }
    ok = ok && cudaDeviceSynchronize() == cudaSuccess;
    if (!ok) {
        *why = "packing the GRU weights failed";
        tc_plan_destroy(p);
        return false;
    }
    ok = ok && encode_2d(fn, &p->a_feat, m.feat, Bp, kBins, kTcBlockM);
    ok = ok && encode_2d(fn, &p->a_e, m.e, Bp, H, TcCfg<kTcGru>::kARowsPiece);
    ok = ok && encode_2d(fn, &p->e128, m.e, Bp, H, kTcBlockM);
    ok = ok && encode_2d(fn, &p->mask_f32, m.mask, Bp, kBins, kTcBlockM, true);
    for (int par = 0; par < 2; par++) {
        for (int l = 0; l < m.L; l++)
            ok = ok && encode_2d(fn, &p->a_hb[par][l], m.hb[par] + (size_t) l * Bp * H, Bp, H, TcCfg<kTcGru>::kARowsPiece);
        ok = ok && encode_2d(fn, &p->a_hb_dec[par], m.hb[par] + (size_t) (m.L - 1) * Bp * H, Bp, H, kTcBlockM);
        for (int l = 0; l < m.L; l++) {
            ok = ok && encode_2d(fn, &p->hf[par][l], m.h[par] + (size_t) l * Bp * H, Bp, H, kTcBlockM, true);
            ok = ok && encode_2d(fn, &p->hb128[par][l], m.hb[par] + (size_t) l * Bp * H, Bp, H, kTcBlockM, false, true);
        }
    }
    ok = ok && encode_2d(fn, &p->b_enc, m.enc_w, H, kBins, TcCfg<kTcEnc>::kBRowsHalf);
    ok = ok && encode_2d(fn, &p->b_dec, m.dec_w, kBins, H, TcCfg<kTcDec>::kBRowsHalf);
    for (int l = 0; l < m.L; l++) {
        ok = ok && encode_2d(fn, &p->b_ih[l], p->wih_p[l], 3 * H, H, TcCfg<kTcGru>::kBRowsPiece);
        ok = ok && encode_2d(fn, &p->b_hh[l], p->whh_p[l], 3 * H, H, TcCfg<kTcGru>::kBRowsPiece);
    }
    if (!ok) {
        *why = "cuTensorMapEncodeTiled failed";
        tc_plan_destroy(p);
        return false;
    }
    ok = cudaFuncSetAttribute(tc_masknet_kernel<kTcEnc>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<kTcEnc>::kSmemBytes) == cudaSuccess &&
         cudaFuncSetAttribute(tc_masknet_kernel<kTcGru>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<kTcGru>::kSmemBytes) == cudaSuccess &&
         cudaFuncSetAttribute(tc_masknet_kernel<kTcDec>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<kTcDec>::kSmemBytes) == cudaSuccess;
    if (!ok) {
        *why = "cudaFuncSetAttribute(MaxDynamicSharedMemorySize) failed";
        tc_plan_destroy(p);
        return false;
    }
    auto occupancy = [&](auto kern, int cluster, int smem_bytes) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned) (cluster * p->num_sms));
        cfg.blockDim = dim3(kTcThreads);
        cfg.dynamicSmemBytes = (size_t) smem_bytes;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned) cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n < 1) { cudaGetLastError(); n = p->num_sms / cluster; }
        return n;
    };
    p->max_clusters[kTcEnc] = occupancy(tc_masknet_kernel<kTcEnc>, TcCfg<kTcEnc>::kCluster, TcCfg<kTcEnc>::kSmemBytes);
    p->max_clusters[kTcGru] = occupancy(tc_masknet_kernel<kTcGru>, TcCfg<kTcGru>::kCluster, TcCfg<kTcGru>::kSmemBytes);
    p->max_clusters[kTcDec] = occupancy(tc_masknet_kernel<kTcDec>, TcCfg<kTcDec>::kCluster, TcCfg<kTcDec>::kSmemBytes);
    const char *tr = getenv("KOALA_TC_TRACE");
    if (tr && *tr >= '1' && *tr <= '3') {
        p->trace_kernel = *tr - '0';
        if (cudaMalloc((void **) &p->trace, 1024 * sizeof(long long)) != cudaSuccess) p->trace = nullptr;
        else cudaMemset(p->trace, 0, 1024 * sizeof(long long));
    }
    *out = p;
    return true;
}

// one mask-estimator step: hb[cur]/h[cur] hold state t-1, results go to hb[cur^1]/h[cur^1]; returns kernels launched
static int tc_masknet_step(TcPlan *p, int cur, cudaStream_t st, KernelProfiler *prof) {
    const TcModel &m = p->m;
    const int nxt = cur ^ 1, H = m.H, mt = m.Bp / kTcPairM;
    const size_t LBH = (size_t) m.Bp * H;
    // persistent grid: one cluster per cluster-tile slot, capped by what the device can keep resident
    auto grid = [&](int mode, int cluster, int ctiles) {
        const int c = ctiles < p->max_clusters[mode] ? ctiles : p->max_clusters[mode];
        return cluster * (c < 1 ? 1 : c);
    };
    {
        using C = TcCfg<kTcEnc>;
        TcArgs a{};
        a.num_m_tiles = mt; a.num_n_tiles = H / C::kLinN; a.kb_per_part = kBins / kTcBlockK; a.H = H;
        a.bias0 = m.enc_b; a.out_bf16 = m.e;
        a.trace = p->trace_kernel == 3 ? p->trace : nullptr;
        if (prof) prof->begin(kKernEnc, st);
        launch_pdl(true, tc_masknet_kernel<kTcEnc>, dim3((unsigned) grid(kTcEnc, C::kCluster, mt * a.num_n_tiles)), dim3(kTcThreads), (size_t) C::kSmemBytes, st, p->a_feat, p->a_feat, p->b_enc, p->b_enc, p->a_feat, p->e128, p->a_feat, a);
        if (prof) prof->end(st);
    }
    for (int l = 0; l < m.L; l++) {
        using C = TcCfg<kTcGru>;
        TcArgs a{};
        a.num_m_tiles = mt; a.num_n_tiles = H / kGruUnits; a.kb_per_part = H / kTcBlockK; a.H = H;
        a.bias0 = m.bih[l]; a.bias1 = m.bhh[l];
        a.h_prev = m.h[cur] + l * LBH; a.h_next = m.h[nxt] + l * LBH; a.out_bf16 = m.hb[nxt] + l * LBH;
        a.a0 = l == 0 ? m.e : m.hb[nxt] + (l - 1) * LBH; a.a1 = m.hb[cur] + l * LBH;
        a.trace = (l == 0 && p->trace_kernel == 1) ? p->trace : nullptr;
        const CUtensorMap &ax = l == 0 ? p->a_e : p->a_hb[nxt][l - 1];
        if (prof) prof->begin(kKernGru, st);
        launch_pdl(true, tc_masknet_kernel<kTcGru>, dim3((unsigned) grid(kTcGru, C::kCluster, (mt / C::kPM) * (a.num_n_tiles / C::kPN))), dim3(kTcThreads), (size_t) C::kSmemBytes, st, ax, p->a_hb[cur][l], p->b_ih[l], p->b_hh[l], p->hf[cur][l], p->hf[nxt][l], p->hb128[nxt][l], a);
        if (prof) prof->end(st);
    }
    {
        using C = TcCfg<kTcDec>;
        TcArgs a{};
        a.num_m_tiles = mt; a.num_n_tiles = kBins / C::kLinN; a.kb_per_part = H / kTcBlockK; a.H = H;
        a.bias0 = m.dec_b; a.out_f32 = m.mask;
        a.trace = p->trace_kernel == 2 ? p->trace : nullptr;
        if (prof) prof->begin(kKernDec, st);
        launch_pdl(true, tc_masknet_kernel<kTcDec>, dim3((unsigned) grid(kTcDec, C::kCluster, mt * a.num_n_tiles)), dim3(kTcThreads), (size_t) C::kSmemBytes, st, p->a_hb_dec[nxt], p->a_hb_dec[nxt], p->b_dec, p->b_dec, p->a_feat, p->mask_f32, p->a_feat, a);
        if (prof) prof->end(st);
    }
    return 2 + m.L;
}

}  // namespace koala
