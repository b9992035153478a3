// koala_b200 -- mask estimator, tensor-core path: ONE persistent kernel per step for encoder -> GRU layers -> decoder.
//
// Middle stage of `pv_koala_process` (/root/reference/include/pv_koala.h:65-80), batched over the stream dimension
// (BASELINE.json configs[2..4]).  The building blocks are those of tcgen05_common.cuh (bf16 operands staged by TMA into
// 128B-swizzled shared memory, tcgen05.mma.cta_group::2 accumulating fp32 in TMEM, gate math fused into the epilogue);
// this file is the schedule.  Launching the four GEMM stages as separate kernels cost, per launch, ~10 k cycles before the
// first MMA (dependent-launch wait, cold operand pipeline) and ~7 k cycles after the last one (final epilogue): a third of
// a GRU launch and most of an encoder / decoder launch (profiles/r01_step_summary.md).  Here every stage is a SEGMENT of
// one global list of pair tiles
//     [encoder tiles | GRU layer 0 tiles | ... | GRU layer L-1 tiles | decoder tiles],  each segment ordered m-major,
// walked round-robin by persistent CTA pairs (2-CTA clusters on all 148 SMs).  A tile of segment s + 1 needs the rows of its
// m tile from ALL n tiles of segment s: every CTA bumps a per-(segment, m tile) counter in global memory once its output
// stores have completed, and the activation producers of a dependent tile test that counter before the first load that
// needs it.  Tiles are taken in list order, so a tile only ever waits for tiles that started earlier: no deadlock as long
// as the grid is co-resident (it is sized by cudaOccupancyMaxActiveClusters).  Counters are never reset: launch number
// `epoch` waits for epoch * (increments per step).  GRU tiles run their h(t-1) part first -- it depends on nothing inside
// the launch -- so the wait only ever holds the second half of a k loop.
//
// Warp roles (23 warps, 736 threads, 80 registers): 0,1 activation (A) producers for even / odd k-blocks, 2 TMEM allocator
// + MMA issuer (pair leader only), 3,4 weight (B) producers (they do not wait for the previous kernel: weights are
// constants), 5..20 epilogue (4 per TMEM lane quarter), 21 GRU state warp, 22 linear-tile store warp.
//   GRU epilogue: two PASSES of 32 units through two staging buffers; the state warp TMA-stores a finished pass (fp32 h(t)
//   in place of h(t-1), plus its bf16 copy) and requests the fp32 h(t-1) box of the pass two ahead into the drained buffer.
//   Linear epilogue: the accumulator is handed back after four TMEM loads; outputs are staged in a separate 32 KB region
//   (encoder: the CTA's whole [128][128] bf16 tile; decoder: [128][128] fp32, half of it in the by then idle GRU boxes) and
//   TMA-stored by warp 22.
// Shared memory: 5 operand stages x 28 KB + 48 KB GRU staging + 32 KB linear staging + barriers / biases = 223 KB.
//
// What bounds it (8192 streams, B200; profiles/r01_step_summary.md): operand delivery.  Skipping every tcgen05.mma leaves
// the kernel time unchanged (66.5 vs 65.9 us); the SMs ingest ~28 KB per k-block per ~500 cycles each (8.3 KB/clk over the
// chip), and a GRU tile with N = 192 needs 73 B/clk/SM at full MMA rate.  TMEM (4 accumulator columns per unit, two
// buffers) caps the tile width, hence the operand bytes per flop.
#pragma once

#include <string.h>

#include <algorithm>
#include <vector>

#include "tcgen05_common.cuh"

namespace koala {

#ifndef KOALA_FU_STAGES
#define KOALA_FU_STAGES 5
#endif
constexpr int kFuStages = KOALA_FU_STAGES;
constexpr int kFuStageBytes = kTcABytes + (kGruRows / 2) * 128;   // 28 KB: A [128][64] + B up to [96][64] bf16 (linear tiles: 64 rows)
#ifndef KOALA_FU_PN
#define KOALA_FU_PN 1
#endif
constexpr int kFuPN = KOALA_FU_PN;                                // CTA pairs per cluster; 2 = neighbouring n tiles of one m tile with activation multicast
                                                                  // (33 clusters = 132 SMs, measured 2 % slower than 74 plain pairs)
constexpr int kFuCluster = 2 * kFuPN;
constexpr int kFuARows = kTcBlockM / kFuPN;                       // rows of A each CTA fetches and multicasts
constexpr int kFuLinN = 128;                                      // outputs per linear pair tile
constexpr int kFuBoxF32 = kTcBlockM * 32 * 4;                     // staging box [128 rows][32 fp32], 128B-swizzled
constexpr int kFuBoxB16 = kTcBlockM * 32 * 2;                     // staging box [128 rows][32 bf16], plain
constexpr int kFuLinBytes = 2 * kFuBoxF32;                        // linear tiles: [128][128] bf16 (encoder) or half of [128][128] fp32 (decoder; other half: the GRU boxes)
constexpr int kFuLinWarp = kTcStateWarp + 1;                      // stores the linear tiles' staged outputs
constexpr int kFuThreads = 32 * (kFuLinWarp + 1);                 // 2 + 1 + 2 + 16 + 1 + 1 warps
constexpr int kFuSmemBytes = kFuStages * kFuStageBytes + 2 * (kFuBoxF32 + kFuBoxB16) + kFuLinBytes + kTcTailBytes;
constexpr int kFuMaxSegs = kMaxLayers + 2;
constexpr int kFuArrivals = kTcEpiWarps;                          // epilogue barriers: one arrival per warp (after __syncwarp), not per thread
enum FuMap : int { kMapA0 = 0, kMapA1, kMapB0, kMapB1, kMapHp, kMapHn, kMapHb, kFuMapsPerSeg };

struct FuSeg {
    int mode;                 // kTcEnc | kTcGru | kTcDec
    int n_ctiles;             // cluster tiles along n (= pair tiles / kFuPN)
    int kb_per_part, parts;   // k-blocks of 64 per operand part; GRU has two parts (run as h(t-1) first, then x)
    int tile_begin;           // index of the segment's first tile in the global list
    int dep, done;            // counter rows this segment waits on / bumps (-1: none)
    unsigned dep_per_step;    // increments per m tile and step of the row it waits on (CTAs that write those rows)
    const float *bias0, *bias1;
    const __nv_bfloat16 *a1;  // GRU: bf16 h(t-1) operand [Bp][H], for the L2 prefetch of the next tile
};
struct FuArgs {
    int nseg, num_m_tiles, total_tiles, H;
    unsigned epoch;
    unsigned *counters;       // [nseg][num_m_tiles]
    const CUtensorMap *maps;  // [nseg][kFuMapsPerSeg], global memory
    long long *trace;
    FuSeg seg[kFuMaxSegs];
};

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned ld_relaxed_gpu(const unsigned *p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_gpu(unsigned *p, unsigned v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void bulk_wait_read_but1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// position of one cluster in the global tile list
struct FuTile {
    int s, m, n0;             // segment, m tile, first n tile of the cluster tile
};
__device__ __forceinline__ FuTile fu_decode(const FuArgs &a, int g) {
    int s = 0;
    while (s + 1 < a.nseg && g >= a.seg[s + 1].tile_begin) ++s;
    const int local = g - a.seg[s].tile_begin, nc = a.seg[s].n_ctiles;
    return FuTile{s, local / nc, (local % nc) * kFuPN};
}

__global__ void __cluster_dims__(kFuCluster, 1, 1) __launch_bounds__(kFuThreads, 1) tc_fused_kernel(const __grid_constant__ FuArgs args) {
    constexpr int kStages = kFuStages, kStageBytes = kFuStageBytes;

    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t) 1023);
    uint8_t *s_f32 = smem + kStages * kStageBytes;       // 2 fp32 staging boxes
    uint8_t *s_b16 = s_f32 + 2 * kFuBoxF32;              // 2 bf16 staging boxes
    uint8_t *s_lin = s_b16 + 2 * kFuBoxB16;              // linear tiles' staging: two 16 KB swizzled boxes
    uint8_t *tail = s_lin + kFuLinBytes;
    uint64_t *bars = reinterpret_cast<uint64_t *>(tail);
    uint64_t *full_bar = bars, *empty_bar = bars + kStages;
    uint64_t *tmem_full = bars + 2 * kStages, *tmem_empty = bars + 2 * kStages + 2;
    uint64_t *box_ready = bars + 2 * kStages + 4;        // [2]: staging buffer is free (linear) / holds h(t-1) (GRU)
    uint64_t *staged = bars + 2 * kStages + 6;           // [2]: the epilogue has staged a pass in the buffer
    uint64_t *lin_free = bars + 2 * kStages + 8, *lin_staged = bars + 2 * kStages + 9;
    uint64_t *gru_done = bars + 2 * kStages + 10;         // the state warp has drained: the GRU staging boxes are free for good
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * kStages + 11);
    float *s_bias = reinterpret_cast<float *>(tail + 256);   // [2 accumulator buffers][256]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t crank = cluster_ctarank();      // rank in the cluster = 2 * pair + position in pair
    const uint32_t rank = crank & 1;               // 0 = pair leader (issues the MMAs), 1 = peer
    const uint32_t leader = crank & ~1u;           // cluster rank of my pair's leader
    const int qn = (int) (crank >> 1);             // my pair's n tile inside the cluster tile
    const uint16_t pair_mask = (uint16_t) (3u << leader), all_mask = (uint16_t) ((1u << kFuCluster) - 1);
    [[maybe_unused]] long long *trace = (args.trace != nullptr && blockIdx.x < 2) ? args.trace + blockIdx.x * 1024 : nullptr;
#ifdef KOALA_FU_TRACE      // clock64 timeline of cluster 0 (tools/gpu_trace.py builds this variant); compiled out of the product
#define KTRACE(slot) do { if (trace && (slot) < 1024) trace[(slot)] = clock64(); } while (0)
#else
#define KTRACE(slot) do { } while (0)
#endif
    if (threadIdx.x == 0) KTRACE(1020);
    pdl_launch_dependents();
    const int cluster_id = blockIdx.x / kFuCluster, num_clusters = gridDim.x / kFuCluster;
    const int total = args.total_tiles;

    if (warp == 0) {
        for (int i = lane; i < args.nseg * kFuMapsPerSeg; i += 32) prefetch_tmap(args.maps + i);
        if (lane == 0) {
            for (int s = 0; s < kStages; ++s) {
                mbar_init(&full_bar[s], 1);          // leader's copy is the one in use: 1 arrive.expect_tx + 2 CTAs' TMA bytes
                mbar_init(&empty_bar[s], kFuPN);     // one multicast tcgen05.commit from every pair leader of the cluster
            }
            for (int b = 0; b < 2; ++b) {
                mbar_init(&tmem_full[b], 1);         // one multicast tcgen05.commit
                mbar_init(&tmem_empty[b], 2 * kFuArrivals);     // leader's copy: the epilogue warps (threads) of both CTAs
                mbar_init(&box_ready[b], 1);
                mbar_init(&staged[b], kFuArrivals);
            }
            mbar_init(lin_free, 1);
            mbar_init(gru_done, 1);
            mbar_init(lin_staged, kFuArrivals);
            fence_mbar_init();
        }
    }
    if (warp == 2) tmem_alloc_pair(tmem_slot, 2 * kTcAccCols);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) KTRACE(1021);

    if (warp == 0 || warp == 1 || warp == 3 || warp == 4) {
        // ===================================================== TMA producers (both CTAs of every pair).  One warp can issue a
        // tensor load only every ~210-380 cycles (tools/micro/tma_rate.cu), so the loads are spread over four warps: warps
        // 0/1 fetch the activation (A) tiles of the even/odd k-blocks, warps 3/4 the weight (B) tiles.  Everything stays
        // warp-uniform (elect_one) so descriptors live in uniform registers.  Weights do not depend on the previous kernel:
        // the B producers fill the pipeline while that kernel is still running, only the A producers wait for it.
        const bool is_a = warp < 2;
        const int par = is_a ? warp : warp - 3;             // my k-block parity
        if (is_a) pdl_wait();
        int stage = par, phase = 0;                         // kStages is even: a warp stays on the stages of its parity
        uint16_t mask_a = 0;                                 // same position in every pair of the cluster
#pragma unroll
        for (int j = 0; j < kFuPN; ++j) mask_a |= (uint16_t) (1u << (2 * j + rank));
        auto arow_of = [&](int m) { return m * kTcPairM + (int) rank * kTcBlockM + qn * kFuARows; };   // first of the 64 rows I fetch
        int pit = 0;
        unsigned seen = 0, seen_this = 0;   // the next tile's dependency counter, sampled one tile early
        bool seen_valid = false, seen_valid_this = false;
        for (int g = cluster_id; g < total; g += num_clusters, ++pit) {
            const FuTile t = fu_decode(args, g);
            const FuSeg &sg = args.seg[t.s];
            const CUtensorMap *maps = args.maps + t.s * kFuMapsPerSeg;
            const bool gru = sg.mode == kTcGru;
            const int num_kb = sg.kb_per_part * sg.parts, n = t.n0 + qn;
            const int brows = gru ? kGruRows / 2 : kFuLinN / 2;         // weight rows each CTA of the pair holds
            const uint32_t pair_tx = 2u * (uint32_t) (kTcABytes + brows * 128);
            if (warp == 0 && lane == 0) KTRACE(pit * 48 + 0);
            bool dep_pending = is_a && sg.dep >= 0;
            seen_valid_this = seen_valid; seen_this = seen;
            seen_valid = false;
            for (int kb = par; kb < num_kb; kb += 2) {
                // GRU tiles take their h(t-1) part FIRST: it does not depend on the previous segment, so the dependency wait
                // below only holds the second half of the k loop and is usually over by the time it is reached
                const bool second = gru && kb < sg.kb_per_part;      // second operand pair (A1 / B1) = the h part
                if (dep_pending && !second) {
                    // the rows of my m tile are written by every n tile of the previous segment: wait until all those CTAs have
                    // signalled (their stores completed before the release).  The counter was already sampled during the
                    // previous tile, so in the steady state nothing is waited for here.  The fast path is a relaxed
                    // L1-bypassing load: the TMA unit reads through L2, where the signalled rows already are, and an acquire
                    // or a proxy fence in this thread waits for its own outstanding TMA loads (~2 k cycles per tile, measured).
                    const unsigned target = args.epoch * sg.dep_per_step;
                    if (!seen_valid_this || (int) (seen_this - target) < 0) {
                        const unsigned *ctr = args.counters + (size_t) sg.dep * args.num_m_tiles + t.m;
                        while ((int) (ld_acquire_gpu(ctr) - target) < 0) __nanosleep(40);
                        fence_proxy_async_all();
                    }
                    dep_pending = false;
                    if (warp == 0 && lane == 0) KTRACE(pit * 48 + 12);
                }
                mbar_wait(&empty_bar[stage], phase ^ 1);     // slot free in every CTA of the cluster
                if (is_a && lane == 0) KTRACE(pit * 48 + 16 + kb);
                const bool elected = elect_one();
                if (elected && is_a && rank == 0) mbar_expect_tx(&full_bar[stage], pair_tx);
                const uint32_t full_leader = map_to_cta(&full_bar[stage], leader);
                uint8_t *sa = smem + stage * kStageBytes, *sb = sa + kTcABytes;
                const int kc = (kb >= sg.kb_per_part ? kb - sg.kb_per_part : kb) * kTcBlockK;
                if (!elected) {
                } else if (is_a) {
                    if (kFuPN > 1) tma_load_2d_pair_mc(maps + (second ? kMapA1 : kMapA0), full_leader, sa + qn * kFuARows * 128, kc, arow_of(t.m), mask_a);
                    else tma_load_2d_pair(maps + (second ? kMapA1 : kMapA0), full_leader, sa, kc, arow_of(t.m));
                } else {
                    tma_load_2d_pair(maps + (second ? kMapB1 : kMapB0), full_leader, sb, kc, n * 2 * brows + (int) rank * brows);
                }
                __syncwarp();
                if (is_a && kb == par) {
                    const int g1 = g + num_clusters;
                    if (g1 < total) {
                        const FuTile t1 = fu_decode(args, g1);
                        if (args.seg[t1.s].dep >= 0) {
                            seen = ld_relaxed_gpu(args.counters + (size_t) args.seg[t1.s].dep * args.num_m_tiles + t1.m);
                            seen_valid = true;
                        }
                    }
                }
                if (warp == 1 && kb == par) {
                    // h(t-1) operands come from HBM (written a whole step ago): my 64 rows of the NEXT tile's (and, for the
                    // cluster's first tile, this tile's) operand are one contiguous range; request it into L2 now, a whole
                    // mainloop ahead -- after this tile's first load so that the request does not queue in front of it
                    const uint32_t bytes = (uint32_t) (kFuARows * args.H * 2);
                    if (elect_one()) {
                        if (pit == 0 && gru) prefetch_l2(sg.a1 + (size_t) arow_of(t.m) * args.H, bytes);
                        const int g1 = g + num_clusters;
                        if (g1 < total) {
                            const FuTile t1 = fu_decode(args, g1);
                            if (args.seg[t1.s].mode == kTcGru) prefetch_l2(args.seg[t1.s].a1 + (size_t) arow_of(t1.m) * args.H, bytes);
                        }
                    }
                    __syncwarp();
                }
                stage += 2;
                if (stage >= kStages) { stage -= kStages; phase ^= 1; }
            }
            if (warp == 0 && lane == 0) KTRACE(pit * 48 + 1);
        }
    } else if (warp == 2) {
        // ===================================================== MMA issuer (one thread of the leader CTA drives both SMs)
        if (rank == 0) {
            int it = 0;
            unsigned kb_total = 0;           // k-blocks issued so far: every lane derives the pipeline position from it
            const uint64_t adesc0 = make_sw128_desc(smem_u32(smem)), bdesc0 = make_sw128_desc(smem_u32(smem) + kTcABytes);
            for (int g = cluster_id; g < total; g += num_clusters, ++it) {
                const FuTile t = fu_decode(args, g);
                const FuSeg &sg = args.seg[t.s];
                const bool gru = sg.mode == kTcGru;
                const int num_kb = sg.kb_per_part * sg.parts, kbp = sg.kb_per_part;
                const uint32_t idesc = gru ? make_idesc(256, kGruRows) : make_idesc(256, kFuLinN);
                const int ab = it & 1, aphase = (it >> 1) & 1;
                mbar_wait(&tmem_empty[ab], aphase);      // both CTAs' epilogues have released (and cleared) this buffer
                tc_fence_after();
                if (lane == 0) KTRACE(it * 48 + 2);
                const uint32_t d = tmem_base + ab * kTcAccCols;
                // The whole k loop runs in ONE elected lane: every instruction between two k-blocks (barrier test, descriptor
                // arithmetic, moves into uniform registers) is serial latency of this thread and shows up one-for-one in the
                // k-block time, so nothing is re-elected, re-synchronised or re-selected per k-block, and the loop is split by
                // operand part to keep its body branch-free.  GRU: the h part comes first and initialises columns [r | z | n_h];
                // the x part accumulates into [n_x | r | z], whose n_x columns the epilogue left zeroed.
                int stage = (int) (kb_total % kStages), phase = (int) ((kb_total / kStages) & 1);
                kb_total += (unsigned) num_kb;
                if (elect_one()) {
                    for (int part = 0; part < sg.parts; ++part) {
                        const uint32_t dk = (gru && part == 0) ? d + kGruUnits : d;
                        const uint32_t first = (gru && part == 1) ? 1u : 0u;    // linear tiles and the GRU h part start from zero
                        for (int kb = 0; kb < kbp; ++kb) {
                            mbar_wait(&full_bar[stage], phase);
                            tc_fence_after();
                            KTRACE(it * 48 + 32 + part * kbp + kb);
                            const uint64_t so = (uint64_t) ((stage * kStageBytes) >> 4);
#pragma unroll
                            for (int k = 0; k < kTcBlockK / 16; ++k)    // +32 B per 16-element k-step
                                umma_bf16_pair(dk, adesc0 + so + 2 * k, bdesc0 + so + 2 * k, idesc, (kb | k) != 0 ? 1u : first);
                            umma_commit_pair(&empty_bar[stage], all_mask);   // partners multicast into my slots, so everyone must know
                            if (++stage == kStages) { stage = 0; phase ^= 1; }
                        }
                    }
                    umma_commit_pair(&tmem_full[ab], pair_mask);
                }
                __syncwarp();
                if (lane == 0) KTRACE(it * 48 + 3);
            }
        }
    } else if (warp == kTcStateWarp) {
        // ===================================================== state warp: one thread moves the staging boxes of the GRU tiles
        // (two passes of 32 units each).  `cur` walks the passes in order (wait until staged, store fp32 + bf16 h(t), commit);
        // `ahead` runs two passes in front and requests the fp32 h(t-1) box of that pass into the buffer `cur` just drained,
        // a whole mainloop before the epilogue needs it.  After a tile's second pass it waits for the stores to COMPLETE and
        // bumps the segment's counter for the m tile, which releases the dependent tiles of the next segment.
        pdl_wait();
        if (elect_one()) {
            const int row0 = (int) rank * kTcBlockM;
            struct PassIter {
                int g, c;
                FuTile t;
            };
            auto start = [&](PassIter &p, int g) {           // first pass of the cluster's next GRU tile at or after g
                for (; g < total; g += num_clusters) {
                    p.t = fu_decode(args, g);
                    if (args.seg[p.t.s].mode == kTcGru) break;
                }
                p.g = g;
                p.c = 0;
            };
            auto advance = [&](PassIter &p) {
                if (++p.c == 2) start(p, p.g + num_clusters);
            };
            auto arm = [&](const PassIter &p, int buf) {     // request the fp32 h(t-1) box of pass p into staging buffer `buf`
                mbar_expect_tx(&box_ready[buf], kFuBoxF32);
                tma_load_2d_local(args.maps + p.t.s * kFuMapsPerSeg + kMapHp, &box_ready[buf], s_f32 + buf * kFuBoxF32,
                                  (p.t.n0 + qn) * kGruUnits + 32 * p.c, p.t.m * kTcPairM + row0);
            };
            PassIter cur, ahead;
            start(cur, cluster_id);
            ahead = cur;
            for (int b = 0; b < 2 && ahead.g < total; ++b) {
                arm(ahead, b);
                advance(ahead);
            }
            unsigned pc = 0;
            int it = 0;
            while (cur.g < total) {
                const int buf = (int) (pc & 1);
                const FuSeg &sg = args.seg[cur.t.s];
                const CUtensorMap *maps = args.maps + cur.t.s * kFuMapsPerSeg;
                const int row = cur.t.m * kTcPairM + row0, col = (cur.t.n0 + qn) * kGruUnits + 32 * cur.c;
                mbar_wait(&staged[buf], (pc >> 1) & 1);
                tma_store_2d(maps + kMapHn, s_f32 + buf * kFuBoxF32, col, row);
                tma_store_2d(maps + kMapHb, s_b16 + buf * kFuBoxB16, col, row);
                bulk_commit();
                KTRACE(it * 48 + 9 + (cur.c & 1) * 2);
                if (ahead.g < total) {
                    bulk_wait_read();                        // the stores have read buffer `buf`
                    arm(ahead, buf);
                    advance(ahead);
                }
                if (cur.c == 1) {
                    if (sg.done >= 0) {
                        bulk_wait_all();                     // this tile's rows are in global memory
                        fence_proxy_async_all();
                        red_release_gpu(args.counters + (size_t) sg.done * args.num_m_tiles + cur.t.m, 1u);
                    }
                    ++it;
                }
                advance(cur);
                ++pc;
            }
            bulk_wait_read();                                // shared memory must outlive the last stores' reads (kernel completion covers the writes)
            mbar_arrive(gru_done);                           // ... and from here on the decoder tiles may stage in the GRU boxes
        }
        __syncwarp();
    } else if (warp == kFuLinWarp) {
        // ===================================================== linear-tile store warp: encoder / decoder outputs leave through
        // their own 32 KB staging region, so they never compete with the GRU passes for staging buffers: the encoder's whole
        // [128][128] bf16 tile of this CTA fits; the decoder's [128][128] fp32 tile takes the two fp32 GRU boxes as well, which
        // are idle by then (a pair's decoder tiles come after all its GRU tiles; `gru_done`).  Per tile: wait until the epilogue
        // warps have staged it, TMA-store its boxes, free the region once they have been read; after an encoder tile, wait
        // for the stores to complete and release the dependent GRU tiles.
        pdl_wait();
        if (elect_one()) {
            const int row0 = (int) rank * kTcBlockM;
            unsigned lc = 0;
            for (int g = cluster_id; g < total; g += num_clusters) {
                const FuTile t = fu_decode(args, g);
                const FuSeg &sg = args.seg[t.s];
                if (sg.mode == kTcGru) continue;
                const CUtensorMap *map = args.maps + t.s * kFuMapsPerSeg + kMapHn;
                const int row = t.m * kTcPairM + row0, col = (t.n0 + qn) * kFuLinN;
                mbar_wait(lin_staged, lc & 1);
                if (sg.mode == kTcEnc) {           // two boxes of 64 bf16 columns
                    tma_store_2d(map, s_lin, col, row);
                    tma_store_2d(map, s_lin + kFuBoxF32, col + 64, row);
                } else {                           // four boxes of 32 fp32 columns: two in the linear region, two in the GRU boxes
                    tma_store_2d(map, s_lin, col, row);
                    tma_store_2d(map, s_lin + kFuBoxF32, col + 32, row);
                    tma_store_2d(map, s_f32, col + 64, row);
                    tma_store_2d(map, s_f32 + kFuBoxF32, col + 96, row);
                }
                bulk_commit();
                bulk_wait_read();
                mbar_arrive(lin_free);
                ++lc;
                if (sg.done >= 0) {
                    bulk_wait_all();                         // this tile's rows are in global memory
                    fence_proxy_async_all();
                    red_release_gpu(args.counters + (size_t) sg.done * args.num_m_tiles + t.m, 1u);
                }
            }
            bulk_wait_read();                                // every store has read its staging box; the grid's completion orders the writes
        }
        __syncwarp();
    } else if (warp >= 5 && warp < kTcStateWarp) {
        // ===================================================== epilogue: warps 5..20.  TMEM lane quarter = warp % 4 (a warp can
        // only touch its own 32 lanes); the 4 warps of a quarter take 8 of a pass's 32 columns each.  16 warps, not 8: the
        // gate math is a long dependent chain (5 MUFU ops per unit) and needs the extra warps per scheduler to hide it.
        const int quarter = warp & 3, part = (warp - 5) >> 2, te = threadIdx.x - 160;
        const uint32_t lane_base = tmem_base + ((uint32_t) (quarter * 32) << 16);
        uint32_t empty_leader[2] = {map_to_cta(&tmem_empty[0], leader), map_to_cta(&tmem_empty[1], leader)};
        // hand both buffers to the MMA issuer for the first time, with the GRU n_x columns cleared (linear tiles start from zero anyway)
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            tmem_zero8(lane_base + b * kTcAccCols + part * 16);
            tmem_zero8(lane_base + b * kTcAccCols + part * 16 + 8);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
            mbar_arrive_cluster(empty_leader[0]);
            mbar_arrive_cluster(empty_leader[1]);
        }

        const int row_in_cta = quarter * 32 + lane;
        const int sw = row_in_cta & 7;
        constexpr float kL2e = 1.4426950408889634f;
        unsigned pc = 0, lc = 0;       // GRU passes / linear rounds so far: staging buffer and barrier phase
        int it = 0;
        for (int g = cluster_id; g < total; g += num_clusters, ++it) {
            const FuTile t = fu_decode(args, g);
            const FuSeg &sg = args.seg[t.s];
            const int mode = sg.mode, n = t.n0 + qn;
            const int ab = it & 1, aphase = (it >> 1) & 1;
            float *sb = s_bias + ab * 256;
            if (mode == kTcGru) {
                // biases of this tile -> smem, ordered like the TMEM columns [n_x | r | z | n_h] x 64; the r and z biases are
                // pre-multiplied by -log2(e) so that the sigmoid argument is one FFMA away from ex2
                if (te < 256) {
                    const int H = args.H, gate = te >> 6, u = n * kGruUnits + (te & 63);
                    sb[te] = gate == 0 ? __ldg(sg.bias0 + 2 * H + u)
                           : gate == 1 ? -kL2e * (__ldg(sg.bias0 + u) + __ldg(sg.bias1 + u))
                           : gate == 2 ? -kL2e * (__ldg(sg.bias0 + H + u) + __ldg(sg.bias1 + H + u))
                                       : __ldg(sg.bias1 + 2 * H + u);
                }
            } else {
                if (te < kFuLinN) sb[te] = __ldg(sg.bias0 + n * kFuLinN + te);
            }
            asm volatile("bar.sync 1, %0;" ::"n"(kTcEpiThreads) : "memory");   // the epilogue threads only
            if (te == 0) KTRACE(it * 48 + 4);
            mbar_wait(&tmem_full[ab], aphase);
            tc_fence_after();
            if (te == 0) KTRACE(it * 48 + 5);
            const uint32_t t0 = lane_base + ab * kTcAccCols;
            if (mode != kTcGru) {
                // linear tile: my warp owns 32 of the 128 outputs (columns part * 32 ..), one row per thread.  The accumulator
                // buffer is handed back after four TMEM loads; the outputs are staged in 128B-swizzled boxes (the linear
                // staging region, plus the idle GRU boxes for the decoder's fp32 tile) and stored by the linear store warp.
                float acc[32];
#pragma unroll
                for (int j = 0; j < 4; ++j) tmem_ld8(t0 + part * 32 + 8 * j, *reinterpret_cast<float(*)[8]>(acc + 8 * j));
                tmem_ld_wait();
                asm volatile("bar.sync 2, %0;" ::"n"(kTcEpiThreads) : "memory");   // everybody has read columns 0..63 ...
                tmem_zero8(t0 + part * 16);                  // ... which a GRU tile that takes this buffer next expects zeroed (its n_x)
                tmem_zero8(t0 + part * 16 + 8);
                tmem_st_wait();
                tc_fence_before();
                if (te == 0) KTRACE(it * 48 + 6);
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(empty_leader[ab]);
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const float v = acc[i] + sb[part * 32 + i];
                    acc[i] = mode == kTcEnc ? fmaxf(v, 0.0f) : sigmoid_f(v);
                }
                mbar_wait(lin_free, (lc & 1) ^ 1);           // the previous linear tile's stores have read the staging boxes
                if (mode == kTcEnc) {                        // box part / 2 holds columns 64 (part / 2) ..; my 32 columns = 4 chunks of 8 bf16
                    uint8_t *rowp = s_lin + (part >> 1) * kFuBoxF32 + row_in_cta * 128;
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        *reinterpret_cast<uint4 *>(rowp + ((((part & 1) * 4 + j) ^ sw) << 4)) = pack_bf16x8(acc + 8 * j);
                } else {                                     // my 32 fp32 columns = one box of 8 chunks of 4: parts 0,1 in the linear region, 2,3 in the GRU boxes
                    if (part >= 2) mbar_wait(gru_done, 0);
                    uint8_t *rowp = (part < 2 ? s_lin : s_f32) + (part & 1) * kFuBoxF32 + row_in_cta * 128;
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        *reinterpret_cast<float4 *>(rowp + ((j ^ sw) << 4)) = make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(lin_staged);
                ++lc;
                continue;
            }
            // GRU tile: two passes of 32 units through the staging buffers
            for (int c = 0; c < 2; ++c, ++pc) {
                const int buf = (int) (pc & 1);
                uint8_t *f32_row = s_f32 + buf * kFuBoxF32 + row_in_cta * 128;
                uint8_t *b16_row = s_b16 + buf * kFuBoxB16 + row_in_cta * 64;
                const int cu = c * 32 + part * 8;            // first of my 8 units inside the tile
                float anx[8], ar[8], az[8], anh[8], hp[8], out[8];
                tmem_ld8(t0 + 0 + cu, anx);
                tmem_ld8(t0 + 64 + cu, ar);
                tmem_ld8(t0 + 128 + cu, az);
                tmem_ld8(t0 + 192 + cu, anh);
                mbar_wait(&box_ready[buf], (pc >> 1) & 1);   // h(t-1) of this pass has landed
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const float4 v = *reinterpret_cast<const float4 *>(f32_row + (((part * 2 + q) ^ sw) << 4));
                    hp[4 * q] = v.x; hp[4 * q + 1] = v.y; hp[4 * q + 2] = v.z; hp[4 * q + 3] = v.w;
                }
                tmem_ld_wait();
                if (te == 0) KTRACE(it * 48 + 8 + c * 2);
                tmem_zero8(t0 + cu);                         // n_x columns must be zero when the next GRU tile's x part accumulates into them
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    // r = 1/(1+er), z = 1/(1+ez) share one reciprocal: 5 MUFU ops per unit.  Arguments are clamped so that
                    // (1+er)(1+ez) cannot overflow: sigmoid(-30) is already 9e-14.
                    const float er = ex2_approx(fminf(fmaf(ar[i], -kL2e, sb[64 + cu + i]), 43.0f));
                    const float ez = ex2_approx(fminf(fmaf(az[i], -kL2e, sb[128 + cu + i]), 43.0f));
                    const float pr = 1.0f + er, pz = 1.0f + ez;
                    const float ip = rcp_approx(pr * pz);
                    const float rg = pz * ip, zg = pr * ip;
                    const float a = fmaf(rg, anh[i] + sb[192 + cu + i], anx[i] + sb[cu + i]);
                    const float ng = fmaf(-2.0f, rcp_approx(1.0f + ex2_approx(a * (2.0f * kL2e))), 1.0f);   // tanh(a)
                    out[i] = fmaf(zg, hp[i] - ng, ng);       // (1 - z) n + z h
                }
                if (c == 1) {                                // last TMEM access of this tile: hand the buffer back
                    tmem_st_wait();
                    tc_fence_before();
                    if (te == 0) KTRACE(it * 48 + 6);
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(empty_leader[ab]);
                }
#pragma unroll
                for (int q = 0; q < 2; ++q)                  // h(t) replaces h(t-1) in place
                    *reinterpret_cast<float4 *>(f32_row + (((part * 2 + q) ^ sw) << 4)) =
                        make_float4(out[4 * q], out[4 * q + 1], out[4 * q + 2], out[4 * q + 3]);
                *reinterpret_cast<uint4 *>(b16_row + part * 16) = pack_bf16x8(out);
                fence_proxy_async();                         // my smem writes -> visible to the TMA engine
                __syncwarp();
                if (lane == 0) mbar_arrive(&staged[buf]);    // the state warp stores the pass once everybody is here
            }
        }
    }
    if (threadIdx.x == 0) KTRACE(1022);
    tc_fence_before();
    cluster_sync_all();      // the peer's smem / TMEM are read and written by the leader's MMAs: leave together
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, 2 * kTcAccCols);
    }
    if (threadIdx.x == 0) KTRACE(1023);
#undef KTRACE
}

// ---------------------------------------------------------------------------------------------------------------
// host side: tensor maps and segment tables for both state parities, the dependency counters, the launch
struct FuPlan {
    int nseg = 0, max_clusters = 0, num_sms = 0;
    __nv_bfloat16 *wih_p[kMaxLayers] = {}, *whh_p[kMaxLayers] = {};   // GRU weights packed per 64-unit tile (pack_gru_weights_kernel)
    unsigned epoch = 0;
    unsigned *counters = nullptr;
    CUtensorMap *d_maps[2] = {};      // [parity][nseg][kFuMapsPerSeg]
    FuArgs args[2];                   // by parity of the buffer that holds h(t-1)
    long long *trace = nullptr;       // 2 x 1024 clock64 slots of CTAs 0 and 1 (allocated when KOALA_FU_TRACE_BUF=1; written by -DKOALA_FU_TRACE=1 builds), else nullptr
};

static void fu_plan_destroy(FuPlan *f) {
    if (!f) return;
    for (int l = 0; l < kMaxLayers; l++) {
        if (f->wih_p[l]) cudaFree(f->wih_p[l]);
        if (f->whh_p[l]) cudaFree(f->whh_p[l]);
    }
    if (f->counters) cudaFree(f->counters);
    for (int i = 0; i < 2; i++)
        if (f->d_maps[i]) cudaFree(f->d_maps[i]);
    if (f->trace) cudaFree(f->trace);
    delete f;
}

static bool fu_plan_create(const TcModel &m, FuPlan **out, std::string *why) {
    if (m.H % 256 != 0 || m.Bp % kTcPairM != 0) {
        *why = "hidden size must be a multiple of 256 and the padded stream count a multiple of 256 for the tensor-core path";
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
    FuPlan *f = new FuPlan();
    const size_t H = m.H, Bp = m.Bp, L = m.L, LBH = Bp * H;
    const int mt = m.Bp / kTcPairM, nseg = m.L + 2;
    f->nseg = nseg;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&f->num_sms, cudaDevAttrMultiProcessorCount, dev);
    bool ok = true;
    for (size_t l = 0; l < L && ok; l++) {
        ok = cudaMalloc((void **) &f->wih_p[l], 3 * H * H * 2) == cudaSuccess && cudaMalloc((void **) &f->whh_p[l], 3 * H * H * 2) == cudaSuccess;
        if (!ok) break;
        pack_gru_weights_kernel<<<(unsigned) (3 * H), 64>>>(m.wih[l], f->wih_p[l], (int) H, 2, 0, 1);   // n | r | z
        pack_gru_weights_kernel<<<(unsigned) (3 * H), 64>>>(m.whh[l], f->whh_p[l], (int) H, 0, 1, 2);   // r | z | n
    }
    ok = ok && cudaDeviceSynchronize() == cudaSuccess;
    ok = ok && cudaMalloc((void **) &f->counters, (size_t) nseg * mt * sizeof(unsigned)) == cudaSuccess &&
              cudaMemset(f->counters, 0, (size_t) nseg * mt * sizeof(unsigned)) == cudaSuccess;
    if (const char *tr = getenv("KOALA_FU_TRACE_BUF")) {
        if (tr[0] == '1' && cudaMalloc((void **) &f->trace, 2048 * sizeof(long long)) == cudaSuccess) cudaMemset(f->trace, 0, 2048 * sizeof(long long));
    }
    for (int cur = 0; cur < 2 && ok; cur++) {
        const int nxt = cur ^ 1;
        std::vector<CUtensorMap> maps((size_t) nseg * kFuMapsPerSeg);
        FuArgs &a = f->args[cur];
        memset(&a, 0, sizeof(a));
        a.nseg = nseg; a.num_m_tiles = mt; a.H = m.H; a.counters = f->counters; a.trace = f->trace;
        int tile = 0;
        for (int s = 0; s < nseg && ok; s++) {
            FuSeg &sg = a.seg[s];
            CUtensorMap *mp = maps.data() + (size_t) s * kFuMapsPerSeg;
            bool used[kFuMapsPerSeg] = {};
            auto put = [&](int k, const void *base, uint64_t rows, uint64_t cols, uint32_t box_rows, bool f32 = false, bool plain32 = false) {
                used[k] = true;
                ok = ok && encode_2d(fn, &mp[k], base, rows, cols, box_rows, f32, plain32);
            };
            sg.tile_begin = tile;
            sg.dep = s == 0 ? -1 : s - 1;
            sg.done = s == nseg - 1 ? -1 : s;
            if (s == 0) {                                  // encoder: e = relu(feat W_enc^T + b)
                sg.mode = kTcEnc; sg.n_ctiles = m.H / kFuLinN / kFuPN; sg.kb_per_part = kBins / kTcBlockK; sg.parts = 1;
                sg.bias0 = m.enc_b;
                put(kMapA0, m.feat, Bp, kBins, kFuARows);
                put(kMapB0, m.enc_w, H, kBins, kFuLinN / 2);
                put(kMapHn, m.e, Bp, H, kTcBlockM);                  // store map: boxes of 64 bf16 columns x 128 rows, 128B swizzle
            } else if (s == nseg - 1) {                    // decoder: mask = sigmoid(h_{L-1}(t) W_dec^T + b)
                sg.mode = kTcDec; sg.n_ctiles = kBins / kFuLinN / kFuPN; sg.kb_per_part = m.H / kTcBlockK; sg.parts = 1;
                sg.bias0 = m.dec_b;
                put(kMapA0, m.hb[nxt] + (L - 1) * LBH, Bp, H, kFuARows);
                put(kMapB0, m.dec_w, kBins, H, kFuLinN / 2);
                put(kMapHn, m.mask, Bp, kBins, kTcBlockM, true);    // store map: boxes of 32 fp32 columns x 128 rows, 128B swizzle
            } else {                                       // GRU layer l
                const size_t l = s - 1;
                sg.mode = kTcGru; sg.n_ctiles = m.H / kGruUnits / kFuPN; sg.kb_per_part = m.H / kTcBlockK; sg.parts = 2;
                sg.bias0 = m.bih[l]; sg.bias1 = m.bhh[l]; sg.a1 = m.hb[cur] + l * LBH;
                put(kMapA0, l == 0 ? m.e : m.hb[nxt] + (l - 1) * LBH, Bp, H, kFuARows);
                put(kMapA1, m.hb[cur] + l * LBH, Bp, H, kFuARows);
                put(kMapB0, f->wih_p[l], 3 * H, H, kGruRows / 2);
                put(kMapB1, f->whh_p[l], 3 * H, H, kGruRows / 2);
                put(kMapHp, m.h[cur] + l * LBH, Bp, H, kTcBlockM, true);
                put(kMapHn, m.h[nxt] + l * LBH, Bp, H, kTcBlockM, true);
                put(kMapHb, m.hb[nxt] + l * LBH, Bp, H, kTcBlockM, false, true);
            }
            // rows of an m tile are written by every CTA (2 per pair tile) of every n tile of the previous segment
            sg.dep_per_step = s == 0 ? 0u : (unsigned) (a.seg[s - 1].n_ctiles * kFuPN * 2);
            for (int k = 0; k < kFuMapsPerSeg; k++)        // unused slots: any valid descriptor (they are only prefetched)
                if (!used[k]) mp[k] = mp[kMapA0];
            tile += mt * sg.n_ctiles;
        }
        a.total_tiles = tile;
        ok = ok && cudaMalloc((void **) &f->d_maps[cur], maps.size() * sizeof(CUtensorMap)) == cudaSuccess &&
             cudaMemcpy(f->d_maps[cur], maps.data(), maps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice) == cudaSuccess;
        a.maps = f->d_maps[cur];
    }
    ok = ok && cudaFuncSetAttribute(tc_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFuSmemBytes) == cudaSuccess;
    if (!ok) {
        *why = "setting up the fused mask-estimator kernel failed (tensor maps / counters / shared memory size)";
        cudaGetLastError();
        fu_plan_destroy(f);
        return false;
    }
    {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned) (kFuCluster * f->num_sms));
        cfg.blockDim = dim3(kFuThreads);
        cfg.dynamicSmemBytes = (size_t) kFuSmemBytes;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = kFuCluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, tc_fused_kernel, &cfg) != cudaSuccess || n < 1) { cudaGetLastError(); n = f->num_sms / kFuCluster - 4; }
        if (const char *e = getenv("KOALA_FU_CLUSTERS")) n = std::max(1, std::min(n, atoi(e)));
        f->max_clusters = n;
    }
    *out = f;
    return true;
}

// one mask-estimator step = one launch: hb[cur] / h[cur] hold state t-1, results go to hb[cur ^ 1] / h[cur ^ 1]
static int fu_masknet_step(FuPlan *f, int cur, cudaStream_t st) {
    FuArgs &a = f->args[cur];
    a.epoch = f->epoch + 1;
    const int clusters = a.total_tiles < f->max_clusters ? a.total_tiles : f->max_clusters;
    // the epoch only advances with a launch that was accepted: the counters then stand at epoch * (increments per step), which
    // is what the next launch waits for (the caller reports the launch error through cudaGetLastError)
    if (launch_pdl(true, tc_fused_kernel, dim3((unsigned) (kFuCluster * clusters)), dim3(kFuThreads), (size_t) kFuSmemBytes, st, a) == cudaSuccess) f->epoch = a.epoch;
    return 1;
}

}  // namespace koala
