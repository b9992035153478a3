// koala_b200 -- mask estimator, tensor-core path: ONE persistent kernel per CHUNK of steps (encoder -> GRU layers -> decoder,
// for every frame of the chunk).
//
// Middle stage of `pv_koala_process` (/root/reference/include/pv_koala.h:65-80), batched over the stream dimension
// (BASELINE.json configs[2..4]) and, inside one launch, walked over TIME: the caller's frame loop
// (/root/reference/demo/c/koala_demo_file.c:466-521) is strictly serial per stream, so a launch per frame pays the launch head
// (dependent-launch wait, cold operand pipeline, encoder tiles before the first GRU tile: ~20 k cycles) and tail (store drain:
// ~8 k) on every frame, and at small stream counts is nothing but dependency latency (profiles/r01_step_summary.md).  Here a
// launch takes `steps` consecutive frames.  The building blocks are those of tcgen05_common.cuh (bf16 operands staged by TMA
// into 128B-swizzled shared memory, tcgen05.mma.cta_group::2 accumulating fp32 in TMEM, gate math fused into the epilogue);
// this file is the schedule: one global list of pair tiles, period p = 0 .. steps + 1, each segment m-major,
//     period p: [GRU layer 0 of step p-1 | decoder of step p-2 | encoder of step p | GRU layer 1 of step p-1 | ... ]
// (tiles whose step falls outside the launch are skipped), walked round-robin by persistent CTA pairs (2-CTA clusters on all
// 148 SMs).  The encoder runs a step ahead and the decoder a step behind so that in the steady state every tile's producers
// are at least four rounds of the grid back in the list: with [enc | GRU 0 | GRU 1 | dec] per step the first GRU tiles of a
// step waited ~6 k cycles for encoder tiles issued just before them and the last decoder tiles a whole GRU tile (r02c trace).  Dependencies are per (segment, 256-stream m
// tile, step slot) COUNTERS in global memory that are never reset: every CTA of a tile of global step g (steps since the
// engine was created, 0-based) adds 1 to slot g % 4 of its row once the tile's output stores have completed, so "segment s
// has finished step g for m tile m" reads counter[s][m][g % 4] >= (g / 4 + 1) * (CTAs per m tile of that segment).  (One
// counter per row would not do: encoder and decoder tiles of step g + 1 do not depend on those of step g and may finish
// first; the (w) waits below bound that run-ahead to 3 steps, hence 4 slots.)  A tile of step g and segment s waits for
//   (x)  segment s-1, step g        the rows of its m tile from every n tile of the previous segment, same step;
//   (h)  segment s,   step g-1      GRU: every n tile of its own layer's previous step (bf16 h(t-1) operand, fp32 h(t-1) slice);
//   (w)  before its first store: the readers of what it overwrites -- the encoder output ring slot (GRU layer 0 of step
//        g - ring), the bf16 state copy of parity g (next segment of step g-2).
// All waits are for tiles EARLIER in the list and every role of a CTA takes its tiles in list order, so nothing can deadlock
// as long as the grid is co-resident (it is sized by cudaOccupancyMaxActiveClusters; engine.cu serialises fused launches of
// different engines).  Waits are relaxed polls of an L2-resident word: the producer's stores are complete in L2 before its
// release-add, and the consumers read through TMA (L2), so no acquire / proxy fence is needed on the consumer side -- those
// cost ~2 k cycles in a thread with TMA loads in flight (measured), which would sit on the recurrence's critical path.
// A GRU tile's k loop has two parts (x: input from the previous segment, h: recurrent); which one runs first is a per-engine
// choice: big batches run h first (at a launch's first step the encoder rows are still being produced), small batches -- whose
// throughput is the latency of the chain GRU_l(t-1) -> GRU_l(t) -- run x first, so that only half a k loop follows the wait.
//
// Warp roles (23 warps, 736 threads): 0,1 activation (A) producers for even / odd k-blocks, 2 TMEM allocator + MMA issuer
// (pair leader only), 3,4 weight (B) producers (they never wait: weights are constants), 5..20 epilogue (4 per TMEM lane
// quarter), 21 GRU state warp, 22 linear-tile store warp.
//   GRU epilogue: two PASSES of 32 units through two staging buffers; the state warp TMA-stores a finished pass (fp32 h(t)
//   in place of h(t-1), plus its bf16 copy) and requests the fp32 h(t-1) box of the pass two ahead into the drained buffer
//   (only if that tile's (h) dependency already holds: it must never block while passes of its own pair are unstored).
//   Linear epilogue: the accumulator is handed back after four TMEM loads; outputs are staged in a separate 32 KB region
//   (encoder: the CTA's whole [128][128] bf16 tile; decoder: two rounds of [128][64] fp32) and TMA-stored by warp 22.
// Shared memory: 5 operand stages x 28 KB + 48 KB GRU staging + 32 KB linear staging + barriers / biases = 223 KB.
//
// What bounds a GRU tile (8192 streams, B200; profiles/r01_step_summary.md): operand delivery -- the SMs ingest ~28 KB per
// k-block per ~500 cycles each, and a tile with N = 192 would need 73 B/clk/SM at full MMA rate.  TMEM (4 accumulator columns
// per unit, two buffers) caps the tile width, hence the operand bytes per flop.
#pragma once

#include <string.h>

#include <algorithm>
#include <vector>

#include "tcgen05_common.cuh"

namespace koala {

#ifndef KOALA_FU_STAGES
#define KOALA_FU_STAGES 5
#endif
constexpr int kFuStageBytes = kTcABytes + (kGruRows / 2) * 128;   // 28 KB: A [128][64] + B up to [96][64] bf16 (linear tiles: 64 rows)
constexpr int kFuCluster = 2;                                     // one CTA pair per cluster (4-CTA clusters with activation multicast: 2 % slower, r01)
constexpr int kFuLinN = 128;                                      // outputs per linear pair tile
constexpr int kFuBoxF32 = kTcBlockM * 32 * 4;                     // staging box [128 rows][32 fp32], 128B-swizzled
constexpr int kFuBoxB16 = kTcBlockM * 32 * 2;                     // staging box [128 rows][32 bf16], plain
constexpr int kFuLinBytes = 2 * kFuBoxF32;                        // linear tiles: [128][128] bf16 (encoder) or one round of [128][64] fp32 (decoder)
constexpr int kFuLinWarp = kTcStateWarp + 1;                      // stores the linear tiles' staged outputs
constexpr int kFuThreads = 32 * (kFuLinWarp + 1);                 // 2 + 1 + 2 + 16 + 1 + 1 warps
// Operand PLANES.  bf16 mode: one plane, the activation rounded to bf16 (SPEC.md section 4, q = bf16_rne).  fp32 mode: three
// planes hi | mid | lo with hi = bf16(v), mid = bf16(v - hi), lo = bf16(v - hi - mid): 3 x 8 significand bits = the whole fp32
// value, and the weights ARE bf16 values, so W v = W hi + W mid + W lo with every product exact in the fp32 accumulator.  An
// activation matrix is then [rows][planes * K] (plane p in columns p K ..), the k loop walks planes * K / 64 k-blocks and the
// weight k-block index wraps: the same tile schedule, dependency counters and epilogues serve both modes.  The planes cost
// 32 KB more staging, paid for with operand stages (3 instead of 5): fp32 mode is for small batches, which are latency-bound.
template <int kPlanes> struct FuCfg {
    static constexpr int kStages = kPlanes == 1 ? KOALA_FU_STAGES : 3;
    static constexpr int kSmemBytes = kStages * kFuStageBytes + 2 * (kFuBoxF32 + kPlanes * kFuBoxB16) + kFuLinBytes + kTcTailBytes;
};
constexpr int kFuMaxSegs = kMaxLayers + 2;
constexpr int kFuArrivals = kTcEpiWarps;                          // epilogue barriers: one arrival per warp (after __syncwarp), not per thread
constexpr int kFuSlots = 4;                                       // step slots per dependency counter row (>= encoder ring, >= 3)
enum FuMap : int { kMapA0 = 0, kMapA1, kMapB0, kMapB1, kMapHp, kMapHn, kMapHb, kFuMapsPerSeg };

struct FuSeg {
    int mode;                 // kTcEnc | kTcGru | kTcDec
    int n_tiles;              // pair tiles along n
    unsigned n_magic;         // ceil(2^32 / n_tiles): index / n_tiles = umulhi(index, n_magic) for index < 65536
    int kb_per_part, parts;   // k-blocks of 64 per operand part (all planes); GRU has two parts (x and h)
    int kb_w;                 // k-blocks of one weight row = kb_per_part / planes
    int step_delta;           // a tile of this segment in period p belongs to step p - 1 + step_delta (encoder +1, decoder -1)
    int x_first;              // GRU: the x part runs before the h part
    unsigned inc;             // what one step adds to this segment's counter of an m tile (2 CTAs per n tile)
    const float *bias0, *bias1;
};
struct FuArgs {
    int nseg, num_m_tiles, tiles_per_step, total_tiles, H, Bp;
    int steps;                // frames in this launch
    int cur0;                 // parity of the state buffers that hold h(t-1) of the launch's first step
    int e_ring;               // slots of the encoder-output ring (slot = global step % e_ring): a power of two <= kFuSlots
    long long epoch0;         // steps completed before this launch = global index of the launch's first step
    unsigned *counters;       // [nseg][num_m_tiles][kFuSlots]
    const CUtensorMap *maps;  // [2 parities][nseg][kFuMapsPerSeg], global memory
    long long *trace;
    int trace_skip;           // trace builds: first tile (of each pair's own sequence) that is recorded
    int pos_seg[kFuMaxSegs];          // list order inside a period: segment at position i ...
    int pos_begin[kFuMaxSegs + 1];    // ... and the index of its first tile
    FuSeg seg[kFuMaxSegs];
};

__device__ __forceinline__ unsigned ld_relaxed_gpu(const unsigned *p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_gpu(unsigned *p, unsigned v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ bool mbar_test(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return done != 0;
}
// blocks until the counter has reached `target` (modular comparison: the counters wrap)
__device__ __forceinline__ void fu_wait_counter(const unsigned *ctr, unsigned target) {
    while ((int) (ld_relaxed_gpu(ctr) - target) < 0) {
    }
}

// position of one pair tile in the global list
struct FuTile {
    int t, s, m, n;           // step inside the launch, segment (-1: list position without a tile in this launch), m tile, n tile
    int par;                  // parity of the state buffers that hold h(t-1) of this step
    long long g;              // global step index (0-based): epoch0 + t
};
// A role's walk over its pair's list positions g = cluster, cluster + stride, ...  Every role of every CTA decodes every tile it
// touches, the MMA issuer between two k loops: with a division per field the decode was ~600 cycles of a ~9 k-cycle GRU tile
// on the tensor pipe's critical path (r02i trace), so the period / offset pair is carried along instead of being recomputed,
// and the one remaining quotient is a multiplication.
struct FuIter {
    int g, p, local;          // list position, its period, its index inside the period
};
__device__ __forceinline__ void fu_iter_init(const FuArgs &a, FuIter &it, int g) {
    it.g = g;
    it.p = g / a.tiles_per_step;
    it.local = g - it.p * a.tiles_per_step;
}
__device__ __forceinline__ void fu_iter_advance(const FuArgs &a, FuIter &it, int stride) {
    it.g += stride;
    it.local += stride;
    while (it.local >= a.tiles_per_step) {
        it.local -= a.tiles_per_step;
        ++it.p;
    }
}
__device__ __forceinline__ FuTile fu_tile_at(const FuArgs &a, const FuIter &it) {
    FuTile r;
    int i = 0;
    while (i + 1 < a.nseg && it.local >= a.pos_begin[i + 1]) ++i;
    const int s = a.pos_seg[i], in_seg = it.local - a.pos_begin[i];
    r.t = it.p - 1 + a.seg[s].step_delta;
    r.s = (r.t >= 0 && r.t < a.steps) ? s : -1;
    r.m = (int) __umulhi((unsigned) in_seg, a.seg[s].n_magic);
    r.n = in_seg - r.m * a.seg[s].n_tiles;
    r.par = (a.cur0 + r.t) & 1;
    r.g = a.epoch0 + r.t;
    return r;
}
// the pair's next tile at or after the iterator's position; false when the list is exhausted
__device__ __forceinline__ bool fu_next(const FuArgs &a, FuIter &it, int stride, FuTile &t) {
    for (; it.g < a.total_tiles; fu_iter_advance(a, it, stride)) {
        t = fu_tile_at(a, it);
        if (t.s >= 0) return true;
    }
    return false;
}
__device__ __forceinline__ int fu_slot(long long g) { return (int) ((unsigned long long) g & (kFuSlots - 1)); }
// "segment s has finished global step g for m tile m": the counter and the value it has reached by then.  False: nothing to
// wait for (before the engine's first step).
__device__ __forceinline__ bool fu_done_ctr(const FuArgs &a, int s, int m, long long g, const unsigned *&ctr, unsigned &target) {
    if (g < 0) return false;
    ctr = a.counters + ((size_t) s * a.num_m_tiles + m) * kFuSlots + fu_slot(g);
    target = ((unsigned) ((unsigned long long) g / kFuSlots) + 1u) * a.seg[s].inc;
    return true;
}
__device__ __forceinline__ void fu_signal_done(const FuArgs &a, const FuTile &t) {
    red_release_gpu(a.counters + ((size_t) t.s * a.num_m_tiles + t.m) * kFuSlots + fu_slot(t.g), 1u);
}
// the two dependencies of a tile's loads: (x) previous segment, same step; (h) own segment, previous step
__device__ __forceinline__ bool fu_dep_x(const FuArgs &a, const FuTile &t, const unsigned *&ctr, unsigned &target) {
    return t.s > 0 && fu_done_ctr(a, t.s - 1, t.m, t.g, ctr, target);
}
__device__ __forceinline__ bool fu_dep_h(const FuArgs &a, const FuTile &t, const unsigned *&ctr, unsigned &target) {
    return a.seg[t.s].mode == kTcGru && fu_done_ctr(a, t.s, t.m, t.g - 1, ctr, target);
}
// (w): the readers of the buffer this tile's stores overwrite must be done
__device__ __forceinline__ bool fu_dep_w(const FuArgs &a, const FuTile &t, const unsigned *&ctr, unsigned &target) {
    const FuSeg &sg = a.seg[t.s];
    if (sg.mode == kTcEnc)         // ring slot g % ring was read by GRU layer 0 of step g - ring
        return fu_done_ctr(a, 1, t.m, t.g - a.e_ring, ctr, target);
    if (sg.mode == kTcGru)         // the bf16 copy this step writes held h(g-2), read as the x operand by the next segment at step g-2
        return fu_done_ctr(a, t.s + 1, t.m, t.g - 2, ctr, target);
    return false;
}
// A dependency sampled ahead of time: the relaxed load is issued when the tile is decoded and only looked at when the tile
// is reached, so its L2 round trip (~1 k cycles under load) stays off the critical path; if it did not hold yet, poll.
struct FuDep {
    const unsigned *ctr;
    unsigned target, seen;
    __device__ __forceinline__ void none() { ctr = nullptr; target = 0; seen = 0; }
    __device__ __forceinline__ void sample() { if (ctr) seen = ld_relaxed_gpu(ctr); }
    __device__ __forceinline__ bool holds() { return ctr == nullptr || (int) (seen - target) >= 0; }
    __device__ __forceinline__ bool poll() { sample(); return holds(); }
    __device__ __forceinline__ void wait() { if (!holds()) { fu_wait_counter(ctr, target); seen = target; } }
};

template <int kPlanes>
__global__ void __cluster_dims__(kFuCluster, 1, 1) __launch_bounds__(kFuThreads, 1) tc_fused_kernel(const __grid_constant__ FuArgs args) {
    constexpr int kStages = FuCfg<kPlanes>::kStages, kStageBytes = kFuStageBytes;

    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t) 1023);
    uint8_t *s_f32 = smem + kStages * kStageBytes;       // 2 fp32 staging boxes
    uint8_t *s_b16 = s_f32 + 2 * kFuBoxF32;              // 2 x kPlanes bf16 staging boxes
    uint8_t *s_lin = s_b16 + 2 * kPlanes * kFuBoxB16;    // linear tiles' staging: two 16 KB swizzled boxes
    uint8_t *tail = s_lin + kFuLinBytes;
    uint64_t *bars = reinterpret_cast<uint64_t *>(tail);
    uint64_t *full_bar = bars, *empty_bar = bars + kStages;
    uint64_t *tmem_full = bars + 2 * kStages, *tmem_empty = bars + 2 * kStages + 2;
    uint64_t *box_ready = bars + 2 * kStages + 4;        // [2]: the staging buffer holds h(t-1) of a pass
    uint64_t *staged = bars + 2 * kStages + 6;           // [2]: the epilogue has staged a pass in the buffer
    uint64_t *lin_free = bars + 2 * kStages + 8;         // the stores of a linear round have read the staging region
    uint64_t *lin_staged16 = bars + 2 * kStages + 9;     // encoder tile staged (all 16 epilogue warps)
    uint64_t *lin_staged8 = bars + 2 * kStages + 10;     // one decoder round staged (the 8 warps that own its columns)
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * kStages + 11);
    float *s_bias = reinterpret_cast<float *>(tail + 256);   // [2 accumulator buffers][256]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();       // 0 = pair leader (issues the MMAs), 1 = peer
    constexpr uint32_t leader = 0;
    constexpr uint16_t pair_mask = 3;
    [[maybe_unused]] long long *trace = (args.trace != nullptr && blockIdx.x < 2) ? args.trace + blockIdx.x * 1024 : nullptr;
#ifdef KOALA_FU_TRACE      // clock64 timeline of cluster 0 (tools/gpu_trace.py builds this variant); compiled out of the product
#define KTRACE(tile, slot) do { const int ti__ = (tile) - args.trace_skip; if (trace && ti__ >= 0 && ti__ < 21) trace[ti__ * 48 + (slot)] = clock64(); } while (0)   /* 21 tiles x 48 slots = 1008; 1012..1015 globaltimer, 1020..1023 clock64 of the header events */
#define KTRACE_HDR(slot) do { if (trace) { trace[(slot)] = clock64(); unsigned long long ns__; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns__)); trace[(slot) - 8] = (long long) ns__; } } while (0)
#else
#define KTRACE(tile, slot) do { } while (0)
#define KTRACE_HDR(slot) do { } while (0)
#endif
    if (threadIdx.x == 0) KTRACE_HDR(1020);
    pdl_launch_dependents();
    const int cluster_id = blockIdx.x / kFuCluster, num_clusters = gridDim.x / kFuCluster;
    const int total = args.total_tiles;   // list positions (valid or not)
    const int ring_mask = args.e_ring - 1;

    if (warp == 0) {
        for (int i = lane; i < 2 * args.nseg * kFuMapsPerSeg; i += 32) prefetch_tmap(args.maps + i);
        if (lane == 0) {
            for (int s = 0; s < kStages; ++s) {
                mbar_init(&full_bar[s], 1);          // leader's copy is the one in use: 1 arrive.expect_tx + 2 CTAs' TMA bytes
                mbar_init(&empty_bar[s], 1);         // one multicast tcgen05.commit from the pair leader
            }
            for (int b = 0; b < 2; ++b) {
                mbar_init(&tmem_full[b], 1);         // one multicast tcgen05.commit
                mbar_init(&tmem_empty[b], 2 * kFuArrivals);     // leader's copy: the epilogue warps of both CTAs
                mbar_init(&box_ready[b], 1);
                mbar_init(&staged[b], kFuArrivals);
            }
            mbar_init(lin_free, 1);
            mbar_init(lin_staged16, kFuArrivals);
            mbar_init(lin_staged8, kFuArrivals / 2);
            fence_mbar_init();
        }
    }
    if (warp == 2) tmem_alloc_pair(tmem_slot, 2 * kTcAccCols);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) KTRACE_HDR(1021);

    if (warp == 0 || warp == 1 || warp == 3 || warp == 4) {
        // ===================================================== TMA producers (both CTAs of every pair).  One warp can issue a
        // tensor load only every ~210-380 cycles (tools/micro/tma_rate.cu), so the loads are spread over four warps: warps
        // 0/1 fetch the activation (A) tiles of the even/odd k-blocks, warps 3/4 the weight (B) tiles.  Everything stays
        // warp-uniform (elect_one) so descriptors live in uniform registers.  Weights do not depend on anything: the B
        // producers run ahead (also into the previous kernel's tail); only the A producers wait, for the previous kernel
        // once and for the dependency counters before the first load of each operand part.
        const bool is_a = warp < 2;
        const int par = is_a ? warp : warp - 3;             // my k-block parity
        if (is_a) pdl_wait();
        int stage = par, phase = 0;                         // kStages is odd or even: see the advance at the bottom of the loop
        const int arow = (int) rank * kTcBlockM;            // my 128 rows inside the pair's 256-stream tile
        int pit = 0;
        unsigned seen_x = 0, seen_h = 0;    // the next tile's dependency counters, sampled one tile early (hides the L2 round trip)
        bool have_seen = false;
        FuTile t, t1;
        FuIter it, it1;
        fu_iter_init(args, it, cluster_id);
        it1 = it;
        bool have = fu_next(args, it, num_clusters, t), have1 = false;
        for (; have; it = it1, t = t1, have = have1, ++pit) {
            const FuSeg &sg = args.seg[t.s];
            const CUtensorMap *maps = args.maps + (size_t) (t.par * args.nseg + t.s) * kFuMapsPerSeg;
            const bool gru = sg.mode == kTcGru;
            const int kbp = sg.kb_per_part, num_kb = kbp * sg.parts, kbw = sg.kb_w;
            const int brows = gru ? kGruRows / 2 : kFuLinN / 2;         // weight rows each CTA of the pair holds
            const uint32_t pair_tx = 2u * (uint32_t) (kTcABytes + brows * 128);
            const int hpart = gru ? (sg.x_first ? 1 : 0) : -1;          // position of the h part in the k loop
            // x operand rows: features of slot t (encoder), encoder-output ring slot (GRU layer 0), else a state copy
            const int a0_row = t.m * kTcPairM + arow +
                               (sg.mode == kTcEnc ? t.t * args.Bp : t.s == 1 ? ((int) t.g & ring_mask) * args.Bp : 0);
            const int a1_row = t.m * kTcPairM + arow;
            const unsigned my_seen_x = seen_x, my_seen_h = seen_h;
            const bool my_have = have_seen;
            have_seen = false;
            if (warp == 0 && lane == 0) KTRACE(pit, 0);
            for (int kb = par; kb < num_kb; kb += 2) {
                const int part = kb >= kbp ? 1 : 0;
                const bool hp = part == hpart;
                if (is_a && (kb == par || kb == kbp + par)) {           // my first load of this operand part: its producers must be done
                    const unsigned *ctr = nullptr;
                    unsigned target = 0;
                    const bool need = hp ? fu_dep_h(args, t, ctr, target) : fu_dep_x(args, t, ctr, target);
                    if (need && !(my_have && (int) ((hp ? my_seen_h : my_seen_x) - target) >= 0)) fu_wait_counter(ctr, target);
                    if (warp == 0 && lane == 0) KTRACE(pit, 12 + part);
                }
                mbar_wait(&empty_bar[stage], phase ^ 1);     // slot free in both CTAs of the pair
                if (is_a && lane == 0) KTRACE(pit, 16 + kb);
                const bool elected = elect_one();
                if (elected && is_a && rank == 0) mbar_expect_tx(&full_bar[stage], pair_tx);
                const uint32_t full_leader = map_to_cta(&full_bar[stage], leader);
                uint8_t *sa = smem + stage * kStageBytes, *sb = sa + kTcABytes;
                int kbi = kb - part * kbp;                  // k-block inside the operand part: activation column block ...
                if (!is_a && kPlanes > 1) {                 // ... the weights repeat for every plane
                    while (kbi >= kbw) kbi -= kbw;
                }
                const int kc = kbi * kTcBlockK;
                if (!elected) {
                } else if (is_a) {
                    tma_load_2d_pair(maps + (hp ? kMapA1 : kMapA0), full_leader, sa, kc, hp ? a1_row : a0_row);
                } else {
                    tma_load_2d_pair(maps + (hp ? kMapB1 : kMapB0), full_leader, sb, kc, t.n * 2 * brows + (int) rank * brows);
                }
                __syncwarp();
                if (kb == par) {         // my first load of the tile is on its way: decode the pair's next tile behind it
                    it1 = it;
                    fu_iter_advance(args, it1, num_clusters);
                    have1 = fu_next(args, it1, num_clusters, t1);
                    if (have1) {
                        if (is_a) {
                            const unsigned *ctr = nullptr;
                            unsigned target = 0;
                            if (fu_dep_x(args, t1, ctr, target)) seen_x = ld_relaxed_gpu(ctr);
                            if (fu_dep_h(args, t1, ctr, target)) seen_h = ld_relaxed_gpu(ctr);
                            have_seen = true;
                        }
                    }
                    __syncwarp();
                }
                stage += 2;
                if (stage >= kStages) { stage -= kStages; phase ^= 1; }
            }
            if (warp == 0 && lane == 0) KTRACE(pit, 1);
        }
    } else if (warp == 2) {
        // ===================================================== MMA issuer (one thread of the leader CTA drives both SMs)
        // The whole tile loop runs in ONE elected lane: every instruction between two k-blocks (barrier test, descriptor
        // arithmetic, moves into uniform registers) and between two tiles (decoding the next tile: ~500 cycles of dependent
        // constant-bank loads) is serial latency of this thread and shows up one-for-one as idle tensor-pipe time, so nothing
        // is re-elected, re-synchronised or broadcast, the k loop is split by operand part to keep its body branch-free, and the
        // NEXT tile is decoded right behind the first k-block of the current one, while its MMAs run (the r02j trace showed
        // ~900 cycles between a tile's last MMA and the next tile's first with the decode at the top of the loop).
        // GRU: the h part writes columns [r | z | n_h], the x part [n_x | r | z]; whichever runs first overwrites its columns,
        // the other accumulates -- its private n columns were left zeroed by the epilogue.
        if (rank == 0 && elect_one()) {
            int it = 0, stage = 0, phase = 0;
            const uint64_t adesc0 = make_sw128_desc(smem_u32(smem)), bdesc0 = make_sw128_desc(smem_u32(smem) + kTcABytes);
            FuTile t, t_next;
            FuIter pos;
            fu_iter_init(args, pos, cluster_id);
            bool have = fu_next(args, pos, num_clusters, t), have_next = false;
            for (; have; ++it) {
                const FuSeg &sg = args.seg[t.s];
                const bool gru = sg.mode == kTcGru;
                const int kbp = sg.kb_per_part, parts = sg.parts;
                const uint32_t idesc = gru ? make_idesc(256, kGruRows) : make_idesc(256, kFuLinN);
                const int hpart = gru ? (sg.x_first ? 1 : 0) : -1;
                const int ab = it & 1, aphase = (it >> 1) & 1;
                mbar_wait(&tmem_empty[ab], aphase);      // both CTAs' epilogues have released (and cleared) this buffer
                tc_fence_after();
                KTRACE(it, 2);
                const uint32_t d = tmem_base + ab * kTcAccCols;
                for (int part = 0; part < parts; ++part) {
                    const uint32_t dk = part == hpart ? d + kGruUnits : d;
                    const uint32_t first = part == 0 ? 0u : 1u;
                    for (int kb = 0; kb < kbp; ++kb) {
                        mbar_wait(&full_bar[stage], phase);
                        tc_fence_after();
                        KTRACE(it, 32 + part * kbp + kb);
                        const uint64_t so = (uint64_t) ((stage * kStageBytes) >> 4);
#pragma unroll
                        for (int k = 0; k < kTcBlockK / 16; ++k)    // +32 B per 16-element k-step
                            umma_bf16_pair(dk, adesc0 + so + 2 * k, bdesc0 + so + 2 * k, idesc, (kb | k) != 0 ? 1u : first);
                        umma_commit_pair(&empty_bar[stage], pair_mask);
                        if (++stage == kStages) { stage = 0; phase ^= 1; }
                        if ((part | kb) == 0) {              // the tensor pipe has a k-block to chew on: find the next tile now
                            fu_iter_advance(args, pos, num_clusters);
                            have_next = fu_next(args, pos, num_clusters, t_next);
                        }
                    }
                }
                umma_commit_pair(&tmem_full[ab], pair_mask);
                KTRACE(it, 3);
                t = t_next;
                have = have_next;
            }
        }
        __syncwarp();
    } else if (warp == kTcStateWarp) {
        // ===================================================== state warp: one thread moves the staging boxes of the GRU tiles
        // (two passes of 32 units each).  `cur` walks the passes in order (wait until staged, store fp32 + bf16 h(t), commit);
        // `ahead` runs up to two passes in front and requests the fp32 h(t-1) box of that pass into the buffer `cur` has
        // drained -- the slice was written by the same tile position one step earlier, so the tile's (h) dependency must hold
        // first.  That test never blocks while a staged pass of this pair may still be unstored (the dependency can be this very
        // pair's previous tile); it blocks only when the pass is the next one to be needed.  After a tile's second pass the warp
        // waits for the stores to COMPLETE and bumps the segment's counter for the m tile, which releases the dependent tiles.
        pdl_wait();
        if (elect_one()) {
            const int row0 = (int) rank * kTcBlockM;
            struct PassIter {
                FuIter it;
                int g, c;
                FuTile t;
                FuDep h, w;      // the tile's (h) and (w) dependencies, sampled when the iterator reaches the tile
            };
            auto start = [&](PassIter &p) {                  // first pass of the pair's next GRU tile at or after p.it
                for (; fu_next(args, p.it, num_clusters, p.t); fu_iter_advance(args, p.it, num_clusters))
                    if (args.seg[p.t.s].mode == kTcGru) break;
                p.g = p.it.g;
                p.c = 0;
                p.h.none();
                p.w.none();
                if (p.g < total) {
                    fu_dep_h(args, p.t, p.h.ctr, p.h.target);
                    fu_dep_w(args, p.t, p.w.ctr, p.w.target);
                    p.h.sample();
                    p.w.sample();
                }
            };
            auto advance = [&](PassIter &p) {
                if (++p.c == 2) {
                    fu_iter_advance(args, p.it, num_clusters);
                    start(p);
                }
            };
            PassIter cur, ahead;
            fu_iter_init(args, cur.it, cluster_id);
            start(cur);
            ahead = cur;
            unsigned armed = 0, pc = 0;      // passes requested / stored so far; pass q uses staging buffer q & 1
            auto try_arm = [&](bool block) {
                if (ahead.g >= total || armed >= pc + 2) return;          // nothing left, or the buffer still holds an unstored pass
                if (!ahead.h.holds() && !ahead.h.poll()) {
                    if (!block) return;
                    ahead.h.wait();
                }
                const int buf = (int) (armed & 1);
                mbar_expect_tx(&box_ready[buf], kFuBoxF32);
                tma_load_2d_local(args.maps + (size_t) (ahead.t.par * args.nseg + ahead.t.s) * kFuMapsPerSeg + kMapHp, &box_ready[buf],
                                  s_f32 + buf * kFuBoxF32, ahead.t.n * kGruUnits + 32 * ahead.c, ahead.t.m * kTcPairM + row0);
                advance(ahead);
                ++armed;
            };
            int it = 0;
            while (cur.g < total) {
                if (armed <= pc) try_arm(true);      // the pass the epilogue needs next: everything before it is stored and signalled
                try_arm(false);
                const int buf = (int) (pc & 1);
                const CUtensorMap *maps = args.maps + (size_t) (cur.t.par * args.nseg + cur.t.s) * kFuMapsPerSeg;
                const int row = cur.t.m * kTcPairM + row0, col = cur.t.n * kGruUnits + 32 * cur.c;
                mbar_wait(&staged[buf], (pc >> 1) & 1);
                if (cur.c == 0) cur.w.wait();
                tma_store_2d(maps + kMapHn, s_f32 + buf * kFuBoxF32, col, row);
#pragma unroll
                for (int pl = 0; pl < kPlanes; ++pl)         // bf16 copy (fp32 mode: its three planes) = the operand of the next tiles
                    tma_store_2d(maps + kMapHb, s_b16 + (buf * kPlanes + pl) * kFuBoxB16, col + pl * args.H, row);
                bulk_commit();
                KTRACE((cur.g - cluster_id) / num_clusters, 9 + (cur.c & 1) * 2);
                bulk_wait_read();                            // the stores have read buffer `buf`
                ++pc;
                if (cur.c == 1) {
                    bulk_wait_all();                         // this tile's rows are in global memory
                    fence_proxy_async_all();
                    fu_signal_done(args, cur.t);
                    ++it;
                }
                advance(cur);
                try_arm(false);                              // the pass two ahead, into the buffer just drained
            }
        }
        __syncwarp();
    } else if (warp == kFuLinWarp) {
        // ===================================================== linear-tile store warp: encoder / decoder outputs leave through
        // their own 32 KB staging region, so they never compete with the GRU passes for staging buffers: the encoder's whole
        // [128][128] bf16 tile of this CTA fits; the decoder's [128][128] fp32 tile takes two rounds of 64 columns.  Per round:
        // wait until the epilogue warps have staged it, TMA-store its boxes, free the region once they have been read; after
        // the tile's last round, wait for the stores to complete and release the dependent tiles.
        pdl_wait();
        if (elect_one()) {
            const int row0 = (int) rank * kTcBlockM;
            unsigned n16 = 0, n8 = 0;
            FuTile t;
            FuIter pos;
            for (fu_iter_init(args, pos, cluster_id); fu_next(args, pos, num_clusters, t); fu_iter_advance(args, pos, num_clusters)) {
                const FuSeg &sg = args.seg[t.s];
                if (sg.mode == kTcGru) continue;
                const CUtensorMap *map = args.maps + (size_t) (t.par * args.nseg + t.s) * kFuMapsPerSeg + kMapHn;
                const int col = t.n * kFuLinN;
                FuDep w;
                w.none();
                fu_dep_w(args, t, w.ctr, w.target);
                w.sample();                        // looked at after the wait for the staged tile
                if (sg.mode == kTcEnc) {           // per plane: two boxes of 64 bf16 columns into ring slot G % ring
                    const int row = t.m * kTcPairM + row0 + ((int) t.g & ring_mask) * args.Bp;
                    for (int pl = 0; pl < kPlanes; ++pl) {
                        mbar_wait(lin_staged16, n16++ & 1);
                        if (pl == 0) w.wait();
                        tma_store_2d(map, s_lin, col + pl * args.H, row);
                        tma_store_2d(map, s_lin + kFuBoxF32, col + pl * args.H + 64, row);
                        bulk_commit();
                        bulk_wait_read();
                        mbar_arrive(lin_free);
                    }
                } else {                           // two rounds of two boxes of 32 fp32 columns into mask slot t
                    const int row = t.m * kTcPairM + row0 + t.t * args.Bp;
                    for (int r = 0; r < 2; ++r) {
                        mbar_wait(lin_staged8, n8++ & 1);
                        tma_store_2d(map, s_lin, col + 64 * r, row);
                        tma_store_2d(map, s_lin + kFuBoxF32, col + 64 * r + 32, row);
                        bulk_commit();
                        bulk_wait_read();
                        mbar_arrive(lin_free);
                    }
                }
                bulk_wait_all();                             // this tile's rows are in global memory
                fence_proxy_async_all();
                fu_signal_done(args, t);
            }
        }
        __syncwarp();
    } else if (warp >= 5 && warp < kTcStateWarp) {
        // ===================================================== epilogue: warps 5..20.  TMEM lane quarter = warp % 4 (a warp can
        // only touch its own 32 lanes); the 4 warps of a quarter take 8 of a pass's 32 columns each.  16 warps, not 8: the
        // gate math is a long dependent chain (5 MUFU ops per unit) and needs the extra warps per scheduler to hide it.
        const int quarter = warp & 3, part = (warp - 5) >> 2, te = threadIdx.x - 160;
        const uint32_t lane_base = tmem_base + ((uint32_t) (quarter * 32) << 16);
        uint32_t empty_leader[2] = {map_to_cta(&tmem_empty[0], leader), map_to_cta(&tmem_empty[1], leader)};
        // hand both buffers to the MMA issuer for the first time, with the GRU n_x and n_h columns cleared
#pragma unroll
        for (int b = 0; b < 2; ++b) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                tmem_zero8(lane_base + b * kTcAccCols + h * 3 * kGruUnits + part * 16);
                tmem_zero8(lane_base + b * kTcAccCols + h * 3 * kGruUnits + part * 16 + 8);
            }
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
        unsigned pc = 0, lf = 0;       // GRU passes / linear rounds so far: staging buffer and barrier phase
        int it = 0;
        FuTile t;
        FuIter pos;
        for (fu_iter_init(args, pos, cluster_id); fu_next(args, pos, num_clusters, t); fu_iter_advance(args, pos, num_clusters), ++it) {
            const FuSeg &sg = args.seg[t.s];
            const int mode = sg.mode, n = t.n;
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
            if (te == 0) KTRACE(it, 4);
            mbar_wait(&tmem_full[ab], aphase);
            tc_fence_after();
            if (te == 0) KTRACE(it, 5);
            const uint32_t t0 = lane_base + ab * kTcAccCols;
            if (mode != kTcGru) {
                // linear tile: my warp owns 32 of the 128 outputs (columns part * 32 ..), one row per thread.  The accumulator
                // buffer is handed back after four TMEM loads; the outputs are staged in 128B-swizzled boxes of the linear
                // staging region and stored by the linear store warp.
                float acc[32];
#pragma unroll
                for (int j = 0; j < 4; ++j) tmem_ld8(t0 + part * 32 + 8 * j, *reinterpret_cast<float(*)[8]>(acc + 8 * j));
                tmem_ld_wait();
                asm volatile("bar.sync 2, %0;" ::"n"(kTcEpiThreads) : "memory");   // everybody has read columns 0..63 ...
                tmem_zero8(t0 + part * 16);                  // ... which a GRU tile that takes this buffer next expects zeroed (its n_x)
                tmem_zero8(t0 + part * 16 + 8);
                tmem_st_wait();
                tc_fence_before();
                if (te == 0) KTRACE(it, 6);
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(empty_leader[ab]);
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const float v = acc[i] + sb[part * 32 + i];
                    acc[i] = mode == kTcEnc ? fmaxf(v, 0.0f) : sigmoid_f(v);
                }
                if (mode == kTcEnc) {                        // box part / 2 holds columns 64 (part / 2) ..; my 32 columns = 4 chunks of 8 bf16
                    uint8_t *rowp = s_lin + (part >> 1) * kFuBoxF32 + row_in_cta * 128;
#pragma unroll
                    for (int pl = 0; pl < kPlanes; ++pl) {   // one round per plane: what is staged is subtracted, the rest goes to the next plane
                        mbar_wait(lin_free, ((lf + (unsigned) pl) & 1) ^ 1);   // the previous linear round's stores have read the staging boxes
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const uint4 u = pack_bf16x8(acc + 8 * j);
                            *reinterpret_cast<uint4 *>(rowp + ((((part & 1) * 4 + j) ^ sw) << 4)) = u;
                            if (pl + 1 < kPlanes) sub_bf16x8(acc + 8 * j, u);
                        }
                        fence_proxy_async();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(lin_staged16);
                    }
                    lf += kPlanes;
                } else {                                     // round part / 2, box part & 1: my 32 fp32 columns = 8 chunks of 4
                    mbar_wait(lin_free, ((lf + (unsigned) (part >> 1)) & 1) ^ 1);
                    uint8_t *rowp = s_lin + (part & 1) * kFuBoxF32 + row_in_cta * 128;
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        *reinterpret_cast<float4 *>(rowp + ((j ^ sw) << 4)) = make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(lin_staged8);
                    lf += 2;
                }
                continue;
            }
            // GRU tile: two passes of 32 units through the staging buffers
            for (int c = 0; c < 2; ++c, ++pc) {
                const int buf = (int) (pc & 1);
                uint8_t *f32_row = s_f32 + buf * kFuBoxF32 + row_in_cta * 128;
                uint8_t *b16_row = s_b16 + buf * kPlanes * kFuBoxB16 + row_in_cta * 64;
                const int cu = c * 32 + part * 8;            // first of my 8 units inside the tile
                float anx[8], ar[8], az[8], anh[8], hp[8], out[8];
                tmem_ld8(t0 + 0 + cu, anx);
                tmem_ld8(t0 + 64 + cu, ar);
                tmem_ld8(t0 + 128 + cu, az);
                tmem_ld8(t0 + 192 + cu, anh);
                if (te == 0) KTRACE(it, 14 + c);
                mbar_wait(&box_ready[buf], (pc >> 1) & 1);   // h(t-1) of this pass has landed
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const float4 v = *reinterpret_cast<const float4 *>(f32_row + (((part * 2 + q) ^ sw) << 4));
                    hp[4 * q] = v.x; hp[4 * q + 1] = v.y; hp[4 * q + 2] = v.z; hp[4 * q + 3] = v.w;
                }
                tmem_ld_wait();
                if (te == 0) KTRACE(it, 8 + c * 2);
                tmem_zero8(t0 + cu);                         // the n columns only one operand part touches must be zero when the
                tmem_zero8(t0 + 3 * kGruUnits + cu);         // next GRU tile's second part accumulates into them
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
                    if (te == 0) KTRACE(it, 6);
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(empty_leader[ab]);
                }
#pragma unroll
                for (int q = 0; q < 2; ++q)                  // h(t) replaces h(t-1) in place
                    *reinterpret_cast<float4 *>(f32_row + (((part * 2 + q) ^ sw) << 4)) =
                        make_float4(out[4 * q], out[4 * q + 1], out[4 * q + 2], out[4 * q + 3]);
#pragma unroll
                for (int pl = 0; pl < kPlanes; ++pl) {
                    const uint4 u = pack_bf16x8(out);
                    *reinterpret_cast<uint4 *>(b16_row + pl * kFuBoxB16 + part * 16) = u;
                    if (pl + 1 < kPlanes) sub_bf16x8(out, u);
                }
                fence_proxy_async();                         // my smem writes -> visible to the TMA engine
                __syncwarp();
                if (lane == 0) mbar_arrive(&staged[buf]);    // the state warp stores the pass once everybody is here
            }
        }
    }
    if (threadIdx.x == 0) KTRACE_HDR(1022);
    tc_fence_before();
    cluster_sync_all();      // the peer's smem / TMEM are read and written by the leader's MMAs: leave together
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, 2 * kTcAccCols);
    }
    if (threadIdx.x == 0) KTRACE_HDR(1023);
#undef KTRACE
#undef KTRACE_HDR
}

// ---------------------------------------------------------------------------------------------------------------
// host side: tensor maps and segment tables for both state parities, the dependency counters, the launch
struct FuPlan {
    int nseg = 0, max_clusters = 0, num_sms = 0, tcap = 1, planes = 1;
    __nv_bfloat16 *wih_p[kMaxLayers] = {}, *whh_p[kMaxLayers] = {};   // GRU weights packed per 64-unit tile (pack_gru_weights_kernel)
    long long epoch = 0;              // steps completed by accepted launches
    unsigned *counters = nullptr;
    CUtensorMap *d_maps = nullptr;    // [parity][nseg][kFuMapsPerSeg]
    FuArgs args;
    long long *trace = nullptr;       // 2 x 1024 clock64 slots of CTAs 0 and 1 (allocated when KOALA_FU_TRACE_BUF=1; written by -DKOALA_FU_TRACE=1 builds), else nullptr
};

static void fu_plan_destroy(FuPlan *f) {
    if (!f) return;
    for (int l = 0; l < kMaxLayers; l++) {
        if (f->wih_p[l]) cudaFree(f->wih_p[l]);
        if (f->whh_p[l]) cudaFree(f->whh_p[l]);
    }
    if (f->counters) cudaFree(f->counters);
    if (f->d_maps) cudaFree(f->d_maps);
    if (f->trace) cudaFree(f->trace);
    delete f;
}

typedef void (*FuKernel)(const FuArgs);
static FuKernel fu_kernel(int planes) { return planes == 1 ? tc_fused_kernel<1> : tc_fused_kernel<3>; }
static int fu_smem_bytes(int planes) { return planes == 1 ? FuCfg<1>::kSmemBytes : FuCfg<3>::kSmemBytes; }

// `m.feat` / `m.mask` hold m.tcap step slots of [Bp] rows, `m.e` holds m.e_ring slots; activation matrices (feat, e, hb) have
// m.planes * K columns
static bool fu_plan_create(const TcModel &m, FuPlan **out, std::string *why) {
    if (m.H % 256 != 0 || m.Bp % kTcPairM != 0 || (m.planes != 1 && m.planes != 3) || (m.e_ring & (m.e_ring - 1)) != 0 ||
        (size_t) (m.Bp / kTcPairM) * (m.H / kGruUnits) >= 65536) {
        *why = "hidden size must be a multiple of 256, the padded stream count a multiple of 256 (at most 2 M streams at H = 512) and the "
               "encoder ring a power of two for the tensor-core path";
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
    f->tcap = m.tcap;
    f->planes = m.planes;
    const uint64_t P = (uint64_t) m.planes;
    FuKernel kernel = fu_kernel(m.planes);
    const int smem_bytes = fu_smem_bytes(m.planes);
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
    ok = ok && cudaMalloc((void **) &f->counters, (size_t) nseg * mt * kFuSlots * sizeof(unsigned)) == cudaSuccess &&
              cudaMemset(f->counters, 0, (size_t) nseg * mt * kFuSlots * sizeof(unsigned)) == cudaSuccess;
    int trace_skip = 0;
    if (const char *tr = getenv("KOALA_FU_TRACE_BUF")) {
        if (tr[0] == '1' && cudaMalloc((void **) &f->trace, 2048 * sizeof(long long)) == cudaSuccess) cudaMemset(f->trace, 0, 2048 * sizeof(long long));
        if (const char *sk = getenv("KOALA_FU_TRACE_SKIP")) trace_skip = atoi(sk);
    }
    ok = ok && cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes) == cudaSuccess;
    int resident_pairs = f->num_sms / kFuCluster;      // CTA pairs the device can hold at once
    if (ok) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned) (kFuCluster * f->num_sms));
        cfg.blockDim = dim3(kFuThreads);
        cfg.dynamicSmemBytes = (size_t) smem_bytes;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = kFuCluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, kernel, &cfg) != cudaSuccess || n < 1) { cudaGetLastError(); n = f->num_sms / kFuCluster - 4; }
        resident_pairs = n;
        if (const char *e = getenv("KOALA_FU_CLUSTERS")) n = std::max(1, std::min(n, atoi(e)));
        f->max_clusters = n;
    }
    std::vector<CUtensorMap> maps((size_t) 2 * nseg * kFuMapsPerSeg);
    FuArgs &a = f->args;
    memset(&a, 0, sizeof(a));
    a.nseg = nseg; a.num_m_tiles = mt; a.H = m.H; a.Bp = m.Bp; a.e_ring = m.e_ring; a.counters = f->counters; a.trace = f->trace; a.trace_skip = trace_skip;
    for (int s = 0; s < nseg; s++) {
        FuSeg &sg = a.seg[s];
        if (s == 0) {                                  // encoder: e = relu(feat W_enc^T + b)
            sg.mode = kTcEnc; sg.n_tiles = m.H / kFuLinN; sg.kb_w = kBins / kTcBlockK; sg.parts = 1; sg.step_delta = 1;
            sg.bias0 = m.enc_b;
        } else if (s == nseg - 1) {                    // decoder: mask = sigmoid(h_{L-1}(t) W_dec^T + b)
            sg.mode = kTcDec; sg.n_tiles = kBins / kFuLinN; sg.kb_w = m.H / kTcBlockK; sg.parts = 1; sg.step_delta = -1;
            sg.bias0 = m.dec_b;
        } else {                                       // GRU layer l
            const size_t l = s - 1;
            sg.mode = kTcGru; sg.n_tiles = m.H / kGruUnits; sg.kb_w = m.H / kTcBlockK; sg.parts = 2; sg.step_delta = 0;
            sg.bias0 = m.bih[l]; sg.bias1 = m.bhh[l];
        }
        sg.kb_per_part = sg.kb_w * m.planes;
        sg.n_magic = (unsigned) ((((uint64_t) 1 << 32) + sg.n_tiles - 1) / sg.n_tiles);
        sg.inc = (unsigned) (2 * sg.n_tiles);          // the rows of an m tile are written by both CTAs of every n tile
    }
    // list order inside a period: GRU layer 0, decoder (a step behind), encoder (a step ahead), GRU layers 1..
    int tile = 0, npos = 0;
    auto place = [&](int s) {
        a.pos_seg[npos] = s;
        a.pos_begin[npos++] = tile;
        tile += mt * a.seg[s].n_tiles;
    };
    place(1);
    place(nseg - 1);
    place(0);
    for (int s = 2; s < nseg - 1; s++) place(s);
    a.pos_begin[npos] = tile;
    a.tiles_per_step = tile;
    // small batches are bound by the latency of the recurrence GRU_l(t-1) -> GRU_l(t): run the x part (whose operands exist
    // earlier) first, so that only the h half of the k loop follows the wait.  KOALA_FU_XFIRST=0/1 overrides.
    bool x_first = tile <= resident_pairs;       // (a property of the model, the stream count and the device: results do not depend on KOALA_FU_CLUSTERS)
    if (const char *e = getenv("KOALA_FU_XFIRST")) x_first = e[0] == '1';
    for (int s = 1; s < nseg - 1; s++) a.seg[s].x_first = x_first ? 1 : 0;
    for (int cur = 0; cur < 2 && ok; cur++) {          // cur = parity of the buffers that hold h(t-1)
        const int nxt = cur ^ 1;
        for (int s = 0; s < nseg && ok; s++) {
            CUtensorMap *mp = maps.data() + (size_t) (cur * nseg + s) * kFuMapsPerSeg;
            bool used[kFuMapsPerSeg] = {};
            auto put = [&](int k, const void *base, uint64_t rows, uint64_t cols, uint32_t box_rows, bool f32 = false, bool plain32 = false) {
                used[k] = true;
                ok = ok && encode_2d(fn, &mp[k], base, rows, cols, box_rows, f32, plain32);
            };
            if (s == 0) {
                put(kMapA0, m.feat, (uint64_t) m.tcap * Bp, P * kBins, kTcBlockM);
                put(kMapB0, m.enc_w, H, kBins, kFuLinN / 2);
                put(kMapHn, m.e, (uint64_t) m.e_ring * Bp, P * H, kTcBlockM);     // store map: boxes of 64 bf16 columns x 128 rows, 128B swizzle
            } else if (s == nseg - 1) {
                put(kMapA0, m.hb[nxt] + (L - 1) * P * LBH, Bp, P * H, kTcBlockM);
                put(kMapB0, m.dec_w, kBins, H, kFuLinN / 2);
                put(kMapHn, m.mask, (uint64_t) m.tcap * Bp, kBins, kTcBlockM, true);    // store map: boxes of 32 fp32 columns x 128 rows, 128B swizzle
            } else {
                const size_t l = s - 1;
                if (l == 0) put(kMapA0, m.e, (uint64_t) m.e_ring * Bp, P * H, kTcBlockM);
                else put(kMapA0, m.hb[nxt] + (l - 1) * P * LBH, Bp, P * H, kTcBlockM);
                put(kMapA1, m.hb[cur] + l * P * LBH, Bp, P * H, kTcBlockM);
                put(kMapB0, f->wih_p[l], 3 * H, H, kGruRows / 2);
                put(kMapB1, f->whh_p[l], 3 * H, H, kGruRows / 2);
                put(kMapHp, m.h[cur] + l * LBH, Bp, H, kTcBlockM, true);
                put(kMapHn, m.h[nxt] + l * LBH, Bp, H, kTcBlockM, true);
                put(kMapHb, m.hb[nxt] + l * P * LBH, Bp, P * H, kTcBlockM, false, true);
            }
            for (int k = 0; k < kFuMapsPerSeg; k++)        // unused slots: any valid descriptor (they are only prefetched)
                if (!used[k]) mp[k] = mp[kMapA0];
        }
    }
    ok = ok && cudaMalloc((void **) &f->d_maps, maps.size() * sizeof(CUtensorMap)) == cudaSuccess &&
         cudaMemcpy(f->d_maps, maps.data(), maps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice) == cudaSuccess;
    a.maps = f->d_maps;
    if (!ok) {
        *why = "setting up the fused mask-estimator kernel failed (tensor maps / counters / shared memory size)";
        cudaGetLastError();
        fu_plan_destroy(f);
        return false;
    }
    *out = f;
    return true;
}

// `steps` mask-estimator steps = one launch: hb[cur] / h[cur] hold the state before the first step; features are read from
// slots 0..steps-1 of `feat`, masks written to the same slots of `mask`; the state ends up in the buffers of parity cur ^ (steps & 1)
static int fu_masknet_steps(FuPlan *f, int cur, int steps, cudaStream_t st) {
    FuArgs &a = f->args;
    a.steps = steps;
    a.cur0 = cur;
    a.epoch0 = f->epoch;
    a.total_tiles = (steps + 2) * a.tiles_per_step;       // periods 0 .. steps + 1 (the first holds only encoder tiles, the last only decoder tiles)
    const int clusters = steps * a.tiles_per_step < f->max_clusters ? steps * a.tiles_per_step : f->max_clusters;
    // the epoch only advances with a launch that was accepted: the counters then stand at epoch * (increments per step), which
    // is what the next launch waits for (the caller reports the launch error through cudaGetLastError)
    if (launch_pdl(true, fu_kernel(f->planes), dim3((unsigned) (kFuCluster * clusters)), dim3(kFuThreads), (size_t) fu_smem_bytes(f->planes), st, a) == cudaSuccess)
        f->epoch += steps;
    return 1;
}

}  // namespace koala
