// koala_b200 -- mask estimator, fixed-point path (SPEC.md section 6).
//
// Middle stage of `pv_koala_process` (/root/reference/include/pv_koala.h:65-80) in the reference engine's own numeric style
// (int8 weights x int16 activations -> int32 with `dp2a` mat-vec kernels taabe36/159/174/225,
// table-driven gates taabe84/26/115, saturating stores), batched over the stream dimension and moved to the integer tensor
// cores: tcgen05.mma kind::i8 (SASS UTCIMMA), s32 accumulators in TMEM.
//
// An int16 activation v travels as two byte PLANES, hi = v >> 8 (signed) and lo = v & 255 (unsigned), v = 256 hi + lo; an
// activation matrix is [rows][2 K] bytes (hi plane | lo plane).  Every k-block stages both planes of the activations and the
// weights ONCE and issues two sets of MMAs -- s8 x s8 for the hi plane into TMEM columns 0..127, u8 x s8 for the lo plane into
// columns 128..255 (walking the planes one after the other loaded every weight tile twice) -- and the epilogue combines them,
// acc = (hi << 8) + lo in wrapping int32 arithmetic: exactly the int32 sum the CPU restatement computes, so everything from the
// quantised features to the mask is BIT-EXACT against oracle/koala_oracle.c (mode 2).  The epilogue then does what SPEC section 6
// says with integers only: 64-bit requantisation multiply, bias, sigmoid / tanh by table + linear interpolation, the GRU blend,
// saturation to int16.
//
// One kernel per layer and step (encoder, GRU layer, decoder), one 128-row x 128-column tile per CTA, cta_group::1:
//   warps 0..7  epilogue (TMEM lane quarter = warp % 4, column half = warp / 4: the integer gate math is a long dependent chain
//   with table lookups, so it wants warps), warp 8 TMA producer, warp 9 TMEM allocation + MMA issuer;
//   2 stages of (A hi + A lo + B, each [128 rows][128 B]), 128B-swizzled, K-major.
// GRU tile = 128 streams x 32 units, TMEM columns [n_x | r | z | n_h] x 32: the x part's weight rows are packed [n | r | z | 0]
// and the h part's [0 | r | z | n] (zero rows instead of the bf16 kernel's column-offset trick: this path is built for parity
// first; see DESIGN.md for what it costs).  Two CTAs fit on an SM (2 x 256 TMEM columns, 2 x 103 KB shared memory), so one CTA's
// epilogue overlaps the other's k loop.
#pragma once

#include <math.h>

#include <vector>

#include "engine.h"
#include "tcgen05_common.cuh"

namespace koala {

constexpr int kI8Stages = 2;
constexpr int kI8Tile = 128;                              // rows per CTA, columns per tile, int8 per k-block
constexpr int kI8TileBytes = kI8Tile * 128;
constexpr int kI8StageBytes = 3 * kI8TileBytes;           // activations hi plane, activations lo plane, weights
constexpr int kI8EpiWarps = 8;                            // two per TMEM lane quarter, each taking half of the tile's columns
constexpr int kI8Threads = 32 * (kI8EpiWarps + 2);
constexpr int kI8Units = 32;                              // GRU units per tile
constexpr int kI8SigN = 2048;                             // sigmoid table intervals over [-8, 8)
constexpr int kI8SigBytes = kI8SigN * 4;                  // device copy: entry i = T[i] | (T[i + 1] - T[i]) << 16, one 32-bit lookup per interpolation
constexpr int kI8SmemBytes = 1024 + kI8Stages * kI8StageBytes + kI8SigBytes + 2 * kI8Tile * 4 + 128;
constexpr int kQF = 14, kQE = 12, kQH = 15, kQP = 12;     // Q formats: features, encoder output, state, pre-activations

enum I8Mode : int { kI8Enc = 0, kI8Gru = 1, kI8Dec = 2 };

struct I8Args {
    CUtensorMap a_x, a_h, b_x, b_h;   // activation planes / packed weights of the x and (GRU) h operand part
    int kb_x, kb_h;                   // k-blocks of 128 per plane and part
    int k_x, H;                       // columns of one plane of the x operand; hidden size
    const int32_t *mult, *bias;       // encoder / decoder: [N]; GRU: [4][H] = r, z, n_x, n_h
    const uint32_t *sig;              // [kI8SigN] packed sigmoid table: T[i] | (T[i + 1] - T[i]) << 16
    const uint8_t *h_prev;            // GRU: planes of h(t-1), [rows][2 H]
    uint8_t *out_planes;              // encoder: planes of e; GRU: planes of h(t)
    float *mask;                      // decoder: [rows][256]
};

__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit_1(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// kind::i8 instruction descriptor: D s32, A s8 (hi plane) or u8 (lo plane), B s8, both K-major
__host__ __device__ constexpr uint32_t make_idesc_i8(int M, int N, bool a_signed) {
    return (2u << 4) | ((a_signed ? 1u : 0u) << 7) | (1u << 10) | ((uint32_t) (N >> 3) << 17) | ((uint32_t) (M >> 4) << 24);
}
__device__ __forceinline__ void tmem_ld8_i32(uint32_t taddr, int32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr));
}

// SPEC.md section 6 integer helpers (the same expressions as oracle/koala_oracle.c)
__device__ __forceinline__ int32_t i8_requant(int32_t acc, int32_t mult) { return (int32_t) (((long long) acc * mult + (1ll << 30)) >> 31); }
__device__ __forceinline__ int32_t i8_sig(const uint32_t *t, int32_t p) {
    const int32_t x = min(max(p, -32768), 32767) + 32768;
    const int32_t i = x >> 5, f = x & 31;
    const uint32_t e = t[i];                                   // T[i] and the step to T[i + 1] (0 .. 64)
    return (int32_t) (e & 0xffffu) + (((int32_t) (e >> 16) * f + 16) >> 5);
}
__device__ __forceinline__ int32_t i8_tanh(const uint32_t *t, int32_t a) { return 2 * i8_sig(t, a < -16384 ? -32768 : a > 16383 ? 32767 : 2 * a) - 32768; }
// the 8 combined accumulators of one gate: columns col .. col + 7 of both planes
__device__ __forceinline__ void i8_load_acc(uint32_t lane_base, int col, int32_t (&acc)[8]) {
    int32_t hi[8], lo[8];
    tmem_ld8_i32(lane_base + col, hi);
    tmem_ld8_i32(lane_base + kI8Tile + col, lo);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = (int32_t) (((uint32_t) hi[i] << 8) + (uint32_t) lo[i]);
}
__device__ __forceinline__ void i8_store_planes(uint8_t *row, int H, int col, const int32_t (&v)[8]) {
    uint32_t h[2] = {0, 0}, l[2] = {0, 0};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        h[i >> 2] |= ((uint32_t) (v[i] >> 8) & 255u) << (8 * (i & 3));
        l[i >> 2] |= ((uint32_t) v[i] & 255u) << (8 * (i & 3));
    }
    *reinterpret_cast<uint2 *>(row + col) = make_uint2(h[0], h[1]);
    *reinterpret_cast<uint2 *>(row + H + col) = make_uint2(l[0], l[1]);
}

template <int MODE>
__global__ void __launch_bounds__(kI8Threads) i8_layer_kernel(const __grid_constant__ I8Args args) {
    extern __shared__ uint8_t i8_smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(i8_smem_raw) + 1023) & ~(uintptr_t) 1023);
    uint8_t *tail = smem + kI8Stages * kI8StageBytes;
    uint32_t *s_sig = reinterpret_cast<uint32_t *>(tail);
    int32_t *s_mult = reinterpret_cast<int32_t *>(tail + kI8SigBytes), *s_bias = s_mult + kI8Tile;   // this tile's 128 columns: multipliers and biases
    uint64_t *bars = reinterpret_cast<uint64_t *>(tail + kI8SigBytes + 2 * kI8Tile * 4);
    uint64_t *full_bar = bars, *empty_bar = bars + kI8Stages, *tmem_full = bars + 2 * kI8Stages;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * kI8Stages + 1);

    const int tid = threadIdx.x, warp = tid >> 5;
    const int m0 = blockIdx.x * kI8Tile, tile = blockIdx.y;
    if (tid == 0) {
        for (int s = 0; s < kI8Stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(tmem_full, 1);
        fence_mbar_init();
    }
    if (warp == kI8EpiWarps + 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (warp < kI8EpiWarps) {
        if (MODE != kI8Enc)
            for (int i = tid; i < kI8SigN / 4; i += 32 * kI8EpiWarps) reinterpret_cast<uint4 *>(s_sig)[i] = __ldg(reinterpret_cast<const uint4 *>(args.sig) + i);
        // column c of the tile: output tile * 128 + c (encoder / decoder), or gate c / 32 of unit tile * 32 + c % 32 (GRU: n_x | r | z | n_h,
        // stored [4][H] in the order r, z, n_x, n_h)
        const int gate = tid >> 5, slot = gate == 0 ? 2 : gate == 1 ? 0 : gate == 2 ? 1 : 3;
        const int idx = MODE == kI8Gru ? slot * args.H + tile * kI8Units + (tid & 31) : tile * kI8Tile + tid;
        if (tid < kI8Tile) {
            s_mult[tid] = __ldg(args.mult + idx);
            s_bias[tid] = __ldg(args.bias + idx);
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int total = args.kb_x + args.kb_h;          // k-blocks: x part, then (GRU) h part

    if (warp == kI8EpiWarps) {
        // ===================================================== TMA producer: both activation planes and the weights of a k-block
        if (elect_one()) {
            for (int i = 0; i < total; ++i) {
                const int s = i % kI8Stages, ph = (i / kI8Stages) & 1;
                const bool hp = i >= args.kb_x;
                const int kb = hp ? i - args.kb_x : i, plane_cols = hp ? args.H : args.k_x;
                mbar_wait(&empty_bar[s], ph ^ 1);
                mbar_expect_tx(&full_bar[s], kI8StageBytes);
                uint8_t *sa = smem + s * kI8StageBytes;
                const CUtensorMap *am = hp ? &args.a_h : &args.a_x;
                tma_load_2d_local(am, &full_bar[s], sa, kb * kI8Tile, m0);
                tma_load_2d_local(am, &full_bar[s], sa + kI8TileBytes, plane_cols + kb * kI8Tile, m0);
                tma_load_2d_local(hp ? &args.b_h : &args.b_x, &full_bar[s], sa + 2 * kI8TileBytes, kb * kI8Tile, tile * kI8Tile);
            }
        }
        __syncwarp();
    } else if (warp == kI8EpiWarps + 1) {
        // ===================================================== MMA issuer
        if (elect_one()) {
            const uint64_t adesc0 = make_sw128_desc(smem_u32(smem)), bdesc0 = make_sw128_desc(smem_u32(smem) + 2 * kI8TileBytes);
            constexpr uint64_t lo_off = (uint64_t) (kI8TileBytes >> 4);
            constexpr uint32_t idesc_hi = make_idesc_i8(kI8Tile, kI8Tile, true), idesc_lo = make_idesc_i8(kI8Tile, kI8Tile, false);
            for (int i = 0; i < total; ++i) {
                const int s = i % kI8Stages, ph = (i / kI8Stages) & 1;
                mbar_wait(&full_bar[s], ph);
                tc_fence_after();
                const uint64_t so = (uint64_t) ((s * kI8StageBytes) >> 4);
#pragma unroll
                for (int k = 0; k < 4; ++k)         // 32 int8 = 32 bytes per instruction; signed high bytes -> columns 0..127
                    umma_i8(tmem_base, adesc0 + so + 2 * k, bdesc0 + so + 2 * k, idesc_hi, (i | k) != 0 ? 1u : 0u);
#pragma unroll
                for (int k = 0; k < 4; ++k)         // unsigned low bytes -> columns 128..255
                    umma_i8(tmem_base + kI8Tile, adesc0 + lo_off + so + 2 * k, bdesc0 + so + 2 * k, idesc_lo, (i | k) != 0 ? 1u : 0u);
                umma_commit_1(&empty_bar[s]);
            }
            umma_commit_1(tmem_full);
        }
        __syncwarp();
    } else {
        // ===================================================== epilogue: thread = stream row = TMEM lane
        const int quarter = warp & 3, half = warp >> 2;
        const size_t row = (size_t) m0 + quarter * 32 + (tid & 31);
        const uint32_t lane_base = tmem_base + ((uint32_t) (quarter * 32) << 16);
        const int H = args.H;
        mbar_wait(tmem_full, 0);
        tc_fence_after();
        if (MODE == kI8Enc) {
            uint8_t *out_row = args.out_planes + row * 2 * H;
            for (int c = half * (kI8Tile / 2); c < (half + 1) * (kI8Tile / 2); c += 8) {
                int32_t acc[8];
                i8_load_acc(lane_base, c, acc);
                const int n = tile * kI8Tile + c;
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    acc[i] = min(max(i8_requant(acc[i], s_mult[c + i]) + s_bias[c + i], 0), 32767);   // saturating relu, Q12
                i8_store_planes(out_row, H, n, acc);
            }
        } else if (MODE == kI8Dec) {
            float *mask_row = args.mask + row * kBins;
            for (int c = half * (kI8Tile / 2); c < (half + 1) * (kI8Tile / 2); c += 8) {
                int32_t acc[8];
                float m[8];
                i8_load_acc(lane_base, c, acc);
                const int n = tile * kI8Tile + c;
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    m[i] = (float) i8_sig(s_sig, i8_requant(acc[i], s_mult[c + i]) + s_bias[c + i]) * (1.0f / 32768.0f);
                *reinterpret_cast<float4 *>(mask_row + n) = make_float4(m[0], m[1], m[2], m[3]);
                *reinterpret_cast<float4 *>(mask_row + n + 4) = make_float4(m[4], m[5], m[6], m[7]);
            }
        } else {
            const uint8_t *prev_row = args.h_prev + row * 2 * H;
            uint8_t *out_row = args.out_planes + row * 2 * H;
            for (int cu = half * (kI8Units / 2); cu < (half + 1) * (kI8Units / 2); cu += 8) {
                int32_t anx[8], ar[8], az[8], anh[8], out[8];
                i8_load_acc(lane_base, cu, anx);
                i8_load_acc(lane_base, kI8Units + cu, ar);
                i8_load_acc(lane_base, 2 * kI8Units + cu, az);
                i8_load_acc(lane_base, 3 * kI8Units + cu, anh);
                const int u = tile * kI8Units + cu;
                const uint2 ph = *reinterpret_cast<const uint2 *>(prev_row + u), pl = *reinterpret_cast<const uint2 *>(prev_row + H + u);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const uint32_t hw = i < 4 ? ph.x : ph.y, lw = i < 4 ? pl.x : pl.y;
                    const int32_t hq = (int32_t) (int8_t) (hw >> (8 * (i & 3))) * 256 + (int32_t) ((lw >> (8 * (i & 3))) & 255u);
                    const int32_t r = i8_sig(s_sig, i8_requant(ar[i], s_mult[kI8Units + cu + i]) + s_bias[kI8Units + cu + i]);
                    const int32_t z = i8_sig(s_sig, i8_requant(az[i], s_mult[2 * kI8Units + cu + i]) + s_bias[2 * kI8Units + cu + i]);
                    const int32_t pnx = i8_requant(anx[i], s_mult[cu + i]) + s_bias[cu + i];
                    const int32_t pnh = i8_requant(anh[i], s_mult[3 * kI8Units + cu + i]) + s_bias[3 * kI8Units + cu + i];
                    const int32_t a = pnx + (int32_t) (((long long) r * pnh + (1 << 14)) >> 15);
                    const int32_t nn = i8_tanh(s_sig, a);
                    const int32_t hn = nn + (int32_t) (((long long) z * (hq - nn) + (1 << 14)) >> 15);
                    out[i] = min(max(hn, -32767), 32767);
                }
                i8_store_planes(out_row, H, u, out);
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == kI8EpiWarps + 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem_base) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------------------
// host side: quantisation of the model (SPEC.md section 6), packed weights, tensor maps, the per-step launches
struct QModelHost {
    int H = 0, L = 0;
    std::vector<int8_t> enc_q, dec_q;                  // [H][256], [256][H]
    std::vector<int32_t> enc_m, enc_b, dec_m, dec_b;
    std::vector<std::vector<int8_t>> wx_p, wh_p;       // per layer, packed per 32-unit tile: [(H/32) * 128][H]
    std::vector<std::vector<int32_t>> g_m, g_b;        // per layer [4][H]: r, z, n_x, n_h
    std::vector<int16_t> sig;
};

static inline float i8_bf16_to_f32(uint16_t b) {
    const uint32_t u = (uint32_t) b << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}
static inline float i8_row_scale(float maxabs) { return maxabs > 0.0f ? maxabs / 127.0f : 1.0f; }
static inline int8_t i8_quant_w(float w, float s) {
    float q = rintf(w / s);
    q = q > 127.0f ? 127.0f : q < -127.0f ? -127.0f : q;
    return (int8_t) q;
}
static inline int32_t i8_quant_mult(float s, int q_in) {
    const long long v = llrint((double) s * ldexp(1.0, kQP - q_in + 31));
    return (int32_t) (v > 2147483647ll ? 2147483647ll : v);
}
static inline int32_t i8_quant_bias(float b) { return (int32_t) lrintf(b * 4096.0f); }
static inline float i8_row_maxabs(const uint16_t *row, int K, float c) {
    float m = 0.0f;
    for (int k = 0; k < K; k++) m = std::max(m, fabsf(c * i8_bf16_to_f32(row[k])));
    return m;
}

static void quantize_model(const ModelHost &m, QModelHost *q) {
    const int H = m.hidden, L = m.layers;
    q->H = H; q->L = L;
    q->enc_q.resize((size_t) H * kBins); q->enc_m.resize(H); q->enc_b.resize(H);
    for (int n = 0; n < H; n++) {
        const uint16_t *row = m.enc_w.data() + (size_t) n * kBins;
        const float s = i8_row_scale(i8_row_maxabs(row, kBins, 1.0f));
        for (int k = 0; k < kBins; k++) q->enc_q[(size_t) n * kBins + k] = i8_quant_w(i8_bf16_to_f32(row[k]), s);
        q->enc_m[n] = i8_quant_mult(s, kQF);
        q->enc_b[n] = i8_quant_bias(m.enc_b[n]);
    }
    q->wx_p.resize(L); q->wh_p.resize(L); q->g_m.resize(L); q->g_b.resize(L);
    for (int l = 0; l < L; l++) {
        const float c = l == 0 ? 8.0f : 1.0f;      // 2^(QH - Q of the layer input)
        const size_t prow = (size_t) (H / kI8Units) * kI8Tile;
        q->wx_p[l].assign(prow * H, 0); q->wh_p[l].assign(prow * H, 0);
        q->g_m[l].resize(4 * (size_t) H); q->g_b[l].resize(4 * (size_t) H);
        auto packed = [&](std::vector<int8_t> &w, int j, int slot) { return w.data() + ((size_t) (j / kI8Units) * kI8Tile + slot * kI8Units + j % kI8Units) * H; };
        for (int j = 0; j < H; j++) {
            for (int g = 0; g < 2; g++) {
                const uint16_t *ri = m.wih[l].data() + (size_t) (g * H + j) * H, *rh = m.whh[l].data() + (size_t) (g * H + j) * H;
                const float s = i8_row_scale(std::max(i8_row_maxabs(ri, H, c), i8_row_maxabs(rh, H, 1.0f)));
                int8_t *px = packed(q->wx_p[l], j, 1 + g), *phh = packed(q->wh_p[l], j, 1 + g);
                for (int k = 0; k < H; k++) {
                    px[k] = i8_quant_w(c * i8_bf16_to_f32(ri[k]), s);
                    phh[k] = i8_quant_w(i8_bf16_to_f32(rh[k]), s);
                }
                q->g_m[l][(size_t) g * H + j] = i8_quant_mult(s, kQH);
                q->g_b[l][(size_t) g * H + j] = i8_quant_bias(m.bih[l][g * H + j] + m.bhh[l][g * H + j]);
            }
            const uint16_t *ri = m.wih[l].data() + (size_t) (2 * H + j) * H, *rh = m.whh[l].data() + (size_t) (2 * H + j) * H;
            const float sx = i8_row_scale(i8_row_maxabs(ri, H, c)), sh = i8_row_scale(i8_row_maxabs(rh, H, 1.0f));
            int8_t *px = packed(q->wx_p[l], j, 0), *phh = packed(q->wh_p[l], j, 3);
            for (int k = 0; k < H; k++) {
                px[k] = i8_quant_w(c * i8_bf16_to_f32(ri[k]), sx);
                phh[k] = i8_quant_w(i8_bf16_to_f32(rh[k]), sh);
            }
            q->g_m[l][2 * (size_t) H + j] = i8_quant_mult(sx, kQH);
            q->g_b[l][2 * (size_t) H + j] = i8_quant_bias(m.bih[l][2 * H + j]);
            q->g_m[l][3 * (size_t) H + j] = i8_quant_mult(sh, kQH);
            q->g_b[l][3 * (size_t) H + j] = i8_quant_bias(m.bhh[l][2 * H + j]);
        }
    }
    q->dec_q.resize((size_t) kBins * H); q->dec_m.resize(kBins); q->dec_b.resize(kBins);
    for (int n = 0; n < kBins; n++) {
        const uint16_t *row = m.dec_w.data() + (size_t) n * H;
        const float s = i8_row_scale(i8_row_maxabs(row, H, 1.0f));
        for (int k = 0; k < H; k++) q->dec_q[(size_t) n * H + k] = i8_quant_w(i8_bf16_to_f32(row[k]), s);
        q->dec_m[n] = i8_quant_mult(s, kQH);
        q->dec_b[n] = i8_quant_bias(m.dec_b[n]);
    }
    q->sig.resize(kI8SigN + 1);
    for (int i = 0; i <= kI8SigN; i++) q->sig[i] = (int16_t) lrint(32768.0 / (1.0 + exp(-(double) (i - kI8SigN / 2) / 128.0)));
}

struct I8Plan {
    int H = 0, L = 0, Bp = 0;
    std::vector<void *> allocs;
    uint8_t *featp = nullptr, *ep = nullptr, *hp[2] = {};     // planes: features [Bp][512], encoder output [Bp][2H], state [L][Bp][2H] x 2 parities
    I8Args enc, gru[2][kMaxLayers], dec[2];                   // [parity of the buffers holding h(t-1)]
};

static void i8_plan_destroy(I8Plan *f) {
    if (!f) return;
    for (void *a : f->allocs) cudaFree(a);
    delete f;
}

// [rows][cols] bytes, box = 128 bytes x 128 rows, 128B swizzle
static bool encode_2d_u8(EncodeTiledFn fn, CUtensorMap *map, const void *base, uint64_t rows, uint64_t cols) {
    const cuuint64_t dims[2] = {cols, rows};
    const cuuint64_t strides[1] = {cols};
    const cuuint32_t box[2] = {128, 128};
    const cuuint32_t estr[2] = {1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static bool i8_plan_create(const ModelHost &model, int Bp, float *mask, I8Plan **out, std::string *why) {
    if (model.hidden % kI8Tile != 0) {
        *why = "hidden size must be a multiple of 128 for the fixed-point path";
        return false;
    }
    void *fnp = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &qres) != cudaSuccess || !fnp || qres != cudaDriverEntryPointSuccess) {
        *why = "cuTensorMapEncodeTiled not available from the driver";
        return false;
    }
    EncodeTiledFn fn = (EncodeTiledFn) fnp;
    QModelHost q;
    quantize_model(model, &q);
    I8Plan *f = new I8Plan();
    const size_t H = model.hidden, L = model.layers;
    f->H = (int) H; f->L = (int) L; f->Bp = Bp;
    bool ok = true;
    auto put = [&](const void *src, size_t bytes) -> void * {
        void *d = nullptr;
        if (!ok || cudaMalloc(&d, bytes) != cudaSuccess) { ok = false; return nullptr; }
        f->allocs.push_back(d);
        ok = (src ? cudaMemcpy(d, src, bytes, cudaMemcpyHostToDevice) : cudaMemset(d, 0, bytes)) == cudaSuccess;
        return d;
    };
    const int8_t *enc_q = (const int8_t *) put(q.enc_q.data(), q.enc_q.size());
    const int8_t *dec_q = (const int8_t *) put(q.dec_q.data(), q.dec_q.size());
    const int32_t *enc_m = (const int32_t *) put(q.enc_m.data(), 4 * q.enc_m.size()), *enc_b = (const int32_t *) put(q.enc_b.data(), 4 * q.enc_b.size());
    const int32_t *dec_m = (const int32_t *) put(q.dec_m.data(), 4 * q.dec_m.size()), *dec_b = (const int32_t *) put(q.dec_b.data(), 4 * q.dec_b.size());
    std::vector<uint32_t> packed_sig(kI8SigN);
    for (int i = 0; i < kI8SigN; i++) packed_sig[i] = (uint32_t) (uint16_t) q.sig[i] | ((uint32_t) (q.sig[i + 1] - q.sig[i]) << 16);
    const uint32_t *sig = (const uint32_t *) put(packed_sig.data(), 4 * packed_sig.size());
    const int8_t *wx[kMaxLayers] = {}, *wh[kMaxLayers] = {};
    const int32_t *gm[kMaxLayers] = {}, *gb[kMaxLayers] = {};
    for (size_t l = 0; l < L; l++) {
        wx[l] = (const int8_t *) put(q.wx_p[l].data(), q.wx_p[l].size());
        wh[l] = (const int8_t *) put(q.wh_p[l].data(), q.wh_p[l].size());
        gm[l] = (const int32_t *) put(q.g_m[l].data(), 4 * q.g_m[l].size());
        gb[l] = (const int32_t *) put(q.g_b[l].data(), 4 * q.g_b[l].size());
    }
    f->featp = (uint8_t *) put(nullptr, (size_t) Bp * 2 * kBins);
    f->ep = (uint8_t *) put(nullptr, (size_t) Bp * 2 * H);
    for (int i = 0; i < 2; i++) f->hp[i] = (uint8_t *) put(nullptr, L * (size_t) Bp * 2 * H);
    const size_t layer_bytes = (size_t) Bp * 2 * H, prow = (H / kI8Units) * kI8Tile;
    if (ok) {
        memset(&f->enc, 0, sizeof(I8Args));
        I8Args &e = f->enc;
        ok = ok && encode_2d_u8(fn, &e.a_x, f->featp, Bp, 2 * kBins) && encode_2d_u8(fn, &e.b_x, enc_q, H, kBins);
        e.a_h = e.a_x; e.b_h = e.b_x;
        e.kb_x = kBins / kI8Tile; e.kb_h = 0; e.k_x = kBins; e.H = (int) H; e.mult = enc_m; e.bias = enc_b; e.sig = sig; e.out_planes = f->ep;
        for (int cur = 0; cur < 2 && ok; cur++) {
            const int nxt = cur ^ 1;
            for (size_t l = 0; l < L && ok; l++) {
                I8Args &g = f->gru[cur][l];
                memset(&g, 0, sizeof(I8Args));
                const uint8_t *x = l == 0 ? f->ep : f->hp[nxt] + (l - 1) * layer_bytes;
                ok = ok && encode_2d_u8(fn, &g.a_x, x, Bp, 2 * H) && encode_2d_u8(fn, &g.a_h, f->hp[cur] + l * layer_bytes, Bp, 2 * H) &&
                     encode_2d_u8(fn, &g.b_x, wx[l], prow, H) && encode_2d_u8(fn, &g.b_h, wh[l], prow, H);
                g.kb_x = g.kb_h = (int) H / kI8Tile; g.k_x = (int) H; g.H = (int) H; g.mult = gm[l]; g.bias = gb[l]; g.sig = sig;
                g.h_prev = f->hp[cur] + l * layer_bytes; g.out_planes = f->hp[nxt] + l * layer_bytes;
            }
            I8Args &d = f->dec[cur];
            memset(&d, 0, sizeof(I8Args));
            ok = ok && encode_2d_u8(fn, &d.a_x, f->hp[nxt] + (L - 1) * layer_bytes, Bp, 2 * H) && encode_2d_u8(fn, &d.b_x, dec_q, kBins, H);
            d.a_h = d.a_x; d.b_h = d.b_x;
            d.kb_x = (int) H / kI8Tile; d.kb_h = 0; d.k_x = (int) H; d.H = (int) H; d.mult = dec_m; d.bias = dec_b; d.sig = sig; d.mask = mask;
        }
    }
    ok = ok && cudaFuncSetAttribute(i8_layer_kernel<kI8Enc>, cudaFuncAttributeMaxDynamicSharedMemorySize, kI8SmemBytes) == cudaSuccess &&
         cudaFuncSetAttribute(i8_layer_kernel<kI8Gru>, cudaFuncAttributeMaxDynamicSharedMemorySize, kI8SmemBytes) == cudaSuccess &&
         cudaFuncSetAttribute(i8_layer_kernel<kI8Dec>, cudaFuncAttributeMaxDynamicSharedMemorySize, kI8SmemBytes) == cudaSuccess;
    if (!ok) {
        *why = "setting up the fixed-point mask-estimator kernels failed (allocation / tensor maps / shared memory size)";
        cudaGetLastError();
        i8_plan_destroy(f);
        return false;
    }
    *out = f;
    return true;
}

// one mask-estimator step: encoder, GRU layers, decoder (2 + L launches); `cur` = parity of the planes that hold h(t-1)
static int i8_masknet_step(I8Plan *f, int cur, cudaStream_t st, KernelProfiler *prof) {
    const dim3 block(kI8Threads);
    const unsigned mt = (unsigned) (f->Bp / kI8Tile);
    if (prof) prof->begin(kKernEnc, st);
    i8_layer_kernel<kI8Enc><<<dim3(mt, f->H / kI8Tile), block, kI8SmemBytes, st>>>(f->enc);
    if (prof) prof->end(st);
    for (int l = 0; l < f->L; l++) {
        if (prof) prof->begin(kKernGru, st);
        i8_layer_kernel<kI8Gru><<<dim3(mt, f->H / kI8Units), block, kI8SmemBytes, st>>>(f->gru[cur][l]);
        if (prof) prof->end(st);
    }
    if (prof) prof->begin(kKernDec, st);
    i8_layer_kernel<kI8Dec><<<dim3(mt, kBins / kI8Tile), block, kI8SmemBytes, st>>>(f->dec[cur]);
    if (prof) prof->end(st);
    return 2 + f->L;
}

}  // namespace koala
