// koala_b200 -- mask estimator, fp32 path (BASELINE.json configs[1]: "fp32 mask path").
//
// Middle stage of `pv_koala_process` (/root/reference/include/pv_koala.h:65-80).  The reference's counterparts are its
// int8 x int16 mat-vec kernels and LUT gate kernels (taabe36/159/174/225, taabe84/26/115), one
// stream at a time; here the batch of streams is the GEMM M dimension.  Operands: activations fp32, weights bf16 in HBM
// (exactly representable, widened on load), fp32 FMA accumulation, gates/state fp32 and fused into the GEMM epilogue.
#pragma once

#include "koala_common.cuh"

namespace koala {

constexpr int kF32Bm = 64;     // streams per CTA
constexpr int kF32Bk = 64;     // k-tile: 4 independent 16-byte loads per thread and operand are in flight per tile
constexpr int kF32Pad = 4;

enum Act : int { kActRelu = 0, kActSigmoid = 1 };

__device__ __forceinline__ void bf16x8_to_f32(const uint4 &u, float (&f)[8]) {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        f[2 * i] = __uint_as_float(w[i] << 16);
        f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
}

// These GEMMs are small (256 streams: 16-128 CTAs) and their k loops are what a CTA spends its time in, so the loop is
// software-pipelined through registers: the global loads of k-tile i + 1 are issued before the FMAs of k-tile i, and a tile is
// wide enough (64) for those FMAs to cover the load latency.  (With 16-wide tiles and no prefetch every iteration paid a full
// global round trip: 36 us for the decoder, 51 us per GRU layer at 256 streams.)
struct ATile { float4 v[kF32Bk / 16]; };      // [64 rows][64 k] fp32: thread t holds row t / 4, k = (t % 4) * 4 + 16 i .. + 3
struct WTile { uint4 v[2]; };                 // up to [64 rows][64 k] bf16: chunk c = t + 256 i holds row c / 8, k = (c % 8) * 8 .. + 7

__device__ __forceinline__ void fetch_a(ATile &r, const float *__restrict__ A, int lda, int m0, int k0, int tid) {
    const float *p = A + (size_t) (m0 + (tid >> 2)) * lda + k0 + (tid & 3) * 4;
#pragma unroll
    for (int i = 0; i < kF32Bk / 16; ++i) r.v[i] = *reinterpret_cast<const float4 *>(p + 16 * i);
}
// ... transposed into As[k][row]
__device__ __forceinline__ void stash_a(float (*As)[kF32Bm + kF32Pad], const ATile &r, int tid) {
    const int row = tid >> 2;
#pragma unroll
    for (int i = 0; i < kF32Bk / 16; ++i) {
        const int kc = (tid & 3) * 4 + 16 * i;
        As[kc + 0][row] = r.v[i].x; As[kc + 1][row] = r.v[i].y; As[kc + 2][row] = r.v[i].z; As[kc + 3][row] = r.v[i].w;
    }
}
// weight rows are addressed through `row_ptr(row)`: pointer to W[that output row][0]
template <int ROWS, typename RowPtr>
__device__ __forceinline__ void fetch_w(WTile &r, RowPtr row_ptr, int k0, int tid) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int c = tid + 256 * i;
        if (c < ROWS * 8) r.v[i] = *reinterpret_cast<const uint4 *>(row_ptr(c >> 3) + k0 + (c & 7) * 8);
    }
}
template <int ROWS, int LD>
__device__ __forceinline__ void stash_w(float (*Ws)[LD], const WTile &r, int tid) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int c = tid + 256 * i;
        if (c < ROWS * 8) {
            float f[8];
            bf16x8_to_f32(r.v[i], f);
#pragma unroll
            for (int j = 0; j < 8; ++j) Ws[(c & 7) * 8 + j][c >> 3] = f[j];
        }
    }
}

// out[m][n] = act(bias[n] + sum_k A[m][k] W[n][k]);  grid = (M/64, N/16), block = 256 (tx = column, ty = 16 x 4 rows).
// The tile is only 16 outputs wide (the GRU kernel's shape): at 256 streams a 64-wide tile left the decoder with 16 CTAs.
constexpr int kF32LinN = 16;
template <int ACT>
__global__ void __launch_bounds__(256)
linear_fp32_kernel(const float *__restrict__ A, const __nv_bfloat16 *__restrict__ W, const float *__restrict__ bias,
                   float *__restrict__ out, int K, int N) {
    __shared__ __align__(16) float As[kF32Bk][kF32Bm + kF32Pad];
    __shared__ __align__(16) float Ws[kF32Bk][kF32LinN + kF32Pad];
    pdl_launch_dependents();
    pdl_wait();
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.x * kF32Bm, n0 = blockIdx.y * kF32LinN;
    auto w_row = [&](int row) { return W + (size_t) (n0 + row) * K; };
    float acc[4] = {};
    ATile ra;
    WTile rw;
    fetch_a(ra, A, K, m0, 0, tid);
    fetch_w<kF32LinN>(rw, w_row, 0, tid);
    for (int k0 = 0; k0 < K; k0 += kF32Bk) {
        stash_a(As, ra, tid);
        stash_w<kF32LinN, kF32LinN + kF32Pad>(Ws, rw, tid);
        __syncthreads();
        if (k0 + kF32Bk < K) {       // next tile's loads fly while this tile is multiplied
            fetch_a(ra, A, K, m0, k0 + kF32Bk, tid);
            fetch_w<kF32LinN>(rw, w_row, k0 + kF32Bk, tid);
        }
#pragma unroll 16
        for (int kk = 0; kk < kF32Bk; ++kk) {
            const float4 a = *reinterpret_cast<const float4 *>(&As[kk][ty * 4]);
            const float w = Ws[kk][tx];
            acc[0] = fmaf(a.x, w, acc[0]);
            acc[1] = fmaf(a.y, w, acc[1]);
            acc[2] = fmaf(a.z, w, acc[2]);
            acc[3] = fmaf(a.w, w, acc[3]);
        }
        __syncthreads();
    }
    const int n = n0 + tx;
    const float bn = bias[n];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const float v = acc[r] + bn;
        out[(size_t) (m0 + ty * 4 + r) * N + n] = ACT == kActRelu ? fmaxf(v, 0.0f) : sigmoid_f(v);
    }
}

// One GRU layer step for a tile of 64 streams x 16 hidden units (PyTorch gate order r | z | n):
//   r = sig(Wir x + bir + Whr h + bhr), z = sig(Wiz x + biz + Whz h + bhz), n = tanh(Win x + bin + r (Whn h + bhn)),
//   h' = (1 - z) n + z h.
// grid = (M/64, H/16), block = 256 (tx = unit, ty = 16 x 4 rows).  h_prev and h_next are different buffers (ping-pong):
// other CTAs still read h_prev rows as their GEMM operand while this one writes its 16 units of h_next.
__global__ void __launch_bounds__(256)
gru_fp32_kernel(const float *__restrict__ x, const float *__restrict__ h_prev, float *__restrict__ h_next,
                const __nv_bfloat16 *__restrict__ Wih, const __nv_bfloat16 *__restrict__ Whh,
                const float *__restrict__ bih, const float *__restrict__ bhh, int H) {
    __shared__ __align__(16) float As[kF32Bk][kF32Bm + kF32Pad];
    __shared__ __align__(16) float Ws[kF32Bk][48 + kF32Pad];
    pdl_launch_dependents();
    pdl_wait();
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.x * kF32Bm, u0 = blockIdx.y * 16;
    float ar[4] = {}, az[4] = {}, anx[4] = {}, anh[4] = {};
    // the two operand parts (x with W_ih, h(t-1) with W_hh) form one pipelined loop of 2 H / 64 k-tiles
    const int tiles_per_part = H / kF32Bk, tiles = 2 * tiles_per_part;
    auto fetch = [&](ATile &ra, WTile &rw, int t) {
        const int part = t >= tiles_per_part, k0 = (part ? t - tiles_per_part : t) * kF32Bk;
        const __nv_bfloat16 *W = part ? Whh : Wih;
        fetch_a(ra, part ? h_prev : x, H, m0, k0, tid);
        fetch_w<48>(rw, [&](int row) { return W + (size_t) ((row >> 4) * H + u0 + (row & 15)) * H; }, k0, tid);   // 3 gates x 16 units
    };
    ATile ra;
    WTile rw;
    fetch(ra, rw, 0);
    for (int t = 0; t < tiles; ++t) {
        stash_a(As, ra, tid);
        stash_w<48, 48 + kF32Pad>(Ws, rw, tid);
        __syncthreads();
        if (t + 1 < tiles) fetch(ra, rw, t + 1);
        const bool xpart = t < tiles_per_part;
#pragma unroll 16
        for (int kk = 0; kk < kF32Bk; ++kk) {
            const float4 a = *reinterpret_cast<const float4 *>(&As[kk][ty * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w};
            const float wr = Ws[kk][tx], wz = Ws[kk][16 + tx], wn = Ws[kk][32 + tx];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                ar[r] = fmaf(av[r], wr, ar[r]);
                az[r] = fmaf(av[r], wz, az[r]);
                if (xpart) anx[r] = fmaf(av[r], wn, anx[r]);
                else anh[r] = fmaf(av[r], wn, anh[r]);
            }
        }
        __syncthreads();
    }
    const int u = u0 + tx;
    const float br = bih[u] + bhh[u], bz = bih[H + u] + bhh[H + u], bnx = bih[2 * H + u], bnh = bhh[2 * H + u];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const size_t idx = (size_t) (m0 + ty * 4 + r) * H + u;
        const float rg = sigmoid_f(ar[r] + br);
        const float zg = sigmoid_f(az[r] + bz);
        const float ng = tanh_f(anx[r] + bnx + rg * (anh[r] + bnh));
        h_next[idx] = (1.0f - zg) * ng + zg * h_prev[idx];
    }
}

}  // namespace koala
