// koala_b200 -- analysis (int16 -> STFT -> features) and synthesis (mask apply -> iSTFT -> overlap-add -> int16).
//
// These are the first and last stage of `pv_koala_process` (/root/reference/include/pv_koala.h:65-80; stage names from
// BASELINE.json north_star; the reference's own versions exist only as sm_61 SASS, SURVEY.md section 2.1 taabe228/151/7).
// One warp owns one stream-frame at a time: a 512-point real FFT is done as a 256-point complex FFT held 8 points per
// lane, 3 radix-2 stages in registers and 5 by warp shuffles, followed by the real-FFT split (also by shuffles).
// HBM-bound stages by bytes, LSU-bound in practice (shuffles): window and twiddles live in registers for all streams a
// warp walks, PCM moves with 16-byte vector loads, outputs leave as full 128-byte lines.
#pragma once

#include "koala_common.cuh"

namespace koala {

constexpr int kStftWarps = 4;         // warps (= streams in flight) per CTA
constexpr int kStftCtasPerSm = 7;     // 28 resident warps per SM: at most 72 registers per thread

template <typename FeatT> __device__ __forceinline__ void store_feat8(FeatT *dst, const float (&f)[8]);
template <> __device__ __forceinline__ void store_feat8<float>(float *dst, const float (&f)[8]) {
    reinterpret_cast<float4 *>(dst)[0] = make_float4(f[0], f[1], f[2], f[3]);
    reinterpret_cast<float4 *>(dst)[1] = make_float4(f[4], f[5], f[6], f[7]);
}
template <> __device__ __forceinline__ void store_feat8<__nv_bfloat16>(__nv_bfloat16 *dst, const float (&f)[8]) {
    __nv_bfloat162 p0 = __floats2bfloat162_rn(f[0], f[1]), p1 = __floats2bfloat162_rn(f[2], f[3]);
    __nv_bfloat162 p2 = __floats2bfloat162_rn(f[4], f[5]), p3 = __floats2bfloat162_rn(f[6], f[7]);
    uint4 u;
    u.x = *reinterpret_cast<uint32_t *>(&p0); u.y = *reinterpret_cast<uint32_t *>(&p1);
    u.z = *reinterpret_cast<uint32_t *>(&p2); u.w = *reinterpret_cast<uint32_t *>(&p3);
    *reinterpret_cast<uint4 *>(dst) = u;
}

// Launch: block = 128, grid = ceil(n / (4 * streams per warp)) with streams per warp chosen by the engine so that the whole
// grid is resident at 7 CTAs per SM (8192 streams: 2 per warp, 1024 CTAs); warp w of CTA b walks streams
// s = b * 4 + w, += gridDim * 4.
// spec: [n][512] fp32 packed (Re, Im of bins 0..255; the Im slot of bin 0 carries Re X[256]).
template <typename FeatT>
__global__ void __launch_bounds__(kStftWarps * 32, kStftCtasPerSm)
frontend_kernel(PcmView v, int n_streams, int16_t *__restrict__ tail, float *__restrict__ spec,
                FeatT *__restrict__ feat, const float2 *__restrict__ lane_tab) {
    __shared__ __align__(16) int16_t s_frame[kStftWarps][kNfft];
    __shared__ __align__(16) float2 s_tab[kLaneTabSmemRows * 32];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    pdl_launch_dependents();
    FftLane c;
    load_fft_lane(c, lane_tab, s_tab, lane);   // constants: may be read before the previous kernel has finished
    pdl_wait();
    const uint32_t *fw = reinterpret_cast<const uint32_t *>(s_frame[warp]);
    const int a = rev5(lane);
    const int src_a = rev5((32 - a) & 31), src_b = 31 - lane;

    for (int s = blockIdx.x * kStftWarps + warp; s < n_streams; s += gridDim.x * kStftWarps) {
        // frame = [previous input frame | this input frame], 1024 bytes: lanes 0-15 fetch 32 B of the tail each,
        // lanes 16-31 32 B of the new samples (two 16-byte vector loads per lane)
        int16_t *tail_s = tail + (size_t) s * kFrame;
        const int16_t *src = lane < 16 ? tail_s + lane * 16
                                       : v.in + (size_t) s * v.stride + (size_t) v.t * kFrame + (lane - 16) * 16;
        const uint4 d0 = reinterpret_cast<const uint4 *>(src)[0];
        const uint4 d1 = reinterpret_cast<const uint4 *>(src)[1];
        __syncwarp();   // previous iteration's reads of s_frame are done
        reinterpret_cast<uint4 *>(&s_frame[warp][lane * 16])[0] = d0;
        reinterpret_cast<uint4 *>(&s_frame[warp][lane * 16])[1] = d1;
        if (lane >= 16) {   // state: tail <- this frame (issued after the loads above have returned for the whole warp)
            reinterpret_cast<uint4 *>(tail_s + (lane - 16) * 16)[0] = d0;
            reinterpret_cast<uint4 *>(tail_s + (lane - 16) * 16)[1] = d1;
        }
        __syncwarp();

        cpx z[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint32_t u = fw[lane + 32 * j];
            const float2 w = c.win(j);
            z[j] = cpx{w.x * (float) (int16_t) (u & 0xffffu), w.y * (float) (int16_t) (u >> 16)};
        }
        warp_fft256_dif(z, c, lane);

        // real-FFT split: X[k] = E + W512^k O,  E = (Z[k] + conj Z[256-k]) / 2,  O = (Z[k] - conj Z[256-k]) / 2i
        cpx X[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int b = rev3c(j);
            const cpx zk = z[j];
            const cpx zp = shfl_c(z[partner_reg(j)], j == 0 ? src_a : src_b);
            const cpx E = {0.5f * (zk.x + zp.x), 0.5f * (zk.y - zp.y)};
            const cpx O = {0.5f * (zk.y + zp.y), -0.5f * (zk.x - zp.x)};
            const float2 w = c.tp(b);
            cpx x = cadd(E, cmul(O, cpx{w.x, w.y}));
            if (j == 0 && lane == 0) x = cpx{zk.x + zk.y, zk.x - zk.y};   // (X[0], X[256]), both real
            X[b] = x;
        }
        float f[8];
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const float im = (b == 0 && lane == 0) ? 0.0f : X[b].y;
            const float p = (X[b].x * X[b].x + im * im) * kFeatPowerScale;
            f[b] = kFeatGain * __logf(p + kFeatEps) + kFeatBias;
        }
        float4 *sp = reinterpret_cast<float4 *>(spec + (size_t) s * kNfft + 16 * a);
#pragma unroll
        for (int q = 0; q < 4; ++q) sp[q] = make_float4(X[2 * q].x, X[2 * q].y, X[2 * q + 1].x, X[2 * q + 1].y);
        store_feat8<FeatT>(feat + (size_t) s * kBins + 8 * a, f);
    }
}

// Same launch shape as frontend_kernel.  mask: [n][256] fp32.  ola: [n][256] fp32 state.
__global__ void __launch_bounds__(kStftWarps * 32, kStftCtasPerSm)
backend_kernel(PcmView v, int n_streams, const float *__restrict__ spec, const float *__restrict__ mask,
               float *__restrict__ ola, const float2 *__restrict__ lane_tab) {
    __shared__ __align__(16) float2 s_tab[kLaneTabSmemRows * 32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    pdl_launch_dependents();
    FftLane c;
    load_fft_lane(c, lane_tab, s_tab, lane);
    pdl_wait();
    const int a = rev5(lane);
    const int src_a = rev5((32 - a) & 31), src_b = 31 - lane;
    constexpr float inv = 1.0f / 256.0f;

    for (int s = blockIdx.x * kStftWarps + warp; s < n_streams; s += gridDim.x * kStftWarps) {
        cpx Y[8];
        float m[8];
        const float4 *sp = reinterpret_cast<const float4 *>(spec + (size_t) s * kNfft + 16 * a);
        const float4 *mp = reinterpret_cast<const float4 *>(mask + (size_t) s * kBins + 8 * a);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float4 t = sp[q];
            Y[2 * q] = cpx{t.x, t.y};
            Y[2 * q + 1] = cpx{t.z, t.w};
        }
        const float4 m0 = mp[0], m1 = mp[1];
        m[0] = m0.x; m[1] = m0.y; m[2] = m0.z; m[3] = m0.w;
        m[4] = m1.x; m[5] = m1.y; m[6] = m1.z; m[7] = m1.w;
        // OLA tail of this stream: issued early so the loads overlap the transform
        float2 *ola2 = reinterpret_cast<float2 *>(ola + (size_t) s * kFrame);
        float2 o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = ola2[lane + 32 * j];

        const float m255 = __shfl_sync(0xffffffffu, m[7], 31);   // lane 31 holds bins 248..255
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            Y[b].x *= m[b];
            Y[b].y *= (b == 0 && lane == 0) ? m255 : m[b];       // lane 0, b 0: Im slot carries X[256], masked by mask[255]
        }
        // inverse split: Z[k] = E + i O,  E = (Y[k] + conj Y[256-k]) / 2,  O = (Y[k] - conj Y[256-k]) / 2 * conj(W512^k)
        cpx z[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int b = rev3c(j), bp = (8 - b) & 7;
            const cpx yk = Y[b];
            const cpx yp = shfl_c(Y[bp], j == 0 ? src_a : src_b);
            const cpx E = {0.5f * (yk.x + yp.x), 0.5f * (yk.y - yp.y)};
            const cpx D = {0.5f * (yk.x - yp.x), 0.5f * (yk.y + yp.y)};
            const float2 w = c.tp(b);
            const cpx O = cmulc(D, cpx{w.x, w.y});
            cpx zz = {E.x - O.y, E.y + O.x};
            if (j == 0 && lane == 0) zz = cpx{0.5f * (yk.x + yk.y), 0.5f * (yk.x - yk.y)};
            z[j] = zz;
        }
        warp_ifft256_dit(z, c, lane);

        // z[j] = 256 (y[2p] + i y[2p+1]), p = lane + 32 j.  j < 4: first half -> output; j >= 4: second half -> new OLA tail.
        uint32_t *out32 = reinterpret_cast<uint32_t *>(v.out + (size_t) s * v.stride + (size_t) v.t * kFrame);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int p = lane + 32 * j;
            const float2 w = c.win(j);
            const float v0 = o[j].x + w.x * (z[j].x * inv), v1 = o[j].y + w.y * (z[j].y * inv);
            short i0, i1;   // round-to-nearest-even + saturate, the oracle's rintf + clamp
            asm("cvt.rni.sat.s16.f32 %0, %1;" : "=h"(i0) : "f"(v0));
            asm("cvt.rni.sat.s16.f32 %0, %1;" : "=h"(i1) : "f"(v1));
            out32[p] = (uint32_t) (uint16_t) i0 | ((uint32_t) (uint16_t) i1 << 16);
        }
#pragma unroll
        for (int j = 4; j < 8; ++j) {
            const float2 w = c.win(j);
            ola2[lane + 32 * (j - 4)] = make_float2(w.x * (z[j].x * inv), w.y * (z[j].y * inv));
        }
    }
}

}  // namespace koala
