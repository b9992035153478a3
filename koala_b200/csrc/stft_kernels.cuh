// koala_b200 -- analysis (int16 -> STFT -> features) and synthesis (mask apply -> iSTFT -> overlap-add -> int16).
//
// These are the first and last stage of `pv_koala_process` (/root/reference/include/pv_koala.h:65-80; stage names from
// BASELINE.json north_star; the reference's own versions exist only as sm_61 SASS, SURVEY.md section 2.1 taabe228/151/7).
// One warp owns one stream-frame at a time: a 512-point real FFT is done as a 256-point complex FFT held 8 points per
// lane (warp_fft256 in koala_common.cuh: 40 shuffles), followed by the real-FFT split, which handles the bins k and 256 - k
// together so that only half of them cross lanes (8 more shuffles).  By bytes these stages are HBM-bound; in practice the
// LSU data pipe (shuffles + global accesses, one wavefront per two cycles) is the limiter, so everything is arranged to
// keep its wavefront count down: no shared memory, PCM read straight into the FFT's register layout with coalesced 128-byte
// rows, window and split twiddles rebuilt from per-lane base values with a few FMAs instead of being looked up.
#pragma once

#include "koala_common.cuh"

namespace koala {

#ifndef KOALA_STFT_CTAS
#define KOALA_STFT_CTAS 6
#endif
constexpr int kStftWarps = 4;         // warps (= streams in flight) per CTA
constexpr int kStftCtasPerSm = KOALA_STFT_CTAS;     // resident CTAs per SM the register budget is sized for

// kPlanes bf16 planes of a feature row of kPlanes * 256 columns (masknet_fused.cuh: 1 = bf16 mode, 3 = hi | mid | lo of fp32 mode)
template <typename FeatT, int kPlanes> __device__ __forceinline__ void store_feat(FeatT *dst, float f);
template <> __device__ __forceinline__ void store_feat<float, 1>(float *dst, float f) { *dst = f; }
template <> __device__ __forceinline__ void store_feat<__nv_bfloat16, 1>(__nv_bfloat16 *dst, float f) { *dst = __float2bfloat16_rn(f); }
template <> __device__ __forceinline__ void store_feat<__nv_bfloat16, 3>(__nv_bfloat16 *dst, float f) {
#pragma unroll
    for (int pl = 0; pl < 3; ++pl) {
        const __nv_bfloat16 b = __float2bfloat16_rn(f);
        dst[pl * kBins] = b;
        f -= __bfloat162float(b);
    }
}

// fixed-point mode (masknet_i8.cuh): the feature as int16 Q14 (round to nearest even, saturating), hi byte plane | lo byte plane
template <> __device__ __forceinline__ void store_feat<uint8_t, 2>(uint8_t *dst, float f) {
    const int q = min(max(__float2int_rn(f * 16384.0f), -32768), 32767);
    dst[0] = (uint8_t) ((q >> 8) & 255);
    dst[kBins] = (uint8_t) (q & 255);
}

__device__ __forceinline__ float feature_of(float re, float im) {
    return kFeatGain * __logf((re * re + im * im) * kFeatPowerScale + kFeatEps) + kFeatBias;
}

// Launch: block = 128; the grid's warps walk the (frame, stream) items of the launch, item = t * n_streams + s, with a stride
// of gridDim * 4 (the engine sizes the grid so that every warp gets about the same number of items).  A launch covers `frames`
// consecutive frames of every stream (the chunk the fused mask-estimator kernel then walks in one launch): frame t's first
// half is frame t - 1 of the caller's buffer, or the stream's tail (the last frame of the previous chunk / call) for t = 0.
// The tail itself is rewritten by backend_kernel, after every reader of this launch is done.
// spec: [frames][slot_rows][512] fp32 packed (Re, Im of bins 0..255; the Im slot of bin 0 carries Re X[256]); feat: [frames][slot_rows][kPlanes * 256].
template <typename FeatT, int kPlanes>
__global__ void __launch_bounds__(kStftWarps * 32, kStftCtasPerSm)
frontend_kernel(PcmView v, int n_streams, int frames, long long slot_rows, const int16_t *__restrict__ tail, float *__restrict__ spec,
                FeatT *__restrict__ feat, const float2 *__restrict__ lane_tab) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    pdl_launch_dependents();
    FftLane c;
    load_fft_lane(c, lane_tab, lane, false);   // constants: may be read before the previous kernel has finished
    const float2 sbase = __ldg(lane_tab + kLaneTabSplit * 32 + lane);
    const float2 we = __ldg(lane_tab + kLaneTabWin * 32 + lane), wo = __ldg(lane_tab + (kLaneTabWin + 1) * 32 + lane);
    pdl_wait();
    const int cb = rev5(lane);                  // my bins: k = cb + 32 m and their partners 256 - k
    const int src = rev5((32 - cb) & 31);       // lane that holds the partner bins
    const bool lane0 = lane == 0;

    const int items = n_streams * frames;
    for (int idx = blockIdx.x * kStftWarps + warp; idx < items; idx += gridDim.x * kStftWarps) {
        const int t = idx / n_streams, s = idx - t * n_streams;
        // frame = [previous input frame | this input frame] as 256 sample pairs; point p = lane + 32 j is pair p: j < 4 comes
        // from the previous frame, j >= 4 from the new samples, every load a coalesced 128-byte row
        const int16_t *in_row = v.in + (size_t) s * v.stride + (size_t) (v.t + t) * v.frame_stride;
        const uint32_t *prev_s = reinterpret_cast<const uint32_t *>(t == 0 ? tail + (size_t) s * kFrame : in_row - v.frame_stride);
        const uint32_t *in_s = reinterpret_cast<const uint32_t *>(in_row);
        uint32_t u[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) u[j] = prev_s[lane + 32 * j];
#pragma unroll
        for (int j = 0; j < 4; ++j) u[4 + j] = in_s[lane + 32 * j];

        cpx z[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float2 w = window_pair(we, wo, j);
            z[j] = cpx{w.x * (float) (int16_t) (u[j] & 0xffffu), w.y * (float) (int16_t) (u[j] >> 16)};
        }
        warp_fft256<false>(z, c, lane);

        // real-FFT split for k = cb + 32 m, m < 4, and 256 - k at once:
        //   E = (Z[k] + conj Z[256-k]) / 2,  O = (Z[k] - conj Z[256-k]) / 2i,  X[k] = E + W512^k O,  X[256-k] = conj(E - W512^k O)
        const size_t row = (size_t) t * slot_rows + s;
        float *spec_s = spec + row * kNfft;
        FeatT *feat_s = feat + row * (kPlanes * kBins);
#pragma unroll
        for (int m = 0; m < 4; ++m) {
            const cpx zk = z[fft_reg_of_m(m)];
            cpx zp = shfl_c(z[fft_reg_of_m(7 - m)], src);
            if (m > 0 && lane0) zp = z[fft_reg_of_m(8 - m)];          // c = 0: the partner of 32 m is 256 - 32 m, in my own registers
            const cpx E = {0.5f * (zk.x + zp.x), 0.5f * (zk.y - zp.y)};
            const cpx O = {0.5f * (zk.y + zp.y), -0.5f * (zk.x - zp.x)};
            const cpx T = cmul(O, split_twiddle(sbase, m));
            cpx xa = cadd(E, T), xb = {E.x - T.x, T.y - E.y};
            int ka = cb + 32 * m, kb = 256 - ka;
            if (m == 0 && lane0) {                                    // bins 0 / 256 (both real, packed) and 128
                const cpx z128 = z[fft_reg_of_m(4)];
                xa = cpx{zk.x + zk.y, zk.x - zk.y};
                xb = cpx{z128.x, -z128.y};
                kb = 128;
            }
            *reinterpret_cast<float2 *>(spec_s + 2 * ka) = make_float2(xa.x, xa.y);
            *reinterpret_cast<float2 *>(spec_s + 2 * kb) = make_float2(xb.x, xb.y);
            store_feat<FeatT, kPlanes>(feat_s + ka, feature_of(xa.x, (m == 0 && lane0) ? 0.0f : xa.y));
            store_feat<FeatT, kPlanes>(feat_s + kb, feature_of(xb.x, xb.y));
        }
    }
}

// Masked spectrum of one frame -> z[j] = 256 (y[2p] + i y[2p+1]), p = lane + 32 j (the 512 synthesis samples before windowing)
__device__ __forceinline__ void synth_frame(const float *__restrict__ spec_s, const float *__restrict__ mask_s, cpx (&z)[8], const FftLane &c,
                                            const float2 &sbase, int lane, int cb, int src, bool lane0) {
    cpx ya[4], yb[4];
    float ma[4], mb[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        const int ka = cb + 32 * m, kb = (m == 0 && lane0) ? 128 : 256 - ka;
        const float2 a = *reinterpret_cast<const float2 *>(spec_s + 2 * ka), b = *reinterpret_cast<const float2 *>(spec_s + 2 * kb);
        ya[m] = cpx{a.x, a.y};
        yb[m] = cpx{b.x, b.y};
        ma[m] = mask_s[ka];
        mb[m] = mask_s[kb];
    }
    const float m255 = __shfl_sync(0xffffffffu, mb[0], 16);       // bin 255 = 256 - 1: partner bin of the lane with c = 1
    cpx zp[4];
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        // mask, then the inverse split: E = (Y[k] + conj Y[256-k]) / 2, O = (Y[k] - conj Y[256-k]) / 2 * conj(W512^k),
        // Z[k] = E + i O, Z[256-k] = conj(E) + i conj(O)
        const cpx yk = {ya[m].x * ma[m], ya[m].y * ((m == 0 && lane0) ? m255 : ma[m])};   // c = 0, m = 0: Im slot carries X[256], masked by mask[255]
        const cpx yp = {yb[m].x * mb[m], yb[m].y * mb[m]};
        const cpx E = {0.5f * (yk.x + yp.x), 0.5f * (yk.y - yp.y)};
        const cpx D = {0.5f * (yk.x - yp.x), 0.5f * (yk.y + yp.y)};
        const cpx O = cmulc(D, split_twiddle(sbase, m));
        cpx zk = {E.x - O.y, E.y + O.x};
        zp[m] = cpx{E.x + O.y, O.x - E.y};
        if (m == 0 && lane0) {
            zk = cpx{0.5f * (yk.x + yk.y), 0.5f * (yk.x - yk.y)};   // Z[0] from the packed real pair (X[0], X[256])
            zp[0] = cpx{yp.x, -yp.y};                               // Z[128] = conj Y[128]
        }
        z[fft_reg_of_m(m)] = zk;
    }
    // the partner halves go home: register m' = 7 - m of lane `src`; c = 0 keeps its own (Z[128] and Z[256 - 32 m])
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        const cpx t = shfl_c(zp[m], src);
        const cpx own = m == 3 ? zp[0] : zp[m + 1];                // lane 0: register 7 - m holds bin 32 (7 - m) = 256 - 32 (m + 1); m = 3: bin 128
        z[fft_reg_of_m(7 - m)] = lane0 ? own : t;
    }
    warp_fft256<true>(z, c, lane);
}

// Same launch shape; items are (run, stream) pairs, run r = frames [r * run, min(frames, (r + 1) * run)) of the launch: a warp
// synthesises them one after the other and carries the overlap-add half in registers.  A run that does not start at the
// launch's first frame first re-synthesises the frame before it to get that half (one extra inverse transform per run;
// the engine makes runs as long as the stream count allows: 8192 streams x 16 frames = one run per stream).  The warp that
// owns a stream's last frame stores the overlap-add state and the new analysis tail (= the last input frame, read before the
// output is written: the caller's buffers may alias).  ola_in != ola_out unless the launch is one run per stream: other
// warps still read ola_in.
// mask: [frames][slot_rows][256] fp32.  ola: [n][256] fp32 state.
__global__ void __launch_bounds__(kStftWarps * 32, kStftCtasPerSm)
backend_kernel(PcmView v, int n_streams, int frames, int run, long long slot_rows, const float *__restrict__ spec,
               const float *__restrict__ mask, const float *ola_in, float *ola_out, int16_t *__restrict__ tail,
               const float2 *__restrict__ lane_tab) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    pdl_launch_dependents();
    FftLane c;
    load_fft_lane(c, lane_tab, lane, true);
    const float2 sbase = __ldg(lane_tab + kLaneTabSplit * 32 + lane);
    const float2 we = __ldg(lane_tab + kLaneTabWin * 32 + lane), wo = __ldg(lane_tab + (kLaneTabWin + 1) * 32 + lane);
    pdl_wait();
    const int cb = rev5(lane);
    const int src = rev5((32 - cb) & 31);
    const bool lane0 = lane == 0;
    constexpr float inv = 1.0f / 256.0f;

    const int runs = (frames + run - 1) / run, items = n_streams * runs;
    for (int idx = blockIdx.x * kStftWarps + warp; idx < items; idx += gridDim.x * kStftWarps) {
        const int r = idx / n_streams, s = idx - r * n_streams;
        const int ta = r * run, tb = min(frames, ta + run);
        float2 o[4];           // second half of the previous synthesis frame
        cpx z[8];
        if (ta == 0) {         // the stream's state: issued early so the loads overlap the first transform
            const float2 *ola2 = reinterpret_cast<const float2 *>(ola_in + (size_t) s * kFrame);
#pragma unroll
            for (int j = 0; j < 4; ++j) o[j] = ola2[lane + 32 * j];
        } else {
            const size_t row = (size_t) (ta - 1) * slot_rows + s;
            synth_frame(spec + row * kNfft, mask + row * kBins, z, c, sbase, lane, cb, src, lane0);
#pragma unroll
            for (int j = 4; j < 8; ++j) {
                const float2 w = window_pair(we, wo, j);
                o[j - 4] = make_float2(w.x * (z[j].x * inv), w.y * (z[j].y * inv));
            }
        }
        for (int t = ta; t < tb; ++t) {
            const size_t row = (size_t) t * slot_rows + s;
            synth_frame(spec + row * kNfft, mask + row * kBins, z, c, sbase, lane, cb, src, lane0);
            if (t == frames - 1) {                                     // state: tail <- the launch's last input frame
                const uint32_t *in_s = reinterpret_cast<const uint32_t *>(v.in + (size_t) s * v.stride + (size_t) (v.t + t) * v.frame_stride);
                uint32_t *tail_s = reinterpret_cast<uint32_t *>(tail + (size_t) s * kFrame);
                uint32_t u[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) u[j] = in_s[lane + 32 * j];
#pragma unroll
                for (int j = 0; j < 4; ++j) tail_s[lane + 32 * j] = u[j];
            }
            // z[j] = 256 (y[2p] + i y[2p+1]), p = lane + 32 j.  j < 4: first half -> output; j >= 4: second half -> next overlap-add half.
            uint32_t *out32 = reinterpret_cast<uint32_t *>(v.out + (size_t) s * v.out_stride + (size_t) (v.t + t) * v.out_frame_stride);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 w = window_pair(we, wo, j);
                const float v0 = o[j].x + w.x * (z[j].x * inv), v1 = o[j].y + w.y * (z[j].y * inv);
                short i0, i1;   // round-to-nearest-even + saturate, the oracle's rintf + clamp
                asm("cvt.rni.sat.s16.f32 %0, %1;" : "=h"(i0) : "f"(v0));
                asm("cvt.rni.sat.s16.f32 %0, %1;" : "=h"(i1) : "f"(v1));
                out32[lane + 32 * j] = (uint32_t) (uint16_t) i0 | ((uint32_t) (uint16_t) i1 << 16);
            }
#pragma unroll
            for (int j = 4; j < 8; ++j) {
                const float2 w = window_pair(we, wo, j);
                o[j - 4] = make_float2(w.x * (z[j].x * inv), w.y * (z[j].y * inv));
            }
        }
        if (tb == frames) {
            float2 *ola2 = reinterpret_cast<float2 *>(ola_out + (size_t) s * kFrame);
#pragma unroll
            for (int j = 0; j < 4; ++j) ola2[lane + 32 * j] = o[j];
        }
    }
}

}  // namespace koala
