// koala_b200 -- shared constants and small device helpers (sm_100a only).
//
// The path implemented here is the inside of `pv_koala_process` (/root/reference/include/pv_koala.h:65-80), which the
// reference ships only as a closed binary; the signal path is SPEC.md of this repository, the contract is the header's.
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

namespace koala {

constexpr int kFrame = 256;      // pv_koala_frame_length()  (pv_koala.h:102-107)
constexpr int kSampleRate = 16000;  // pv_sample_rate()      (picovoice.h:33-36)
constexpr int kNfft = 512;
constexpr int kBins = 256;       // bins fed to the mask estimator; Nyquist bin reuses mask[255]
constexpr int kDelay = 256;      // pv_koala_delay_sample()  (pv_koala.h:92-100)
constexpr int kMaxLayers = 8;

constexpr float kFeatPowerScale = 9.31322574615478515625e-10f;  // 2^-30
constexpr float kFeatEps = 1e-6f;
constexpr float kFeatGain = 0.125f;
constexpr float kFeatBias = 0.25f;

enum Precision : int { kFp32 = 0, kBf16 = 1 };

// kernel classes of one step, in launch order (per-class timing for bench.py's roofline object)
enum KernelClass : int { kKernFrontend = 0, kKernEnc = 1, kKernGru = 2, kKernDec = 3, kKernBackend = 4, kKernClasses = 5 };

// Optional per-launch CUDA-event timing on the launching stream (off by default: events perturb back-to-back launches).
struct KernelProfiler {
    struct Span { int cls; cudaEvent_t a, b; };
    Span *spans = nullptr;
    int n = 0, cap = 0;
    void begin(int cls, cudaStream_t st) {
        if (n == cap) {
            const int ncap = cap ? 2 * cap : 1024;
            Span *ns = new Span[ncap];
            for (int i = 0; i < n; i++) ns[i] = spans[i];
            delete[] spans;
            spans = ns;
            cap = ncap;
        }
        spans[n].cls = cls;
        cudaEventCreate(&spans[n].a);
        cudaEventCreate(&spans[n].b);
        cudaEventRecord(spans[n].a, st);
    }
    void end(cudaStream_t st) { cudaEventRecord(spans[n++].b, st); }
    // sums elapsed ms per class and releases the events; the caller must have synchronised the stream
    void drain(double *ms, long long *count) {
        for (int i = 0; i < n; i++) {
            float t = 0.0f;
            if (cudaEventElapsedTime(&t, spans[i].a, spans[i].b) == cudaSuccess) {
                ms[spans[i].cls] += t;
                count[spans[i].cls] += 1;
            }
            cudaEventDestroy(spans[i].a);
            cudaEventDestroy(spans[i].b);
        }
        n = 0;
    }
    ~KernelProfiler() {
        double ms[kKernClasses] = {};
        long long c[kKernClasses] = {};
        drain(ms, c);
        delete[] spans;
    }
};

// Where one step's PCM lives: frame t of stream s starts at pcm + s * stride + t * 256 (samples).
struct PcmView {
    const int16_t *in;
    int16_t *out;
    long long stride;   // samples between consecutive streams
    int t;              // frame index inside the caller's buffer
};

// ---------------------------------------------------------------------------------------------------------------
// complex helpers
struct cpx { float x, y; };
__device__ __forceinline__ cpx cadd(cpx a, cpx b) { return {a.x + b.x, a.y + b.y}; }
__device__ __forceinline__ cpx csub(cpx a, cpx b) { return {a.x - b.x, a.y - b.y}; }
__device__ __forceinline__ cpx cmul(cpx a, cpx w) { return {a.x * w.x - a.y * w.y, a.x * w.y + a.y * w.x}; }
__device__ __forceinline__ cpx cmulc(cpx a, cpx w) { return {a.x * w.x + a.y * w.y, a.y * w.x - a.x * w.y}; }  // a * conj(w)
__device__ __forceinline__ cpx shfl_xor_c(cpx v, int m) {
    return {__shfl_xor_sync(0xffffffffu, v.x, m), __shfl_xor_sync(0xffffffffu, v.y, m)};
}
__device__ __forceinline__ cpx shfl_c(cpx v, int src) {
    return {__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src)};
}
__device__ __forceinline__ int rev5(int v) { return (int) (__brev((unsigned) v) >> 27); }

// Per-lane constants of the warp FFT.  The host lays them out lane-major (float2 [kLaneTabRows][32]): lane l's copy of
// constant i is tab[i * 32 + l], so a row is one coalesced 256-byte request from global memory and a conflict-free access
// from shared memory (the first version indexed twiddles by bin, with up to 16-way bank conflicts: profiles/r01_step_summary.md).
// The twiddles of the butterfly stages (rows 0-10) are used by every stage and live in registers for all streams a warp
// walks; the real-FFT split factors and the window (rows 11-26) are used once per frame and stay in shared memory, which
// keeps the kernels under 72 registers = 28 resident warps per SM (these kernels are latency-bound: occupancy is speed).
constexpr int kLaneTabRows = 27;
constexpr int kLaneTabRegRows = 11;                               // t4, t2, t1, tx
constexpr int kLaneTabSmemRows = kLaneTabRows - kLaneTabRegRows;  // tp[8], win[8]
struct FftLane {
    float2 t4[4];    // span 128: W512^{2 (lane + 32 q)},  q = j & 3
    float2 t2[2];    // span  64: W512^{4 (lane + 32 q)},  q = j & 1
    float2 t1;       // span  32: W512^{8 lane}
    float2 tx[4];    // spans 16, 8, 4, 2 (lane exchange): W512^{(lane & (h-1)) * 256 / h}
    const float2 *sm;  // shared-memory rows of this lane: tp(b) = real-FFT split W512^{8 rev5(lane) + b}; win(j) = sqrt-Hann
                       // window pair (w[2p], w[2p+1]), p = lane + 32 j
    __device__ __forceinline__ float2 tp(int b) const { return sm[32 * b]; }
    __device__ __forceinline__ float2 win(int j) const { return sm[32 * (8 + j)]; }
};
// s_tab: float2 [kLaneTabSmemRows * 32] of shared memory, filled by the whole CTA (ends with __syncthreads)
__device__ __forceinline__ void load_fft_lane(FftLane &c, const float2 *__restrict__ tab, float2 *s_tab, int lane) {
    for (int i = threadIdx.x; i < kLaneTabSmemRows * 32; i += blockDim.x) s_tab[i] = __ldg(tab + kLaneTabRegRows * 32 + i);
    const float2 *t = tab + lane;
#pragma unroll
    for (int i = 0; i < 4; ++i) c.t4[i] = __ldg(t + 32 * i);
#pragma unroll
    for (int i = 0; i < 2; ++i) c.t2[i] = __ldg(t + 32 * (4 + i));
    c.t1 = __ldg(t + 32 * 6);
#pragma unroll
    for (int i = 0; i < 4; ++i) c.tx[i] = __ldg(t + 32 * (7 + i));
    c.sm = s_tab + lane;
    __syncthreads();
}

// One warp = one 256-point complex FFT.  Lane l, register j hold element p = l + 32 j.
// Forward: radix-2 DIF, natural order in -> bit-reversed out: after the call z[j] = Z[8 * rev5(lane) + rev3(j)].
__device__ __forceinline__ void warp_fft256_dif(cpx (&z)[8], const FftLane &c, int lane) {
#pragma unroll
    for (int dj = 4; dj >= 1; dj >>= 1) {   // spans 128, 64, 32: partner is another register of the same lane
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (j & dj) continue;
            const float2 w2 = dj == 4 ? c.t4[j & 3] : dj == 2 ? c.t2[j & 1] : c.t1;
            const cpx a = z[j], b = z[j + dj];
            z[j] = cadd(a, b);
            z[j + dj] = cmul(csub(a, b), cpx{w2.x, w2.y});
        }
    }
    // spans 16..1: partner is lane ^ h.  Branch-free butterfly: both lanes compute (o + sgn * v) * w' with sgn = -1 and
    // w' = twiddle in the upper lane, sgn = +1 and w' = 1 in the lower lane (6 arithmetic instructions + 2 shuffles per point)
#pragma unroll
    for (int s = 0; s < 5; ++s) {
        const int h = 16 >> s;
        const bool up = (lane & h) != 0;
        const float sgn = up ? -1.0f : 1.0f;
        const cpx w = (s < 4 && up) ? cpx{c.tx[s < 4 ? s : 0].x, c.tx[s < 4 ? s : 0].y} : cpx{1.0f, 0.0f};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const cpx v = z[j], o = shfl_xor_c(v, h);
            const cpx t = {fmaf(sgn, v.x, o.x), fmaf(sgn, v.y, o.y)};
            z[j] = s < 4 ? cmul(t, w) : t;   // span 1: twiddle is 1 for everyone
        }
    }
}

// Inverse: radix-2 DIT with conjugate twiddles, bit-reversed in (layout produced by warp_fft256_dif) -> natural out.
// No 1/256 scaling is applied.
__device__ __forceinline__ void warp_ifft256_dit(cpx (&z)[8], const FftLane &c, int lane) {
    // spans 1, 2, 4, 8, 16: upper lane first scales by conj(twiddle) (lower lane by 1), then out = o + sgn * t
#pragma unroll
    for (int s = 4; s >= 0; --s) {
        const int h = 16 >> s;
        const bool up = (lane & h) != 0;
        const float sgn = up ? -1.0f : 1.0f;
        const cpx w = (s < 4 && up) ? cpx{c.tx[s < 4 ? s : 0].x, c.tx[s < 4 ? s : 0].y} : cpx{1.0f, 0.0f};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const cpx t = s < 4 ? cmulc(z[j], w) : z[j];
            const cpx o = shfl_xor_c(t, h);
            z[j] = cpx{fmaf(sgn, t.x, o.x), fmaf(sgn, t.y, o.y)};
        }
    }
#pragma unroll
    for (int dj = 1; dj <= 4; dj <<= 1) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (j & dj) continue;
            const float2 w2 = dj == 4 ? c.t4[j & 3] : dj == 2 ? c.t2[j & 1] : c.t1;
            const cpx a = z[j], b = cmulc(z[j + dj], cpx{w2.x, w2.y});
            z[j] = cadd(a, b);
            z[j + dj] = csub(a, b);
        }
    }
}

// register index holding the partner bin 256 - k of register j (k = 8 a + rev3(j)); see DESIGN.md "FFT layout"
__device__ __forceinline__ constexpr int partner_reg(int j) {
    return j == 0 ? 0 : j == 1 ? 1 : j == 2 ? 3 : j == 3 ? 2 : j == 4 ? 7 : j == 5 ? 6 : j == 6 ? 5 : 4;
}
__device__ __forceinline__ constexpr int rev3c(int j) { return ((j & 1) << 2) | (j & 2) | ((j >> 2) & 1); }

// ---------------------------------------------------------------------------------------------------------------
// mbarrier helpers shared by the TMA / tcgen05 pipelines
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!done);
}
// Programmatic dependent launch: every kernel of the step is launched with programmatic stream serialisation, calls
// pdl_launch_dependents() first thing (so the next kernel's launch, CTA placement and prologue overlap this kernel) and
// pdl_wait() before its first access to data produced by the previous kernel (blocks until that grid has fully finished).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// Process-wide switch (KOALA_B200_PDL=0 turns it off).  The fp32 CUDA-core path launches without it: its many small grids ran
// 50 % slower with every future kernel's CTAs parked on the SMs.
static inline bool pdl_enabled() {
    static const bool on = [] { const char *e = getenv("KOALA_B200_PDL"); return !(e && e[0] == '0'); }();
    return on;
}

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(bool pdl, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args &&...args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (pdl && pdl_enabled()) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// accurate-enough transcendental forms shared by every epilogue (abs error ~1e-7, see SPEC.md "numerics")
__device__ __forceinline__ float sigmoid_f(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float tanh_f(float x) {
    // 1 - 2 / (1 + e^{2x}); saturates cleanly for |x| large (e^{2x} -> inf => 1, -> 0 => -1)
    return 1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * x));
}

}  // namespace koala
