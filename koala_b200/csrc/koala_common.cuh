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

enum Precision : int { kFp32 = 0, kBf16 = 1, kInt8 = 2 };   // kInt8: the fixed-point variant (SPEC.md section 6, masknet_i8.cuh)

// kernel classes of one step, in launch order (per-class timing for bench.py's roofline object)
// (kKernMasknet: the fused encoder -> GRU -> decoder kernel of the bf16 path; Enc / Gru / Dec: the fp32 path's separate kernels)
enum KernelClass : int { kKernFrontend = 0, kKernEnc = 1, kKernGru = 2, kKernDec = 3, kKernBackend = 4, kKernMasknet = 5, kKernClasses = 6 };

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

// Where a launch's PCM lives: frame t of stream s starts at in + s * stride + t * frame_stride and
// out + s * out_stride + t * out_frame_stride (samples).  Stream-major buffers: frame stride 256; time-major: stride 256,
// frame stride = streams * 256.
struct PcmView {
    const int16_t *in;
    int16_t *out;
    long long stride;       // samples between consecutive streams of `in`
    long long out_stride;   // ... of `out` (differs only inside the host ingest path, whose output blocks are wider)
    long long frame_stride, out_frame_stride;   // samples between consecutive frames of one stream
    int t;                  // index of the launch's first frame inside the caller's buffer
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

// One warp = one 256-point complex FFT, 8 points per lane, radix-2 decimation in frequency.
//
// A butterfly stage needs both elements of a pair in the same lane.  The three high index bits start out as the register
// index; each of the five low bits lives in the lane id and is first SWAPPED with an already-processed register bit (lane
// l and lane l ^ h exchange half of their registers: 8 shuffles per stage, half of what a butterfly across lanes costs)
// and then processed in registers.  These kernels are bound by the LSU data pipe that executes shuffles and shared-memory
// accesses at one wavefront per two cycles (profiles/r01_step_summary.md), so shuffles are what is being minimised: 40 per
// transform, no shared memory.
//
// The same routine runs both directions; only the bit bookkeeping and the twiddle tables differ:
//   forward: in  z[j] = x[lane + 32 j]                       out z[j] = Z[k], k = rev5(lane) + 32 m, m = (j1 j2 j0)
//   inverse: in  that frequency layout                        out z[j] = 256 x[lane + 32 j]
// (j2 j1 j0 = bits of the register index j; "m = (j1 j2 j0)" means m bit 2 = j bit 1, m bit 1 = j bit 2, m bit 0 = j bit 0).
// Twiddles are per-lane constants, laid out lane-major by the host (float2 [kLaneTabRows][32], engine.cu): a row is one
// coalesced 256-byte load, and they stay in registers for all streams a warp walks.
constexpr int kLaneTabRows = 25;
constexpr int kLaneTabInv = 11;      // first row of the inverse set
constexpr int kLaneTabSplit = 22;    // W512^c, c = rev5(lane)
constexpr int kLaneTabWin = 23;      // (sin, cos) of pi (2 lane) / 512 and of pi (2 lane + 1) / 512
struct FftLane {
    float2 t4[4];    // first stage:  W256^(b + 32 q), q = the two register bits still to come;  b = low five index bits of the lane
    float2 t2[2];    // second stage: W128^(b + 32 q), q = the last register bit
    float2 t1;       // third stage:  W64^b
    float2 tx[4];    // swapped stages t = 4..1: W_{2^(t+1)}^(b mod 2^t), negated in the lanes whose swapped bit is set
};
__device__ __forceinline__ void load_fft_lane(FftLane &c, const float2 *__restrict__ tab, int lane, bool inverse) {
    const float2 *t = tab + (inverse ? kLaneTabInv * 32 : 0) + lane;
#pragma unroll
    for (int i = 0; i < 4; ++i) c.t4[i] = __ldg(t + 32 * i);
#pragma unroll
    for (int i = 0; i < 2; ++i) c.t2[i] = __ldg(t + 32 * (4 + i));
    c.t1 = __ldg(t + 32 * 6);
#pragma unroll
    for (int i = 0; i < 4; ++i) c.tx[i] = __ldg(t + 32 * (7 + i));
}

template <bool INV>
__device__ __forceinline__ void warp_fft256(cpx (&z)[8], const FftLane &c, int lane) {
    // register bits of the three in-register stages, in processing order (index bits 7, 6, 5)
    constexpr int RB0 = INV ? 2 : 4, RB1 = INV ? 4 : 2, RB2 = 1;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        if (j & RB0) continue;
        const float2 w = c.t4[((j & RB1) ? 2 : 0) | ((j & RB2) ? 1 : 0)];
        const cpx a = z[j], b = z[j | RB0];
        z[j] = cadd(a, b);
        z[j | RB0] = cmul(csub(a, b), cpx{w.x, w.y});
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        if (j & RB1) continue;
        const float2 w = c.t2[(j & RB2) ? 1 : 0];
        const cpx a = z[j], b = z[j | RB1];
        z[j] = cadd(a, b);
        z[j | RB1] = cmul(csub(a, b), cpx{w.x, w.y});
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        if (j & RB2) continue;
        const cpx a = z[j], b = z[j | RB2];
        z[j] = cadd(a, b);
        z[j | RB2] = cmul(csub(a, b), cpx{c.t1.x, c.t1.y});
    }
    // index bits 4..0: lane bit h <-> register bit d, then the stage on register bit d.  The lane with bit h clear keeps its
    // d = 0 registers and receives the partner's, the other lane keeps d = 1: `kept` and `recv` are then the two elements of a
    // pair, in the order (a, b) in the lower lane and (b, a) in the upper one; a + b is symmetric and the sign of a - b is
    // folded into the upper lanes' twiddle.
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const int h = INV ? (1 << i) : (16 >> i);
        const int d = INV ? (i == 0 ? 2 : i == 1 ? 4 : i == 2 ? 1 : i == 3 ? 2 : 4) : (i == 0 ? 4 : i == 1 ? 2 : i == 2 ? 1 : i == 3 ? 4 : 2);
        const bool up = (lane & h) != 0;
        const float sgn = up ? -1.0f : 1.0f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (j & d) continue;
            const cpx kept = up ? z[j | d] : z[j], send = up ? z[j] : z[j | d];
            const cpx recv = shfl_xor_c(send, h);
            z[j] = cadd(kept, recv);
            const cpx df = csub(kept, recv);
            z[j | d] = i < 4 ? cmul(df, cpx{c.tx[i < 4 ? i : 0].x, c.tx[i < 4 ? i : 0].y}) : cpx{sgn * df.x, sgn * df.y};
        }
    }
}

// Frequency layout shared by the two kernels (output of the forward transform): lane holds the bins k = c + 32 m,
// c = rev5(lane), m = 0..7, in register kFftRegOfM[m].  The partner bin 256 - k of the real-FFT split is bin (32 - c) + 32 (7 - m):
// always the same lane rev5((32 - c) & 31), register kFftRegOfM[7 - m] -- except c = 0, whose partners are its own registers.
__device__ __forceinline__ constexpr int fft_reg_of_m(int m) { return m == 0 ? 0 : m == 1 ? 1 : m == 2 ? 4 : m == 3 ? 5 : m == 4 ? 2 : m == 5 ? 3 : m == 6 ? 6 : 7; }

// sqrt-Hann window pair (w[2p], w[2p + 1]) of complex point p = lane + 32 j from the lane's base angles: sin(t + j pi / 8)
__device__ __forceinline__ float2 window_pair(const float2 &we, const float2 &wo, int j) {
    constexpr float cs[8] = {1.0f, 0.92387953251128674f, 0.70710678118654752f, 0.38268343236508977f,
                             0.0f, -0.38268343236508977f, -0.70710678118654752f, -0.92387953251128674f};
    constexpr float sn[8] = {0.0f, 0.38268343236508977f, 0.70710678118654752f, 0.92387953251128674f,
                             1.0f, 0.92387953251128674f, 0.70710678118654752f, 0.38268343236508977f};
    return make_float2(fmaf(we.x, cs[j], we.y * sn[j]), fmaf(wo.x, cs[j], wo.y * sn[j]));
}
// W512^(c + 32 m) = W512^c * exp(-i pi m / 8), m = 0..3
__device__ __forceinline__ cpx split_twiddle(const float2 &base, int m) {
    constexpr float cs[4] = {1.0f, 0.92387953251128674f, 0.70710678118654752f, 0.38268343236508977f};
    constexpr float sn[4] = {0.0f, -0.38268343236508977f, -0.70710678118654752f, -0.92387953251128674f};
    return m == 0 ? cpx{base.x, base.y} : cmul(cpx{base.x, base.y}, cpx{cs[m], sn[m]});
}

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
