// koala_b200 -- Engine implementation: state layout in HBM, model upload, per-step launch sequence.
//
// Per-stream state rows (all zero after reset, pv_koala.h:82-90):
//   tail [Bp][256] int16   previous input frame (analysis overlap)
//   ola  [Bp][256] fp32    second half of the previous synthesis frame (bf16 path: [2][...] ping-pong by chunk, see backend_kernel)
//   h    [L][Bp][H] fp32     recurrent state; bf16 path: updated in place, fp32 path: [2][...] ping-pong by step parity
//   hb   [2][L][Bp][P H] bf16  the same state as GEMM operand of the tensor-core path, ping-pong by step parity (every unit tile
//                            reads all of h(t-1)): P = 1 plane, h rounded to bf16, in bf16 mode; P = 3 planes hi | mid | lo that sum
//                            to the fp32 value exactly in fp32 mode (masknet_fused.cuh)
// The tensor-core path (both precisions, H a multiple of 256) keeps all of the above in one arena (Engine::create); fp32 mode
// with another hidden size runs the CUDA-core kernels of masknet_fp32.cuh frame by frame.
// Scratch: feat [T][Bp][256] (fp32 | bf16), spec [T][Bp][512] fp32, mask [T][Bp][256] fp32 for the T frames one chunk of the
// bf16 path takes through its three launches (analysis of T frames -> fused mask estimator walking T steps -> synthesis of T
// frames; T = chunk_frames() <= 64, sized so that the scratch stays under 512 MB), e [ring][Bp][H] (encoder output ring of
// the fused kernel).  The fp32 path works frame by frame in slot 0.
// Bp = B rounded up to 256 (one CTA-pair tile) so that every GEMM tile is full; padding rows stay zero-input and are
// never copied out.
#include "engine.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <mutex>

#include "koala_common.cuh"
#include "masknet_fp32.cuh"
#include "masknet_fused.cuh"
#include "masknet_i8.cuh"
#include "stft_kernels.cuh"

namespace koala {

// ------------------------------------------------------------------------------------------------ model file
static uint32_t crc32_bytes(const uint8_t *p, size_t n) {
    static uint32_t table[256];
    static bool init = false;
    if (!init) {
        for (uint32_t i = 0; i < 256; i++) {
            uint32_t c = i;
            for (int j = 0; j < 8; j++) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
            table[i] = c;
        }
        init = true;
    }
    uint32_t c = 0xFFFFFFFFu;
    for (size_t i = 0; i < n; i++) c = table[(c ^ p[i]) & 0xFF] ^ (c >> 8);
    return c ^ 0xFFFFFFFFu;
}

Status load_model_file(const char *path, ModelHost *out, std::vector<std::string> *errors) {
    FILE *f = fopen(path, "rb");
    if (!f) {
        errors->push_back(std::string("Failed to open file `") + path + "`.");
        return kIoError;
    }
    fseek(f, 0, SEEK_END);
    const long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::vector<uint8_t> blob(n > 0 ? n : 0);
    const bool ok = n > 0 && fread(blob.data(), 1, n, f) == (size_t) n;
    fclose(f);
    if (!ok) {
        errors->push_back(std::string("Failed to read file `") + path + "`.");
        return kIoError;
    }
    if (n < 48 || memcmp(blob.data(), "koala_b200\0\0", 12) != 0) {
        // the reference's own blob starts with "koala3.0.0" (SURVEY.md F6); it is a different product's format
        char product[13] = {0};
        for (int i = 0; i < 12 && i < n; i++) product[i] = (blob[i] >= 32 && blob[i] < 127) ? (char) blob[i] : '.';
        errors->push_back(std::string("Model file product is `") + product + "` but library product is `koala_b200`.");
        return kInvalidArgument;
    }
    uint32_t hd[8];
    memcpy(hd, blob.data() + 12, 32);
    if (hd[0] != 1 || hd[1] != (uint32_t) kNfft || hd[2] != (uint32_t) kFrame || hd[3] != (uint32_t) kBins || hd[6] != 1 ||
        hd[5] < 1 || hd[5] > (uint32_t) kMaxLayers || hd[4] < 64 || hd[4] > 4096 || hd[4] % 64 != 0) {
        errors->push_back("Model file belongs to a different version of the library.");
        return kInvalidArgument;
    }
    const size_t H = hd[4], L = hd[5];
    const size_t need = 44 + 2 * H * kBins + 4 * H + L * (2 * 2 * 3 * H * H + 2 * 4 * 3 * H) + 2 * kBins * H + 4 * kBins + 4;
    uint32_t crc;
    memcpy(&crc, blob.data() + n - 4, 4);
    if ((size_t) n != need || crc != crc32_bytes(blob.data(), n - 4)) {
        errors->push_back("Model file is corrupt (size or checksum mismatch).");
        return kInvalidArgument;
    }
    out->hidden = (int) H;
    out->layers = (int) L;
    const uint8_t *p = blob.data() + 44;
    auto take16 = [&](std::vector<uint16_t> &v, size_t cnt) { v.resize(cnt); memcpy(v.data(), p, 2 * cnt); p += 2 * cnt; };
    auto take32 = [&](std::vector<float> &v, size_t cnt) { v.resize(cnt); memcpy(v.data(), p, 4 * cnt); p += 4 * cnt; };
    take16(out->enc_w, H * kBins);
    take32(out->enc_b, H);
    out->wih.resize(L); out->whh.resize(L); out->bih.resize(L); out->bhh.resize(L);
    for (size_t l = 0; l < L; l++) {
        take16(out->wih[l], 3 * H * H);
        take16(out->whh[l], 3 * H * H);
        take32(out->bih[l], 3 * H);
        take32(out->bhh[l], 3 * H);
    }
    take16(out->dec_w, kBins * H);
    take32(out->dec_b, kBins);
    return kSuccess;
}

// ------------------------------------------------------------------------------------------------ engine
// Batches larger than this many streams are run as PARTITIONS of at most this size, one after the other inside every chunk: a
// partition's state and scratch (34 MB + the chunk's features / spectra / masks at 4096 streams) then stay in the 126 MB L2 for the
// three launches of its chunk.  Measured per frame of all streams (r02r): 8192 streams 75.8 us as one batch, 2 x 34.9 as two
// partitions; 16 384 streams 160.4 vs 4 x 34.9; 2048-stream partitions are slower again (19.6 us each).  KOALA_PARTITION_STREAMS
// overrides (0: never partition).
constexpr int kPartStreams = 4096;
constexpr int kHostRing = 3;           // device input staging buffers of the host ingest path
constexpr int kHostOutRing = 2;        // device output staging buffers (blocks)
constexpr int kHostChunkFrames = 8;    // frames per input chunk
constexpr int kHostOutFrames = 32;     // frames per output block
constexpr int kHostBlockMinFrames = 128;   // shorter calls send their output chunk by chunk
constexpr int kHostChunkFramesTm = 4;  // frames per chunk of a time-major call (ramped schedules 1, 2, 4 .. 8 .. 4, 2, 1 and 1, 2, 4 .. 4 .. 2, 1 were
                                       // tried for long calls of 8192 streams: 81.8 and 90.5 M frames/s against 91.4 M with uniform chunks, r02z)

struct Engine::Impl {
    cudaStream_t stream = nullptr;
    int H = 0, L = 0, num_sms = 148, stft_per_warp = 0;   // 0: derive from the stream count
    int parity = 0;   // h[parity] holds h(t-1)
    int tcap = 1;     // frames per chunk of the bf16 path (slots of feat / spec / mask)
    int e_ring = 1;   // slots of the encoder output ring
    int planes = 1;   // bf16 planes per activation operand of the tensor-core path (3 in fp32 mode)
    int ola_par = 0;  // ola[ola_par] holds the overlap-add state
    int last_slot = 0;   // scratch slot of the last finished step (debug_read)
    // model
    __nv_bfloat16 *enc_w = nullptr, *dec_w = nullptr, *wih[kMaxLayers] = {}, *whh[kMaxLayers] = {};
    float *enc_b = nullptr, *dec_b = nullptr, *bih[kMaxLayers] = {}, *bhh[kMaxLayers] = {};
    float2 *tables = nullptr;   // lane-major FFT constants
    // state
    int16_t *tail = nullptr;
    float *ola[2] = {};          // [Bp][256]; fp32 path: both entries are the same buffer
    float *h[2] = {};            // [L][Bp][H]
    __nv_bfloat16 *hb[2] = {};   // [L][Bp][H]   (bf16 path)
    // scratch
    void *feat = nullptr;        // fp32 | bf16 [tcap][Bp][256]
    float *spec = nullptr;       // [tcap][Bp][512]
    void *e = nullptr;           // fp32 | bf16 [e_ring][Bp][H]
    float *mask = nullptr;       // [tcap][Bp][256]
    // staging for host-buffer calls
    int16_t *d_in[kHostRing] = {}, *d_out[kHostOutRing] = {};   // [B][staging_frames][256] / [B][staging_out_frames][256] each
    size_t staging_frames = 0, staging_out_frames = 0;
    int16_t *d_small = nullptr;  // in | out staging of calls small enough to be one chunk (process_host)
    size_t small_samples = 0;
    cudaStream_t copy_in = nullptr, copy_out = nullptr;
    cudaEvent_t ev_in[kHostRing] = {}, ev_comp[kHostRing] = {}, ev_out[kHostOutRing] = {};
    FuPlan *fu = nullptr;        // tcgen05 path: packed weights, tensor maps and the tile schedule of the fused mask-estimator kernel
    I8Plan *i8 = nullptr;        // fixed-point path: quantised weights, byte-plane activations and state, per-layer launches
    uint8_t *arena = nullptr;    // bf16 path: all per-stream state in one allocation
    size_t arena_bytes = 0;
    KernelProfiler *prof = nullptr;
    std::vector<void *> allocs;
    // Ordering of state-mutating work across streams: every step reads and writes the per-stream state rows (and the fused
    // kernel's dependency counters), so work enqueued on a new stream must run after what was enqueued on the previous one.
    // The event is recorded lazily, on the PREVIOUS stream at the moment the stream changes (an event record between two
    // kernels of the same stream would break their programmatic-dependent-launch overlap).
    cudaStream_t last_stream = nullptr;
    bool has_last = false;
    cudaEvent_t ev_order = nullptr;
    // A big batch is a composite: `parts` are engines of <= kPartStreams streams each (streams part_first[k] ..), which own all
    // device state; the composite keeps only the host ingest path's streams and staging buffers.
    std::vector<Engine *> parts;
    std::vector<int> part_first;
};

// Makes `st` wait for everything enqueued so far on `prev` (falls back to a device synchronisation if `prev` is no longer a
// valid stream, e.g. a caller's stream that has since been destroyed).
static cudaError_t chain_streams(cudaStream_t prev, cudaStream_t st, cudaEvent_t *ev) {
    if (prev == st) return cudaSuccess;
    if (!*ev) {
        cudaError_t e = cudaEventCreateWithFlags(ev, cudaEventDisableTiming);
        if (e != cudaSuccess) return e;
    }
    if (cudaEventRecord(*ev, prev) != cudaSuccess) {
        cudaGetLastError();
        return cudaDeviceSynchronize();
    }
    return cudaStreamWaitEvent(st, *ev, 0);
}

// The fused mask-estimator kernel spins on dependency counters and relies on its whole grid being co-resident
// (masknet_fused.cuh).  Two such grids from different engines (every pv_koala_t / pv_koala_batch_t owns one) running
// concurrently on one device could each hold part of the SMs and wait for clusters that cannot be scheduled.  Fused launches
// of different engines of this process are therefore serialised per device: an engine that launches after another one makes its
// stream wait for the other engine's stream first.  One engine on one stream (the steady state) never pays for this.
// (Other PROCESSES sharing the device through MPS are outside what a library can order; see include/pv_koala_b200.h.)
namespace {
struct FusedOwner {
    const void *impl = nullptr;
    cudaStream_t stream = nullptr;
};
std::mutex g_fused_mu;
FusedOwner g_fused_last[64];
}  // namespace

static cudaError_t serialize_fused_launches(int device, const void *impl, cudaStream_t st, cudaEvent_t *ev) {
    if (device < 0 || device >= 64) return cudaSuccess;
    std::lock_guard<std::mutex> lock(g_fused_mu);
    FusedOwner &o = g_fused_last[device];
    cudaError_t e = cudaSuccess;
    if (o.impl && o.impl != impl) e = chain_streams(o.stream, st, ev);
    o.impl = impl;
    o.stream = st;
    return e;
}
static void forget_fused_owner(int device, const void *impl) {
    if (device < 0 || device >= 64) return;
    std::lock_guard<std::mutex> lock(g_fused_mu);
    if (g_fused_last[device].impl == impl) g_fused_last[device] = FusedOwner();
}

#define KCHECK(expr)                                                                                      \
    do {                                                                                                  \
        cudaError_t e__ = (expr);                                                                         \
        if (e__ != cudaSuccess) {                                                                         \
            if (errors) errors->push_back(std::string(#expr) + " failed: " + cudaGetErrorString(e__));    \
            return e__ == cudaErrorMemoryAllocation ? kOutOfMemory : kRuntimeError;                       \
        }                                                                                                 \
    } while (0)

template <typename T>
static cudaError_t dev_alloc(std::vector<void *> &allocs, T **ptr, size_t count, bool zero = true) {
    cudaError_t e = cudaMalloc((void **) ptr, count * sizeof(T));
    if (e != cudaSuccess) return e;
    allocs.push_back(*ptr);
    return zero ? cudaMemset(*ptr, 0, count * sizeof(T)) : cudaSuccess;
}

template <typename T, typename S>
static cudaError_t upload(std::vector<void *> &allocs, T **ptr, const std::vector<S> &src) {
    static_assert(sizeof(T) == sizeof(S), "element size");
    cudaError_t e = dev_alloc(allocs, ptr, src.size(), false);
    if (e != cudaSuccess) return e;
    return cudaMemcpy(*ptr, src.data(), src.size() * sizeof(S), cudaMemcpyHostToDevice);
}

Status Engine::create(const ModelHost &model, int device, int num_streams, int precision, Engine **out,
                      std::vector<std::string> *errors) {
    if (num_streams < 1 || (precision != kFp32 && precision != kBf16 && precision != kInt8)) {
        errors->push_back("Invalid number of streams or precision.");
        return kInvalidArgument;
    }
    KCHECK(cudaSetDevice(device));
    cudaDeviceProp prop;
    KCHECK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        // the reference probes its device with `compatibility_test_kernel` and fails the same way
        errors->push_back("Selected GPU device is incompatible with the library (needs compute capability 10.x).");
        return kRuntimeError;
    }
    Engine *eng = new Engine();
    Impl *p = eng->p_ = new Impl();
    eng->n_ = num_streams;
    eng->npad_ = (num_streams + 255) / 256 * 256;
    eng->device_ = device;
    eng->precision_ = precision;
    const size_t Bp = eng->npad_, H = model.hidden, L = model.layers;
    p->H = (int) H;
    p->L = (int) L;
    p->num_sms = prop.multiProcessorCount;
    {
        const char *pe = getenv("KOALA_PARTITION_STREAMS");
        const int part = pe ? std::max(0, atoi(pe)) / 256 * 256 : kPartStreams;
        const char *cc = getenv("KOALA_FP32_CUDA_CORES");
        const bool tensor_path = precision == kBf16 || (precision == kFp32 && H % 256 == 0 && !(cc && cc[0] == '1'));
        if (tensor_path && part > 0 && num_streams > part) {
            Status st = [&]() -> Status {
                KCHECK(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
                return kSuccess;
            }();
            p->tcap = 1 << 30;
            for (int s0 = 0; s0 < num_streams && st == kSuccess; s0 += part) {
                Engine *sub = nullptr;
                st = Engine::create(model, device, std::min(part, num_streams - s0), precision, &sub, errors);
                if (st == kSuccess) {
                    p->parts.push_back(sub);
                    p->part_first.push_back(s0);
                    p->tcap = std::min(p->tcap, sub->chunk_frames());
                }
            }
            if (st != kSuccess) {
                delete eng;
                return st;
            }
            *out = eng;
            return kSuccess;
        }
    }
    if (const char *e = getenv("KOALA_STFT_PER_WARP")) p->stft_per_warp = std::max(1, atoi(e));
    Status st = [&]() -> Status {
        KCHECK(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
        KCHECK(upload(p->allocs, &p->enc_w, model.enc_w));
        KCHECK(upload(p->allocs, &p->enc_b, model.enc_b));
        KCHECK(upload(p->allocs, &p->dec_w, model.dec_w));
        KCHECK(upload(p->allocs, &p->dec_b, model.dec_b));
        for (size_t l = 0; l < L; l++) {
            KCHECK(upload(p->allocs, &p->wih[l], model.wih[l]));
            KCHECK(upload(p->allocs, &p->whh[l], model.whh[l]));
            KCHECK(upload(p->allocs, &p->bih[l], model.bih[l]));
            KCHECK(upload(p->allocs, &p->bhh[l], model.bhh[l]));
        }
        // lane-major constants of the warp FFT (FftLane / warp_fft256 in koala_common.cuh): float2 [25][32], computed in double.
        // Rows 0-10: forward set, rows 11-21: inverse set (conjugate twiddles, bit-reversed lane index), row 22: split base,
        // rows 23-24: window base angles.
        std::vector<float2> tab((size_t) kLaneTabRows * 32);
        auto rev5h = [](int v) { int r = 0; for (int b = 0; b < 5; b++) r |= ((v >> b) & 1) << (4 - b); return r; };
        for (int dir = 0; dir < 2; dir++) {
            auto w256 = [dir](int e, double sign) {       // W256^e (forward) or its conjugate (inverse), times sign
                const double a = -2.0 * M_PI * (double) (e & 255) / 256.0;
                return make_float2((float) (sign * cos(a)), (float) (sign * (dir == 0 ? sin(a) : -sin(a))));
            };
            float2 *t = tab.data() + (size_t) (dir == 0 ? 0 : kLaneTabInv) * 32;
            for (int lane = 0; lane < 32; lane++) {
                const int base = dir == 0 ? lane : rev5h(lane);    // low five index bits held by this lane
                for (int q = 0; q < 4; q++) t[(0 + q) * 32 + lane] = w256(base + 32 * q, 1.0);
                for (int q = 0; q < 2; q++) t[(4 + q) * 32 + lane] = w256(2 * (base + 32 * q), 1.0);
                t[6 * 32 + lane] = w256(4 * base, 1.0);
                for (int i = 0; i < 4; i++) {                      // swapped stages on index bit 4 - i
                    const int bit = 4 - i, h = dir == 0 ? (16 >> i) : (1 << i);
                    t[(7 + i) * 32 + lane] = w256((base & ((1 << bit) - 1)) << (7 - bit), (lane & h) ? -1.0 : 1.0);
                }
            }
        }
        for (int lane = 0; lane < 32; lane++) {
            const double a = -2.0 * M_PI * (double) rev5h(lane) / kNfft;
            tab[(size_t) kLaneTabSplit * 32 + lane] = make_float2((float) cos(a), (float) sin(a));
            for (int o = 0; o < 2; o++) {
                const double th = M_PI * (double) (2 * lane + o) / kNfft;
                tab[(size_t) (kLaneTabWin + o) * 32 + lane] = make_float2((float) sin(th), (float) cos(th));
            }
        }
        KCHECK(upload(p->allocs, &p->tables, tab));
        // fp32 mode runs on the tensor cores too when the model fits the fused kernel's tiles (three bf16 planes per operand);
        // KOALA_FP32_CUDA_CORES=1 keeps the CUDA-core kernels (comparison runs)
        const char *cc = getenv("KOALA_FP32_CUDA_CORES");
        const bool fused = precision == kBf16 || (precision == kFp32 && H % 256 == 0 && !(cc && cc[0] == '1'));
        p->planes = precision == kFp32 ? 3 : 1;
        if (fused) {
            // frames per chunk: as many as keep feat + spec + mask under 512 MB, at most 64 (KOALA_CHUNK_FRAMES overrides)
            const size_t slot_bytes = Bp * (kNfft * 4 + kBins * 4 + kBins * 2 * p->planes);
            int cap = 64;
            while (cap > 1 && (size_t) cap * slot_bytes > ((size_t) 512 << 20)) cap >>= 1;
            if (const char *e = getenv("KOALA_CHUNK_FRAMES")) cap = std::max(1, std::min(256, atoi(e)));
            p->tcap = cap;
            p->e_ring = cap >= 4 ? 4 : cap >= 2 ? 2 : 1;      // a power of two <= kFuSlots
            if (const char *e = getenv("KOALA_E_RING")) p->e_ring = atoi(e) >= 2 && p->e_ring >= 2 ? 2 : 1;
        }
        const size_t T = p->tcap, P = p->planes;
        KCHECK(dev_alloc(p->allocs, &p->spec, T * Bp * kNfft));
        KCHECK(dev_alloc(p->allocs, &p->mask, T * Bp * kBins));
        if (!fused) {
            KCHECK(dev_alloc(p->allocs, &p->tail, Bp * kFrame));
            KCHECK(dev_alloc(p->allocs, &p->ola[0], Bp * kFrame));
            p->ola[1] = p->ola[0];
            if (precision == kFp32)
                for (int i = 0; i < 2; i++) KCHECK(dev_alloc(p->allocs, &p->h[i], L * Bp * H));
        } else {
            // One arena for everything that survives a step: fp32 h (updated IN PLACE: a GRU tile reads and writes only its own
            // [256 streams x 64 units] slice; the other tiles read the bf16 copies, which stay ping-pong), both bf16 copies,
            // the overlap-add halves and the analysis tail.  In-place h takes 32 MB off the ~160 MB a step of 8192 streams touches
            // (126 MB L2): 90.4 -> 83.7 us per step.  (A persisting L2 access-policy window over the arena was tried and made
            // the step 44 % SLOWER -- the set-aside starves the per-step scratch -- so the arena uses the normal policy.)
            const size_t h_bytes = L * Bp * H * sizeof(float), hb_bytes = L * Bp * P * H * sizeof(__nv_bfloat16);
            const size_t ola_bytes = Bp * kFrame * sizeof(float), tail_bytes = Bp * kFrame * sizeof(int16_t);
            p->arena_bytes = h_bytes + 2 * hb_bytes + 2 * ola_bytes + tail_bytes;
            KCHECK(dev_alloc(p->allocs, &p->arena, p->arena_bytes));
            uint8_t *a = p->arena;
            p->h[0] = p->h[1] = (float *) a; a += h_bytes;
            for (int i = 0; i < 2; i++) { p->hb[i] = (__nv_bfloat16 *) a; a += hb_bytes; }
            for (int i = 0; i < 2; i++) { p->ola[i] = (float *) a; a += ola_bytes; }
            p->tail = (int16_t *) a;
        }
        if (precision == kInt8) {
            std::string why;
            if (!i8_plan_create(model, (int) Bp, p->mask, &p->i8, &why)) {
                errors->push_back("Failed to set up the fixed-point mask path: " + why);
                return kRuntimeError;
            }
        } else if (!fused) {
            KCHECK(dev_alloc(p->allocs, (float **) &p->feat, Bp * kBins));
            KCHECK(dev_alloc(p->allocs, (float **) &p->e, Bp * H));
        } else {
            KCHECK(dev_alloc(p->allocs, (__nv_bfloat16 **) &p->feat, T * Bp * P * kBins));
            KCHECK(dev_alloc(p->allocs, (__nv_bfloat16 **) &p->e, (size_t) p->e_ring * Bp * P * H));
            TcModel tm;
            tm.H = (int) H; tm.L = (int) L; tm.Bp = (int) Bp; tm.tcap = p->tcap; tm.e_ring = p->e_ring; tm.planes = p->planes;
            tm.enc_w = p->enc_w; tm.dec_w = p->dec_w; tm.enc_b = p->enc_b; tm.dec_b = p->dec_b;
            for (size_t l = 0; l < L; l++) { tm.wih[l] = p->wih[l]; tm.whh[l] = p->whh[l]; tm.bih[l] = p->bih[l]; tm.bhh[l] = p->bhh[l]; }
            tm.feat = (__nv_bfloat16 *) p->feat; tm.e = (__nv_bfloat16 *) p->e; tm.mask = p->mask;
            for (int i = 0; i < 2; i++) { tm.h[i] = p->h[i]; tm.hb[i] = p->hb[i]; }
            std::string why;
            if (!fu_plan_create(tm, &p->fu, &why)) {
                errors->push_back("Failed to set up the tensor-core mask path: " + why);
                return kRuntimeError;
            }
        }
        KCHECK(cudaDeviceSynchronize());
        return kSuccess;
    }();
    if (st != kSuccess) {
        delete eng;
        return st;
    }
    *out = eng;
    return kSuccess;
}

Engine::~Engine() {
    if (!p_) return;
    cudaSetDevice(device_);
    for (Engine *sub : p_->parts) delete sub;
    if (p_->has_last && p_->last_stream != p_->stream) cudaDeviceSynchronize();   // work may still be queued on a caller's stream
    cudaGetLastError();
    if (p_->stream) cudaStreamSynchronize(p_->stream);
    forget_fused_owner(device_, p_);
    if (p_->ev_order) cudaEventDestroy(p_->ev_order);
    if (p_->fu) fu_plan_destroy(p_->fu);
    if (p_->i8) i8_plan_destroy(p_->i8);
    delete p_->prof;
    for (void *a : p_->allocs) cudaFree(a);
    for (int i = 0; i < kHostRing; i++) {
        if (p_->d_in[i]) cudaFree(p_->d_in[i]);
        if (p_->ev_in[i]) cudaEventDestroy(p_->ev_in[i]);
        if (p_->ev_comp[i]) cudaEventDestroy(p_->ev_comp[i]);
    }
    for (int i = 0; i < kHostOutRing; i++) {
        if (p_->d_out[i]) cudaFree(p_->d_out[i]);
        if (p_->ev_out[i]) cudaEventDestroy(p_->ev_out[i]);
    }
    if (p_->d_small) cudaFree(p_->d_small);
    if (p_->copy_in) cudaStreamDestroy(p_->copy_in);
    if (p_->copy_out) cudaStreamDestroy(p_->copy_out);
    if (p_->stream) cudaStreamDestroy(p_->stream);
    delete p_;
}

// STFT launches: the grid's warps walk the launch's items with a grid stride.
// * Many items per warp (the chunked launches: 4096 streams x 32 frames = 37 per resident warp): exactly ONE wave of CTAs, so
//   every warp gets the same count to within one.  With the count rounded down the grid was 911 CTAs where 888 fit at once, and the
//   23 left over ran a second, almost empty wave as long as the first (frontend 5.15 -> 3.78 us per 4096-stream frame, r02x).
// * A few items per warp (one frame per launch): items / resident warps rounded DOWN, i.e. a grid slightly larger than what fits at
//   once, whose CTAs are short and let the next kernel of the programmatic-dependent-launch chain move in earlier (8192 streams:
//   2 per warp, 82.4 vs 83.6 us per step).
static int stft_grid_for(int items, int num_sms, int per_warp_override) {
    const int resident_ctas = num_sms * kStftCtasPerSm, resident_warps = resident_ctas * kStftWarps;
    if (per_warp_override <= 0 && items > 3 * resident_warps) return resident_ctas;
    const int per_warp = per_warp_override > 0 ? per_warp_override : std::max(1, items / resident_warps);
    return std::max(1, (items + kStftWarps * per_warp - 1) / (kStftWarps * per_warp));
}

// Synthesis of a chunk of `frames` frames of B streams: an item is a RUN of consecutive frames of one stream (the overlap-add half
// stays in registers inside a run; a run that does not start the chunk re-synthesises the frame before it: one extra transform).
// Picks the run length that minimises the transforms the busiest warp does -- ceil(items / resident warps) runs of run (+ 1)
// transforms -- e.g. 4096 streams x 32 frames: 4 runs of 8 frames (5 items x 9 = 45 transforms per warp; 2 runs of 16 in a grid
// of 1024 CTAs, as before, was two waves of 2 x 17 = 68).
static int synthesis_run_for(int B, int frames, int num_sms) {
    const long long resident_warps = (long long) num_sms * kStftCtasPerSm * kStftWarps;
    int best_run = frames;
    long long best_cost = -1;
    for (int runs = 1; runs <= frames; runs++) {
        const int run = (frames + runs - 1) / runs, actual = (frames + run - 1) / run;
        if (actual != runs) continue;                       // the same run length as a smaller count: already seen
        const long long items = (long long) B * actual, per_warp = std::max(1ll, (items + resident_warps - 1) / resident_warps);
        const long long cost = per_warp * (run + (actual > 1 ? 1 : 0));
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_run = run; }
    }
    return best_run;
}

Status Engine::process_device(const int16_t *pcm, int16_t *out, int frames, long long stride, void *stream_,
                              std::vector<std::string> *errors, long long out_stride, long long frame_stride, long long out_frame_stride) {
    if (out_stride == 0) out_stride = stride;
    if (frame_stride == 0) frame_stride = kFrame;
    if (out_frame_stride == 0) out_frame_stride = frame_stride;
    if (!pcm || !out || frames < 0 || stride < kFrame || (stride & 7) || out_stride < kFrame || (out_stride & 7) || frame_stride < kFrame ||
        (frame_stride & 7) || out_frame_stride < kFrame || (out_frame_stride & 7) || ((uintptr_t) pcm & 15) || ((uintptr_t) out & 15)) {
        if (errors) errors->push_back("PCM buffers must be non-NULL, 16-byte aligned, with stream and frame strides >= 256 and multiples of 8.");
        return kInvalidArgument;
    }
    Impl *p = p_;
    KCHECK(cudaSetDevice(device_));
    cudaStream_t st = (cudaStream_t) stream_;
    if (frames == 0) return kSuccess;
    if (p->has_last) KCHECK(chain_streams(p->last_stream, st, &p->ev_order));
    p->last_stream = st;
    p->has_last = true;
    if (!p->parts.empty()) {       // chunk by chunk, partition after partition
        for (int t0 = 0; t0 < frames; t0 += p->tcap) {
            const int tc = std::min(p->tcap, frames - t0);
            for (size_t k = 0; k < p->parts.size(); k++) {
                const long long s0 = p->part_first[k];
                const Status r = p->parts[k]->process_device(pcm + s0 * stride + (long long) t0 * frame_stride, out + s0 * out_stride + (long long) t0 * out_frame_stride,
                                                              tc, stride, stream_, errors, out_stride, frame_stride, out_frame_stride);
                if (r != kSuccess) return r;
            }
        }
        launches_ = 0;
        for (Engine *sub : p->parts) launches_ += sub->kernel_launches();
        return kSuccess;
    }
    if (p->fu) KCHECK(serialize_fused_launches(device_, p, st, &p->ev_order));
    const int B = n_, Bp = npad_, H = p->H, L = p->L;
    const size_t LBH = (size_t) Bp * H;
    KernelProfiler *prof = p->prof;
    static const bool only_masknet = [] { const char *e = getenv("KOALA_B200_ONLY_MASKNET"); return e && e[0] == '1'; }();
    if (p->i8) {           // fixed-point mode: analysis, encoder, GRU layers, decoder, synthesis -- one launch each per frame
        const int grid = stft_grid_for(B, p->num_sms, p->stft_per_warp);
        for (int t = 0; t < frames; t++) {
            PcmView v{pcm, out, stride, out_stride, frame_stride, out_frame_stride, t};
            if (prof) prof->begin(kKernFrontend, st);
            launch_pdl(false, frontend_kernel<uint8_t, 2>, dim3(grid), dim3(kStftWarps * 32), 0, st, v, B, 1, (long long) Bp, p->tail, p->spec, p->i8->featp, p->tables);
            if (prof) prof->end(st);
            launches_ += 2 + i8_masknet_step(p->i8, p->parity, st, prof);
            if (prof) prof->begin(kKernBackend, st);
            launch_pdl(false, backend_kernel, dim3(grid), dim3(kStftWarps * 32), 0, st, v, B, 1, 1, (long long) Bp, p->spec, p->mask, p->ola[0], p->ola[0],
                                                                            p->tail, p->tables);
            if (prof) prof->end(st);
            p->parity ^= 1;
        }
        KCHECK(cudaGetLastError());
        return kSuccess;
    }
    if (!p->fu) {          // fp32 mode, hidden size the fused kernel's tiles do not fit: CUDA-core kernels, frame by frame
        const int grid = stft_grid_for(B, p->num_sms, p->stft_per_warp);
        float *feat = (float *) p->feat, *e = (float *) p->e;
        for (int t = 0; t < frames; t++) {
            PcmView v{pcm, out, stride, out_stride, frame_stride, out_frame_stride, t};
            const int cur = p->parity, nxt = cur ^ 1;
            if (prof) prof->begin(kKernFrontend, st);
            launch_pdl(false, frontend_kernel<float, 1>, dim3(grid), dim3(kStftWarps * 32), 0, st, v, B, 1, (long long) Bp, p->tail, p->spec, feat, p->tables);
            if (prof) { prof->end(st); prof->begin(kKernEnc, st); }
            launch_pdl(false, linear_fp32_kernel<kActRelu>, dim3(Bp / kF32Bm, H / kF32LinN), dim3(256), 0, st, feat, p->enc_w, p->enc_b, e, kBins, H);
            if (prof) prof->end(st);
            const float *x = e;
            for (int l = 0; l < L; l++) {
                if (prof) prof->begin(kKernGru, st);
                launch_pdl(false, gru_fp32_kernel, dim3(Bp / kF32Bm, H / 16), dim3(256), 0, st, x, p->h[cur] + l * LBH, p->h[nxt] + l * LBH,
                                                                            p->wih[l], p->whh[l], p->bih[l], p->bhh[l], H);
                if (prof) prof->end(st);
                x = p->h[nxt] + l * LBH;
            }
            if (prof) prof->begin(kKernDec, st);
            launch_pdl(false, linear_fp32_kernel<kActSigmoid>, dim3(Bp / kF32Bm, kBins / kF32LinN), dim3(256), 0, st, x, p->dec_w, p->dec_b, p->mask, H, kBins);
            if (prof) { prof->end(st); prof->begin(kKernBackend, st); }
            launch_pdl(false, backend_kernel, dim3(grid), dim3(kStftWarps * 32), 0, st, v, B, 1, 1, (long long) Bp, p->spec, p->mask, p->ola[0], p->ola[0],
                                                                            p->tail, p->tables);
            if (prof) prof->end(st);
            launches_ += 4 + L;
            p->parity = nxt;
        }
        KCHECK(cudaGetLastError());
        return kSuccess;
    }
    // tensor-core path: chunks of up to tcap frames, three launches per chunk chained with programmatic dependent launch
    for (int t0 = 0; t0 < frames; t0 += p->tcap) {
        const int tc = std::min(p->tcap, frames - t0);
        PcmView v{pcm, out, stride, out_stride, frame_stride, out_frame_stride, t0};
        if (only_masknet) {      // tuning aid (KOALA_B200_ONLY_MASKNET=1): the fused kernel alone, back to back, on stale features
            if (prof) prof->begin(kKernMasknet, st);
            launches_ += fu_masknet_steps(p->fu, p->parity, tc, st);
            if (prof) prof->end(st);
            p->parity ^= tc & 1;
            continue;
        }
        if (prof) prof->begin(kKernFrontend, st);
        launch_pdl(true, p->planes == 1 ? frontend_kernel<__nv_bfloat16, 1> : frontend_kernel<__nv_bfloat16, 3>,
                   dim3(stft_grid_for(B * tc, p->num_sms, p->stft_per_warp)), dim3(kStftWarps * 32), 0, st, v, B, tc,
                   (long long) Bp, p->tail, p->spec, (__nv_bfloat16 *) p->feat, p->tables);
        if (prof) { prof->end(st); prof->begin(kKernMasknet, st); }
        launches_ += 1 + fu_masknet_steps(p->fu, p->parity, tc, st);
        if (prof) { prof->end(st); prof->begin(kKernBackend, st); }
        const int run = synthesis_run_for(B, tc, p->num_sms), runs = (tc + run - 1) / run;
        launch_pdl(true, backend_kernel, dim3(stft_grid_for(B * runs, p->num_sms, p->stft_per_warp)), dim3(kStftWarps * 32), 0, st, v, B, tc, run,
                   (long long) Bp, p->spec, p->mask, p->ola[p->ola_par], p->ola[p->ola_par ^ 1], p->tail, p->tables);
        if (prof) prof->end(st);
        launches_ += 1;
        p->ola_par ^= 1;
        p->parity ^= tc & 1;
        p->last_slot = tc - 1;
    }
    KCHECK(cudaGetLastError());
    return kSuccess;
}

// Host ingest path.  The caller's [B][frames][256] host buffers are cut into INPUT chunks of
// kHostChunkFrames frames (host -> device, then computed) and OUTPUT blocks of kHostOutFrames frames (device -> host), each on
// its own stream through rings of device staging buffers, so that chunk c+1 travels in and the previous block travels out
// while chunk c is being processed.  The two sizes differ because of what the copy engine does under load (tools/d2h_probe.py,
// B200, PCIe 5): both directions are pitched 2-D copies whose contiguous run is one stream's frames of the chunk; host ->
// device keeps 51 GiB/s with 4 KB runs whatever the SMs do, device -> host drops to 31 GiB/s with 4-8 KB runs while the step's
// kernels are running and needs >= 16 KB runs (32 frames) for 44 GiB/s.  The last block is sent chunk by chunk instead: nothing
// competes with those copies once the compute has drained, and the call ends one chunk, not one block, after the last step.
// Short calls (< kHostBlockMinFrames) keep chunk-sized output copies: a block would leave too late to be hidden.
// TIME-MAJOR host buffers ([frames][B][256], what a caller that collects one frame per stream per tick has anyway) avoid the
// problem altogether: a chunk is one contiguous range on both sides, every copy runs at ~50 GiB/s in both directions under
// load, chunks can be short (kHostChunkFramesTm frames: 0.3 ms to fill and to drain the pipeline) and the call is bound by
// max(compute, PCIe): 4 MiB in + 4 MiB out per step of 8192 streams is ~83 us at 47 GiB/s each way, the step itself 83 us.
// With pinned host memory the copies are true DMA; pageable memory still works (the copies just serialise).
Status Engine::process_host(const int16_t *pcm, int16_t *out, int frames, std::vector<std::string> *errors, bool time_major) {
    if (!pcm || !out || frames < 0) {
        if (errors) errors->push_back("PCM buffers must be non-NULL.");
        return kInvalidArgument;
    }
    if (frames == 0) return kSuccess;
    Impl *p = p_;
    KCHECK(cudaSetDevice(device_));
    const int chunk_forced = [] { const char *e = getenv("KOALA_HOST_CHUNK"); return e ? std::max(0, atoi(e)) : 0; }();      // tests / tuning: exactly this many frames per input chunk
    // A call that is one chunk anyway (the reference-shaped use: pv_koala_process, one frame of one stream) has nothing to overlap:
    // copy in, step, copy out on the engine's own stream, one synchronisation -- no copy streams, no events (95 -> ~75 us per call).
    if (chunk_forced <= 0 && (size_t) n_ * frames * kFrame * sizeof(int16_t) <= ((size_t) 64 << 10)) {
        const size_t samples = (size_t) n_ * frames * kFrame;
        if (samples > p->small_samples) {
            if (p->d_small) cudaFree(p->d_small);
            p->d_small = nullptr;
            p->small_samples = 0;
            KCHECK(cudaMalloc((void **) &p->d_small, 2 * samples * sizeof(int16_t)));
            p->small_samples = samples;
        }
        int16_t *d_in = p->d_small, *d_out = p->d_small + p->small_samples;
        KCHECK(cudaMemcpyAsync(d_in, pcm, samples * sizeof(int16_t), cudaMemcpyHostToDevice, p->stream));
        const Status st = time_major ? process_device(d_in, d_out, frames, kFrame, p->stream, errors, kFrame, (long long) n_ * kFrame, (long long) n_ * kFrame)
                                     : process_device(d_in, d_out, frames, (long long) frames * kFrame, p->stream, errors);
        if (st != kSuccess) return st;
        KCHECK(cudaMemcpyAsync(out, d_out, samples * sizeof(int16_t), cudaMemcpyDeviceToHost, p->stream));
        KCHECK(cudaStreamSynchronize(p->stream));
        return kSuccess;
    }
    static const int out_env = [] { const char *e = getenv("KOALA_HOST_OUT_CHUNK"); const int v = e ? atoi(e) : 0; return v > 0 ? v : kHostOutFrames; }();
    // Small batches: a chunk of a few frames is a few hundred KB and the call becomes bound by the host's per-chunk work (two copies,
    // three launches, four events: ~40 us) -- scale the chunk to ~4 MiB of PCM, up to what one fused launch walks (128 streams: 64 frames)
    // ... but never so long that the call has fewer than ~4 chunks to overlap its copies with its compute
    const int by_bytes = (int) std::min<size_t>((size_t) std::max(p->tcap, 1), ((size_t) 4 << 20) / ((size_t) n_ * kFrame * sizeof(int16_t)));
    const int base_chunk = time_major ? kHostChunkFramesTm : kHostChunkFrames;
    const int chunk_frames = chunk_forced > 0 ? chunk_forced : std::max(base_chunk, std::min(by_bytes, std::max(base_chunk, frames / 4)));
    const int Tc = frames < chunk_frames ? frames : chunk_frames;                 // frames per input chunk
    const int per_block = (time_major || frames < kHostBlockMinFrames) ? 1 : std::max(1, std::min(out_env, frames) / Tc);   // input chunks per output block
    const int To = per_block * Tc;                                                // frames per output block
    // staging is sized for the layout's full chunk / block, not for this call's length: a short call must not make the next long one reallocate
    const int full_chunk = chunk_forced > 0 ? chunk_forced : std::max(base_chunk, by_bytes);
    const int need_in = full_chunk, need_out = std::max(To, (time_major ? 1 : std::max(1, out_env / full_chunk)) * full_chunk);
    if ((size_t) need_in > p->staging_frames || (size_t) need_out > p->staging_out_frames) {
        for (int i = 0; i < kHostRing; i++) {
            if (p->d_in[i]) cudaFree(p->d_in[i]);
            p->d_in[i] = nullptr;
        }
        for (int i = 0; i < kHostOutRing; i++) {
            if (p->d_out[i]) cudaFree(p->d_out[i]);
            p->d_out[i] = nullptr;
        }
        p->staging_frames = p->staging_out_frames = 0;
        for (int i = 0; i < kHostRing; i++) KCHECK(cudaMalloc((void **) &p->d_in[i], (size_t) n_ * need_in * kFrame * sizeof(int16_t)));
        for (int i = 0; i < kHostOutRing; i++) KCHECK(cudaMalloc((void **) &p->d_out[i], (size_t) n_ * need_out * kFrame * sizeof(int16_t)));
        p->staging_frames = need_in;
        p->staging_out_frames = need_out;
    }
    if (!p->copy_in) {
        KCHECK(cudaStreamCreateWithFlags(&p->copy_in, cudaStreamNonBlocking));
        KCHECK(cudaStreamCreateWithFlags(&p->copy_out, cudaStreamNonBlocking));
        for (int i = 0; i < kHostRing; i++) {
            KCHECK(cudaEventCreateWithFlags(&p->ev_in[i], cudaEventDisableTiming));
            KCHECK(cudaEventCreateWithFlags(&p->ev_comp[i], cudaEventDisableTiming));
        }
        for (int i = 0; i < kHostOutRing; i++) KCHECK(cudaEventCreateWithFlags(&p->ev_out[i], cudaEventDisableTiming));
    }
    const size_t in_pitch = (size_t) p->staging_frames * kFrame * sizeof(int16_t);        // device input staging: [B][Tc][256]
    const size_t out_pitch = (size_t) p->staging_out_frames * kFrame * sizeof(int16_t);   // device output staging: [B][To][256]
    const size_t hpitch = (size_t) frames * kFrame * sizeof(int16_t);                     // host: [B][frames][256]
    const int chunks = (frames + Tc - 1) / Tc, blocks = (chunks + per_block - 1) / per_block;
    for (int c = 0; c < chunks; c++) {
        const int ib = c % kHostRing, blk = c / per_block, ob = blk % kHostOutRing, slot = c % per_block;
        const int t0 = c * Tc, tc = frames - t0 < Tc ? frames - t0 : Tc;
        const bool block_end = slot == per_block - 1 || c == chunks - 1, last_block = blk == blocks - 1;
        // host -> device once the compute that last used this input buffer is done
        if (c >= kHostRing) KCHECK(cudaStreamWaitEvent(p->copy_in, p->ev_comp[ib], 0));
        if (time_major)
            KCHECK(cudaMemcpyAsync(p->d_in[ib], pcm + (size_t) t0 * n_ * kFrame, (size_t) tc * n_ * kFrame * sizeof(int16_t), cudaMemcpyHostToDevice, p->copy_in));
        else
            KCHECK(cudaMemcpy2DAsync(p->d_in[ib], in_pitch, pcm + (size_t) t0 * kFrame, hpitch, (size_t) tc * kFrame * sizeof(int16_t), n_,
                                     cudaMemcpyHostToDevice, p->copy_in));
        KCHECK(cudaEventRecord(p->ev_in[ib], p->copy_in));
        // compute once the input has landed and (first chunk of a block) the block that last used this output buffer has left
        KCHECK(cudaStreamWaitEvent(p->stream, p->ev_in[ib], 0));
        if (slot == 0 && blk >= kHostOutRing) KCHECK(cudaStreamWaitEvent(p->stream, p->ev_out[ob], 0));
        Status st = kSuccess;
        if (time_major) {       // staging buffers are [tc][B][256]: streams 256 samples apart, frames B * 256
            st = process_device(p->d_in[ib], p->d_out[ob], tc, kFrame, p->stream, errors, kFrame, (long long) n_ * kFrame, (long long) n_ * kFrame);
        } else {
            st = process_device(p->d_in[ib], p->d_out[ob] + (size_t) slot * Tc * kFrame, tc, (long long) p->staging_frames * kFrame, p->stream,
                                errors, (long long) p->staging_out_frames * kFrame);
        }
        if (st != kSuccess) return st;
        KCHECK(cudaEventRecord(p->ev_comp[ib], p->stream));
        // device -> host: whole blocks, except the last block, which leaves chunk by chunk
        if (block_end || last_block) {
            const int first = last_block ? slot : 0;                                  // first input chunk of the block in this copy
            const int f0 = blk * To + first * Tc, nf = t0 + tc - f0;                  // frames [f0, f0 + nf) of the call
            KCHECK(cudaStreamWaitEvent(p->copy_out, p->ev_comp[ib], 0));
            if (time_major)
                KCHECK(cudaMemcpyAsync(out + (size_t) f0 * n_ * kFrame, p->d_out[ob], (size_t) nf * n_ * kFrame * sizeof(int16_t), cudaMemcpyDeviceToHost, p->copy_out));
            else
                KCHECK(cudaMemcpy2DAsync(out + (size_t) f0 * kFrame, hpitch, p->d_out[ob] + (size_t) first * Tc * kFrame, out_pitch,
                                         (size_t) nf * kFrame * sizeof(int16_t), n_, cudaMemcpyDeviceToHost, p->copy_out));
            if (block_end) KCHECK(cudaEventRecord(p->ev_out[ob], p->copy_out));
        }
    }
    KCHECK(cudaStreamSynchronize(p->copy_out));
    KCHECK(cudaStreamSynchronize(p->stream));
    return kSuccess;
}

__global__ void reset_streams_kernel(const int32_t *__restrict__ ids, int n_ids, int n_streams, int16_t *tail, float *ola, float *ola1,
                                     float *h0, float *h1, __nv_bfloat16 *hb0, __nv_bfloat16 *hb1, int H, int L, size_t LBH, int planes) {
    const int i = blockIdx.x;
    if (i >= n_ids) return;
    const int s = ids[i];
    if (s < 0 || s >= n_streams) return;
    for (int k = threadIdx.x; k < kFrame; k += blockDim.x) {
        tail[(size_t) s * kFrame + k] = 0;
        ola[(size_t) s * kFrame + k] = 0.0f;
        ola1[(size_t) s * kFrame + k] = 0.0f;
    }
    if (h0)
        for (int l = 0; l < L; l++)
            for (int k = threadIdx.x; k < H; k += blockDim.x) {
                const size_t idx = l * LBH + (size_t) s * H + k;
                h0[idx] = 0.0f;
                h1[idx] = 0.0f;
            }
    if (hb0)
        for (int l = 0; l < L; l++)
            for (int k = threadIdx.x; k < planes * H; k += blockDim.x) {
                const size_t idx = (l * LBH + (size_t) s * H) * planes + k;
                hb0[idx] = __float2bfloat16(0.0f);
                hb1[idx] = __float2bfloat16(0.0f);
            }
}

Status Engine::reset(const int32_t *stream_ids, int n, std::vector<std::string> *errors) {
    Impl *p = p_;
    KCHECK(cudaSetDevice(device_));
    if (!p->parts.empty()) {
        if (stream_ids) {
            if (n < 0) {
                if (errors) errors->push_back("Negative stream count.");
                return kInvalidArgument;
            }
            for (int i = 0; i < n; i++)
                if (stream_ids[i] < 0 || stream_ids[i] >= n_) {
                    if (errors) errors->push_back("Stream id out of range.");
                    return kInvalidArgument;
                }
        }
        for (size_t k = 0; k < p->parts.size(); k++) {
            Engine *sub = p->parts[k];
            if (!stream_ids) {
                const Status r = sub->reset(nullptr, 0, errors);
                if (r != kSuccess) return r;
                continue;
            }
            std::vector<int32_t> local;
            for (int i = 0; i < n; i++)
                if (stream_ids[i] >= p->part_first[k] && stream_ids[i] < p->part_first[k] + sub->num_streams()) local.push_back(stream_ids[i] - p->part_first[k]);
            if (!local.empty()) {
                const Status r = sub->reset(local.data(), (int) local.size(), errors);
                if (r != kSuccess) return r;
            }
        }
        return kSuccess;
    }
    const size_t Bp = npad_, H = p->H, L = p->L;
    // the state rows may still be in use by steps enqueued on a caller's stream (process_device): run after them
    if (p->has_last) KCHECK(chain_streams(p->last_stream, p->stream, &p->ev_order));
    p->last_stream = p->stream;
    p->has_last = true;
    if (!stream_ids) {
        KCHECK(cudaMemsetAsync(p->tail, 0, Bp * kFrame * sizeof(int16_t), p->stream));
        if (p->i8)
            for (int i = 0; i < 2; i++) KCHECK(cudaMemsetAsync(p->i8->hp[i], 0, L * Bp * 2 * H, p->stream));
        for (int i = 0; i < 2; i++) {
            KCHECK(cudaMemsetAsync(p->ola[i], 0, Bp * kFrame * sizeof(float), p->stream));
            if (p->h[i]) KCHECK(cudaMemsetAsync(p->h[i], 0, L * Bp * H * sizeof(float), p->stream));
            if (p->hb[i]) KCHECK(cudaMemsetAsync(p->hb[i], 0, L * Bp * p->planes * H * sizeof(__nv_bfloat16), p->stream));
        }
    } else {
        if (n < 0) {
            if (errors) errors->push_back("Negative stream count.");
            return kInvalidArgument;
        }
        for (int i = 0; i < n; i++)
            if (stream_ids[i] < 0 || stream_ids[i] >= n_) {
                if (errors) errors->push_back("Stream id out of range.");
                return kInvalidArgument;
            }
        if (n > 0) {
            int32_t *d_ids = nullptr;
            KCHECK(cudaMalloc((void **) &d_ids, n * sizeof(int32_t)));
            cudaError_t e1 = cudaMemcpyAsync(d_ids, stream_ids, n * sizeof(int32_t), cudaMemcpyHostToDevice, p->stream);
            // (fixed-point mode: the state is the two byte planes of h, 2 H bytes per stream and layer = H 2-byte elements)
            __nv_bfloat16 *hb0 = p->i8 ? (__nv_bfloat16 *) p->i8->hp[0] : p->hb[0], *hb1 = p->i8 ? (__nv_bfloat16 *) p->i8->hp[1] : p->hb[1];
            reset_streams_kernel<<<n, 128, 0, p->stream>>>(d_ids, n, n_, p->tail, p->ola[0], p->ola[1], p->h[0], p->h[1], hb0, hb1,
                                                          (int) H, (int) L, Bp * H, p->i8 ? 1 : p->planes);
            cudaError_t e2 = cudaStreamSynchronize(p->stream);
            cudaFree(d_ids);
            KCHECK(e1);
            KCHECK(e2);
        }
    }
    KCHECK(cudaStreamSynchronize(p->stream));
    return kSuccess;
}

void Engine::set_profile(bool on) {
    for (Engine *sub : p_->parts) sub->set_profile(on);
    if (!p_->parts.empty()) return;
    if (on && !p_->prof) p_->prof = new KernelProfiler();
    if (!on) {
        delete p_->prof;
        p_->prof = nullptr;
    }
}

Status Engine::profile_read(double *ms, long long *count, int n_classes, std::vector<std::string> *errors) {
    if (!p_->parts.empty()) {
        std::vector<double> m(std::max(n_classes, 0));
        std::vector<long long> c(std::max(n_classes, 0));
        for (int i = 0; i < n_classes; i++) { ms[i] = 0.0; count[i] = 0; }
        for (Engine *sub : p_->parts) {
            const Status r = sub->profile_read(m.data(), c.data(), n_classes, errors);
            if (r != kSuccess) return r;
            for (int i = 0; i < n_classes; i++) { ms[i] += m[i]; count[i] += c[i]; }
        }
        return kSuccess;
    }
    if (n_classes < kKernClasses) {
        if (errors) errors->push_back("`num_classes` must be at least 6 (analysis, encoder, GRU, decoder, synthesis, fused mask estimator).");
        return kInvalidArgument;
    }
    if (!p_->prof) {
        if (errors) errors->push_back("Profiling is not enabled.");
        return kInvalidState;
    }
    KCHECK(cudaSetDevice(device_));
    KCHECK(cudaDeviceSynchronize());
    for (int i = 0; i < n_classes; i++) { ms[i] = 0.0; count[i] = 0; }
    p_->prof->drain(ms, count);
    return kSuccess;
}

void *Engine::own_stream() const { return p_->stream; }
int Engine::chunk_frames() const { return p_->tcap; }

Status Engine::synchronize(std::vector<std::string> *errors) {
    KCHECK(cudaSetDevice(device_));
    KCHECK(cudaDeviceSynchronize());
    return kSuccess;
}

Status Engine::debug_read(const char *name, void *dst, size_t bytes, std::vector<std::string> *errors) {
    Impl *p = p_;
    KCHECK(cudaSetDevice(device_));
    if (!p->parts.empty()) {       // whole-batch reads only: every partition contributes its rows
        if (name && std::string(name) == "trace") return p->parts[0]->debug_read(name, dst, bytes, errors);
        if (bytes % (size_t) n_ != 0) {
            if (errors) errors->push_back("A partitioned engine hands out whole tensors only (size must be a multiple of the stream count).");
            return kInvalidArgument;
        }
        const size_t row = bytes / (size_t) n_;
        for (size_t k = 0; k < p->parts.size(); k++) {
            const Status r = p->parts[k]->debug_read(name, (uint8_t *) dst + (size_t) p->part_first[k] * row, (size_t) p->parts[k]->num_streams() * row, errors);
            if (r != kSuccess) return r;
        }
        return kSuccess;
    }
    KCHECK(cudaDeviceSynchronize());
    const size_t B = n_, Bp = npad_, H = p->H;
    const void *src = nullptr;
    size_t avail = 0;
    const size_t esz = precision_ == kFp32 ? 4 : 2;
    const std::string nm(name ? name : "");
    // scratch of the last finished step: slot last_slot of the chunk buffers, ring slot (steps so far) % ring of e
    const size_t slot = p->fu ? (size_t) p->last_slot : 0;
    if (p->i8) {
        // fixed-point mode: activations and state are byte planes [rows][hi K | lo K]; hand out int16 ("feat" Q14, "e" Q12) and
        // the state as fp32 = q / 32768 (exact), as the oracle keeps it
        const bool is_h = nm.size() == 2 && nm[0] == 'h' && nm[1] >= '0' && nm[1] < '0' + p->L;
        if (nm == "feat" || nm == "e" || is_h) {
            const size_t cols = nm == "feat" ? (size_t) kBins : H, esz_out = is_h ? 4 : 2;
            const uint8_t *base = nm == "feat" ? p->i8->featp : nm == "e" ? p->i8->ep : p->i8->hp[p->parity] + (size_t) (nm[1] - '0') * Bp * 2 * H;
            if (bytes > B * cols * esz_out) {
                if (errors) errors->push_back("Unknown tensor name or size too large.");
                return kInvalidArgument;
            }
            const size_t rows = (bytes / esz_out + cols - 1) / cols;
            std::vector<uint8_t> raw(rows * 2 * cols);
            KCHECK(cudaMemcpy(raw.data(), base, raw.size(), cudaMemcpyDeviceToHost));
            std::vector<uint8_t> outb(rows * cols * esz_out);
            for (size_t r = 0; r < rows; r++)
                for (size_t c = 0; c < cols; c++) {
                    const int16_t q = (int16_t) (((int) (int8_t) raw[r * 2 * cols + c]) * 256 + raw[r * 2 * cols + cols + c]);
                    if (is_h) { const float f = (float) q * (1.0f / 32768.0f); memcpy(&outb[(r * cols + c) * 4], &f, 4); }
                    else memcpy(&outb[(r * cols + c) * 2], &q, 2);
                }
            memcpy(dst, outb.data(), bytes);
            return kSuccess;
        }
    }
    if (p->fu && p->planes > 1 && (nm == "feat" || nm == "e")) {
        // fp32 mode on the tensor-core path keeps these as three bf16 planes per row: hand out their sum, the fp32 value
        const size_t cols = nm == "feat" ? (size_t) kBins : H, P = p->planes;
        const size_t e_slot = p->fu->epoch > 0 ? (size_t) ((p->fu->epoch - 1) % p->e_ring) : 0;
        const __nv_bfloat16 *base = nm == "feat" ? (const __nv_bfloat16 *) p->feat + slot * Bp * P * cols : (const __nv_bfloat16 *) p->e + e_slot * Bp * P * cols;
        if (bytes > B * cols * 4) {
            if (errors) errors->push_back("Unknown tensor name or size too large.");
            return kInvalidArgument;
        }
        const size_t rows = (bytes / 4 + cols - 1) / cols;
        std::vector<uint16_t> raw(rows * P * cols);
        KCHECK(cudaMemcpy(raw.data(), base, raw.size() * 2, cudaMemcpyDeviceToHost));
        std::vector<float> sum(rows * cols);
        for (size_t r = 0; r < rows; r++)
            for (size_t c = 0; c < cols; c++) {
                float v = 0.0f;
                for (size_t pl = P; pl-- > 0;) {          // small planes first: the sum is exact either way
                    const uint32_t u = (uint32_t) raw[(r * P + pl) * cols + c] << 16;
                    float f;
                    memcpy(&f, &u, 4);
                    v += f;
                }
                sum[r * cols + c] = v;
            }
        memcpy(dst, sum.data(), bytes);
        return kSuccess;
    }
    const size_t e_slot = (p->fu && p->fu->epoch > 0) ? (size_t) ((p->fu->epoch - 1) % p->e_ring) : 0;
    if (nm == "feat") { src = (const uint8_t *) p->feat + slot * Bp * kBins * esz; avail = B * kBins * esz; }
    else if (nm == "spec") { src = p->spec + slot * Bp * kNfft; avail = B * kNfft * 4; }
    else if (nm == "mask") { src = p->mask + slot * Bp * kBins; avail = B * kBins * 4; }
    else if (nm == "e") { src = (const uint8_t *) p->e + e_slot * Bp * H * esz; avail = B * H * esz; }
    else if (nm == "ola") { src = p->ola[p->ola_par]; avail = B * kFrame * 4; }
    else if (nm == "tail") { src = p->tail; avail = B * kFrame * 2; }
    else if (nm == "trace" && p->fu && p->fu->trace) { src = p->fu->trace; avail = 2048 * sizeof(long long); }
    else if (nm.size() == 2 && nm[0] == 'h' && nm[1] >= '0' && nm[1] < '0' + p->L) {
        src = p->h[p->parity] + (size_t) (nm[1] - '0') * Bp * H;   // h(t) of the last finished step
        avail = B * H * 4;
    }
    if (!src || bytes > avail) {
        if (errors) errors->push_back("Unknown tensor name or size too large.");
        return kInvalidArgument;
    }
    KCHECK(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
    return kSuccess;
}

}  // namespace koala
