// koala_b200 -- host-side engine: owns the per-stream state rows in HBM and launches one step of the hot path.
//
// One Engine == B independent streams on one GPU (the reference's pv_koala_t is the B == 1 case,
// /root/reference/include/pv_koala.h:27-35).  Plain C++ interface; the C ABI in koala_abi.cu is the only exported surface.
#pragma once

#include <stdint.h>

#include <string>
#include <vector>

namespace koala {

enum Status : int {   // mirrors pv_status_t, /root/reference/include/picovoice.h:41-54
    kSuccess = 0, kOutOfMemory, kIoError, kInvalidArgument, kStopIteration, kKeyError, kInvalidState, kRuntimeError,
    kActivationError, kActivationLimitReached, kActivationThrottled, kActivationRefused
};

struct ModelHost {
    int hidden = 0, layers = 0;
    std::vector<uint16_t> enc_w, dec_w;                 // bf16 bit patterns, [H][256] and [256][H]
    std::vector<float> enc_b, dec_b;
    std::vector<std::vector<uint16_t>> wih, whh;        // [3H][H]
    std::vector<std::vector<float>> bih, bhh;           // [3H]
};

// Parses the koala_b200 model file (format: koala_b200/spec.py).  Pushes human-readable reasons on `errors`.
Status load_model_file(const char *path, ModelHost *out, std::vector<std::string> *errors);

class Engine {
public:
    // precision: 0 fp32 mask path (tcgen05 with three bf16 planes per activation when the hidden size is a multiple of 256, else
    // CUDA-core kernels), 1 bf16 tcgen05 mask path
    static Status create(const ModelHost &model, int device, int num_streams, int precision, Engine **out,
                         std::vector<std::string> *errors);
    ~Engine();

    int num_streams() const { return n_; }
    int device() const { return device_; }
    int precision() const { return precision_; }

    // pcm / out: frame t of stream s at base + s * stride + t * frame_stride (int16 samples; frame_stride 256 for stream-major
    // buffers, streams * 256 with stride 256 for time-major ones).  Device pointers, 16-byte aligned, strides multiples of 8.
    // Enqueues `frames` consecutive steps on `stream` (a cudaStream_t taken literally: nullptr is the legacy default stream;
    // own_stream() is the engine's private one) and returns without synchronising.  The tensor-core path takes the frames in chunks of
    // up to chunk_frames(): analysis of the chunk, ONE fused mask-estimator launch walking its steps, synthesis of the chunk.
    Status process_device(const int16_t *pcm, int16_t *out, int frames, long long stride, void *stream,
                          std::vector<std::string> *errors, long long out_stride = 0 /* 0: same as stride */,
                          long long frame_stride = 0 /* 0: 256 */, long long out_frame_stride = 0 /* 0: same as frame_stride */);
    int chunk_frames() const;
    // Host buffers, [B][frames][256] or (time_major) [frames][B][256]: H2D copies, steps and D2H copies overlapped, synchronise.
    Status process_host(const int16_t *pcm, int16_t *out, int frames, std::vector<std::string> *errors, bool time_major = false);
    Status reset(const int32_t *stream_ids, int n, std::vector<std::string> *errors);   // ids == nullptr: all streams
    Status synchronize(std::vector<std::string> *errors);   // waits for everything queued on the device
    void *own_stream() const;

    // test hooks: copy an internal tensor to the host ("feat", "spec", "mask", "h0", "h1", ..., "ola", "tail")
    Status debug_read(const char *name, void *dst, size_t bytes, std::vector<std::string> *errors);
    long long kernel_launches() const { return launches_; }
    // per-kernel-class CUDA-event timing (bench.py roofline): enable, run steps, then read (synchronises)
    void set_profile(bool on);
    Status profile_read(double *ms_per_class, long long *launches_per_class, int n_classes, std::vector<std::string> *errors);

private:
    Engine() = default;
    struct Impl;
    Impl *p_ = nullptr;
    int n_ = 0, npad_ = 0, device_ = 0, precision_ = 0;
    long long launches_ = 0;
};

}  // namespace koala
