// koala_b200 -- the C ABI (include/pv_koala_b200.h).  Everything exported from libpv_koala_b200.so lives here.
//
// Behaviour mirrors what the reference binary does for the same calls (observations recorded by tools/make_abi_kat.py
// into tests/golden/abi_kat.json): status codes of picovoice.h:41-54, a per-thread error stack that is readable once
// (picovoice.h:64-79), validation order of pv_koala_init = NULL arguments -> device string -> model file -> AccessKey.
// There is no licence client and no CPU engine: ACTIVATION_* is never returned and `cpu` devices fail loudly.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <string>
#include <vector>

#include "../../include/pv_koala_b200.h"
#include "engine.h"
#include "koala_common.cuh"

using koala::Engine;
using koala::ModelHost;
using koala::Status;

namespace {

thread_local std::vector<std::string> tl_stack;
thread_local bool tl_pending = false;
bool g_log = false;
std::mutex g_sdk_mutex;
std::string g_sdk = "c";

enum : unsigned { kCodeNullArg = 0x64, kCodeGeneric = 0x12C, kCodeDevice = 0x322, kCodeComm = 0x334, kCodeFile = 0xC9, kCodeKey = 0x190 };

std::string tag(unsigned code, const std::string &text) {
    char head[32];
    snprintf(head, sizeof(head), "kb200 %08X: ", code);
    return std::string(head) + text;
}

pv_status_t fail(pv_status_t st, std::vector<std::string> msgs) {
    if (msgs.size() > 7) msgs.resize(7);   // the reference keeps its stack below 8 (binding/python/test_koala.py:182-183)
    tl_stack = std::move(msgs);
    tl_pending = true;
    if (g_log)
        for (const auto &m : tl_stack) fprintf(stderr, "[koala_b200] %s\n", m.c_str());
    return st;
}

pv_status_t fail_null(const char *arg) {
    return fail(PV_STATUS_INVALID_ARGUMENT, {tag(kCodeNullArg, std::string("Argument `") + arg + "` is NULL.")});
}

pv_status_t fail_engine(Status st, const std::vector<std::string> &errs, unsigned code = kCodeGeneric) {
    std::vector<std::string> msgs;
    for (const auto &e : errs) msgs.push_back(tag(code, e));
    if (msgs.empty()) msgs.push_back(tag(kCodeGeneric, "Picovoice Error."));
    return fail((pv_status_t) st, msgs);
}

// "best" | "gpu" | "gpu:K" | "cpu" | "cpu:N".  Returns false for anything else.
bool parse_device(const char *s, bool *is_cpu, int *index) {
    *is_cpu = false;
    *index = -1;   // -1: pick automatically
    if (strcmp(s, "best") == 0 || strcmp(s, "gpu") == 0 || strcmp(s, "gpu:") == 0) return true;
    if (strcmp(s, "cpu") == 0) { *is_cpu = true; return true; }
    const bool g = strncmp(s, "gpu:", 4) == 0, c = strncmp(s, "cpu:", 4) == 0;
    if (!g && !c) return false;
    const char *d = s + 4;
    if (!*d) return false;
    for (const char *q = d; *q; ++q)
        if (*q < '0' || *q > '9') return false;
    char *end = nullptr;
    const long v = strtol(d, &end, 10);
    if (*end || v < 0 || v > 4096) return false;   // digits only (checked above), bounded: no int overflow
    *is_cpu = c;
    *index = (int) v;
    return true;
}

// resolves the device string to a CUDA ordinal of a compute-capability-10.x device
pv_status_t resolve_device(const char *device, int *ordinal) {
    bool is_cpu;
    int index;
    if (!parse_device(device, &is_cpu, &index))
        return fail(PV_STATUS_INVALID_ARGUMENT, {tag(kCodeDevice, std::string(device) + " is not a valid device string"),
                                                 tag(kCodeGeneric, "Picovoice Error.")});
    if (is_cpu)
        return fail(PV_STATUS_RUNTIME_ERROR,
                    {tag(kCodeDevice, "CPU inference is not part of koala_b200 (GPU-only build); use `best`, `gpu` or `gpu:K`."),
                     tag(kCodeGeneric, "Picovoice Error.")});
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        cudaGetLastError();
        return fail(PV_STATUS_RUNTIME_ERROR, {tag(kCodeComm, "Failed to communicate with device."), tag(kCodeGeneric, "Picovoice Error.")});
    }
    if (index >= 0) {
        if (index >= count)
            return fail(PV_STATUS_RUNTIME_ERROR, {tag(kCodeComm, "Failed to communicate with device."), tag(kCodeGeneric, "Picovoice Error.")});
        *ordinal = index;
        return PV_STATUS_SUCCESS;
    }
    for (int i = 0; i < count; i++) {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, i) == cudaSuccess && prop.major == 10) {
            *ordinal = i;
            return PV_STATUS_SUCCESS;
        }
    }
    return fail(PV_STATUS_RUNTIME_ERROR, {tag(kCodeComm, "Selected GPU device is incompatible with the library."), tag(kCodeGeneric, "Picovoice Error.")});
}

// Syntax check only (there is no licence server): base64 alphabet, '=' padding at the end, length >= 16 and % 4 == 0.
bool access_key_parses(const char *k) {
    const size_t n = strlen(k);
    if (n < 16 || n % 4 != 0) return false;
    size_t pad = 0;
    for (size_t i = 0; i < n; i++) {
        const char ch = k[i];
        const bool b64 = (ch >= 'A' && ch <= 'Z') || (ch >= 'a' && ch <= 'z') || (ch >= '0' && ch <= '9') || ch == '+' || ch == '/';
        if (ch == '=') { if (i < n - 2) return false; pad++; }
        else if (!b64 || pad) return false;
    }
    return true;
}

int precision_from(const char *s, bool *ok) {
    *ok = true;
    if (!s || !*s) {
        s = getenv("KOALA_B200_PRECISION");
        if (!s || !*s) return koala::kBf16;
    }
    if (strcmp(s, "bf16") == 0) return koala::kBf16;
    if (strcmp(s, "fp32") == 0) return koala::kFp32;
    if (strcmp(s, "int8") == 0) return koala::kInt8;
    *ok = false;
    return koala::kBf16;
}

}  // namespace

struct pv_koala {
    Engine *engine = nullptr;
    int16_t *pin_in = nullptr, *pin_out = nullptr;   // pinned staging for the caller's 512-byte frames
};

struct pv_koala_batch {
    Engine *engine = nullptr;
};

extern "C" {

PV_API int32_t pv_sample_rate(void) { return koala::kSampleRate; }

PV_API const char *pv_status_to_string(pv_status_t status) {
    static const char *const names[] = {"SUCCESS", "OUT_OF_MEMORY", "IO_ERROR", "INVALID_ARGUMENT", "STOP_ITERATION", "KEY_ERROR",
                                        "INVALID_STATE", "RUNTIME_ERROR", "ACTIVATION_ERROR", "ACTIVATION_LIMIT_REACHED",
                                        "ACTIVATION_THROTTLED", "ACTIVATION_REFUSED"};
    const int s = (int) status;
    return (s >= 0 && s < 12) ? names[s] : NULL;
}

PV_API pv_status_t pv_get_error_stack(char ***message_stack, int32_t *message_stack_depth) {
    if (!message_stack || !message_stack_depth) return PV_STATUS_INVALID_ARGUMENT;
    const size_t n = tl_pending ? tl_stack.size() : 0;
    char **out = (char **) calloc(n + 1, sizeof(char *));
    if (!out) return PV_STATUS_OUT_OF_MEMORY;
    for (size_t i = 0; i < n; i++) out[i] = strdup(tl_stack[i].c_str());
    *message_stack = out;
    *message_stack_depth = (int32_t) n;
    const bool had = tl_pending;
    tl_pending = false;
    tl_stack.clear();
    return had ? PV_STATUS_SUCCESS : PV_STATUS_INVALID_STATE;
}

PV_API void pv_free_error_stack(char **message_stack) {
    if (!message_stack) return;
    for (char **p = message_stack; *p; ++p) free(*p);
    free(message_stack);
}

PV_API void pv_set_sdk(const char *sdk) {
    if (!sdk) return;
    std::lock_guard<std::mutex> lock(g_sdk_mutex);
    g_sdk = sdk;
}

PV_API const char *pv_get_sdk(void) {
    static thread_local std::string copy;
    std::lock_guard<std::mutex> lock(g_sdk_mutex);
    copy = g_sdk;
    return copy.c_str();
}

PV_API void pv_free(void *ptr) { free(ptr); }
PV_API void pv_log_enable(void) { g_log = true; }
PV_API void pv_log_disable(void) { g_log = false; }

PV_API int32_t pv_koala_frame_length(void) { return koala::kFrame; }
PV_API const char *pv_koala_version(void) { return "1.0.0"; }

static pv_status_t create_engine(const char *model_path, const char *device, int num_streams, int precision, Engine **out) {
    int ordinal = 0;
    pv_status_t st = resolve_device(device, &ordinal);
    if (st != PV_STATUS_SUCCESS) return st;
    ModelHost model;
    std::vector<std::string> errs;
    Status ms = koala::load_model_file(model_path, &model, &errs);
    if (ms != koala::kSuccess) return fail_engine(ms, errs, kCodeFile);
    ms = Engine::create(model, ordinal, num_streams, precision, out, &errs);
    if (ms != koala::kSuccess) return fail_engine(ms, errs);
    return PV_STATUS_SUCCESS;
}

PV_API pv_status_t pv_koala_init(const char *access_key, const char *model_path, const char *device, pv_koala_t **object) {
    if (!access_key) return fail_null("access_key");
    if (!model_path) return fail_null("model_path");
    if (!device) return fail_null("device");   // the reference segfaults here
    if (!object) return fail_null("object");
    *object = NULL;
    // observed order on the reference: device string (and device reachability) -> model file -> AccessKey
    int ordinal = 0;
    pv_status_t dst = resolve_device(device, &ordinal);
    if (dst != PV_STATUS_SUCCESS) return dst;
    FILE *f = fopen(model_path, "rb");
    if (!f) return fail(PV_STATUS_IO_ERROR, {tag(kCodeFile, std::string("Failed to open file `") + model_path + "`."), tag(kCodeGeneric, "Picovoice Error.")});
    fclose(f);
    if (!access_key_parses(access_key))
        return fail(PV_STATUS_INVALID_ARGUMENT, {tag(0x6F, "Picovoice Error."), tag(kCodeKey, std::string("Failed to parse AccessKey `") + access_key + "`."),
                                                 tag(kCodeGeneric, "Picovoice Error.")});
    bool ok;
    const int precision = precision_from(NULL, &ok);
    if (!ok) return fail(PV_STATUS_INVALID_ARGUMENT, {tag(kCodeGeneric, "KOALA_B200_PRECISION must be `bf16`, `fp32` or `int8`.")});
    Engine *eng = nullptr;
    pv_status_t st = create_engine(model_path, device, 1, precision, &eng);
    if (st != PV_STATUS_SUCCESS) return st;
    pv_koala_t *o = new pv_koala();
    o->engine = eng;
    if (cudaMallocHost((void **) &o->pin_in, koala::kFrame * sizeof(int16_t)) != cudaSuccess ||
        cudaMallocHost((void **) &o->pin_out, koala::kFrame * sizeof(int16_t)) != cudaSuccess) {
        pv_koala_delete(o);
        return fail(PV_STATUS_OUT_OF_MEMORY, {tag(kCodeGeneric, "Failed to allocate pinned staging memory.")});
    }
    *object = o;
    return PV_STATUS_SUCCESS;
}

PV_API void pv_koala_delete(pv_koala_t *object) {
    if (!object) return;
    delete object->engine;
    if (object->pin_in) cudaFreeHost(object->pin_in);
    if (object->pin_out) cudaFreeHost(object->pin_out);
    delete object;
}

PV_API pv_status_t pv_koala_process(pv_koala_t *object, const int16_t *pcm, int16_t *enhanced_pcm) {
    if (!object) return fail_null("object");
    if (!pcm) return fail_null("pcm");
    if (!enhanced_pcm) return fail_null("enhanced_pcm");
    memcpy(object->pin_in, pcm, koala::kFrame * sizeof(int16_t));
    std::vector<std::string> errs;
    Status st = object->engine->process_host(object->pin_in, object->pin_out, 1, &errs);
    if (st != koala::kSuccess) return fail_engine(st, errs);
    memcpy(enhanced_pcm, object->pin_out, koala::kFrame * sizeof(int16_t));
    return PV_STATUS_SUCCESS;
}

PV_API pv_status_t pv_koala_reset(pv_koala_t *object) {
    if (!object) return PV_STATUS_INVALID_ARGUMENT;   // the reference pushes no message for this one
    std::vector<std::string> errs;
    Status st = object->engine->reset(nullptr, 0, &errs);
    if (st != koala::kSuccess) return fail_engine(st, errs);
    return PV_STATUS_SUCCESS;
}

PV_API pv_status_t pv_koala_delay_sample(const pv_koala_t *object, int32_t *delay_sample) {
    if (!object) return fail_null("object");
    if (!delay_sample) return fail_null("delay_sample");
    *delay_sample = koala::kDelay;
    return PV_STATUS_SUCCESS;
}

PV_API pv_status_t pv_koala_list_hardware_devices(char ***hardware_devices, int32_t *num_hardware_devices) {
    if (!hardware_devices) return fail_null("hardware_devices");
    if (!num_hardware_devices) return fail_null("num_hardware_devices");
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess) {
        cudaGetLastError();
        count = 0;
    }
    std::vector<std::string> names;
    for (int i = 0; i < count; i++) {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, i) == cudaSuccess && prop.major == 10) {
            char line[320];
            snprintf(line, sizeof(line), "gpu:%d - %s", i, prop.name);
            names.push_back(line);
        }
    }
    char **out = (char **) calloc(names.size() + 1, sizeof(char *));
    if (!out) return fail(PV_STATUS_OUT_OF_MEMORY, {tag(kCodeGeneric, "Out of memory.")});
    for (size_t i = 0; i < names.size(); i++) out[i] = strdup(names[i].c_str());
    *hardware_devices = out;
    *num_hardware_devices = (int32_t) names.size();
    return PV_STATUS_SUCCESS;
}

PV_API void pv_koala_free_hardware_devices(char **hardware_devices, int32_t num_hardware_devices) {
    if (!hardware_devices) return;
    for (int32_t i = 0; i < num_hardware_devices; i++) free(hardware_devices[i]);
    free(hardware_devices);
}

// ------------------------------------------------------------------------------------------------ batched extension
PV_API pv_status_t pv_koala_batch_init(const char *model_path, const char *device, int32_t num_streams, const char *precision,
                                       pv_koala_batch_t **object) {
    if (!model_path) return fail_null("model_path");
    if (!device) return fail_null("device");
    if (!object) return fail_null("object");
    *object = NULL;
    if (num_streams < 1) return fail(PV_STATUS_INVALID_ARGUMENT, {tag(kCodeGeneric, "`num_streams` must be positive.")});
    bool ok;
    const int prec = precision_from(precision, &ok);
    if (!ok) return fail(PV_STATUS_INVALID_ARGUMENT, {tag(kCodeGeneric, "`precision` must be `bf16`, `fp32` or `int8`.")});
    Engine *eng = nullptr;
    pv_status_t st = create_engine(model_path, device, num_streams, prec, &eng);
    if (st != PV_STATUS_SUCCESS) return st;
    pv_koala_batch_t *o = new pv_koala_batch();
    o->engine = eng;
    *object = o;
    return PV_STATUS_SUCCESS;
}

PV_API void pv_koala_batch_delete(pv_koala_batch_t *object) {
    if (!object) return;
    delete object->engine;
    delete object;
}

static int pointer_kind(const void *p) {   // 0 host, 1 device
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged) ? 1 : 0;
}

static pv_status_t batch_process(pv_koala_batch_t *object, const int16_t *pcm, int16_t *enhanced_pcm, int32_t num_frames, bool time_major) {
    if (!object) return fail_null("object");
    if (!pcm) return fail_null("pcm");
    if (!enhanced_pcm) return fail_null("enhanced_pcm");
    if (num_frames < 0) return fail(PV_STATUS_INVALID_ARGUMENT, {tag(kCodeGeneric, "`num_frames` must not be negative.")});
    if (num_frames == 0) return PV_STATUS_SUCCESS;   // nothing to do, host or device buffers alike
    cudaSetDevice(object->engine->device());
    const int kin = pointer_kind(pcm), kout = pointer_kind(enhanced_pcm);
    if (kin != kout) return fail(PV_STATUS_INVALID_ARGUMENT, {tag(kCodeGeneric, "`pcm` and `enhanced_pcm` must both be host or both be device memory.")});
    std::vector<std::string> errs;
    Status st = koala::kSuccess;
    if (kin == 1) {
        if (time_major) {   // [num_frames][num_streams][256]: streams 256 samples apart, frames num_streams * 256
            const long long frame = (long long) object->engine->num_streams() * koala::kFrame;
            st = object->engine->process_device(pcm, enhanced_pcm, num_frames, koala::kFrame, object->engine->own_stream(), &errs, koala::kFrame, frame, frame);
        } else {
            st = object->engine->process_device(pcm, enhanced_pcm, num_frames, (long long) num_frames * koala::kFrame,
                                                object->engine->own_stream(), &errs);
        }
        if (st == koala::kSuccess) st = object->engine->synchronize(&errs);
    } else {
        st = object->engine->process_host(pcm, enhanced_pcm, num_frames, &errs, time_major);
    }
    if (st != koala::kSuccess) return fail_engine(st, errs);
    return PV_STATUS_SUCCESS;
}

PV_API pv_status_t pv_koala_batch_process(pv_koala_batch_t *object, const int16_t *pcm, int16_t *enhanced_pcm, int32_t num_frames) {
    return batch_process(object, pcm, enhanced_pcm, num_frames, false);
}

PV_API pv_status_t pv_koala_batch_process_time_major(pv_koala_batch_t *object, const int16_t *pcm, int16_t *enhanced_pcm, int32_t num_frames) {
    return batch_process(object, pcm, enhanced_pcm, num_frames, true);
}

PV_API pv_status_t pv_koala_batch_process_async(pv_koala_batch_t *object, const int16_t *pcm, int16_t *enhanced_pcm,
                                                int32_t num_frames, int64_t stream_stride, void *cuda_stream) {
    if (!object) return fail_null("object");
    if (!pcm) return fail_null("pcm");
    if (!enhanced_pcm) return fail_null("enhanced_pcm");
    std::vector<std::string> errs;
    Status st = object->engine->process_device(pcm, enhanced_pcm, num_frames, stream_stride, cuda_stream, &errs);
    if (st != koala::kSuccess) return fail_engine(st, errs);
    return PV_STATUS_SUCCESS;
}

PV_API pv_status_t pv_koala_batch_process_async_strided(pv_koala_batch_t *object, const int16_t *pcm, int16_t *enhanced_pcm,
                                                        int32_t num_frames, int64_t stream_stride, int64_t frame_stride, void *cuda_stream) {
    if (!object) return fail_null("object");
    if (!pcm) return fail_null("pcm");
    if (!enhanced_pcm) return fail_null("enhanced_pcm");
    std::vector<std::string> errs;
    Status st = object->engine->process_device(pcm, enhanced_pcm, num_frames, stream_stride, cuda_stream, &errs, stream_stride, frame_stride, frame_stride);
    if (st != koala::kSuccess) return fail_engine(st, errs);
    return PV_STATUS_SUCCESS;
}

PV_API pv_status_t pv_koala_batch_chunk_frames(const pv_koala_batch_t *object, int32_t *chunk_frames) {
    if (!object) return fail_null("object");
    if (!chunk_frames) return fail_null("chunk_frames");
    *chunk_frames = object->engine->chunk_frames();
    return PV_STATUS_SUCCESS;
}

PV_API pv_status_t pv_koala_batch_synchronize(pv_koala_batch_t *object) {
    if (!object) return fail_null("object");
    std::vector<std::string> errs;
    Status st = object->engine->synchronize(&errs);
    if (st != koala::kSuccess) return fail_engine(st, errs);
    return PV_STATUS_SUCCESS;
}

PV_API pv_status_t pv_koala_batch_reset(pv_koala_batch_t *object, const int32_t *stream_ids, int32_t num_ids) {
    if (!object) return fail_null("object");
    std::vector<std::string> errs;
    Status st = object->engine->reset(stream_ids, num_ids, &errs);
    if (st != koala::kSuccess) return fail_engine(st, errs);
    return PV_STATUS_SUCCESS;
}

PV_API pv_status_t pv_koala_batch_num_streams(const pv_koala_batch_t *object, int32_t *num_streams) {
    if (!object) return fail_null("object");
    if (!num_streams) return fail_null("num_streams");
    *num_streams = object->engine->num_streams();
    return PV_STATUS_SUCCESS;
}

PV_API pv_status_t pv_koala_batch_device(const pv_koala_batch_t *object, int32_t *device_index) {
    if (!object) return fail_null("object");
    if (!device_index) return fail_null("device_index");
    *device_index = object->engine->device();
    return PV_STATUS_SUCCESS;
}

PV_API pv_status_t pv_koala_batch_delay_sample(const pv_koala_batch_t *object, int32_t *delay_sample) {
    if (!object) return fail_null("object");
    if (!delay_sample) return fail_null("delay_sample");
    *delay_sample = koala::kDelay;
    return PV_STATUS_SUCCESS;
}

PV_API pv_status_t pv_koala_batch_kernel_launches(const pv_koala_batch_t *object, int64_t *launches) {
    if (!object) return fail_null("object");
    if (!launches) return fail_null("launches");
    *launches = object->engine->kernel_launches();
    return PV_STATUS_SUCCESS;
}

PV_API pv_status_t pv_koala_batch_profile(pv_koala_batch_t *object, int32_t enable) {
    if (!object) return fail_null("object");
    object->engine->set_profile(enable != 0);
    return PV_STATUS_SUCCESS;
}

PV_API pv_status_t pv_koala_batch_profile_read(pv_koala_batch_t *object, double *ms_per_class, int64_t *launches_per_class, int32_t num_classes) {
    if (!object) return fail_null("object");
    if (!ms_per_class) return fail_null("ms_per_class");
    if (!launches_per_class) return fail_null("launches_per_class");
    std::vector<std::string> errs;
    std::vector<long long> cnt(num_classes > 0 ? num_classes : 0);
    Status st = object->engine->profile_read(ms_per_class, cnt.data(), num_classes, &errs);
    if (st != koala::kSuccess) return fail_engine(st, errs);
    for (int i = 0; i < num_classes; i++) launches_per_class[i] = cnt[i];
    return PV_STATUS_SUCCESS;
}

PV_API pv_status_t pv_koala_batch_debug_read(pv_koala_batch_t *object, const char *name, void *dst, int64_t bytes) {
    if (!object) return fail_null("object");
    if (!name) return fail_null("name");
    if (!dst) return fail_null("dst");
    std::vector<std::string> errs;
    Status st = object->engine->debug_read(name, dst, (size_t) bytes, &errs);
    if (st != koala::kSuccess) return fail_engine(st, errs);
    return PV_STATUS_SUCCESS;
}

}  // extern "C"
