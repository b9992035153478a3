/*
 * pv_koala_b200.h -- C ABI of libpv_koala_b200.so, a B200-native drop-in for the per-frame noise-suppression path of
 * Picovoice Koala (`pv_koala_process`).  Plain C: opaque handles, plain pointers and sizes, no CUDA or torch types.
 *
 * Part 1 re-declares, with identical names, signatures, status codes and ownership rules, every symbol the reference
 * engine exports (`nm -D lib/linux/x86_64/libpv_koala.so`: 18 functions) so that the reference's own bindings
 * (binding/python/_koala.py:154-222, demo/c/koala_demo_file.c:265-333) bind to this library unmodified.  Each
 * declaration cites the reference interface it replaces (paths relative to /root/reference).
 * Part 2 is an additive batched extension (not in the reference): many independent streams per call, host or device
 * buffers.  See INTEGRATION.md for the binding stubs.
 */
#ifndef PV_KOALA_B200_H
#define PV_KOALA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PV_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------ Part 1: the reference surface ---------------- */

/* include/picovoice.h:41-54 -- same enumerators, same values */
typedef enum {
    PV_STATUS_SUCCESS = 0,
    PV_STATUS_OUT_OF_MEMORY,
    PV_STATUS_IO_ERROR,
    PV_STATUS_INVALID_ARGUMENT,
    PV_STATUS_STOP_ITERATION,
    PV_STATUS_KEY_ERROR,
    PV_STATUS_INVALID_STATE,
    PV_STATUS_RUNTIME_ERROR,
    PV_STATUS_ACTIVATION_ERROR,
    PV_STATUS_ACTIVATION_LIMIT_REACHED,
    PV_STATUS_ACTIVATION_THROTTLED,
    PV_STATUS_ACTIVATION_REFUSED
} pv_status_t;

/* include/picovoice.h:33-36 -- 16000 */
PV_API int32_t pv_sample_rate(void);
/* include/picovoice.h:56-62 -- "SUCCESS" ... "ACTIVATION_REFUSED"; NULL outside the enum (observed on the reference .so) */
PV_API const char *pv_status_to_string(pv_status_t status);
/* include/picovoice.h:64-79 -- per-thread stack of the last failure, readable once; INVALID_STATE + depth 0 if none pending */
PV_API pv_status_t pv_get_error_stack(char ***message_stack, int32_t *message_stack_depth);
/* include/picovoice.h:81-86 */
PV_API void pv_free_error_stack(char **message_stack);

/* include/pv_koala.h:27-35 -- one handle == one 16 kHz mono stream with its analysis tail, OLA tail and recurrent state */
typedef struct pv_koala pv_koala_t;

/* include/pv_koala.h:37-56.  access_key: syntax-checked only (no licence server).  model_path: a koala_b200 .kpv file.
 * device: "best" | "gpu" | "gpu:K" (B200 only); "cpu" / "cpu:N" parse but fail with PV_STATUS_RUNTIME_ERROR -- this
 * library has no CPU engine by design. */
PV_API pv_status_t pv_koala_init(const char *access_key, const char *model_path, const char *device, pv_koala_t **object);
/* include/pv_koala.h:58-63 -- NULL is a no-op */
PV_API void pv_koala_delete(pv_koala_t *object);
/* include/pv_koala.h:65-80 -- pcm and enhanced_pcm: caller-allocated host buffers of pv_koala_frame_length() samples */
PV_API pv_status_t pv_koala_process(pv_koala_t *object, const int16_t *pcm, int16_t *enhanced_pcm);
/* include/pv_koala.h:82-90 */
PV_API pv_status_t pv_koala_reset(pv_koala_t *object);
/* include/pv_koala.h:92-100 -- 256 */
PV_API pv_status_t pv_koala_delay_sample(const pv_koala_t *object, int32_t *delay_sample);
/* include/pv_koala.h:102-107 -- 256 */
PV_API int32_t pv_koala_frame_length(void);
/* include/pv_koala.h:109-114 */
PV_API const char *pv_koala_version(void);
/* include/pv_koala.h:116-128 -- entries "gpu:K - <name>" for every compute-capability-10.x device */
PV_API pv_status_t pv_koala_list_hardware_devices(char ***hardware_devices, int32_t *num_hardware_devices);
/* include/pv_koala.h:130-138 */
PV_API void pv_koala_free_hardware_devices(char **hardware_devices, int32_t num_hardware_devices);

/* exported by the reference binary but absent from its headers; binding/python/_koala.py:156-160 calls pv_set_sdk */
PV_API void pv_set_sdk(const char *sdk);
PV_API const char *pv_get_sdk(void);
PV_API void pv_free(void *ptr);
PV_API void pv_log_enable(void);
PV_API void pv_log_disable(void);

/* ------------------------------------------------------------------ Part 2: batched extension -------------------- */

/* B independent streams resident on one GPU; stream s keeps the same state a pv_koala_t would. */
typedef struct pv_koala_batch pv_koala_batch_t;

/* precision: "bf16" (tcgen05 mask estimator, bf16 operands; default when NULL), "fp32" (the same kernel with every activation
 * split into three bf16 planes; CUDA-core kernels when the hidden size is not a multiple of 256) or "int8" (the fixed-point
 * variant: int8 weights x int16 activations on tcgen05 kind::i8, integer gates; SPEC.md section 6). */
PV_API pv_status_t pv_koala_batch_init(const char *model_path, const char *device, int32_t num_streams, const char *precision,
                                       pv_koala_batch_t **object);
PV_API void pv_koala_batch_delete(pv_koala_batch_t *object);

/* pcm / enhanced_pcm: [num_streams][num_frames][256] int16, both host or both device memory (detected).  Frame t of
 * every stream is one step of the hot path; state carries across calls exactly as across pv_koala_process calls.
 * Host buffers: copies + compute + copy back, returns when enhanced_pcm is valid.  Device buffers: synchronous too. */
PV_API pv_status_t pv_koala_batch_process(pv_koala_batch_t *object, const int16_t *pcm, int16_t *enhanced_pcm, int32_t num_frames);

/* Same call with TIME-MAJOR buffers, [num_frames][num_streams][256]: frame t of every stream is one contiguous block, which
 * is what a caller driving many pv_koala_process-style streams in lock step (one 256-sample frame per stream per 16 ms tick,
 * /root/reference/demo/c/koala_demo_file.c:466-521 run for many files at once) has in hand.  For host buffers this is the
 * fast path: every chunk of the ingest pipeline is one contiguous copy in each direction. */
PV_API pv_status_t pv_koala_batch_process_time_major(pv_koala_batch_t *object, const int16_t *pcm, int16_t *enhanced_pcm,
                                                     int32_t num_frames);

/* Device buffers only, enqueue-only: frame t of stream s at base + s * stream_stride + t * 256 samples (16-byte aligned,
 * stride % 8 == 0).  cuda_stream: a cudaStream_t passed as void* and taken literally (NULL = CUDA's legacy default stream).
 * Work is ordered on that stream like any other kernel; pv_koala_batch_synchronize waits for the whole device.
 * A handle's steps mutate its per-stream state, so the library orders them itself when the stream changes between calls (or
 * between this call and pv_koala_batch_reset / a host-buffer call): the new stream waits for what was enqueued on the previous
 * one.  Handles are not thread-safe: serialise the calls on one handle (as the reference's bindings do for pv_koala_t).
 * Several handles on one device may be used concurrently from one process; the library serialises their mask-estimator launches
 * on the device.  Sharing the device with OTHER processes through MPS while a "bf16" handle is stepping is not supported (the
 * mask-estimator kernel needs its whole grid resident). */
PV_API pv_status_t pv_koala_batch_process_async(pv_koala_batch_t *object, const int16_t *pcm, int16_t *enhanced_pcm,
                                                int32_t num_frames, int64_t stream_stride, void *cuda_stream);
/* Same with an explicit distance between consecutive frames of one stream: frame t of stream s at
 * base + s * stream_stride + t * frame_stride samples (time-major device buffers: stream_stride 256, frame_stride
 * num_streams * 256).  The call's frames are taken in chunks of up to pv_koala_batch_chunk_frames() frames: three kernel
 * launches per chunk (analysis of the chunk, one persistent mask-estimator launch walking its steps, synthesis), so the
 * caller's serial frame loop (reference: demo/c/koala_demo_file.c:466-521) costs launches per chunk, not per frame.
 * Handles of more than 4096 streams run every chunk as partitions of 4096 streams, one after the other (each partition's
 * state then stays in the L2 for its chunk): three launches per partition and chunk; invisible otherwise.  Feed
 * pv_koala_batch_chunk_frames() frames per call when they are available ("int8" handles and "fp32" handles whose hidden size is
 * not a multiple of 256 work frame by frame: their chunk is 1). */
PV_API pv_status_t pv_koala_batch_process_async_strided(pv_koala_batch_t *object, const int16_t *pcm, int16_t *enhanced_pcm,
                                                        int32_t num_frames, int64_t stream_stride, int64_t frame_stride, void *cuda_stream);
PV_API pv_status_t pv_koala_batch_chunk_frames(const pv_koala_batch_t *object, int32_t *chunk_frames);
PV_API pv_status_t pv_koala_batch_synchronize(pv_koala_batch_t *object);

/* stream_ids == NULL resets every stream (pv_koala_reset semantics per stream). */
PV_API pv_status_t pv_koala_batch_reset(pv_koala_batch_t *object, const int32_t *stream_ids, int32_t num_ids);
PV_API pv_status_t pv_koala_batch_num_streams(const pv_koala_batch_t *object, int32_t *num_streams);
/* CUDA ordinal of the device the handle's streams live on (what "best" / "gpu" / "gpu:K" resolved to) */
PV_API pv_status_t pv_koala_batch_device(const pv_koala_batch_t *object, int32_t *device_index);
PV_API pv_status_t pv_koala_batch_delay_sample(const pv_koala_batch_t *object, int32_t *delay_sample);
/* kernels launched so far by this handle (bench.py's gpu_launches) */
PV_API pv_status_t pv_koala_batch_kernel_launches(const pv_koala_batch_t *object, int64_t *launches);
/* Per-kernel-class CUDA-event timing on the launching stream.  SIX classes: 0 analysis/STFT, 1 encoder GEMM, 2 GRU layer,
 * 3 decoder GEMM, 4 synthesis/iSTFT (1-3: "int8" handles and the CUDA-core "fp32" path), 5 fused mask estimator (encoder + GRU
 * layers + decoder in one kernel: "bf16" and "fp32" handles).  Enable, run steps, then read: read synchronises and returns summed milliseconds and launch counts per class
 * since the previous read; num_classes must be >= 6 (INVALID_ARGUMENT otherwise; INVALID_STATE if profiling is off).
 * Off by default (events perturb back-to-back launches). */
PV_API pv_status_t pv_koala_batch_profile(pv_koala_batch_t *object, int32_t enable);
PV_API pv_status_t pv_koala_batch_profile_read(pv_koala_batch_t *object, double *ms_per_class, int64_t *launches_per_class,
                                               int32_t num_classes);
/* test hook: copy an internal tensor of the last step to the host: "feat" "spec" "mask" "e" "h0".."h7" "ola" "tail"
 * ([num_streams][...] rows; "feat" / "e": fp32 for "fp32" handles, bf16 bits for "bf16", int16 for "int8"; "h*": fp32) */
PV_API pv_status_t pv_koala_batch_debug_read(pv_koala_batch_t *object, const char *name, void *dst, int64_t bytes);

#ifdef __cplusplus
}
#endif

#endif /* PV_KOALA_B200_H */
