"""Batched front door over `pv_koala_batch_*` (additive to the reference surface, include/pv_koala_b200.h part 2).

`BatchKoala.process` is the throughput path: B independent streams advance one or more frames per call.  Inputs may be
numpy int16 arrays (host; copied inside the call) or torch CUDA int16 tensors (device; zero-copy, launched on torch's
current stream).  torch is optional and only used for its tensors / streams.
"""
from ctypes import POINTER, Structure, byref, c_char_p, c_int, c_int16, c_int32, c_int64, c_void_p, cast
from typing import Optional, Sequence

import numpy as np

from ._koala import KoalaInvalidArgumentError, check, load_library
from ._util import default_library_path, default_model_path

FRAME_LENGTH = 256
ANY_ACCESS_KEY = "a29hbGFfYjIwMF9uby1saWNlbmNlLXNlcnZlcg=="   # syntactically valid; nothing validates keys here


class BatchKoala(object):
    """`num_streams` independent 16 kHz streams on one B200 behind `pv_koala_batch_*` (include/pv_koala_b200.h, part 2); stream s keeps
    the state a `Koala` handle would (/root/reference/binding/python/_koala.py:224-254 is the one-stream call this batches).

    precision: "bf16" (tcgen05 mask estimator, bf16 operands), "fp32" (the same kernel, every activation as three bf16 planes) or
    "int8" (fixed-point variant: int8 weights x int16 activations on the integer tensor cores; SPEC.md section 6).
    `process(pcm)`: int16 `[num_streams][frames][256]` (or `[frames][num_streams][256]` with `time_major=True`), numpy / pinned
    torch host tensors or CUDA tensors; state carries across calls.  Feed `chunk_frames` frames per call when they are available:
    a call's frames are walked by one persistent mask-estimator launch per chunk (and per 4096-stream partition of a bigger batch).
    A handle is not thread-safe; CUDA-tensor calls are enqueued on torch's current stream and ordered against earlier work on
    other streams by the library."""

    class CBatch(Structure):
        pass

    def __init__(self, num_streams: int, model_path: Optional[str] = None, device: str = "best",
                 precision: str = "bf16", library_path: Optional[str] = None) -> None:
        library = load_library(default_library_path() if library_path is None else library_path)
        self._library = library
        H = POINTER(self.CBatch)
        library.pv_koala_batch_init.argtypes = [c_char_p, c_char_p, c_int32, c_char_p, POINTER(H)]
        library.pv_koala_batch_init.restype = c_int
        library.pv_koala_batch_delete.argtypes = [H]
        library.pv_koala_batch_delete.restype = None
        library.pv_koala_batch_process.argtypes = [H, c_void_p, c_void_p, c_int32]
        library.pv_koala_batch_process.restype = c_int
        library.pv_koala_batch_process_time_major.argtypes = [H, c_void_p, c_void_p, c_int32]
        library.pv_koala_batch_process_time_major.restype = c_int
        library.pv_koala_batch_process_async.argtypes = [H, c_void_p, c_void_p, c_int32, c_int64, c_void_p]
        library.pv_koala_batch_process_async.restype = c_int
        library.pv_koala_batch_process_async_strided.argtypes = [H, c_void_p, c_void_p, c_int32, c_int64, c_int64, c_void_p]
        library.pv_koala_batch_process_async_strided.restype = c_int
        library.pv_koala_batch_chunk_frames.argtypes = [H, POINTER(c_int32)]
        library.pv_koala_batch_chunk_frames.restype = c_int
        library.pv_koala_batch_synchronize.argtypes = [H]
        library.pv_koala_batch_synchronize.restype = c_int
        library.pv_koala_batch_reset.argtypes = [H, POINTER(c_int32), c_int32]
        library.pv_koala_batch_reset.restype = c_int
        library.pv_koala_batch_kernel_launches.argtypes = [H, POINTER(c_int64)]
        library.pv_koala_batch_kernel_launches.restype = c_int
        library.pv_koala_batch_profile.argtypes = [H, c_int32]
        library.pv_koala_batch_profile.restype = c_int
        library.pv_koala_batch_profile_read.argtypes = [H, c_void_p, c_void_p, c_int32]
        library.pv_koala_batch_profile_read.restype = c_int
        library.pv_koala_batch_debug_read.argtypes = [H, c_char_p, c_void_p, c_int64]
        library.pv_koala_batch_debug_read.restype = c_int
        self._handle = H()
        model_path = default_model_path() if model_path is None else model_path
        check(library, library.pv_koala_batch_init(model_path.encode(), device.encode(), int(num_streams),
                                                   precision.encode(), byref(self._handle)), 'Initialization failed')
        self.num_streams = int(num_streams)
        self.precision = precision
        dev = c_int32(-1)
        library.pv_koala_batch_device.argtypes = [H, POINTER(c_int32)]
        library.pv_koala_batch_device.restype = c_int
        check(library, library.pv_koala_batch_device(self._handle, byref(dev)), 'device query failed')
        self.device_index = dev.value                 # CUDA ordinal the handle's streams live on
        self.frame_length = library.pv_koala_frame_length()
        self.sample_rate = library.pv_sample_rate()
        self.delay_sample = 256
        cf = c_int32(1)
        check(library, library.pv_koala_batch_chunk_frames(self._handle, byref(cf)), 'chunk_frames failed')
        self.chunk_frames = cf.value                  # frames one mask-estimator launch walks (bf16 path); 1 for fp32

    def delete(self) -> None:
        if self._handle:
            self._library.pv_koala_batch_delete(self._handle)
            self._handle = POINTER(self.CBatch)()

    def __del__(self):
        try:
            self.delete()
        except Exception:
            pass

    def _shape(self, shape, time_major: bool = False) -> int:
        if len(shape) == 2:
            shape = (1, shape[0], shape[1]) if time_major else (shape[0], 1, shape[1])
        want = "[frames][%d][%d]" % (self.num_streams, self.frame_length) if time_major else "[%d][frames][%d]" % (self.num_streams, self.frame_length)
        if len(shape) != 3:
            raise KoalaInvalidArgumentError("expected pcm of shape %s, got %s" % (want, tuple(shape)))
        streams, frames = (shape[1], shape[0]) if time_major else (shape[0], shape[1])
        if streams != self.num_streams or shape[2] != self.frame_length:
            raise KoalaInvalidArgumentError("expected pcm of shape %s, got %s" % (want, tuple(shape)))
        return frames

    def process(self, pcm, out=None, time_major: bool = False):
        """pcm: int16 [B][256] or [B][T][256] (time_major: [T][B][256]); numpy (host), pinned torch tensor (host) or torch.cuda
        tensor (device).  Returns the same kind and layout.  Time-major host buffers are the fast way in: every chunk of the
        ingest pipeline is then one contiguous copy."""
        entry = self._library.pv_koala_batch_process_time_major if time_major else self._library.pv_koala_batch_process
        if isinstance(pcm, np.ndarray):
            frames = self._shape(pcm.shape, time_major)
            pcm = np.ascontiguousarray(pcm, dtype=np.int16)
            if out is None:
                out = np.empty_like(pcm)
            elif not (isinstance(out, np.ndarray) and out.dtype == np.int16 and out.shape == pcm.shape and out.flags.c_contiguous
                      and out.flags.writeable):
                raise KoalaInvalidArgumentError("out must be a writeable C-contiguous int16 numpy array of pcm's shape")
            check(self._library, entry(self._handle, pcm.ctypes.data, out.ctypes.data, frames), 'Processing failed')
            return out
        import torch  # device tensors (or pinned host tensors) only
        if not isinstance(pcm, torch.Tensor) or pcm.dtype != torch.int16 or not pcm.is_contiguous():
            raise KoalaInvalidArgumentError("pcm must be a contiguous int16 numpy array or torch tensor")
        frames = self._shape(tuple(pcm.shape), time_major)
        if out is None:
            out = torch.empty_like(pcm)
        elif not (isinstance(out, torch.Tensor) and out.dtype == torch.int16 and out.shape == pcm.shape and out.is_contiguous()
                  and out.device == pcm.device):
            raise KoalaInvalidArgumentError("out must be a contiguous int16 torch tensor of pcm's shape on pcm's device")
        if pcm.is_cuda and pcm.device.index != self.device_index:
            raise KoalaInvalidArgumentError("pcm lives on cuda:%d but this engine runs on cuda:%d" % (pcm.device.index, self.device_index))
        if pcm.is_cuda and not time_major:
            stream = torch.cuda.current_stream(pcm.device).cuda_stream
            check(self._library, self._library.pv_koala_batch_process_async(
                self._handle, pcm.data_ptr(), out.data_ptr(), frames, frames * self.frame_length, c_void_p(stream)),
                'Processing failed')
        elif pcm.is_cuda:
            stream = torch.cuda.current_stream(pcm.device).cuda_stream
            # streams 256 samples apart, frames num_streams * 256
            check(self._library, self._library.pv_koala_batch_process_async_strided(
                self._handle, pcm.data_ptr(), out.data_ptr(), frames, self.frame_length, self.num_streams * self.frame_length,
                c_void_p(stream)), 'Processing failed')
        else:
            check(self._library, entry(self._handle, pcm.data_ptr(), out.data_ptr(), frames), 'Processing failed')
        return out

    def synchronize(self) -> None:
        check(self._library, self._library.pv_koala_batch_synchronize(self._handle), 'Synchronize failed')

    def reset(self, stream_ids: Optional[Sequence[int]] = None) -> None:
        if stream_ids is None:
            check(self._library, self._library.pv_koala_batch_reset(self._handle, None, 0), 'Reset failed')
        else:
            ids = (c_int32 * len(stream_ids))(*stream_ids)
            check(self._library, self._library.pv_koala_batch_reset(self._handle, ids, len(stream_ids)), 'Reset failed')

    @property
    def kernel_launches(self) -> int:
        n = c_int64()
        check(self._library, self._library.pv_koala_batch_kernel_launches(self._handle, byref(n)), 'launch count failed')
        return n.value

    KERNEL_CLASSES = ("frontend", "enc", "gru", "dec", "backend", "masknet")

    def profile(self, enable: bool) -> None:
        check(self._library, self._library.pv_koala_batch_profile(self._handle, 1 if enable else 0), 'profile failed')

    def profile_read(self):
        """{class: (total_ms, launches)} since the last read; synchronises the device."""
        ms = np.zeros(8, np.float64)
        cnt = np.zeros(8, np.int64)
        check(self._library, self._library.pv_koala_batch_profile_read(self._handle, ms.ctypes.data, cnt.ctypes.data, 8),
              'profile_read failed')
        return {name: (float(ms[i]), int(cnt[i])) for i, name in enumerate(self.KERNEL_CLASSES)}

    def debug_read(self, name: str, shape, dtype) -> np.ndarray:
        arr = np.empty(shape, dtype=dtype)
        check(self._library, self._library.pv_koala_batch_debug_read(self._handle, name.encode(), arr.ctypes.data, arr.nbytes),
              'debug_read failed')
        return arr


__all__ = ['BatchKoala', 'ANY_ACCESS_KEY']
