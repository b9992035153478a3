"""ctypes wrapper over oracle/libkoala_oracle.so (C restatement of SPEC.md).  Test infrastructure only."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
FRAME = 256
BINS = 256


def oracle_lib_path() -> str:
    return os.path.join(_HERE, "libkoala_oracle.so")


def build_oracle(force: bool = False) -> str:
    so, src = oracle_lib_path(), os.path.join(_HERE, "koala_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return so


_lib = None


def _load():
    global _lib
    if _lib is None:
        lib = C.CDLL(build_oracle())
        vp, ip = C.c_void_p, C.c_int
        i16p, f32p = C.POINTER(C.c_int16), C.POINTER(C.c_float)
        lib.ko_model_load.argtypes = [C.c_char_p, C.POINTER(vp)]; lib.ko_model_load.restype = ip
        lib.ko_model_free.argtypes = [vp]; lib.ko_model_free.restype = None
        lib.ko_model_hidden.argtypes = [vp]; lib.ko_model_hidden.restype = ip
        lib.ko_model_layers.argtypes = [vp]; lib.ko_model_layers.restype = ip
        lib.ko_stream_new.argtypes = [vp, ip]; lib.ko_stream_new.restype = vp
        lib.ko_stream_free.argtypes = [vp]; lib.ko_stream_free.restype = None
        lib.ko_stream_reset.argtypes = [vp]; lib.ko_stream_reset.restype = None
        lib.ko_stream_process.argtypes = [vp, i16p, i16p]; lib.ko_stream_process.restype = None
        lib.ko_frontend.argtypes = [vp, i16p, f32p, f32p]; lib.ko_frontend.restype = None
        lib.ko_masknet.argtypes = [vp, f32p, f32p]; lib.ko_masknet.restype = None
        lib.ko_masknet_q.argtypes = [vp, i16p, f32p]; lib.ko_masknet_q.restype = None
        lib.ko_quantize_feat.argtypes = [f32p, i16p]; lib.ko_quantize_feat.restype = None
        lib.ko_backend.argtypes = [vp, f32p, f32p, i16p]; lib.ko_backend.restype = None
        lib.ko_stream_h.argtypes = [vp]; lib.ko_stream_h.restype = f32p
        lib.ko_stream_ola.argtypes = [vp]; lib.ko_stream_ola.restype = f32p
        lib.ko_stream_tail.argtypes = [vp]; lib.ko_stream_tail.restype = i16p
        lib.ko_stream_last_mask.argtypes = [vp]; lib.ko_stream_last_mask.restype = f32p
        lib.ko_stream_last_feat.argtypes = [vp]; lib.ko_stream_last_feat.restype = f32p
        lib.ko_batch_new.argtypes = [vp, ip, ip]; lib.ko_batch_new.restype = vp
        lib.ko_batch_free.argtypes = [vp]; lib.ko_batch_free.restype = None
        lib.ko_batch_reset.argtypes = [vp]; lib.ko_batch_reset.restype = None
        lib.ko_batch_stream.argtypes = [vp, ip]; lib.ko_batch_stream.restype = vp
        lib.ko_batch_process.argtypes = [vp, i16p, i16p, ip, ip]; lib.ko_batch_process.restype = None
        _lib = lib
    return _lib


def _i16(a):
    return a.ctypes.data_as(C.POINTER(C.c_int16))


def _f32(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


MODES = {"fp32": 0, "bf16": 1, "int8": 2}


class OracleModel:
    def __init__(self, path: str):
        self._lib = _load()
        self._h = C.c_void_p()
        rc = self._lib.ko_model_load(path.encode(), C.byref(self._h))
        if rc != 0:
            raise IOError(f"ko_model_load({path}) failed with code {rc}")
        self.hidden = self._lib.ko_model_hidden(self._h)
        self.layers = self._lib.ko_model_layers(self._h)

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.ko_model_free(self._h)
            self._h = None


class _StreamView:
    """State accessors shared by Oracle and the streams inside an OracleBatch."""

    def __init__(self, lib, handle, model: OracleModel):
        self._lib, self._s, self.model = lib, handle, model

    @property
    def h(self) -> np.ndarray:
        n = self.model.layers * self.model.hidden
        return np.ctypeslib.as_array(self._lib.ko_stream_h(self._s), shape=(n,)).reshape(self.model.layers, self.model.hidden)

    @property
    def ola(self) -> np.ndarray:
        return np.ctypeslib.as_array(self._lib.ko_stream_ola(self._s), shape=(FRAME,))

    @property
    def tail(self) -> np.ndarray:
        return np.ctypeslib.as_array(self._lib.ko_stream_tail(self._s), shape=(FRAME,))

    @property
    def last_mask(self) -> np.ndarray:
        return np.ctypeslib.as_array(self._lib.ko_stream_last_mask(self._s), shape=(BINS,)).copy()

    @property
    def last_feat(self) -> np.ndarray:
        return np.ctypeslib.as_array(self._lib.ko_stream_last_feat(self._s), shape=(BINS,)).copy()


class Oracle(_StreamView):
    """One stream; mirrors the reference handle: process / reset / delay_sample (pv_koala.h:65-100)."""

    frame_length = FRAME
    sample_rate = 16000
    delay_sample = 256

    def __init__(self, model: OracleModel, mode: str = "fp32"):
        lib = _load()
        super().__init__(lib, C.c_void_p(lib.ko_stream_new(model._h, MODES[mode])), model)

    def __del__(self):
        if getattr(self, "_s", None):
            self._lib.ko_stream_free(self._s)
            self._s = None

    def reset(self):
        self._lib.ko_stream_reset(self._s)

    def process(self, pcm) -> np.ndarray:
        pcm = np.ascontiguousarray(pcm, dtype=np.int16)
        assert pcm.shape == (FRAME,)
        out = np.empty(FRAME, np.int16)
        self._lib.ko_stream_process(self._s, _i16(pcm), _i16(out))
        return out

    def frontend(self, pcm):
        pcm = np.ascontiguousarray(pcm, dtype=np.int16)
        spec, feat = np.empty(512, np.float32), np.empty(BINS, np.float32)
        self._lib.ko_frontend(self._s, _i16(pcm), _f32(spec), _f32(feat))
        return spec, feat

    def masknet(self, feat):
        feat = np.ascontiguousarray(feat, dtype=np.float32)
        mask = np.empty(BINS, np.float32)
        self._lib.ko_masknet(self._s, _f32(feat), _f32(mask))
        return mask

    def masknet_q(self, feat_q):
        """Fixed-point mode only: the integer mask network on already quantised features (int16 Q14)."""
        feat_q = np.ascontiguousarray(feat_q, dtype=np.int16)
        mask = np.empty(BINS, np.float32)
        self._lib.ko_masknet_q(self._s, _i16(feat_q), _f32(mask))
        return mask

    def backend(self, spec, mask):
        spec = np.ascontiguousarray(spec, dtype=np.float32)
        mask = np.ascontiguousarray(mask, dtype=np.float32)
        out = np.empty(FRAME, np.int16)
        self._lib.ko_backend(self._s, _f32(spec), _f32(mask), _i16(out))
        return out


class OracleBatch:
    """B independent streams; pcm [B][T][256] int16 -> enhanced [B][T][256] int16, `threads` host threads."""

    def __init__(self, model: OracleModel, n_streams: int, mode: str = "fp32"):
        self._lib = _load()
        self.model, self.n = model, n_streams
        self._b = C.c_void_p(self._lib.ko_batch_new(model._h, n_streams, MODES[mode]))

    def __del__(self):
        if getattr(self, "_b", None):
            self._lib.ko_batch_free(self._b)
            self._b = None

    def reset(self):
        self._lib.ko_batch_reset(self._b)

    def stream(self, i: int) -> _StreamView:
        return _StreamView(self._lib, C.c_void_p(self._lib.ko_batch_stream(self._b, i)), self.model)

    def process(self, pcm: np.ndarray, threads: int = 1) -> np.ndarray:
        pcm = np.ascontiguousarray(pcm, dtype=np.int16)
        if pcm.ndim == 2:
            pcm = pcm[:, None, :]
            squeeze = True
        else:
            squeeze = False
        assert pcm.shape[0] == self.n and pcm.shape[2] == FRAME
        out = np.empty_like(pcm)
        self._lib.ko_batch_process(self._b, _i16(pcm), _i16(out), pcm.shape[1], threads)
        return out[:, 0, :] if squeeze else out
