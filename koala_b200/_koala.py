"""Single-stream handle over the C ABI -- same class, method and exception names as the reference Python binding
(/root/reference/binding/python/_koala.py:19-358), so code written against `pvkoala.Koala` runs against this engine.

The hot call is `Koala.process` (_koala.py:224-254 in the reference): length check -> ctypes frame -> pv_koala_process
-> list.  There is no CPU implementation behind it: if libpv_koala_b200.so is missing or no B200 is visible the
constructor raises.
"""
import os
from array import array
from ctypes import CDLL, POINTER, Structure, byref, c_char_p, c_int, c_int32, c_short
from enum import Enum
from typing import Sequence


class KoalaError(Exception):
    """Base class of every engine error.  Besides the summary line it carries the engine's message stack
    (pv_get_error_stack), innermost message first, as the reference binding's exception does (same attribute names)."""

    def __init__(self, message: str = '', message_stack: Sequence[str] = None):
        super().__init__(message)
        self._message = message
        self._message_stack = tuple(message_stack or ())

    message = property(lambda self: self._message)
    message_stack = property(lambda self: self._message_stack)

    def __str__(self):
        head = self._message + (':' if self._message_stack else '')
        return '\n'.join([head] + ['  [%d] %s' % entry for entry in enumerate(self._message_stack)])


class KoalaMemoryError(KoalaError):
    pass


class KoalaIOError(KoalaError):
    pass


class KoalaInvalidArgumentError(KoalaError):
    pass


class KoalaStopIterationError(KoalaError):
    pass


class KoalaKeyError(KoalaError):
    pass


class KoalaInvalidStateError(KoalaError):
    pass


class KoalaRuntimeError(KoalaError):
    pass


class KoalaActivationError(KoalaError):
    pass


class KoalaActivationLimitError(KoalaError):
    pass


class KoalaActivationThrottledError(KoalaError):
    pass


class KoalaActivationRefusedError(KoalaError):
    pass


class PicovoiceStatuses(Enum):
    SUCCESS = 0
    OUT_OF_MEMORY = 1
    IO_ERROR = 2
    INVALID_ARGUMENT = 3
    STOP_ITERATION = 4
    KEY_ERROR = 5
    INVALID_STATE = 6
    RUNTIME_ERROR = 7
    ACTIVATION_ERROR = 8
    ACTIVATION_LIMIT_REACHED = 9
    ACTIVATION_THROTTLED = 10
    ACTIVATION_REFUSED = 11


PICOVOICE_STATUS_TO_EXCEPTION = {
    PicovoiceStatuses.OUT_OF_MEMORY: KoalaMemoryError,
    PicovoiceStatuses.IO_ERROR: KoalaIOError,
    PicovoiceStatuses.INVALID_ARGUMENT: KoalaInvalidArgumentError,
    PicovoiceStatuses.STOP_ITERATION: KoalaStopIterationError,
    PicovoiceStatuses.KEY_ERROR: KoalaKeyError,
    PicovoiceStatuses.INVALID_STATE: KoalaInvalidStateError,
    PicovoiceStatuses.RUNTIME_ERROR: KoalaRuntimeError,
    PicovoiceStatuses.ACTIVATION_ERROR: KoalaActivationError,
    PicovoiceStatuses.ACTIVATION_LIMIT_REACHED: KoalaActivationLimitError,
    PicovoiceStatuses.ACTIVATION_THROTTLED: KoalaActivationThrottledError,
    PicovoiceStatuses.ACTIVATION_REFUSED: KoalaActivationRefusedError,
}


def load_library(library_path: str) -> CDLL:
    """dlopen the engine and declare the error-stack entry points; raises (never falls back) if it is absent."""
    if not os.path.exists(library_path):
        raise KoalaIOError(
            "Could not find Koala's dynamic library at `%s` (build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'`)." % library_path)
    library = CDLL(library_path)
    library.pv_set_sdk.argtypes = [c_char_p]
    library.pv_set_sdk.restype = None
    library.pv_get_error_stack.argtypes = [POINTER(POINTER(c_char_p)), POINTER(c_int)]
    library.pv_get_error_stack.restype = c_int
    library.pv_free_error_stack.argtypes = [POINTER(c_char_p)]
    library.pv_free_error_stack.restype = None
    return library


def pop_error_stack(library: CDLL) -> Sequence[str]:
    stack_ref = POINTER(c_char_p)()
    depth = c_int()
    status = PicovoiceStatuses(library.pv_get_error_stack(byref(stack_ref), byref(depth)))
    if status is not PicovoiceStatuses.SUCCESS:
        raise PICOVOICE_STATUS_TO_EXCEPTION[status](message='Unable to get Koala error state')
    messages = [stack_ref[i].decode('utf-8') for i in range(depth.value)]
    library.pv_free_error_stack(stack_ref)
    return messages


def check(library: CDLL, status: int, message: str) -> None:
    status = PicovoiceStatuses(status)
    if status is not PicovoiceStatuses.SUCCESS:
        raise PICOVOICE_STATUS_TO_EXCEPTION[status](message=message, message_stack=pop_error_stack(library))


class Koala(object):
    """Python binding for the koala_b200 noise-suppression engine (one 16 kHz mono stream per instance)."""

    PicovoiceStatuses = PicovoiceStatuses
    _PICOVOICE_STATUS_TO_EXCEPTION = PICOVOICE_STATUS_TO_EXCEPTION

    class CKoala(Structure):
        pass

    def __init__(self, access_key: str, model_path: str, device: str, library_path: str) -> None:
        """
        :param access_key: AccessKey string.  Only its syntax is checked; there is no licence server behind this engine.
        :param model_path: Absolute path to a koala_b200 parameter file (`.kpv`).
        :param device: `best`, `gpu` or `gpu:${GPU_INDEX}`.  `cpu` / `cpu:${N}` are rejected: this engine is GPU-only.
        :param library_path: Absolute path to libpv_koala_b200.so.
        """
        if not isinstance(access_key, str) or len(access_key) == 0:
            raise KoalaInvalidArgumentError("`access_key` should be a non-empty string.")
        if not os.path.exists(model_path):
            raise KoalaIOError("Could not find model file at `%s`." % model_path)
        if not isinstance(device, str) or len(device) == 0:
            raise KoalaInvalidArgumentError("`device` should be a non-empty string.")

        library = load_library(library_path)
        library.pv_set_sdk('python'.encode('utf-8'))
        self._library = library

        library.pv_koala_init.argtypes = [c_char_p, c_char_p, c_char_p, POINTER(POINTER(self.CKoala))]
        library.pv_koala_init.restype = c_int
        self._handle = POINTER(self.CKoala)()
        check(library, library.pv_koala_init(access_key.encode(), model_path.encode(), device.encode(), byref(self._handle)),
              'Initialization failed')

        self._delete_func = library.pv_koala_delete
        self._delete_func.argtypes = [POINTER(self.CKoala)]
        self._delete_func.restype = None

        library.pv_koala_delay_sample.argtypes = [POINTER(self.CKoala), POINTER(c_int32)]
        library.pv_koala_delay_sample.restype = c_int
        delay_sample = c_int32()
        status = library.pv_koala_delay_sample(self._handle, delay_sample)
        if PicovoiceStatuses(status) is not PicovoiceStatuses.SUCCESS:
            self.delete()
            check(library, status, 'Failed to get delay samples')
        self._delay_sample = delay_sample.value

        self._process_func = library.pv_koala_process
        self._process_func.argtypes = [POINTER(self.CKoala), POINTER(c_short), POINTER(c_short)]
        self._process_func.restype = c_int

        self._reset_func = library.pv_koala_reset
        self._reset_func.argtypes = [POINTER(self.CKoala)]
        self._reset_func.restype = c_int

        self._sample_rate = library.pv_sample_rate()
        self._frame_length = library.pv_koala_frame_length()
        library.pv_koala_version.argtypes = []
        library.pv_koala_version.restype = c_char_p
        self._version = library.pv_koala_version().decode('utf-8')

    def process(self, pcm: Sequence[int]) -> Sequence[int]:
        """
        Processes a frame of audio and returns delayed enhanced audio.

        :param pcm: `.frame_length` 16-bit samples at `.sample_rate`; consecutive calls must carry consecutive frames of
        the same source unless `.reset()` was called in between.
        :return: `.frame_length` enhanced samples belonging to input given `.delay_sample` samples earlier.
        """
        if len(pcm) != self.frame_length:
            raise KoalaInvalidArgumentError(
                "Length of input frame %d does not match required frame length %d" % (len(pcm), self.frame_length))
        frame_type = c_short * self.frame_length
        try:
            # same bytes as `frame_type(*pcm)` for in-range ints, 5x cheaper to build (the call itself is ~65 us)
            frame = frame_type.from_buffer(array('h', pcm))
        except (OverflowError, TypeError):
            frame = frame_type(*pcm)       # whatever the reference binding's marshalling does with such input (wrap / raise)
        enhanced_pcm = frame_type()
        check(self._library, self._process_func(self._handle, frame, enhanced_pcm), 'Processing failed')
        return enhanced_pcm[:]

    def reset(self) -> None:
        """Resets Koala into a state as if it had just been newly created."""
        check(self._library, self._reset_func(self._handle), 'Reset failed')

    def delete(self) -> None:
        """Releases resources acquired by Koala."""
        self._delete_func(self._handle)

    @property
    def sample_rate(self) -> int:
        return self._sample_rate

    @property
    def frame_length(self) -> int:
        return self._frame_length

    @property
    def delay_sample(self) -> int:
        return self._delay_sample

    @property
    def version(self) -> str:
        return self._version

    def _get_error_stack(self) -> Sequence[str]:
        return pop_error_stack(self._library)


def list_hardware_devices(library_path: str) -> Sequence[str]:
    library = load_library(library_path)
    library.pv_koala_list_hardware_devices.argtypes = [POINTER(POINTER(c_char_p)), POINTER(c_int32)]
    library.pv_koala_list_hardware_devices.restype = c_int
    devices = POINTER(c_char_p)()
    count = c_int32()
    check(library, library.pv_koala_list_hardware_devices(byref(devices), byref(count)),
          '`pv_koala_list_hardware_devices` failed.')
    res = [devices[i].decode() for i in range(count.value)]
    library.pv_koala_free_hardware_devices.argtypes = [POINTER(c_char_p), c_int32]
    library.pv_koala_free_hardware_devices.restype = None
    library.pv_koala_free_hardware_devices(devices, count.value)
    return res


__all__ = [
    'Koala',
    'KoalaActivationError',
    'KoalaActivationLimitError',
    'KoalaActivationRefusedError',
    'KoalaActivationThrottledError',
    'KoalaError',
    'KoalaIOError',
    'KoalaInvalidArgumentError',
    'KoalaInvalidStateError',
    'KoalaKeyError',
    'KoalaMemoryError',
    'KoalaRuntimeError',
    'KoalaStopIterationError',
    'list_hardware_devices',
]
