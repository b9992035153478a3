"""A minimal ctypes consumer with the prototypes the reference Python binding declares
(/root/reference/binding/python/_koala.py:154-222: LoadLibrary, pv_set_sdk('python'), pv_get_error_stack / pv_free_error_stack,
pv_koala_init(c_char_p x3, POINTER(POINTER(CKoala))), pv_koala_delay_sample, pv_koala_process(POINTER(c_short) x2),
pv_koala_reset, pv_sample_rate, pv_koala_frame_length, pv_koala_version; _koala.py:315-340 for the device list).
The reference package itself does not travel to the GPU box; this replica (test code, not product) binds the engine the same
way so that the GPU tests exercise exactly the calls the unmodified binding makes."""
from ctypes import POINTER, Structure, byref, c_char_p, c_int, c_int32, c_short, cdll


class CKoala(Structure):
    pass


class EngineFailure(Exception):
    def __init__(self, status, stack):
        super().__init__("status %d: %s" % (status, " | ".join(stack)))
        self.status, self.stack = status, stack


class ReplicaKoala:
    def __init__(self, access_key, model_path, device, library_path):
        lib = self.lib = cdll.LoadLibrary(library_path)
        lib.pv_set_sdk.argtypes = [c_char_p]
        lib.pv_set_sdk.restype = None
        lib.pv_set_sdk('python'.encode('utf-8'))
        lib.pv_get_error_stack.argtypes = [POINTER(POINTER(c_char_p)), POINTER(c_int)]
        lib.pv_get_error_stack.restype = c_int
        lib.pv_free_error_stack.argtypes = [POINTER(c_char_p)]
        lib.pv_free_error_stack.restype = None
        lib.pv_koala_init.argtypes = [c_char_p, c_char_p, c_char_p, POINTER(POINTER(CKoala))]
        lib.pv_koala_init.restype = c_int
        self.handle = POINTER(CKoala)()
        status = lib.pv_koala_init(access_key.encode(), model_path.encode(), device.encode(), byref(self.handle))
        if status != 0:
            raise EngineFailure(status, self.error_stack())
        lib.pv_koala_delete.argtypes = [POINTER(CKoala)]
        lib.pv_koala_delete.restype = None
        lib.pv_koala_delay_sample.argtypes = [POINTER(CKoala), POINTER(c_int32)]
        lib.pv_koala_delay_sample.restype = c_int
        d = c_int32()
        status = lib.pv_koala_delay_sample(self.handle, d)
        if status != 0:
            raise EngineFailure(status, self.error_stack())
        self.delay_sample = d.value
        lib.pv_koala_process.argtypes = [POINTER(CKoala), POINTER(c_short), POINTER(c_short)]
        lib.pv_koala_process.restype = c_int
        lib.pv_koala_reset.argtypes = [POINTER(CKoala)]
        lib.pv_koala_reset.restype = c_int
        self.sample_rate = lib.pv_sample_rate()
        self.frame_length = lib.pv_koala_frame_length()
        lib.pv_koala_version.argtypes = []
        lib.pv_koala_version.restype = c_char_p
        self.version = lib.pv_koala_version().decode('utf-8')

    def error_stack(self):
        stack, depth = POINTER(c_char_p)(), c_int()
        if self.lib.pv_get_error_stack(byref(stack), byref(depth)) != 0:
            return []
        out = [stack[i].decode('utf-8') for i in range(depth.value)]
        self.lib.pv_free_error_stack(stack)
        return out

    def process(self, pcm):
        assert len(pcm) == self.frame_length
        frame = (c_short * len(pcm))(*pcm)
        enhanced = (c_short * len(pcm))()
        status = self.lib.pv_koala_process(self.handle, frame, enhanced)
        if status != 0:
            raise EngineFailure(status, self.error_stack())
        return list(enhanced)

    def reset(self):
        status = self.lib.pv_koala_reset(self.handle)
        if status != 0:
            raise EngineFailure(status, self.error_stack())

    def delete(self):
        self.lib.pv_koala_delete(self.handle)


def list_hardware_devices(library_path):
    lib = cdll.LoadLibrary(library_path)
    lib.pv_koala_list_hardware_devices.argtypes = [POINTER(POINTER(c_char_p)), POINTER(c_int32)]
    lib.pv_koala_list_hardware_devices.restype = c_int
    lib.pv_koala_free_hardware_devices.argtypes = [POINTER(c_char_p), c_int32]
    lib.pv_koala_free_hardware_devices.restype = None
    devs, n = POINTER(c_char_p)(), c_int32()
    status = lib.pv_koala_list_hardware_devices(byref(devs), byref(n))
    if status != 0:
        raise EngineFailure(status, [])
    out = [devs[i].decode('utf-8') for i in range(n.value)]
    lib.pv_koala_free_hardware_devices(devs, n)
    return out
