#!/usr/bin/env python
"""Record key-less known answers of the REFERENCE binary (/root/reference/lib/linux/x86_64/libpv_koala.so) into
tests/golden/abi_kat.json.  Runs only in the build container (the reference tree does not travel to the GPU box).

These are the only outputs of the reference engine obtainable without a licence key: constants,
status strings, error-stack semantics, argument validation.  tests/test_abi_cpu.py checks our library against them.
"""
import json
import os
import sys
from ctypes import POINTER, byref, c_char_p, c_int32, c_short, c_void_p, cdll

REF = "/root/reference/lib/linux/x86_64/libpv_koala.so"
MODEL = "/root/reference/lib/common/koala_params.pv"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "abi_kat.json")


def main():
    lib = cdll.LoadLibrary(REF)
    lib.pv_koala_version.restype = c_char_p
    lib.pv_status_to_string.restype = c_char_p
    lib.pv_get_sdk.restype = c_char_p

    def stack():
        ms, d = POINTER(c_char_p)(), c_int32(-7)
        st = lib.pv_get_error_stack(byref(ms), byref(d))
        out = [ms[i].decode() for i in range(d.value)] if d.value > 0 else []
        if ms:
            lib.pv_free_error_stack(ms)
        return {"status": st, "depth": d.value, "texts": [m.split(": ", 1)[1] if ": " in m else m for m in out]}

    kat = {
        "source": "nm -D + ctypes probes of " + REF,
        "exports": sorted(l.split()[-1] for l in os.popen("nm -D --defined-only " + REF).read().splitlines()
                          if l.split()[-1].startswith("pv_")),
        "frame_length": lib.pv_koala_frame_length(),
        "sample_rate": lib.pv_sample_rate(),
        "reference_version": lib.pv_koala_version().decode(),
        "default_sdk": lib.pv_get_sdk().decode(),
        "status_strings": {str(i): (lib.pv_status_to_string(i).decode() if lib.pv_status_to_string(i) else None)
                           for i in range(-1, 14)},
        "empty_stack": stack(),
    }
    h = c_void_p()
    cases = {
        "null_access_key": (None, MODEL.encode(), b"cpu", byref(h)),
        "null_model_path": (b"k", None, b"cpu", byref(h)),
        "null_object": (b"k", MODEL.encode(), b"cpu", None),
        "bad_device_tpu": (b"invalid", MODEL.encode(), b"tpu", byref(h)),
        "bad_device_upper": (b"invalid", MODEL.encode(), b"CPU", byref(h)),
        "gpu_unreachable": (b"invalid", MODEL.encode(), b"gpu:3", byref(h)),
        "missing_model": (b"invalid", b"/nonexistent.pv", b"cpu", byref(h)),
        "invalid_key": (b"invalid", MODEL.encode(), b"cpu", byref(h)),
        "empty_key": (b"", MODEL.encode(), b"cpu", byref(h)),
    }
    kat["init"] = {}
    for name, args in cases.items():
        st = lib.pv_koala_init(*args)
        kat["init"][name] = {"status": st, "stack": stack()}
    st1 = lib.pv_koala_init(b"invalid", MODEL.encode(), b"cpu", byref(h))
    st2 = lib.pv_koala_init(None, MODEL.encode(), b"cpu", byref(h))
    kat["two_failures_keep_latest"] = {"statuses": [st1, st2], "stack": stack(), "then": stack()}
    buf = (c_short * 256)()
    kat["process_null"] = {"status": lib.pv_koala_process(None, buf, buf), "stack": stack()}
    kat["reset_null"] = {"status": lib.pv_koala_reset(None), "stack": stack()}
    d = c_int32(-5)
    kat["delay_null"] = {"status": lib.pv_koala_delay_sample(None, byref(d)), "value_after": d.value, "stack": stack()}
    lib.pv_koala_delete(None)
    kat["delete_null"] = "no-op"
    n = c_int32()
    devs = POINTER(c_char_p)()
    kat["list_null_devices"] = {"status": lib.pv_koala_list_hardware_devices(None, byref(n)), "last_text": stack()["texts"][-1]}
    kat["list_null_count"] = {"status": lib.pv_koala_list_hardware_devices(byref(devs), None), "last_text": stack()["texts"][-1]}
    import importlib.util
    spec_ = importlib.util.spec_from_file_location("pvkoala_ref", "/root/reference/binding/python/__init__.py",
                                                   submodule_search_locations=["/root/reference/binding/python"])
    ref = importlib.util.module_from_spec(spec_)   # the reference binding, imported only to record its public names
    sys.modules["pvkoala_ref"] = ref
    spec_.loader.exec_module(ref)
    ref_koala = sys.modules["pvkoala_ref._koala"]
    kat["python_all"] = sorted(set(ref_koala.__all__) | set(sys.modules["pvkoala_ref._factory"].__all__)
                               | set(sys.modules["pvkoala_ref._util"].__all__))
    kat["python_koala_members"] = sorted(m for m in dir(ref_koala.Koala) if not m.startswith("_"))
    kat["python_status_enum"] = {m.name: m.value for m in ref_koala.Koala.PicovoiceStatuses}
    with open(OUT, "w") as f:
        json.dump(kat, f, indent=1, sort_keys=True)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
