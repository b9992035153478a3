"""Drop-in proof with the reference's own consumers ("boundary proof is by restatement only").

CPU part (this container; skipped where /root/reference does not exist, i.e. on the GPU box): the UNMODIFIED reference C file
demo (`demo/c/koala_demo_file.c`, compiled with gcc against its own dr_libs) and the UNMODIFIED reference Python package
(`binding/python`, imported from where it lies) are pointed at libpv_koala_b200.so.  Without a GPU every path that needs
one must fail the way the reference reports failures (status + message stack), and the key-less paths must work.

GPU part (`-m gpu`): nothing is vendored, so the same calls are made by two stand-ins that bind exactly what those consumers
bind -- tests/c/dropin_consumer.c (dlopen + the demo's eleven symbols + the demo's delay-trim loop) and
tests/ref_binding_replica.py (the binding's ctypes prototypes) -- and the enhanced audio is checked against the oracle."""
import importlib.util
import os
import subprocess
import sys

import numpy as np
import pytest

import koala_b200 as kb

from conftest import GOLDEN, ROOT

REF = "/root/reference"
needs_ref = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "demo", "c", "koala_demo_file.c")),
                               reason="the reference checkout is only present in the build container")


@pytest.fixture(scope="module")
def reference_demo(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("refdemo") / "koala_demo_file")
    subprocess.check_call(["gcc", "-std=c99", "-O1", "-I", os.path.join(REF, "include"), "-I", os.path.join(REF, "demo", "c", "dr_libs"),
                           os.path.join(REF, "demo", "c", "koala_demo_file.c"), "-ldl", "-lm", "-o", exe])
    return exe


@pytest.fixture(scope="module")
def reference_package():
    pkg = os.path.join(REF, "binding", "python")
    spec = importlib.util.spec_from_file_location("pvkoala_reference_under_test", os.path.join(pkg, "__init__.py"),
                                                  submodule_search_locations=[pkg])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = mod
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="module")
def consumer(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("consumer") / "dropin_consumer")
    subprocess.check_call(["gcc", "-std=c99", "-D_DEFAULT_SOURCE", "-O1", "-Wall", os.path.join(ROOT, "tests", "c", "dropin_consumer.c"),
                           "-ldl", "-o", exe])
    return exe


def _has_gpu():
    try:
        return len(kb.available_devices()) > 0
    except Exception:
        return False


@needs_ref
def test_unmodified_reference_c_demo_binds_our_library(reference_demo, library_path, shipped_model_path, tmp_path):
    # demo/c/test/test_koala_c.py:73-84: `-z` exits 0 with empty stderr (all eleven dlsym's resolved on the way)
    p = subprocess.run([reference_demo, "-l", library_path, "-z"], capture_output=True, text=True)
    assert p.returncode == 0 and p.stderr == ""
    assert all(line.startswith("gpu:") for line in p.stdout.split("\n") if line)
    # an init failure surfaces through the demo's own error path (koala_demo_file.c:336-360): status string + message stack
    out = str(tmp_path / "out.wav")
    p = subprocess.run([reference_demo, "-l", library_path, "-m", shipped_model_path, "-a", "invalid", "-y", "gpu",
                        "-i", os.path.join(GOLDEN, "test.wav"), "-o", out], capture_output=True, text=True)
    if _has_gpu():
        assert p.returncode != 0 and "INVALID_ARGUMENT" in p.stderr and "Failed to parse AccessKey" in p.stderr
        # ... and with a parseable key the whole demo runs: demo/c/test/test_koala_c.py:56-71
        p = subprocess.run([reference_demo, "-l", library_path, "-m", shipped_model_path, "-a", kb.ANY_ACCESS_KEY, "-y", "gpu",
                            "-i", os.path.join(GOLDEN, "test.wav"), "-o", out], capture_output=True, text=True)
        assert p.returncode == 0 and p.stderr == "" and "Real time factor" in p.stdout
    else:
        assert p.returncode != 0 and "Failed to init with 'RUNTIME_ERROR'" in p.stderr
        assert "Failed to communicate with device." in p.stderr
    p = subprocess.run([reference_demo, "-l", library_path, "-m", shipped_model_path, "-a", kb.ANY_ACCESS_KEY, "-y", "tpu",
                        "-i", os.path.join(GOLDEN, "test.wav"), "-o", out], capture_output=True, text=True)
    assert p.returncode != 0 and "INVALID_ARGUMENT" in p.stderr and "tpu is not a valid device string" in p.stderr


@needs_ref
def test_unmodified_reference_python_binding_binds_our_library(reference_package, library_path, shipped_model_path):
    ref = reference_package
    # test_koala.py:187-192 (device list: key-less)
    devices = ref.available_devices(library_path=library_path)
    assert isinstance(devices, list) and all(isinstance(d, str) and d.startswith("gpu:") for d in devices)
    # test_koala.py:136-162: a failing init raises the binding's own exception type with a non-empty, repeatable stack
    def failing_init():
        with pytest.raises(ref.KoalaError) as e:
            ref.create(access_key="invalid", model_path=shipped_model_path, device="gpu", library_path=library_path)
        return e.value
    first, second = failing_init(), failing_init()
    assert 0 < len(first.message_stack) < 8 and list(first.message_stack) == list(second.message_stack)
    if _has_gpu():
        assert isinstance(first, ref.KoalaInvalidArgumentError)          # "Failed to parse AccessKey"
        k = ref.create(access_key=kb.ANY_ACCESS_KEY, model_path=shipped_model_path, device="gpu", library_path=library_path)
        assert k.frame_length == 256 and k.sample_rate == 16000 and k.delay_sample >= 0 and len(k.version) > 0   # :58-62, :131-134
        frame = k.process([0] * k.frame_length)
        assert len(frame) == k.frame_length
        k.delete()
    else:
        assert isinstance(first, ref.KoalaRuntimeError)                  # no device: the reference's own status for that
    with pytest.raises(ref.KoalaInvalidArgumentError):
        ref.create(access_key="x", model_path=shipped_model_path, device="tpu", library_path=library_path)
    with pytest.raises(ref.KoalaIOError):
        ref.create(access_key="x", model_path="/nonexistent.kpv", device="gpu", library_path=library_path)


def test_c_consumer_and_binding_replica_without_gpu(consumer, library_path, shipped_model_path, tmp_path):
    """Runs everywhere: the stand-ins resolve every symbol; on a box without a GPU the init failure comes back as status + stack."""
    from ref_binding_replica import EngineFailure, ReplicaKoala, list_hardware_devices
    p = subprocess.run([consumer, library_path, "-z"], capture_output=True, text=True)
    assert p.returncode == 0 and p.stderr == ""
    assert list_hardware_devices(library_path) == [l for l in p.stdout.split("\n") if l]
    with pytest.raises(EngineFailure) as e:
        ReplicaKoala("invalid", shipped_model_path, "tpu", library_path)
    assert e.value.status == 3 and "tpu is not a valid device string" in e.value.stack[0]
    if not _has_gpu():
        with pytest.raises(EngineFailure) as e:
            ReplicaKoala(kb.ANY_ACCESS_KEY, shipped_model_path, "gpu", library_path)
        assert e.value.status == 7 and 0 < len(e.value.stack) < 8


@pytest.mark.gpu
def test_c_consumer_file_loop_on_gpu_matches_oracle(consumer, library_path, shipped_model_path, test_pcm, noise_pcm, tmp_path):
    """The demo loop (koala_demo_file.c:466-521) in C through dlopen, end to end on the fixture WAV: output length == input length,
    delay removed, samples within +-1 LSB of the oracle driven through the same loop; 'Real time factor' printed."""
    from oracle import Oracle, OracleModel
    mixed = np.clip(test_pcm.astype(np.int32) + noise_pcm.astype(np.int32), -32768, 32767).astype(np.int16)
    fin, fout = str(tmp_path / "in.raw"), str(tmp_path / "out.raw")
    mixed.astype("<i2").tofile(fin)
    p = subprocess.run([consumer, library_path, shipped_model_path, kb.ANY_ACCESS_KEY, "gpu", fin, fout], capture_output=True, text=True)
    assert p.returncode == 0 and p.stderr == "" and "Real time factor" in p.stdout, p.stderr
    got = np.fromfile(fout, dtype="<i2")
    assert got.size == mixed.size
    o = Oracle(OracleModel(shipped_model_path), "bf16")
    delay, fl, total = o.delay_sample, 256, mixed.size
    ref, start = [], 0
    while start < total + delay:
        frame = np.zeros(fl, np.int16)
        seg = mixed[start:start + fl]
        frame[:len(seg)] = seg
        ref.append(o.process(frame))
        start += fl
    ref = np.concatenate(ref)[delay:delay + total]
    assert np.abs(got.astype(np.int32) - ref.astype(np.int32)).max() <= 1
    p = subprocess.run([consumer, library_path, shipped_model_path, "invalid", "gpu", fin, fout], capture_output=True, text=True)
    assert p.returncode == 1 and "INVALID_ARGUMENT" in p.stderr and "Failed to parse AccessKey" in p.stderr
    p = subprocess.run([consumer, library_path, "-z"], capture_output=True, text=True)
    assert p.returncode == 0 and p.stdout.startswith("gpu:0 - ")


@pytest.mark.gpu
def test_binding_replica_replays_reference_tests_on_gpu(library_path, shipped_model_path, test_pcm, noise_pcm):
    """/root/reference/binding/python/test_koala.py:58-62, 71-134, 164-185 through the binding's exact prototypes."""
    import math
    from ref_binding_replica import EngineFailure, ReplicaKoala

    def rms(x):
        return math.sqrt(sum((v / 32768.0) ** 2 for v in x) / len(x))

    k = ReplicaKoala(kb.ANY_ACCESS_KEY, shipped_model_path, "gpu", library_path)
    assert k.frame_length > 0 and k.delay_sample >= 0 and len(k.version) > 0
    fl, delay = k.frame_length, k.delay_sample
    noisy = [int(a) + int(b) for a, b in zip(test_pcm[:60 * fl], noise_pcm[:60 * fl])]
    clean = test_pcm.tolist()
    for start in range(0, len(noisy) - fl + 1, fl):                              # test_mixed, first 60 frames
        frame = k.process(noisy[start:start + fl])
        dev = rms(frame) if start < delay else abs(rms(frame) - rms(clean[start - delay:start - delay + fl]))
        assert dev < 0.02
    k.reset()
    first = [k.process(clean[s:s + fl]) for s in range(0, 30 * fl, fl)]          # test_reset
    k.reset()
    assert all(k.process(clean[s:s + fl]) == first[i] for i, s in enumerate(range(0, 30 * fl, fl)))
    handle, k.handle = k.handle, None                                             # test_process_message_stack
    with pytest.raises(EngineFailure) as e:
        k.process([0] * fl)
    assert 0 < len(e.value.stack) < 8
    k.handle = handle
    k.delete()
