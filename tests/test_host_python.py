"""Host-side mirror of the reference Python package: same public names, same argument validation
(/root/reference/binding/python/__init__.py:12-14, _koala.py:122-152,239-241), model-file format, stream partition."""
import json
import os

import numpy as np
import pytest

import koala_b200 as kb
from koala_b200 import spec

from conftest import GOLDEN


@pytest.fixture(scope="module")
def kat():
    with open(os.path.join(GOLDEN, "abi_kat.json")) as f:
        return json.load(f)


def test_public_names_cover_the_reference_package(kat):
    for name in kat["python_all"]:
        assert hasattr(kb, name), name
    for member in kat["python_koala_members"]:
        assert hasattr(kb.Koala, member), member
    assert {m.name: m.value for m in kb.Koala.PicovoiceStatuses} == kat["python_status_enum"]
    assert issubclass(kb.KoalaActivationLimitError, kb.KoalaError)


def test_constructor_argument_validation(library_path, random_model_path):
    with pytest.raises(kb.KoalaInvalidArgumentError):
        kb.Koala("", random_model_path, "gpu", library_path)
    with pytest.raises(kb.KoalaIOError):
        kb.Koala("k", "/nonexistent/model.kpv", "gpu", library_path)
    with pytest.raises(kb.KoalaInvalidArgumentError):
        kb.Koala("k", random_model_path, "", library_path)
    with pytest.raises(kb.KoalaIOError):                                     # missing engine: loud, no fallback
        kb.Koala("k", random_model_path, "gpu", "/nonexistent/libpv_koala_b200.so")
    with pytest.raises(kb.KoalaError) as e:                                  # test_koala.py:136-162 (invalid key)
        kb.create("invalid", model_path=random_model_path, device="tpu")
    assert len(e.value.message_stack) > 0 and "[0]" in str(e.value)
    assert kb.default_library_path().endswith("libpv_koala_b200.so")
    assert os.path.exists(kb.default_model_path())
    assert isinstance(kb.available_devices(), list)


def test_model_file_roundtrip_and_rejection(tmp_path):
    t = spec.random_model(seed=3)
    p = str(tmp_path / "a.kpv")
    spec.save_model(p, t)
    m = spec.load_model(p)
    assert m.hidden == 512 and m.layers == 2
    for k, v in t.items():
        np.testing.assert_array_equal(m[k], v)                               # weights are bf16-exact, biases fp32
    assert (spec.bf16_round(m["gru0.weight_hh"]) == m["gru0.weight_hh"]).all()
    blob = bytearray(open(p, "rb").read())
    blob[1000] ^= 0x40
    bad = str(tmp_path / "bad.kpv")
    open(bad, "wb").write(bytes(blob))
    with pytest.raises(ValueError):
        spec.load_model(bad)
    from oracle import OracleModel
    with pytest.raises(IOError):
        OracleModel(bad)
    ref_like = str(tmp_path / "ref.pv")
    open(ref_like, "wb").write(b"koala3.0.0\x01\x01\x01\x00\x00\x11" + bytes(100))   # magic of the reference blob
    with pytest.raises(ValueError):
        spec.load_model(ref_like)


def test_bf16_rounding_is_nearest_even():
    x = np.array([1.0, 1.00390625, 1.01171875, -1.00390625, 3.0e38, 1e-40], np.float32)
    r = spec.bf16_round(x)
    assert r[0] == 1.0 and r[1] == 1.0 and r[2] == np.float32(1.015625) and r[3] == -1.0   # ties to even
    import torch
    ref = torch.from_numpy(x).to(torch.bfloat16).to(torch.float32).numpy()
    np.testing.assert_array_equal(r[:4], ref[:4])


def test_tables():
    w = spec.window()
    np.testing.assert_allclose(w[:256] ** 2 + w[256:] ** 2, 1.0, atol=1e-6)  # perfect reconstruction at hop 256
    tw = spec.twiddles()
    np.testing.assert_allclose(tw[:, 0] ** 2 + tw[:, 1] ** 2, 1.0, atol=1e-6)
    assert tw[0, 0] == 1.0 and tw[128, 1] == -1.0


def test_shard_streams_partitions_exactly():
    for total in (0, 1, 7, 1024, 65536, 65537):
        for world in (1, 2, 3, 8):
            spans = [kb.shard_streams(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == total
            for (s0, c0), (s1, _) in zip(spans, spans[1:]):
                assert s0 + c0 == s1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
            for s in {0, total // 2, total - 1} - {-1}:
                if total:
                    r = kb.owner_of(s, total, world)
                    assert spans[r][0] <= s < spans[r][0] + spans[r][1]
    with pytest.raises(ValueError):
        kb.shard_streams(8, 2, 2)


def test_batch_layout_validation_without_a_device():
    """BatchKoala.process accepts [B][T][256] or, time-major, [T][B][256]; the shape check runs before any library call."""
    from koala_b200._batch import BatchKoala
    from koala_b200 import KoalaInvalidArgumentError
    b = object.__new__(BatchKoala)
    b.num_streams, b.frame_length = 6, 256
    assert b._shape((6, 10, 256)) == 10 and b._shape((6, 256)) == 1
    assert b._shape((10, 6, 256), time_major=True) == 10 and b._shape((6, 256), time_major=True) == 1
    for shape, tm in (((10, 6, 256), False), ((6, 10, 256), True), ((6, 10, 255), False), ((6,), False)):
        with pytest.raises(KoalaInvalidArgumentError):
            b._shape(shape, time_major=tm)
    b._handle = None      # nothing to release
