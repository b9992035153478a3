"""Parity of the CUDA path against the CPU oracle, through the C ABI (run on a B200: pytest -m gpu).

Tolerances (BASELINE.json north_star): enhanced int16 within +-1 LSB of the oracle on identical frames; internal mask path
within 1e-3 relative.  Stage tensors (features, spectrum, recurrent state) are compared with tolerances stated inline."""
import math

import numpy as np
import pytest

import koala_b200 as kb
from oracle import Oracle, OracleBatch, OracleModel

from conftest import synth_pcm

pytestmark = pytest.mark.gpu

LSB_TOL = 1          # int16 output
MASK_RTOL = 1e-3     # internal floating-point mask path


def bf16_to_f32(a):
    return (a.astype(np.uint32) << 16).view(np.float32)


def run_oracle(model_path, mode, pcm):
    ob = OracleBatch(OracleModel(model_path), pcm.shape[0], mode)
    return ob, ob.process(pcm, threads=8)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("n_streams", [1, 37, 130])
def test_batch_parity_against_oracle(library_path, random_model_path, precision, n_streams):
    frames = 24
    pcm = synth_pcm(n_streams, frames, seed=100 + n_streams)
    eng = kb.BatchKoala(n_streams, model_path=random_model_path, precision=precision)
    out = eng.process(pcm)
    ob, ref = run_oracle(random_model_path, precision, pcm)
    diff = np.abs(out.astype(np.int32) - ref.astype(np.int32))
    assert diff.max() <= LSB_TOL, (diff.max(), np.argwhere(diff > LSB_TOL)[:5])
    # internal tensors of the last step
    mask = eng.debug_read("mask", (n_streams, 256), np.float32)
    ref_mask = np.stack([ob.stream(s).last_mask for s in range(n_streams)])
    np.testing.assert_allclose(mask, ref_mask, rtol=MASK_RTOL, atol=1e-6)
    for l in range(2):
        h = eng.debug_read(f"h{l}", (n_streams, 512), np.float32)
        ref_h = np.stack([ob.stream(s).h[l] for s in range(n_streams)])
        np.testing.assert_allclose(h, ref_h, atol=2e-4)
    ola = eng.debug_read("ola", (n_streams, 256), np.float32)
    np.testing.assert_allclose(ola, np.stack([ob.stream(s).ola for s in range(n_streams)]), atol=0.25, rtol=1e-3)
    tail = eng.debug_read("tail", (n_streams, 256), np.int16)
    assert (tail == pcm[:, -1, :]).all()
    eng.delete()


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_frontend_stage_parity(library_path, random_model_path, precision):
    """STFT + features of one step: spectrum within 1e-5 of its scale, features within 1e-4 (fp32) / one bf16 ulp."""
    n = 9
    pcm = synth_pcm(n, 2, seed=5)
    eng = kb.BatchKoala(n, model_path=random_model_path, precision=precision)
    eng.process(pcm)
    om = OracleModel(random_model_path)
    spec = eng.debug_read("spec", (n, 512), np.float32)
    feat = eng.debug_read("feat", (n, 256), np.float32 if precision == "fp32" else np.uint16)
    for s in range(n):
        o = Oracle(om, precision)
        o.frontend(pcm[s, 0])
        rspec, rfeat = o.frontend(pcm[s, 1])
        np.testing.assert_allclose(spec[s], rspec, atol=1e-5 * np.abs(rspec).max())
        if precision == "fp32":
            np.testing.assert_allclose(feat[s], rfeat, atol=1e-4)
        else:
            np.testing.assert_allclose(bf16_to_f32(feat[s]), rfeat, atol=2 ** -7)   # |feat| < 2 -> bf16 ulp <= 2^-7
    eng.delete()


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_state_carry_reset_and_subset_reset(library_path, random_model_path, precision):
    """T frames in one call == T calls of one frame (config 5 state-carry); reset == fresh (pv_koala.h:82-90)."""
    n, frames = 6, 10
    pcm = synth_pcm(n, frames, seed=77)
    eng = kb.BatchKoala(n, model_path=random_model_path, precision=precision)
    whole = eng.process(pcm).copy()
    eng.reset()
    stepped = np.stack([eng.process(np.ascontiguousarray(pcm[:, t, :])) for t in range(frames)], axis=1)
    assert (stepped == whole).all()
    eng.reset()
    again = eng.process(pcm)
    assert (again == whole).all()                                   # bit-exact after reset (test_koala.py:116-129)
    # subset reset: streams 1 and 4 restart, the others keep their state
    eng.reset([1, 4])
    cont = eng.process(pcm)
    assert (cont[[1, 4]] == whole[[1, 4]]).all()
    assert not (cont[[0, 2, 3, 5]] == whole[[0, 2, 3, 5]]).all()
    with pytest.raises(kb.KoalaInvalidArgumentError):
        eng.reset([6])
    eng.delete()


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("hidden,layers", [(256, 1), (768, 3)])
def test_other_model_shapes(library_path, tmp_path, hidden, layers, precision):
    """The model file fixes the hidden size (a multiple of 256 on the tensor-core path) and the number of GRU layers; the fused
    kernel's tile list (L + 2 segments, H / 64 unit tiles, H / 64 k-blocks per part) must follow it.  300 streams = two
    256-stream tiles, so the per-tile dependency counters of both tiles are exercised."""
    from koala_b200 import spec
    path = str(tmp_path / f"model_{hidden}_{layers}.kpv")
    spec.save_model(path, spec.random_model(seed=hidden + layers, hidden=hidden, layers=layers), hidden=hidden, layers=layers)
    n, frames = 300, 6
    pcm = synth_pcm(n, frames, seed=hidden)
    eng = kb.BatchKoala(n, model_path=path, precision=precision)
    out = eng.process(pcm)
    ob, ref = run_oracle(path, precision, pcm)
    diff = np.abs(out.astype(np.int32) - ref.astype(np.int32))
    assert diff.max() <= LSB_TOL, (diff.max(), np.argwhere(diff > LSB_TOL)[:5])
    mask = eng.debug_read("mask", (n, 256), np.float32)
    np.testing.assert_allclose(mask, np.stack([ob.stream(s).last_mask for s in range(n)]), rtol=MASK_RTOL, atol=1e-6)
    for l in range(layers):
        h = eng.debug_read(f"h{l}", (n, hidden), np.float32)
        np.testing.assert_allclose(h, np.stack([ob.stream(s).h[l] for s in range(n)]), atol=2e-4)
    eng.delete()


def test_long_host_call_layouts_agree(library_path, random_model_path):
    """A 140-frame host call exercises the ingest pipeline's output blocks (32-frame blocks, a short last block, a 4-frame last
    chunk); the time-major entry point (host and device buffers) and frame-by-frame calls must give the same samples."""
    import os
    import torch
    n, frames = 6, 140
    pcm = synth_pcm(n, frames, seed=5)
    eng = kb.BatchKoala(n, model_path=random_model_path, precision="bf16")
    whole = eng.process(pcm).copy()          # few streams: the ingest path scales its chunks up to 64 frames (64 + 64 + 12)
    eng.reset()
    stepped = np.stack([eng.process(np.ascontiguousarray(pcm[:, t, :])) for t in range(frames)], axis=1)
    assert (stepped == whole).all()
    tm = np.ascontiguousarray(pcm.transpose(1, 0, 2))
    for chunk in ("8", "4"):                  # the chunk sizes a big batch gets: 8-frame input chunks in 32-frame output blocks; 4-frame time-major chunks
        os.environ["KOALA_HOST_CHUNK"] = chunk
        try:
            eng.reset()
            assert (eng.process(pcm) == whole).all(), chunk
            eng.reset()
            assert (eng.process(tm, time_major=True).transpose(1, 0, 2) == whole).all(), chunk
        finally:
            del os.environ["KOALA_HOST_CHUNK"]
    eng.reset()
    assert (eng.process(tm, time_major=True).transpose(1, 0, 2) == whole).all()
    eng.reset()
    d_out = eng.process(torch.from_numpy(tm).cuda(), time_major=True)
    torch.cuda.synchronize()
    assert (d_out.cpu().numpy().transpose(1, 0, 2) == whole).all()
    with pytest.raises(kb.KoalaInvalidArgumentError):
        eng.process(pcm, time_major=True)                           # wrong layout for the entry point
    eng.delete()


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_edge_inputs(library_path, random_model_path, precision):
    n = 4
    eng = kb.BatchKoala(n, model_path=random_model_path, precision=precision)
    zeros = np.zeros((n, 3, 256), np.int16)
    assert (eng.process(zeros) == 0).all()                          # silence in -> silence out
    full = np.empty((n, 6, 256), np.int16)
    full[0] = 32767
    full[1] = -32768
    full[2] = np.where(np.arange(256) % 2 == 0, 32767, -32768)
    full[3] = np.where(np.arange(256) % 64 < 32, 32767, -32768)
    eng.reset()
    out = eng.process(full)
    _, ref = run_oracle(random_model_path, precision, full)
    diff = np.abs(out.astype(np.int32) - ref.astype(np.int32))
    if precision == "fp32":
        assert diff.max() <= LSB_TOL                                 # saturating, full-scale cases
    else:
        # bf16 operands: GPU and oracle agree on the pre-rounding fp32 value to ~1e-6, so once in a while an operand
        # lands on the other side of a bf16 rounding boundary (2^-9 relative).  That moves the mask by ~3e-5, i.e. by
        # 1 LSB at |x| = 32767 -- the only level where it can show.  Bound of SPEC.md section 4: 1 + 1e-3 |x| LSB (measured: up to 6 LSB at the rails), quantified over
        # 10^6 samples in tests/test_gpu_baseline_sizes.py; after only 6 frames it is 2 LSB on under 10 % of the samples.
        assert diff.max() <= 2 and (diff > LSB_TOL).mean() < 0.10, (diff.max(), (diff > LSB_TOL).mean())
    assert eng.process(np.zeros((n, 0, 256), np.int16)).shape == (n, 0, 256)      # empty call is a no-op
    with pytest.raises(kb.KoalaInvalidArgumentError):
        eng.process(np.zeros((n + 1, 1, 256), np.int16))
    with pytest.raises(kb.KoalaInvalidArgumentError):
        eng.process(np.zeros((n, 1, 255), np.int16))
    eng.delete()


def test_device_tensors_and_full_size_properties(library_path, random_model_path):
    """BASELINE config sizes (8192 streams / GPU, bf16): properties that do not need the oracle at full size --
    duplicated streams agree bit for bit, silent streams stay silent -- plus an oracle spot check of 24 streams."""
    import torch
    n, frames = 8192, 6
    base = synth_pcm(64, frames, seed=9)
    pcm = np.tile(base, (n // 64, 1, 1))
    pcm[5::64] = 0
    eng = kb.BatchKoala(n, model_path=random_model_path, precision="bf16")
    d_in = torch.from_numpy(pcm).cuda()
    d_out = eng.process(d_in)
    torch.cuda.synchronize()
    out = d_out.cpu().numpy()
    assert (out.reshape(n // 64, 64, frames, 256) == out[:64][None]).all()        # replicas are bit-identical
    assert (out[5::64] == 0).all()
    pick = [0, 1, 2, 3, 6, 7, 17, 63]
    _, ref = run_oracle(random_model_path, "bf16", np.ascontiguousarray(pcm[pick]))
    assert np.abs(out[pick].astype(np.int32) - ref.astype(np.int32)).max() <= LSB_TOL
    host = eng.process(pcm) if False else None                                    # host path covered elsewhere
    # analysis, fused mask estimator, synthesis per chunk and per partition (batches above 4096 streams run as 4096-stream partitions)
    assert eng.kernel_launches == 3 * 2 * ((frames + eng.chunk_frames - 1) // eng.chunk_frames)
    eng.delete()


def _rms(x):
    return math.sqrt(float(np.mean((np.asarray(x, np.float64) / 32768.0) ** 2)))


@pytest.mark.parametrize("case", ["speech", "noise", "mixed"])
def test_reference_behaviour_through_single_stream_abi(library_path, shipped_model_path, test_pcm, noise_pcm, case):
    """/root/reference/binding/python/test_koala.py:71-114 restated against `Koala` (pv_koala_init/process/delete)."""
    if case == "speech":
        inp, ref = test_pcm, test_pcm
    elif case == "noise":
        inp, ref = noise_pcm, None
    else:
        inp = np.clip(test_pcm.astype(np.int32) + noise_pcm.astype(np.int32), -32768, 32767).astype(np.int16)
        ref = test_pcm
    o = kb.create(kb.ANY_ACCESS_KEY, model_path=shipped_model_path, device="gpu")
    oracle = Oracle(OracleModel(shipped_model_path), "bf16")
    try:
        assert o.frame_length == 256 and o.sample_rate == 16000 and o.delay_sample == 256 and len(o.version) > 0
        fl, delay = o.frame_length, o.delay_sample
        for start in range(0, len(inp) - fl + 1, fl):
            frame = o.process(inp[start:start + fl].tolist())
            assert len(frame) == fl
            want = oracle.process(inp[start:start + fl])
            assert np.abs(np.asarray(frame, np.int32) - want.astype(np.int32)).max() <= LSB_TOL
            energy = _rms(frame)
            if ref is None or start < delay:
                dev = energy
            else:
                dev = abs(energy - _rms(ref[start - delay:start - delay + fl]))
            assert dev < 0.02
    finally:
        o.delete()


def test_single_stream_reset_and_errors(library_path, shipped_model_path, test_pcm):
    o = kb.create(kb.ANY_ACCESS_KEY, model_path=shipped_model_path)
    frames = [test_pcm[i * 256:(i + 1) * 256].tolist() for i in range(40)]
    o.reset()
    first = [o.process(f) for f in frames]
    o.reset()
    assert all(o.process(f) == a for f, a in zip(frames, first))            # test_koala.py:116-129
    with pytest.raises(kb.KoalaInvalidArgumentError):
        o.process([0] * 255)
    handle, o._handle = o._handle, None                                      # test_koala.py:164-185
    with pytest.raises(kb.KoalaError) as e:
        o.process([0] * 256)
    assert 0 < len(e.value.message_stack) < 8
    o._handle = handle
    o.delete()
    assert len(kb.available_devices()) > 0                                   # test_koala.py:187-192
    assert all(d.startswith("gpu:") for d in kb.available_devices())


def test_init_errors_on_gpu_box_match_reference_kat(library_path, shipped_model_path):
    import json
    import os
    from conftest import GOLDEN
    kat = json.load(open(os.path.join(GOLDEN, "abi_kat.json")))
    with pytest.raises(kb.KoalaError) as e:                                  # test_koala.py:136-162
        kb.create("invalid", model_path=shipped_model_path, device="gpu")
    texts = [m.split(": ", 1)[1] for m in e.value.message_stack]
    assert kat["init"]["invalid_key"]["stack"]["texts"][1] in texts          # "Failed to parse AccessKey `invalid`."
    assert isinstance(e.value, kb.KoalaInvalidArgumentError)
    with pytest.raises(kb.KoalaError) as e2:
        kb.create("invalid", model_path=shipped_model_path, device="gpu")
    assert list(e2.value.message_stack) == list(e.value.message_stack)      # repeatable
    with pytest.raises(kb.KoalaRuntimeError):
        kb.create(kb.ANY_ACCESS_KEY, model_path=shipped_model_path, device="gpu:99")


def test_long_state_carry_does_not_drift(library_path, shipped_model_path, test_pcm, noise_pcm):
    """BASELINE configs[4] in miniature: state carried over 1400 frames (22 s) in chunks of 64 frames per call; the CUDA
    path must still sit within +-1 LSB of the oracle at the end, i.e. rounding differences do not accumulate in the
    recurrent state (shipped weights, real speech + noise, bf16 path)."""
    n, frames, chunk = 6, 1400, 64
    mixed = np.clip(test_pcm.astype(np.int32) + noise_pcm.astype(np.int32), -32768, 32767).astype(np.int16)
    src = [test_pcm, noise_pcm, mixed, test_pcm[5000:], noise_pcm[::-1].copy(), mixed[20000:]]
    pcm = np.stack([np.resize(s, frames * 256) for s in src]).reshape(n, frames, 256)
    eng = kb.BatchKoala(n, model_path=shipped_model_path, precision="bf16")
    outs = [eng.process(np.ascontiguousarray(pcm[:, t:t + chunk])) for t in range(0, frames, chunk)]
    out = np.concatenate(outs, axis=1)
    ob, ref = run_oracle(shipped_model_path, "bf16", pcm)
    diff = np.abs(out.astype(np.int32) - ref.astype(np.int32))
    assert diff.max() <= LSB_TOL, (diff.max(), np.argwhere(diff > LSB_TOL)[:5])
    assert diff[:, -100:].max() <= LSB_TOL
    # internal tensors after 1400 steps with trained weights: bf16 operand re-rounding (an operand on a rounding boundary
    # may round the other way on the GPU) shows up as ~1e-3 in h and ~3e-4 in the mask; it does not grow with time and never
    # reaches the int16 output (asserted above).  Measured: h 1.0e-3, mask 2.6e-4 abs (tools/gpu_long_carry.py).
    mask = eng.debug_read("mask", (n, 256), np.float32)
    np.testing.assert_allclose(mask, np.stack([ob.stream(s).last_mask for s in range(n)]), rtol=MASK_RTOL, atol=5e-4)
    for l in range(2):
        h = eng.debug_read(f"h{l}", (n, 512), np.float32)
        np.testing.assert_allclose(h, np.stack([ob.stream(s).h[l] for s in range(n)]), atol=3e-3)
    eng.delete()


def test_long_state_carry_fp32_is_tight(library_path, shipped_model_path, test_pcm, noise_pcm):
    """Same run on the fp32 path: no operand re-rounding, so state and mask stay within 1e-4 / 1e-3 relative."""
    n, frames = 3, 700
    mixed = np.clip(test_pcm.astype(np.int32) + noise_pcm.astype(np.int32), -32768, 32767).astype(np.int16)
    pcm = np.stack([np.resize(s, frames * 256) for s in (test_pcm, noise_pcm, mixed)]).reshape(n, frames, 256)
    eng = kb.BatchKoala(n, model_path=shipped_model_path, precision="fp32")
    out = eng.process(pcm)
    ob, ref = run_oracle(shipped_model_path, "fp32", pcm)
    assert np.abs(out.astype(np.int32) - ref.astype(np.int32)).max() <= LSB_TOL
    mask = eng.debug_read("mask", (n, 256), np.float32)
    np.testing.assert_allclose(mask, np.stack([ob.stream(s).last_mask for s in range(n)]), rtol=MASK_RTOL, atol=1e-6)
    for l in range(2):
        h = eng.debug_read(f"h{l}", (n, 512), np.float32)
        np.testing.assert_allclose(h, np.stack([ob.stream(s).h[l] for s in range(n)]), atol=1e-4)
    eng.delete()


def test_model_file_rejections_on_gpu_box(library_path, shipped_model_path, tmp_path):
    """pv_koala_init validation order on a box with a GPU: device ok -> model file -> AccessKey."""
    ref_like = str(tmp_path / "ref.pv")
    open(ref_like, "wb").write(b"koala3.0.0\x01\x01\x01\x00\x00\x11" + bytes(200))
    with pytest.raises(kb.KoalaInvalidArgumentError) as e:
        kb.Koala(kb.ANY_ACCESS_KEY, ref_like, "gpu", library_path)
    assert "library product is `koala_b200`" in e.value.message_stack[0]
    blob = bytearray(open(shipped_model_path, "rb").read())
    blob[5000] ^= 1
    bad = str(tmp_path / "bad.kpv")
    open(bad, "wb").write(bytes(blob))
    with pytest.raises(kb.KoalaInvalidArgumentError) as e:
        kb.Koala(kb.ANY_ACCESS_KEY, bad, "gpu", library_path)
    assert "corrupt" in e.value.message_stack[0]
    with pytest.raises(kb.KoalaInvalidArgumentError):                         # key is looked at after the model opened
        kb.Koala("invalid", shipped_model_path, "gpu", library_path)
