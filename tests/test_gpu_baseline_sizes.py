"""Oracle-checked parity at the BASELINE.json config sizes (run on a B200: pytest -m gpu).

Every 256-stream tile of the fused kernel's m dimension carries DISTINCT data and has streams compared with the CPU oracle
(>= 1 per tile, 32 tiles at 8192 streams), over 64 frames -- long enough for the dependency-counter epochs, the ping-pong
state buffers and the TMEM / staging rings to wrap many times at the full grid.  Each test also writes the histogram of
|cuda - oracle| in LSB over >= 10^6 compared samples to gpurun_out/parity_hist_*.json."""
import json
import os

import numpy as np
import pytest

import koala_b200 as kb
from oracle import OracleBatch, OracleModel

from conftest import ROOT, load_wav, synth_pcm

pytestmark = pytest.mark.gpu


def distinct_pcm(n_streams, n_frames, seed, pool=64):
    """[n_streams][n_frames][256] int16 with no two streams alike: a pool of synthetic speech-like / noise streams
     plus the two fixture WAVs, each further stream a copy rolled by a stream-specific number of samples
    and scaled by a stream-specific gain."""
    base = synth_pcm(pool, n_frames, seed=seed).reshape(pool, -1).astype(np.float32)
    wavs = [np.resize(load_wav(w), n_frames * 256).astype(np.float32) for w in ("test.wav", "noise.wav")]
    base[0], base[1] = wavs[0], wavs[1]
    base[2] = np.clip(wavs[0] + wavs[1], -32768, 32767)
    out = np.empty((n_streams, n_frames * 256), np.int16)
    for s in range(n_streams):
        rep = s // pool
        gain = 1.0 / (1.0 + 0.07 * (rep % 13))
        out[s] = np.rint(np.roll(base[s % pool], 131 * rep) * gain).astype(np.int16)
    return out.reshape(n_streams, n_frames, 256)


def lsb_histogram(out, ref):
    d = np.abs(out.astype(np.int32) - ref.astype(np.int32)).ravel()
    return {"samples": int(d.size), "max": int(d.max()), "hist": {str(k): int((d == k).sum()) for k in range(int(d.max()) + 1)}}


def dump(name, payload):
    d = os.path.join(ROOT, "gpurun_out")
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, name), "w") as f:
        json.dump(payload, f, indent=1)


def picked_streams(n_streams, per_tile, seed):
    rng = np.random.default_rng(seed)
    picks = []
    for tile in range((n_streams + 255) // 256):
        lo, hi = tile * 256, min(n_streams, tile * 256 + 256)
        picks += sorted(rng.choice(np.arange(lo, hi), size=min(per_tile, hi - lo), replace=False).tolist())
    return picks


@pytest.mark.parametrize("n_streams,precision,per_tile", [(256, "fp32", 256), (4096, "bf16", 4), (8192, "bf16", 2)])
def test_baseline_config_sizes_against_oracle(library_path, random_model_path, n_streams, precision, per_tile):
    """BASELINE.json configs[1] (256 streams, fp32 mask path), configs[2] (4096, bf16) and the per-GPU partition of
    configs[3] (8192, bf16): 64 frames, int16 within +-1 LSB of the oracle on the compared streams."""
    frames = 64
    pcm = distinct_pcm(n_streams, frames, seed=n_streams)
    assert len({pcm[s, 3].tobytes() for s in range(0, n_streams, 61)}) == len(range(0, n_streams, 61))
    eng = kb.BatchKoala(n_streams, model_path=random_model_path, precision=precision)
    out = eng.process(pcm)
    picks = picked_streams(n_streams, per_tile, seed=7)
    assert len(picks) * frames * 256 >= 10 ** 6
    ref = OracleBatch(OracleModel(random_model_path), len(picks), precision).process(np.ascontiguousarray(pcm[picks]), threads=os.cpu_count() or 8)
    hist = lsb_histogram(out[picks], ref)
    dump(f"parity_hist_{n_streams}_{precision}.json", {"streams": n_streams, "precision": precision, "frames": frames,
                                                       "compared_streams": len(picks), **hist})
    assert hist["max"] <= 1, hist
    # the state after 64 steps, on the compared streams
    h = [eng.debug_read(f"h{l}", (n_streams, 512), np.float32) for l in range(2)]
    ob = OracleBatch(OracleModel(random_model_path), len(picks), precision)
    ob.process(np.ascontiguousarray(pcm[picks]), threads=os.cpu_count() or 8)
    for l in range(2):
        np.testing.assert_allclose(h[l][picks], np.stack([ob.stream(i).h[l] for i in range(len(picks))]), atol=2e-4 if precision == "fp32" else 2e-3)
    eng.delete()


def full_scale_pcm(n, frames, seed=11):
    """+-32767 adversarial inputs: square waves, clipped sines, full-scale binary noise, DC rails."""
    rng = np.random.default_rng(seed)
    k = np.arange(frames * 256)
    pcm = np.empty((n, frames * 256), np.int16)
    for s in range(n):
        period = int(rng.integers(2, 200))
        kind = s % 4
        if kind == 0:
            pcm[s] = np.where((k // period) % 2 == 0, 32767, -32768)
        elif kind == 1:
            pcm[s] = np.clip(np.rint(40000.0 * np.sin(2 * np.pi * k / (period + 2.5))), -32768, 32767)
        elif kind == 2:
            pcm[s] = np.where(rng.random(k.size) < 0.5, 32767, -32768)
        else:
            pcm[s] = 32767 if s % 8 == 3 else -32768
    return pcm.reshape(n, frames, 256)


def test_full_scale_inputs_residual_is_quantified(library_path, random_model_path):
    """What the +-1 LSB bar means at +-32767 full scale (SPEC.md section 4), measured over > 10^6 samples per mode.

    fp32 mode: GPU and oracle masks agree to ~1e-6, so the int16 output stays within +-1 LSB at any level -- asserted.
    bf16 mode: both sides round every GEMM operand to bf16, but from fp32 values that differ in the last bits (summation
    order, FFT schedule, transcendental approximations), so now and then an operand lands on the other side of a bf16 rounding
    boundary (~0.5 per stream-step).  One such flip moves the mask by ~3e-5; carried through the recurrent state over 64 steps the
    masks differ by up to ~3e-4 absolute -- well inside the 1e-3 mask tolerance, invisible (< 1 LSB) for |x| < ~3000 (the synthetic
    BASELINE workloads: max 1 LSB, 99.7 % exact, test above), but up to 6 LSB measured at the rails.  SPEC.md section 4 states the
    general bound 1 + 1e-3 |x|; asserted here, tighter: |cuda - oracle| <= 9 LSB, and at most 1 LSB on at least 80 % of the samples."""
    n, frames = 64, 64
    pcm = full_scale_pcm(n, frames)
    for precision, max_lsb in (("fp32", 1), ("bf16", 9)):
        eng = kb.BatchKoala(n, model_path=random_model_path, precision=precision)
        out = eng.process(pcm)
        ob = OracleBatch(OracleModel(random_model_path), n, precision)
        ref = ob.process(pcm, threads=os.cpu_count() or 8)
        hist = lsb_histogram(out, ref)
        frac2 = sum(v for k_, v in hist["hist"].items() if int(k_) >= 2) / hist["samples"]
        mask = eng.debug_read("mask", (n, 256), np.float32)
        dmask = float(np.abs(mask - np.stack([ob.stream(s).last_mask for s in range(n)])).max())
        dump(f"parity_hist_full_scale_{precision}.json", {"streams": n, "frames": frames, "fraction_ge_2_lsb": frac2,
                                                          "max_abs_mask_difference_last_step": dmask, **hist})
        assert hist["samples"] >= 10 ** 6
        assert hist["max"] <= max_lsb and frac2 < 0.20, (precision, hist, frac2)
        assert dmask < (1e-5 if precision == "fp32" else 1e-3), (precision, dmask)
        eng.delete()


def test_config5_shape_state_carry_in_chunks(library_path, shipped_model_path):
    """BASELINE.json configs[4] shape: 1024 streams (the 8-GPU total; here on one GPU = four 256-stream tiles), state carried
    across 32 process() calls of 64 frames each (2048 frames = 33 s per stream), trained weights, distinct streams.  16 streams
    (4 per tile) are followed by the oracle over the whole run."""
    n, frames, chunk = 1024, 2048, 64
    pcm = distinct_pcm(n, frames, seed=5, pool=32)
    eng = kb.BatchKoala(n, model_path=shipped_model_path, precision="bf16")
    out = np.concatenate([eng.process(np.ascontiguousarray(pcm[:, t:t + chunk])) for t in range(0, frames, chunk)], axis=1)
    picks = picked_streams(n, 4, seed=3)
    ref = OracleBatch(OracleModel(shipped_model_path), len(picks), "bf16").process(np.ascontiguousarray(pcm[picks]), threads=os.cpu_count() or 8)
    hist = lsb_histogram(out[picks], ref)
    d = np.abs(out[picks].astype(np.int32) - ref.astype(np.int32))
    frac_le1 = float((d <= 1).mean())
    # bf16 mode (SPEC.md section 4): +-1 LSB except where a re-rounded GEMM operand moved the mask -- the mask tolerance is 1e-3
    # relative, so the output bound is 1 + 1e-3 |x| LSB; with the trained weights the loud streams (|x| ~ 3000-4000) show 2 LSB on
    # ~3e-5 of the samples (same histogram before and after the time-persistent kernel: the chunked drive is bit-identical to
    # the one-frame-per-call drive, tests/test_gpu_chunked.py)
    bound = 1 + np.ceil(1e-3 * np.abs(ref.astype(np.int32))).astype(np.int32)
    dump("parity_hist_cfg5_1024x2048_bf16.json", {"streams": n, "frames": frames, "chunk": chunk, "compared_streams": len(picks),
                                                  "fraction_within_1_lsb": frac_le1, **hist})
    assert (d <= bound).all(), hist
    assert hist["max"] <= 2 and frac_le1 >= 0.9999, (hist, frac_le1)
    first, last = d[:, :frames // 4], d[:, -frames // 4:]
    assert (last >= 1).mean() <= 2.0 * (first >= 1).mean() + 1e-3      # no drift: the end of the run looks like its beginning
    eng.delete()


def test_every_stream_of_the_full_batch_against_oracle(library_path, random_model_path):
    """The per-GPU partition of BASELINE configs[3] with NO sampling: all 8192 streams (distinct data, both 4096-stream partitions,
    all 32 stream tiles) x 32 frames = 67 M samples compared with the oracle; +-1 LSB everywhere, histogram written out."""
    n_streams, frames = 8192, 32
    pcm = distinct_pcm(n_streams, frames, seed=81920)
    eng = kb.BatchKoala(n_streams, model_path=random_model_path, precision="bf16")
    out = eng.process(np.ascontiguousarray(pcm.transpose(1, 0, 2)), time_major=True).transpose(1, 0, 2)
    eng.delete()
    ref = OracleBatch(OracleModel(random_model_path), n_streams, "bf16").process(pcm, threads=os.cpu_count() or 8)
    hist = lsb_histogram(out, ref)
    dump("parity_hist_8192_bf16_every_stream.json", {"streams": n_streams, "precision": "bf16", "frames": frames, "compared_streams": n_streams, **hist})
    assert hist["samples"] == n_streams * frames * 256 and hist["max"] <= 1, hist


def test_every_stream_of_a_fixed_point_batch_against_oracle(library_path, random_model_path):
    """Fixed-point mode at BASELINE configs[2] size: all 4096 streams x 16 frames against the oracle's mode 2 (+-1 LSB; the integer mask
    network itself is compared for equality in tests/test_gpu_fixed_point.py)."""
    n_streams, frames = 4096, 16
    pcm = distinct_pcm(n_streams, frames, seed=40960)
    eng = kb.BatchKoala(n_streams, model_path=random_model_path, precision="int8")
    out = eng.process(pcm)
    eng.delete()
    ref = OracleBatch(OracleModel(random_model_path), n_streams, "int8").process(pcm, threads=os.cpu_count() or 8)
    hist = lsb_histogram(out, ref)
    dump("parity_hist_4096_int8_every_stream.json", {"streams": n_streams, "precision": "int8", "frames": frames, "compared_streams": n_streams, **hist})
    assert hist["max"] <= 1, hist
