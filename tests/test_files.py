"""File front door: the batched clip loop reproduces the reference demo's delay-trim / zero-flush
semantics (/root/reference/demo/python/koala_demo_file.py:96-116), checked against a literal restatement of that loop."""
import os

import numpy as np
import pytest

import koala_b200 as kb
from oracle import Oracle, OracleBatch, OracleModel

from conftest import GOLDEN, synth_pcm


class OracleEngine:
    """Adapter: the CPU oracle behind the engine interface enhance_clips() expects (test infrastructure)."""
    frame_length, delay_sample, sample_rate = 256, 256, 16000

    def __init__(self, model_path, n, mode):
        self.num_streams = n
        self._b = OracleBatch(OracleModel(model_path), n, mode)

    def process(self, pcm):
        return self._b.process(pcm, threads=4)


def demo_loop(engine, pcm):
    """The reference demo loop, one stream, restated sample by sample."""
    fl, delay, n = engine.frame_length, engine.delay_sample, len(pcm)
    out, start = [], 0
    while start < n + delay:
        end = start + fl
        frame = np.zeros(fl, np.int16)
        seg = pcm[start:end]
        frame[:len(seg)] = seg
        o = engine.process(frame)
        if end > delay:
            if end > n + delay:
                o = o[:n + delay - start]
            if start < delay:
                o = o[delay - start:]
            out.append(np.asarray(o, np.int16))
        start = end
    return np.concatenate(out) if out else np.zeros(0, np.int16)


def test_batched_clip_loop_equals_reference_demo_loop(random_model_path):
    lengths = [1000, 256, 255, 4096 + 17, 1]                       # ragged: shorter than a frame, exact multiple, long
    clips = [synth_pcm(1, (n + 255) // 256, seed=40 + i)[0].reshape(-1)[:n] for i, n in enumerate(lengths)]
    got = kb.enhance_clips(OracleEngine(random_model_path, 8, "bf16"), clips, chunk_frames=5)
    om = OracleModel(random_model_path)
    for clip, g in zip(clips, got):
        want = demo_loop(Oracle(om, "bf16"), clip)
        assert len(g) == len(clip) == len(want)
        assert (g == want).all()
    assert kb.enhance_clips(OracleEngine(random_model_path, 2, "bf16"), []) == []
    with pytest.raises(kb.KoalaInvalidArgumentError):
        kb.enhance_clips(OracleEngine(random_model_path, 1, "bf16"), clips)


def test_identity_alignment_and_wav_roundtrip(tmp_path, shipped_model_path, test_pcm):
    """Clean speech through the shipped weights comes back aligned sample for sample (energy deviation < 0.02 per frame,
    the reference's own criterion) and WAV I/O round-trips."""
    eng = OracleEngine(shipped_model_path, 1, "bf16")
    out = kb.enhance_clips(eng, [test_pcm])[0]
    assert len(out) == len(test_pcm)
    rms = lambda x: float(np.sqrt(np.mean((x.astype(np.float64) / 32768.0) ** 2)))
    for s in range(0, len(out) - 255, 256):
        assert abs(rms(out[s:s + 256]) - rms(test_pcm[s:s + 256])) < 0.02     # already delay-compensated
    p = str(tmp_path / "o.wav")
    kb.write_wav(p, out)
    assert (kb.read_wav(p) == out).all()
    assert (kb.read_wav(os.path.join(GOLDEN, "test.wav")) == test_pcm).all()
    stereo = str(tmp_path / "s.wav")
    import wave
    with wave.open(stereo, "wb") as f:
        f.setnchannels(2); f.setsampwidth(2); f.setframerate(16000); f.writeframes(bytes(8))
    with pytest.raises(kb.KoalaInvalidArgumentError):
        kb.read_wav(stereo)


@pytest.mark.gpu
def test_enhance_files_on_gpu_matches_oracle(tmp_path, shipped_model_path, test_pcm, noise_pcm):
    ins = [os.path.join(GOLDEN, "test.wav"), os.path.join(GOLDEN, "noise.wav")]
    outs = [str(tmp_path / "a.wav"), str(tmp_path / "b.wav")]
    stats = kb.enhance_files(ins, outs, model_path=shipped_model_path, device="gpu", precision="bf16")
    assert stats["files"] == 2 and stats["real_time_factor"] > 0 and stats["rtf_x"] > 1
    want = kb.enhance_clips(OracleEngine(shipped_model_path, 2, "bf16"), [test_pcm, noise_pcm])
    for path, w in zip(outs, want):
        g = kb.read_wav(path)
        assert len(g) == len(w) and np.abs(g.astype(np.int32) - w.astype(np.int32)).max() <= 1
    with pytest.raises(kb.KoalaInvalidArgumentError):
        kb.enhance_files(ins, ins)
