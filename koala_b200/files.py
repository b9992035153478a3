"""File / clip front door: enhance whole 16 kHz mono clips, many at a time.

Semantics are those of the reference file demos (/root/reference/demo/c/koala_demo_file.c:466-521,
/root/reference/demo/python/koala_demo_file.py:96-116): frames of `frame_length` samples are fed until
`start < length + delay_sample`, the last frames zero-padded (feeding zeros flushes the delay line); the first
`delay_sample` output samples are dropped and the output is cut to the input length, so output[i] lines up with input[i].
Here every clip is one stream of a batch and all clips advance together; shorter clips simply see zero frames (their state
keeps evolving, their surplus output is cut away).  The engine object only needs `.process(int16 [B][T][256]) -> same`,
`.frame_length`, `.delay_sample`, `.sample_rate`, `.num_streams` -- `BatchKoala` in production, an oracle adapter in tests.
"""
from __future__ import annotations

import time
import wave
from typing import List, Optional, Sequence

import numpy as np

from ._koala import KoalaInvalidArgumentError


def read_wav(path: str, sample_rate: int = 16000) -> np.ndarray:
    """16-bit mono PCM at `sample_rate` -> int16 array; the same three checks as koala_demo_file.py:80-88."""
    with wave.open(path, 'rb') as f:
        if f.getframerate() != sample_rate:
            raise KoalaInvalidArgumentError('Invalid sample rate of `%d`. Koala only accepts `%d`' % (f.getframerate(), sample_rate))
        if f.getnchannels() != 1:
            raise KoalaInvalidArgumentError('Only single-channel WAV files can be processed')
        if f.getsampwidth() != 2:
            raise KoalaInvalidArgumentError('Only WAV files with 16-bit PCM encoding can be processed')
        return np.frombuffer(f.readframes(f.getnframes()), dtype='<i2').astype(np.int16)


def write_wav(path: str, pcm: np.ndarray, sample_rate: int = 16000) -> None:
    with wave.open(path, 'wb') as f:
        f.setnchannels(1)
        f.setsampwidth(2)
        f.setframerate(sample_rate)
        f.writeframes(np.ascontiguousarray(pcm, dtype='<i2').tobytes())


def enhance_clips(engine, clips: Sequence[np.ndarray], chunk_frames: int = 64) -> List[np.ndarray]:
    """Enhance `len(clips) <= engine.num_streams` clips (int16 arrays of any lengths) in lock-step; the engine must be freshly
    created or reset.  Returns one int16 array per clip, the same length as its input, delay already removed."""
    B, fl, delay = engine.num_streams, engine.frame_length, engine.delay_sample
    if len(clips) > B:
        raise KoalaInvalidArgumentError('more clips (%d) than streams (%d)' % (len(clips), B))
    lengths = [int(len(c)) for c in clips]
    if not lengths:
        return []
    # demo loop: while start < length + delay -> number of frames per clip; all clips run for the longest count
    n_frames = max((n + delay + fl - 1) // fl for n in lengths)
    outs = [np.empty(n_frames * fl, np.int16) for _ in clips]
    for t0 in range(0, n_frames, chunk_frames):
        tc = min(chunk_frames, n_frames - t0)
        block = np.zeros((B, tc, fl), np.int16)
        lo, hi = t0 * fl, (t0 + tc) * fl
        for s, c in enumerate(clips):
            seg = np.asarray(c[lo:hi], dtype=np.int16)                  # zero padding past the end = the flush
            block[s].reshape(-1)[:len(seg)] = seg
        enhanced = engine.process(block)
        for s in range(len(clips)):
            outs[s][lo:hi] = np.asarray(enhanced[s]).reshape(-1)
    return [outs[s][delay:delay + lengths[s]].copy() for s in range(len(clips))]


def enhance_files(input_paths: Sequence[str], output_paths: Sequence[str], engine=None, model_path: Optional[str] = None,
                  device: str = 'best', precision: str = 'bf16', library_path: Optional[str] = None) -> dict:
    """WAV in -> enhanced WAV out for a list of files, all processed as one batch on the GPU.  Returns timing in both
    conventions: `real_time_factor` = compute / audio as printed by the reference demo (koala_demo_file.c:526-527, lower
    is faster) and `rtf_x` = audio / compute."""
    if len(input_paths) != len(output_paths):
        raise KoalaInvalidArgumentError('need one output path per input path')
    for i, o in zip(input_paths, output_paths):
        if i == o:
            raise KoalaInvalidArgumentError('cannot overwrite an input path')
    own = engine is None
    if own:
        from ._batch import BatchKoala
        engine = BatchKoala(max(1, len(input_paths)), model_path=model_path, device=device, precision=precision,
                            library_path=library_path)
    try:
        clips = [read_wav(p, engine.sample_rate) for p in input_paths]
        t0 = time.perf_counter()
        enhanced = enhance_clips(engine, clips)
        compute_s = time.perf_counter() - t0
        for p, pcm in zip(output_paths, enhanced):
            write_wav(p, pcm, engine.sample_rate)
        audio_s = sum(len(c) for c in clips) / float(engine.sample_rate)
        return {'files': len(clips), 'audio_seconds': audio_s, 'compute_seconds': compute_s,
                'real_time_factor': compute_s / audio_s if audio_s else 0.0, 'rtf_x': audio_s / compute_s if compute_s else 0.0}
    finally:
        if own:
            engine.delete()


def main(argv=None) -> int:
    """`python -m koala_b200.files -i in1.wav in2.wav -o out1.wav out2.wav` -- batched twin of the reference file demo."""
    import argparse
    ap = argparse.ArgumentParser(description='Enhance 16 kHz mono 16-bit WAV files with koala_b200 (all files in one batch).')
    ap.add_argument('-i', '--input_paths', nargs='+', required=True)
    ap.add_argument('-o', '--output_paths', nargs='+', required=True)
    ap.add_argument('-m', '--model_path', default=None)
    ap.add_argument('-y', '--device', default='best')
    ap.add_argument('--precision', default='bf16', choices=['bf16', 'fp32', 'int8'])
    ap.add_argument('-l', '--library_path', default=None)
    args = ap.parse_args(argv)
    stats = enhance_files(args.input_paths, args.output_paths, model_path=args.model_path, device=args.device,
                          precision=args.precision, library_path=args.library_path)
    print('Processed %d file(s), %.2f s of audio' % (stats['files'], stats['audio_seconds']))
    print('Real time factor : %.6f' % stats['real_time_factor'])
    return 0


__all__ = ['enhance_clips', 'enhance_files', 'read_wav', 'write_wav']

if __name__ == '__main__':
    raise SystemExit(main())
