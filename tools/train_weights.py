#!/usr/bin/env python
"""Train the mask estimator of SPEC.md on CPU and write `koala_b200/lib/koala_b200_params.kpv`.

Why this exists: the reference's parameters (`lib/common/koala_params.pv`) are an undocumented blob for a closed
engine, so this repository owns its weights.  The reference's behavioural tests
(`binding/python/test_koala.py:71-114`) need weights that really suppress `noise.wav` and pass `test.wav`; there is
no dataset offline, so training data = synthetic speech-like signals + the reference's two fixture WAVs, mixed with
synthetic noises + the noise fixture at random SNRs.  Deterministic given --seed.  Quality beyond the contract tests
is not claimed.

Usage: python tools/train_weights.py [--iters 1500] [--out koala_b200/lib/koala_b200_params.kpv]
"""
from __future__ import annotations

import argparse
import os
import sys
import time
import wave

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from koala_b200 import spec  # noqa: E402


def load_wav(path):
    with wave.open(path, "rb") as f:
        return np.frombuffer(f.readframes(f.getnframes()), dtype="<i2").astype(np.float32)


class MaskNet(nn.Module):
    def __init__(self, hidden=spec.HIDDEN, layers=spec.LAYERS, bins=spec.N_BINS):
        super().__init__()
        self.enc = nn.Linear(bins, hidden)
        self.gru = nn.GRU(hidden, hidden, num_layers=layers, batch_first=True)
        self.dec = nn.Linear(hidden, bins)

    def forward(self, feat, h=None):
        e = F.relu(self.enc(feat))
        y, h = self.gru(e, h)
        return torch.sigmoid(self.dec(y)), h


def stft(x, win):
    """x: [B, T*256] raw int16-scale float -> complex [B, T, 257]; frame t = [x[(t-1)*256 : (t+1)*256]] * win, zero history."""
    xp = F.pad(x, (spec.HOP, 0))
    frames = xp.unfold(1, spec.N_FFT, spec.HOP) * win
    return torch.fft.rfft(frames, dim=-1)


def features(X):
    p = (X.real ** 2 + X.imag ** 2)[..., : spec.N_BINS] * spec.FEAT_POWER_SCALE
    return spec.FEAT_GAIN * torch.log(p + spec.FEAT_EPS) + spec.FEAT_BIAS


def synth_speech(rng, n):
    """AM-modulated harmonic stack with drifting f0 and 3 formant-like resonances + fricative bursts + pauses."""
    t = np.arange(n) / spec.SAMPLE_RATE
    f0 = rng.uniform(90, 260) * (1 + 0.15 * np.sin(2 * np.pi * rng.uniform(0.3, 2.0) * t + rng.uniform(0, 6.28)))
    ph = 2 * np.pi * np.cumsum(f0) / spec.SAMPLE_RATE
    formants = rng.uniform([300, 900, 2200], [900, 2200, 3400])
    bw = rng.uniform(80, 300, 3)
    y = np.zeros(n)
    for h in range(1, 40):
        fh = h * np.mean(f0)
        if fh > 7000:
            break
        g = sum(1.0 / (1.0 + ((fh - fc) / b) ** 2) for fc, b in zip(formants, bw)) + 0.02
        y += g * np.sin(h * ph + rng.uniform(0, 6.28)) / h ** 0.3
    syl = rng.uniform(2.5, 6.0)
    env = np.clip(np.sin(2 * np.pi * syl * t + rng.uniform(0, 6.28)) + rng.uniform(-0.2, 0.5), 0, None) ** 1.5
    gate = (np.sin(2 * np.pi * rng.uniform(0.15, 0.5) * t + rng.uniform(0, 6.28)) > rng.uniform(-0.9, 0.2)).astype(float)
    gate = np.convolve(gate, np.ones(800) / 800, mode="same")
    fric = rng.standard_normal(n) * (np.clip(np.sin(2 * np.pi * syl * t + rng.uniform(0, 6.28)) - 0.7, 0, None) * 2)
    fric = np.diff(fric, prepend=0.0)
    y = (y / (np.abs(y).max() + 1e-9) + 0.15 * fric) * env * gate
    return y / (np.sqrt(np.mean(y ** 2)) + 1e-9)


def synth_noise(rng, n):
    kind = rng.integers(0, 5)
    w = rng.standard_normal(n + 512)
    if kind == 0:
        y = w
    elif kind == 1:      # pink-ish
        spec_ = np.fft.rfft(w)
        f = np.arange(len(spec_)) + 1.0
        y = np.fft.irfft(spec_ / f ** rng.uniform(0.3, 1.0), n=len(w))
    elif kind == 2:      # band-limited
        spec_ = np.fft.rfft(w)
        f = np.linspace(0, 8000, len(spec_))
        lo, hi = sorted(rng.uniform(50, 7000, 2))
        y = np.fft.irfft(spec_ * ((f > lo) & (f < hi + 300)), n=len(w))
    elif kind == 3:      # slowly modulated (babble-like)
        spec_ = np.fft.rfft(w)
        f = np.linspace(0, 8000, len(spec_))
        y = np.fft.irfft(spec_ / (1 + (f / 1000.0) ** 2), n=len(w))
        tt = np.arange(len(w)) / spec.SAMPLE_RATE
        y = y * (1 + 0.6 * np.sin(2 * np.pi * rng.uniform(0.5, 3) * tt))
    else:                # hum + hiss
        tt = np.arange(len(w)) / spec.SAMPLE_RATE
        y = 0.3 * w + sum(np.sin(2 * np.pi * rng.uniform(50, 400) * k * tt) / k for k in range(1, 5))
    y = y[:n]
    return y / (np.sqrt(np.mean(y ** 2)) + 1e-9)


def make_batch(rng, batch, frames, fix_speech, fix_noise):
    n = frames * spec.HOP
    S = np.zeros((batch, n), np.float32)
    N = np.zeros((batch, n), np.float32)
    for b in range(batch):
        u = rng.uniform()
        # speech
        if u < 0.45:       # fixture speech, random offset (offset 0 often: the tests start there)
            off = 0 if rng.uniform() < 0.5 else int(rng.integers(0, len(fix_speech) - n)) if len(fix_speech) > n else 0
            seg = fix_speech[off:off + n]
            seg = np.pad(seg, (0, n - len(seg)))
            s = seg * (1.0 if rng.uniform() < 0.6 else 10 ** rng.uniform(-0.7, 0.3))
        else:
            s = synth_speech(rng, n) * 32768 * 10 ** rng.uniform(-2.0, -0.7)
        mode = rng.uniform()
        if mode < 0.15:
            s = np.zeros(n)
        # noise
        v = rng.uniform()
        if v < 0.45:
            off = 0 if rng.uniform() < 0.5 else int(rng.integers(0, max(1, len(fix_noise) - n)))
            seg = fix_noise[off:off + n]
            seg = np.pad(seg, (0, n - len(seg)))
            nz = seg * (1.0 if rng.uniform() < 0.6 else 10 ** rng.uniform(-0.5, 0.4))
        else:
            nz = synth_noise(rng, n) * 32768 * 10 ** rng.uniform(-2.6, -1.2)
        if 0.15 <= mode < 0.35:
            nz = np.zeros(n)
        S[b], N[b] = s, nz
    return torch.from_numpy(S), torch.from_numpy(N)


def quantised_state_dict(model):
    sd = model.state_dict()
    t = {"enc.weight": sd["enc.weight"], "enc.bias": sd["enc.bias"],
         "dec.weight": sd["dec.weight"], "dec.bias": sd["dec.bias"]}
    for l in range(spec.LAYERS):
        t[f"gru{l}.weight_ih"] = sd[f"gru.weight_ih_l{l}"]
        t[f"gru{l}.weight_hh"] = sd[f"gru.weight_hh_l{l}"]
        t[f"gru{l}.bias_ih"] = sd[f"gru.bias_ih_l{l}"]
        t[f"gru{l}.bias_hh"] = sd[f"gru.bias_hh_l{l}"]
    return {k: v.detach().cpu().numpy().astype(np.float32) for k, v in t.items()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=1500)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--frames", type=int, default=128)
    ap.add_argument("--lr", type=float, default=1e-3)
    ap.add_argument("--seed", type=int, default=20261017)
    ap.add_argument("--out", default=os.path.join(ROOT, "koala_b200", "lib", "koala_b200_params.kpv"))
    ap.add_argument("--fixtures", default=os.path.join(ROOT, "tests", "golden"))
    args = ap.parse_args()

    torch.manual_seed(args.seed)
    rng = np.random.default_rng(args.seed)
    fix_speech = load_wav(os.path.join(args.fixtures, "test.wav"))
    fix_noise = load_wav(os.path.join(args.fixtures, "noise.wav"))
    win = torch.from_numpy(spec.window())
    model = MaskNet()
    opt = torch.optim.Adam(model.parameters(), lr=args.lr)
    sched = torch.optim.lr_scheduler.OneCycleLR(opt, max_lr=args.lr, total_steps=args.iters, pct_start=0.1)
    t0 = time.time()
    for it in range(args.iters):
        S, N = make_batch(rng, args.batch, args.frames, fix_speech, fix_noise)
        Xs, Xn = stft(S, win), stft(S + N, win)
        mask, _ = model(features(Xn))
        mag_n = Xn.abs()[..., : spec.N_BINS] / 32768.0
        mag_s = Xs.abs()[..., : spec.N_BINS] / 32768.0
        est = mask * mag_n
        loss = ((est + 1e-4) ** 0.5 - (mag_s + 1e-4) ** 0.5).pow(2).mean() * 10 + (est - mag_s).pow(2).mean()
        opt.zero_grad()
        loss.backward()
        nn.utils.clip_grad_norm_(model.parameters(), 5.0)
        opt.step()
        sched.step()
        if it % 25 == 0 or it == args.iters - 1:
            print(f"it {it:5d} loss {loss.item():.5f}  {time.time() - t0:.0f}s", flush=True)
        if (it % 250 == 0 and it > 0) or it == args.iters - 1:
            spec.save_model(args.out, quantised_state_dict(model))
    print("saved", args.out)


if __name__ == "__main__":
    main()
