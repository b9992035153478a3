"""numpy restatement of SPEC.md, independent of the C oracle (cross-check; small cases only).  Test infrastructure.

Follows the same contract citations as koala_oracle.c: frame geometry pv_koala.h:65-80, delay :92-100, reset :82-90.
"""
from __future__ import annotations

import numpy as np

from koala_b200 import spec as S


class NumpyOracle:
    def __init__(self, model: S.Model, mode: str = "fp32"):
        self.m, self.mode = model, mode
        self.win = S.window()
        self.reset()

    def reset(self):
        self.tail = np.zeros(S.FRAME_LENGTH, np.int16)
        self.ola = np.zeros(S.FRAME_LENGTH, np.float32)
        self.h = np.zeros((self.m.layers, self.m.hidden), np.float32)

    def _q(self, x):
        return S.bf16_round(x) if self.mode == "bf16" else x.astype(np.float32)

    @staticmethod
    def _lin(x, W, b):
        # sequential-k accumulation in fp32, matching gemm_rows() of the C oracle
        acc = b.astype(np.float32).copy()
        Wt = np.ascontiguousarray(W.T)
        for k in range(Wt.shape[0]):
            acc += x[k] * Wt[k]
        return acc

    def frontend(self, pcm):
        frame = np.concatenate([self.tail, np.asarray(pcm, np.int16)]).astype(np.float32) * self.win
        X = np.fft.rfft(frame.astype(np.float64))
        self.tail = np.asarray(pcm, np.int16).copy()
        Xr, Xi = X.real.astype(np.float32), X.imag.astype(np.float32)
        Xi[0] = 0.0
        p = (Xr[:256] * Xr[:256] + Xi[:256] * Xi[:256]) * np.float32(S.FEAT_POWER_SCALE)
        feat = np.float32(S.FEAT_GAIN) * np.log(p + np.float32(S.FEAT_EPS)) + np.float32(S.FEAT_BIAS)
        return X, feat.astype(np.float32)

    def masknet(self, feat):
        m = self.m
        sig = lambda v: (1.0 / (1.0 + np.exp(-v, dtype=np.float32))).astype(np.float32)
        e = np.maximum(self._lin(self._q(feat), m["enc.weight"], m["enc.bias"]), 0.0).astype(np.float32)
        H = m.hidden
        for l in range(m.layers):
            gi = self._lin(self._q(e), m[f"gru{l}.weight_ih"], m[f"gru{l}.bias_ih"])
            gh = self._lin(self._q(self.h[l]), m[f"gru{l}.weight_hh"], m[f"gru{l}.bias_hh"])
            r = sig(gi[:H] + gh[:H])
            z = sig(gi[H:2 * H] + gh[H:2 * H])
            n = np.tanh(gi[2 * H:] + r * gh[2 * H:], dtype=np.float32)
            hn = ((np.float32(1.0) - z) * n + z * self.h[l]).astype(np.float32)
            self.h[l] = hn
            e = hn
        return sig(self._lin(self._q(e), m["dec.weight"], m["dec.bias"]))

    def backend(self, X, mask):
        mk = np.concatenate([mask, mask[-1:]]).astype(np.float64)
        Y = X * mk
        Y[0] = Y[0].real
        Y[256] = Y[256].real
        y = np.fft.irfft(Y, n=S.N_FFT).astype(np.float32) * self.win
        v = self.ola + y[:256]
        out = np.clip(np.rint(v), -32768, 32767).astype(np.int16)
        self.ola = y[256:].astype(np.float32)
        return out, v

    def process(self, pcm):
        X, feat = self.frontend(pcm)
        mask = self.masknet(feat)
        out, _ = self.backend(X, mask)
        return out
