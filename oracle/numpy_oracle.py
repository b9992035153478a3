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


class NumpyFixedPoint:
    """SPEC.md section 6 (fixed-point mask network: int8 weights x int16 activations, integer gates), written independently of the C
    oracle's mode 2 with numpy integer arrays.  One stream; `masknet_q` maps int16 Q14 features to the Q15 mask and updates `hq`."""

    QF, QE, QH, QP, SIG_N = 14, 12, 15, 12, 2048

    def __init__(self, model: S.Model):
        self.m = model
        H, L = model.hidden, model.layers
        f32 = np.float32

        def scale(maxabs):
            return np.where(maxabs > 0, (maxabs / f32(127.0)).astype(f32), f32(1.0)).astype(f32)

        def quant(w, s):
            return np.clip(np.rint((w / s[:, None]).astype(f32)), -127, 127).astype(np.int64)

        def mult(s, q_in):
            return np.minimum(np.rint(s.astype(np.float64) * 2.0 ** (self.QP - q_in + 31)), 2 ** 31 - 1).astype(np.int64)

        def bias(b):
            return np.rint((b.astype(f32) * f32(4096.0)).astype(f32)).astype(np.int64)

        w = model["enc.weight"].astype(f32)
        s = scale(np.abs(w).max(axis=1))
        self.enc = (quant(w, s), mult(s, self.QF), bias(model["enc.bias"]))
        self.gru = []
        for l in range(L):
            c = f32(8.0 if l == 0 else 1.0)
            wi, wh = model[f"gru{l}.weight_ih"].astype(f32) * c, model[f"gru{l}.weight_hh"].astype(f32)
            bi, bh = model[f"gru{l}.bias_ih"].astype(f32), model[f"gru{l}.bias_hh"].astype(f32)
            s_rz = scale(np.maximum(np.abs(wi[:2 * H]).max(axis=1), np.abs(wh[:2 * H]).max(axis=1)))
            s_nx, s_nh = scale(np.abs(wi[2 * H:]).max(axis=1)), scale(np.abs(wh[2 * H:]).max(axis=1))
            self.gru.append(dict(
                qi=np.concatenate([quant(wi[:2 * H], s_rz), quant(wi[2 * H:], s_nx)]), qh=np.concatenate([quant(wh[:2 * H], s_rz), quant(wh[2 * H:], s_nh)]),
                m_rz=mult(s_rz, self.QH), b_rz=bias((bi[:2 * H] + bh[:2 * H]).astype(f32)),
                m_nx=mult(s_nx, self.QH), b_nx=bias(bi[2 * H:]), m_nh=mult(s_nh, self.QH), b_nh=bias(bh[2 * H:])))
        w = model["dec.weight"].astype(f32)
        s = scale(np.abs(w).max(axis=1))
        self.dec = (quant(w, s), mult(s, self.QH), bias(model["dec.bias"]))
        i = np.arange(self.SIG_N + 1, dtype=np.float64)
        self.sig_t = np.rint(32768.0 / (1.0 + np.exp(-(i - self.SIG_N / 2) / 128.0))).astype(np.int64)
        self.hq = np.zeros((L, H), np.int64)

    @staticmethod
    def wrap32(a):
        return ((a + 2 ** 31) % 2 ** 32) - 2 ** 31

    @staticmethod
    def requant(acc, m):
        return (acc * m + (1 << 30)) >> 31

    def sig(self, p):
        x = np.clip(p, -32768, 32767) + 32768
        i, f = x >> 5, x & 31
        return self.sig_t[i] + (((self.sig_t[i + 1] - self.sig_t[i]) * f + 16) >> 5)

    def tanh(self, a):
        return 2 * self.sig(np.where(a < -16384, -32768, np.where(a > 16383, 32767, 2 * a))) - 32768

    @staticmethod
    def quantize_feat(feat):
        return np.clip(np.rint((feat.astype(np.float32) * np.float32(16384.0)).astype(np.float32)), -32768, 32767).astype(np.int64)

    def masknet_q(self, fq):
        H = self.m.hidden
        q, m, b = self.enc
        x = np.clip(self.requant(self.wrap32(q @ np.asarray(fq, np.int64)), m) + b, 0, 32767)
        for l, g in enumerate(self.gru):
            h = self.hq[l]
            ai, ah = g["qi"] @ x, g["qh"] @ h
            p_rz = self.requant(self.wrap32(ai[:2 * H] + ah[:2 * H]), g["m_rz"]) + g["b_rz"]
            r, z = self.sig(p_rz[:H]), self.sig(p_rz[H:])
            pnx = self.requant(self.wrap32(ai[2 * H:]), g["m_nx"]) + g["b_nx"]
            pnh = self.requant(self.wrap32(ah[2 * H:]), g["m_nh"]) + g["b_nh"]
            n = self.tanh(pnx + ((r * pnh + (1 << 14)) >> 15))
            hn = np.clip(n + ((z * (h - n) + (1 << 14)) >> 15), -32767, 32767)
            self.hq[l] = hn
            x = hn
        q, m, b = self.dec
        return self.sig(self.requant(self.wrap32(q @ x), m) + b)       # Q15: mask = value / 32768
