"""Signal-path constants, lookup tables and the model-file format of koala_b200.

The reference engine (Picovoice Koala v3.0.0) is closed source: `/root/reference` ships only
`include/pv_koala.h`, bindings and prebuilt binaries.  The contract that IS published
-- frame_length 256, sample_rate 16000, int16 in / int16 out, a fixed positive delay, causal,
reset == fresh (`include/pv_koala.h:26-34,65-100`) -- is kept; everything inside it is specified by
this repository (SPEC.md) and restated here as numbers.  Three consumers share this file: the
trainer (`tools/train_weights.py`), the numpy restatement of the oracle (`oracle/numpy_oracle.py`)
and the tests.  The CUDA engine and the C oracle carry the same constants in C.
"""
from __future__ import annotations

import struct
import zlib
from dataclasses import dataclass
from typing import Dict

import numpy as np

SAMPLE_RATE = 16000          # pv_sample_rate(), picovoice.h:33-36 (value read from the reference .so)
FRAME_LENGTH = 256           # pv_koala_frame_length(), pv_koala.h:102-107 (value read from the reference .so)
N_FFT = 512                  # analysis/synthesis frame = previous frame + current frame
HOP = FRAME_LENGTH
N_BINS = 256                 # bins 0..255 feed the mask estimator; Nyquist bin 256 reuses mask[255]
HIDDEN = 512
LAYERS = 2
DELAY_SAMPLE = N_FFT - HOP   # 256: output of call t is the enhanced input of call t-1
VERSION = "1.0.0"

FEAT_POWER_SCALE = 2.0 ** -30    # |X|^2 of raw int16-scaled spectrum -> unit-scale power
FEAT_EPS = 1e-6
FEAT_GAIN = 0.125
FEAT_BIAS = 0.25

MAGIC = b"koala_b200\x00\x00"    # 12 bytes
FORMAT_VERSION = 1

# tensor name -> shape, in file order.  Weights are bf16 (stored as the upper 16 bits of the fp32
# pattern), biases fp32.  Gate order inside the GRU tensors is r | z | n (rows 0..H-1, H..2H-1, 2H..3H-1).
def tensor_layout(hidden: int = HIDDEN, layers: int = LAYERS, bins: int = N_BINS):
    out = [("enc.weight", (hidden, bins), "bf16"), ("enc.bias", (hidden,), "f32")]
    for l in range(layers):
        out += [
            (f"gru{l}.weight_ih", (3 * hidden, hidden), "bf16"),
            (f"gru{l}.weight_hh", (3 * hidden, hidden), "bf16"),
            (f"gru{l}.bias_ih", (3 * hidden,), "f32"),
            (f"gru{l}.bias_hh", (3 * hidden,), "f32"),
        ]
    out += [("dec.weight", (bins, hidden), "bf16"), ("dec.bias", (bins,), "f32")]
    return out


def window() -> np.ndarray:
    """sqrt-Hann (periodic) == sin(pi n / N); w[n]^2 + w[n+256]^2 == 1 (perfect reconstruction at hop 256)."""
    n = np.arange(N_FFT, dtype=np.float64)
    return np.sin(np.pi * n / N_FFT).astype(np.float32)


def twiddles() -> np.ndarray:
    """W_512^k = exp(-2 pi i k / 512), k = 0..255, as float32 (re, im) pairs computed in double."""
    k = np.arange(N_FFT // 2, dtype=np.float64)
    ang = -2.0 * np.pi * k / N_FFT
    return np.stack([np.cos(ang), np.sin(ang)], axis=1).astype(np.float32)


def bf16_round(x: np.ndarray) -> np.ndarray:
    """Round fp32 -> bf16 (round-to-nearest-even) and return as fp32 (NaN not expected)."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    r = ((u >> 16) & 1) + np.uint32(0x7FFF)
    return ((u + r) & np.uint32(0xFFFF0000)).view(np.float32)


def bf16_bits(x: np.ndarray) -> np.ndarray:
    return (bf16_round(x).view(np.uint32) >> 16).astype(np.uint16)


@dataclass
class Model:
    hidden: int
    layers: int
    bins: int
    tensors: Dict[str, np.ndarray]   # all fp32; weights already bf16-representable

    def __getitem__(self, k: str) -> np.ndarray:
        return self.tensors[k]


def save_model(path: str, tensors: Dict[str, np.ndarray], hidden: int = HIDDEN, layers: int = LAYERS,
               bins: int = N_BINS) -> None:
    layout = tensor_layout(hidden, layers, bins)
    body = bytearray()
    body += MAGIC
    body += struct.pack("<8I", FORMAT_VERSION, N_FFT, HOP, bins, hidden, layers, 1, len(layout))
    for name, shape, dt in layout:
        t = np.asarray(tensors[name], dtype=np.float32)
        if t.shape != shape:
            raise ValueError(f"{name}: shape {t.shape} != {shape}")
        body += (bf16_bits(t).astype("<u2").tobytes() if dt == "bf16" else t.astype("<f4").tobytes())
    body += struct.pack("<I", zlib.crc32(bytes(body)) & 0xFFFFFFFF)
    with open(path, "wb") as f:
        f.write(bytes(body))


def load_model(path: str) -> Model:
    with open(path, "rb") as f:
        blob = f.read()
    if len(blob) < 48 or blob[:12] != MAGIC:
        raise ValueError("not a koala_b200 model file")
    ver, n_fft, hop, bins, hidden, layers, dtype, n_t = struct.unpack("<8I", blob[12:44])
    if ver != FORMAT_VERSION or n_fft != N_FFT or hop != HOP or dtype != 1:
        raise ValueError("unsupported koala_b200 model version/geometry")
    (crc,) = struct.unpack("<I", blob[-4:])
    if crc != (zlib.crc32(blob[:-4]) & 0xFFFFFFFF):
        raise ValueError("model file checksum mismatch")
    off = 44
    tensors = {}
    layout = tensor_layout(hidden, layers, bins)
    if n_t != len(layout):
        raise ValueError("tensor count mismatch")
    for name, shape, dt in layout:
        n = int(np.prod(shape))
        if dt == "bf16":
            raw = np.frombuffer(blob, dtype="<u2", count=n, offset=off)
            tensors[name] = (raw.astype(np.uint32) << 16).view(np.float32).reshape(shape).copy()
            off += 2 * n
        else:
            tensors[name] = np.frombuffer(blob, dtype="<f4", count=n, offset=off).reshape(shape).copy()
            off += 4 * n
    if off != len(blob) - 4:
        raise ValueError("model file size mismatch")
    return Model(hidden, layers, bins, tensors)


def random_model(seed: int = 0x4B4F414C, hidden: int = HIDDEN, layers: int = LAYERS,
                 bins: int = N_BINS) -> Dict[str, np.ndarray]:
    """Seeded random-init weights of the spec's architecture (used by tests and as the synthetic-bench model)."""
    rng = np.random.default_rng(seed)
    out = {}
    for name, shape, dt in tensor_layout(hidden, layers, bins):
        fan_in = shape[-1] if len(shape) == 2 else hidden
        bound = 1.0 / np.sqrt(fan_in)
        t = rng.uniform(-bound, bound, size=shape).astype(np.float32)
        out[name] = bf16_round(t) if dt == "bf16" else t
    return out
