"""The fixed-point variant on the GPU (run on a B200: pytest -m gpu): int8 weights x int16 activations on tcgen05 kind::i8,
integer requantisation and table-driven gates (SPEC.md section 6; SURVEY.md section 8f row 3 -- the reference engine's numeric
style, /root/reference/include/pv_koala.h:65-80 being the contract it runs behind).

Between the quantised features and the mask the path is pure integer arithmetic, so the CUDA kernels must reproduce the CPU
oracle (mode 2) BIT FOR BIT there; analysis and synthesis are the fp32 stages of the other modes (+-1 LSB)."""
import os

import numpy as np
import pytest

import koala_b200 as kb
from oracle import Oracle, OracleBatch, OracleModel

from conftest import synth_pcm

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]


@pytest.mark.parametrize("n_streams", [5, 130, 300])
def test_integer_mask_network_is_bit_exact(library_path, random_model_path, n_streams):
    """Every step: take the features the GPU quantised (int16 Q14), run the oracle's integer mask network on exactly those, and
    compare the mask (Q15) and both layers' state (Q15) with what the GPU computed -- equality, not tolerance."""
    frames = 8
    pcm = synth_pcm(n_streams, frames, seed=900 + n_streams)
    eng = kb.BatchKoala(n_streams, model_path=random_model_path, precision="int8")
    om = OracleModel(random_model_path)
    streams = [Oracle(om, "int8") for _ in range(n_streams)]
    for t in range(frames):
        eng.process(np.ascontiguousarray(pcm[:, t]))
        fq = eng.debug_read("feat", (n_streams, 256), np.int16)
        mask = eng.debug_read("mask", (n_streams, 256), np.float32)
        h = [eng.debug_read(f"h{l}", (n_streams, 512), np.float32) for l in range(2)]
        for s, o in enumerate(streams):
            ref_mask = o.masknet_q(fq[s])
            assert (mask[s] == ref_mask).all(), (t, s, np.abs(mask[s] - ref_mask).max())
            for l in range(2):
                assert (h[l][s] == o.h[l]).all(), (t, s, l)
    eng.delete()


def test_fixed_point_end_to_end_against_oracle(library_path, random_model_path):
    """Whole path against the oracle in the same mode: the fp32 analysis differs in the last bits (FFT schedule, log), so now and
    then a feature lands on the other side of a Q14 rounding step (6e-5); everything downstream is exact.  Enhanced samples within
    +-1 LSB, mask within 1e-3."""
    n, frames = 260, 64
    pcm = synth_pcm(n, frames, seed=31)
    eng = kb.BatchKoala(n, model_path=random_model_path, precision="int8")
    out = eng.process(pcm)
    ob = OracleBatch(OracleModel(random_model_path), n, "int8")
    ref = ob.process(pcm, threads=os.cpu_count() or 8)
    diff = np.abs(out.astype(np.int32) - ref.astype(np.int32))
    assert diff.max() <= 1, (diff.max(), (diff > 0).mean())
    mask = eng.debug_read("mask", (n, 256), np.float32)
    np.testing.assert_allclose(mask, np.stack([ob.stream(s).last_mask for s in range(n)]), rtol=1e-3, atol=1e-4)
    fq = eng.debug_read("feat", (n, 256), np.int16).astype(np.int32)
    ref_fq = np.clip(np.rint(np.stack([ob.stream(s).last_feat for s in range(n)]) * 16384.0), -32768, 32767).astype(np.int32)
    assert np.abs(fq - ref_fq).max() <= 1 and (fq != ref_fq).mean() < 0.05
    os.makedirs("gpurun_out", exist_ok=True)
    import json
    with open("gpurun_out/parity_hist_fixed_point_260.json", "w") as f:
        json.dump({"streams": n, "frames": frames, "samples": int(diff.size), "max": int(diff.max()),
                   "hist": {str(k): int((diff == k).sum()) for k in range(int(diff.max()) + 1)},
                   "features_off_by_one_fraction": float((fq != ref_fq).mean())}, f, indent=1)
    eng.delete()


def test_fixed_point_state_carry_and_reset(library_path, random_model_path):
    """T frames in one call == T calls; reset == fresh handle; per-stream reset (pv_koala.h:82-90) -- bit for bit."""
    n, frames = 6, 10
    pcm = synth_pcm(n, frames, seed=77)
    eng = kb.BatchKoala(n, model_path=random_model_path, precision="int8")
    whole = eng.process(pcm).copy()
    eng.reset()
    stepped = np.stack([eng.process(np.ascontiguousarray(pcm[:, t, :])) for t in range(frames)], axis=1)
    assert (stepped == whole).all()
    eng.reset([1, 4])
    cont = eng.process(pcm)
    assert (cont[[1, 4]] == whole[[1, 4]]).all()
    assert not (cont[[0, 2, 3, 5]] == whole[[0, 2, 3, 5]]).all()
    assert (eng.process(np.zeros((n, 0, 256), np.int16)).shape == (n, 0, 256))
    eng.reset()
    assert (eng.process(np.zeros((n, 3, 256), np.int16)) == 0).all()      # silence in -> silence out
    eng.delete()


def test_fixed_point_tracks_the_floating_point_modes_on_speech(library_path, shipped_model_path, test_pcm):
    """Shipped weights, fixture speech: the fixed-point engine's output stays within a few LSB of the fp32 engine's."""
    n = len(test_pcm) // 256
    pcm = test_pcm[: n * 256].reshape(1, n, 256)
    outs = {}
    for precision in ("fp32", "int8"):
        eng = kb.BatchKoala(1, model_path=shipped_model_path, precision=precision)
        outs[precision] = eng.process(pcm)
        eng.delete()
    d = np.abs(outs["fp32"].astype(np.int32) - outs["int8"].astype(np.int32))
    assert d.max() <= 16 and d.mean() < 1.0, (d.max(), d.mean())
