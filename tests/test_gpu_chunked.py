"""The time-persistent mask-estimator launch (run on a B200: pytest -m gpu).

A multi-frame call is taken in chunks: analysis of the chunk's frames, ONE fused kernel launch that walks the chunk's steps
with per-(segment, stream tile) step counters, synthesis of the chunk.  The caller's serial frame loop is the reference's
(/root/reference/demo/c/koala_demo_file.c:466-521; state carried inside the handle, /root/reference/include/pv_koala.h:65-80).
Bit-for-bit checks against the same engine driven one frame per call; +-1 LSB checks against the CPU oracle."""
import os

import numpy as np
import pytest

import koala_b200 as kb
from oracle import OracleBatch, OracleModel

from conftest import synth_pcm

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]


class env:
    """Engine-creation knobs are read from the environment by Engine::create / fu_plan_create."""

    def __init__(self, **kv):
        self.kv = {k: str(v) for k, v in kv.items()}

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kv}
        os.environ.update(self.kv)

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def frame_by_frame(model, n, pcm, precision="bf16"):
    eng = kb.BatchKoala(n, model_path=model, precision=precision)
    out = np.empty_like(pcm)
    for t in range(pcm.shape[1]):
        out[:, t] = eng.process(np.ascontiguousarray(pcm[:, t]))
    h = [eng.debug_read(f"h{l}", (n, 512), np.float32) for l in range(2)]
    eng.delete()
    return out, h


@pytest.mark.parametrize("n_streams,chunk,precision", [(300, 8, "bf16"), (70, 5, "bf16"), (300, 64, "bf16"), (300, 8, "fp32"), (70, 64, "fp32")])
def test_chunked_call_is_bit_identical_to_frame_by_frame(library_path, random_model_path, n_streams, chunk, precision):
    """37 frames in one call (chunks of `chunk` frames = steps per fused launch) == 37 calls of one frame, output and state.
    fp32 mode runs the same kernel with three bf16 operand planes (hi | mid | lo) per activation."""
    import torch
    frames = 37
    pcm = synth_pcm(n_streams, frames, seed=700 + n_streams)
    ref, ref_h = frame_by_frame(random_model_path, n_streams, pcm, precision)
    with env(KOALA_CHUNK_FRAMES=chunk):
        eng = kb.BatchKoala(n_streams, model_path=random_model_path, precision=precision)
    assert eng.chunk_frames == chunk
    # host buffers, stream-major
    out = eng.process(pcm)
    assert (out == ref).all()
    for l in range(2):
        assert (eng.debug_read(f"h{l}", (n_streams, 512), np.float32) == ref_h[l]).all()
    launches = eng.kernel_launches
    # device buffers, both layouts, after a reset
    eng.reset()
    d = torch.from_numpy(pcm).cuda()
    out_d = eng.process(d).cpu().numpy()
    assert eng.kernel_launches - launches == 3 * ((frames + chunk - 1) // chunk)      # analysis, fused mask estimator, synthesis per chunk
    assert (out_d == ref).all()
    eng.reset()
    d_tm = torch.from_numpy(np.ascontiguousarray(pcm.transpose(1, 0, 2))).cuda()
    out_tm = eng.process(d_tm, time_major=True).cpu().numpy().transpose(1, 0, 2)
    assert (out_tm == ref).all()
    # in place: the caller's output buffer may be its input buffer
    eng.reset()
    d2 = torch.from_numpy(pcm).cuda()
    eng.process(d2, out=d2)
    assert (d2.cpu().numpy() == ref).all()
    eng.delete()


def test_chunked_call_against_oracle_small_batch(library_path, random_model_path):
    """configs[4]-shaped: 128 streams, 64-frame calls with the state carried across calls; +-1 LSB of the oracle."""
    n, frames, calls = 128, 64, 3
    pcm = synth_pcm(n, frames * calls, seed=41)
    eng = kb.BatchKoala(n, model_path=random_model_path, precision="bf16")
    out = np.concatenate([eng.process(np.ascontiguousarray(pcm[:, c * frames:(c + 1) * frames])) for c in range(calls)], axis=1)
    ob = OracleBatch(OracleModel(random_model_path), n, "bf16")
    ref = ob.process(pcm, threads=8)
    diff = np.abs(out.astype(np.int32) - ref.astype(np.int32))
    assert diff.max() <= 1, diff.max()
    for l in range(2):
        h = eng.debug_read(f"h{l}", (n, 512), np.float32)
        np.testing.assert_allclose(h, np.stack([ob.stream(s).h[l] for s in range(n)]), atol=1e-3)
    eng.delete()


def test_few_clusters_do_not_deadlock(library_path, random_model_path):
    """The dependency waits are only ever for tiles earlier in the list, so any grid size must finish and give the same bits:
    3 CTA pairs walk 24 steps of 300 streams (every wait is then for a tile of the same or a neighbouring pair)."""
    n, frames = 300, 24
    pcm = synth_pcm(n, frames, seed=77)
    eng = kb.BatchKoala(n, model_path=random_model_path, precision="bf16")
    ref = eng.process(pcm)
    eng.delete()
    for clusters in (1, 3):
        with env(KOALA_FU_CLUSTERS=clusters):
            eng = kb.BatchKoala(n, model_path=random_model_path, precision="bf16")
        out = eng.process(pcm)
        assert (out == ref).all(), clusters
        eng.delete()


def test_both_part_orders_agree_with_the_oracle(library_path, random_model_path):
    """x part first (small batches) and h part first (big batches) differ only in fp32 summation order."""
    n, frames = 130, 20
    pcm = synth_pcm(n, frames, seed=5)
    ref = OracleBatch(OracleModel(random_model_path), n, "bf16").process(pcm, threads=8)
    for order in (0, 1):
        with env(KOALA_FU_XFIRST=order):
            eng = kb.BatchKoala(n, model_path=random_model_path, precision="bf16")
        out = eng.process(pcm)
        assert np.abs(out.astype(np.int32) - ref.astype(np.int32)).max() <= 1, order
        eng.delete()


def test_full_size_chunk_boundary(library_path, random_model_path):
    """8192 streams x 40 frames in one call (two partitions of 4096 streams, chunks of 32 frames: one boundary): replicas
    bit-identical, silent streams silent, equal to the one-frame-per-call drive of a second engine, oracle spot check."""
    import torch
    n, frames = 8192, 40
    base = synth_pcm(64, frames, seed=19)
    pcm = np.tile(base, (n // 64, 1, 1))
    pcm[5::64] = 0
    eng = kb.BatchKoala(n, model_path=random_model_path, precision="bf16")
    assert eng.chunk_frames >= 2
    d_in = torch.from_numpy(np.ascontiguousarray(pcm.transpose(1, 0, 2))).cuda()       # time-major
    out = eng.process(d_in, time_major=True).cpu().numpy().transpose(1, 0, 2)
    eng.delete()
    assert (out.reshape(n // 64, 64, frames, 256) == out[:64][None]).all()
    assert (out[5::64] == 0).all()
    eng1 = kb.BatchKoala(n, model_path=random_model_path, precision="bf16")
    one = np.empty_like(pcm)
    for t in range(frames):
        one[:, t] = eng1.process(d_in[t]).cpu().numpy()
    eng1.delete()
    assert (one == out).all()
    pick = [0, 1, 2, 3, 6, 7, 17, 63, 4096 + 9, 8191]
    ref = OracleBatch(OracleModel(random_model_path), len(pick), "bf16").process(np.ascontiguousarray(pcm[pick]), threads=8)
    assert np.abs(out[pick].astype(np.int32) - ref.astype(np.int32)).max() <= 1


def test_fp32_tensor_core_path_against_cuda_core_path_and_oracle(library_path, random_model_path):
    """fp32 mode has two implementations: the fused tensor-core kernel with every activation split into three bf16 planes (the
    default whenever the hidden size fits its tiles) and the CUDA-core kernels (other hidden sizes; KOALA_FP32_CUDA_CORES=1).
    Both must sit within the fp32 tolerances of the oracle -- and therefore of each other -- over a 200-frame state-carried run
    fed in 64-frame calls."""
    n, frames = 260, 200
    pcm = synth_pcm(n, frames, seed=23)
    ob = OracleBatch(OracleModel(random_model_path), n, "fp32")
    ref = ob.process(pcm, threads=os.cpu_count() or 8)
    outs = {}
    for cores in (0, 1):
        with env(KOALA_FP32_CUDA_CORES=cores):
            eng = kb.BatchKoala(n, model_path=random_model_path, precision="fp32")
        out = np.concatenate([eng.process(np.ascontiguousarray(pcm[:, t:t + 64])) for t in range(0, frames, 64)], axis=1)
        # CUDA cores: 6 launches per frame; tensor cores: 3 per chunk (host buffers travel in chunks of a few frames)
        assert (eng.kernel_launches == 6 * frames) if cores else (eng.kernel_launches % 3 == 0 and eng.kernel_launches <= 3 * frames // 4), eng.kernel_launches
        assert np.abs(out.astype(np.int32) - ref.astype(np.int32)).max() <= 1, cores
        for l in range(2):
            h = eng.debug_read(f"h{l}", (n, 512), np.float32)
            np.testing.assert_allclose(h, np.stack([ob.stream(s).h[l] for s in range(n)]), atol=1e-4)
        mask = eng.debug_read("mask", (n, 256), np.float32)
        np.testing.assert_allclose(mask, np.stack([ob.stream(s).last_mask for s in range(n)]), rtol=1e-3, atol=1e-6)
        outs[cores] = out
        eng.delete()
    assert np.abs(outs[0].astype(np.int32) - outs[1].astype(np.int32)).max() <= 1


def test_fp32_hidden_size_outside_the_fused_tiles(library_path, tmp_path):
    """H = 320 is not a multiple of 256: fp32 mode falls to the CUDA-core kernels (bf16 mode refuses such a model)."""
    from koala_b200 import spec
    path = str(tmp_path / "model_320_2.kpv")
    spec.save_model(path, spec.random_model(seed=320, hidden=320, layers=2), hidden=320, layers=2)
    n, frames = 37, 12
    pcm = synth_pcm(n, frames, seed=320)
    eng = kb.BatchKoala(n, model_path=path, precision="fp32")
    out = eng.process(pcm)
    assert eng.kernel_launches == 6 * frames
    ref = OracleBatch(OracleModel(path), n, "fp32").process(pcm, threads=8)
    assert np.abs(out.astype(np.int32) - ref.astype(np.int32)).max() <= 1
    eng.delete()
    with pytest.raises(kb.KoalaError):
        kb.BatchKoala(n, model_path=path, precision="bf16")


def test_partitioned_batch_equals_one_batch(library_path, random_model_path):
    """A batch above 4096 streams runs as partitions of 4096 (engine.cu kPartStreams) so that each partition's working set stays in
    the L2.  Partitioning must be invisible: 600 streams as partitions of 256 (a ragged last one) == the same 600 streams as one
    batch, bit for bit, through host and device buffers, with state carry, whole and per-stream reset and the state read-back."""
    import torch
    n, frames = 600, 21
    pcm = synth_pcm(n, frames, seed=4096)
    one = kb.BatchKoala(n, model_path=random_model_path, precision="bf16")
    ref = one.process(pcm)
    ref_h = [one.debug_read(f"h{l}", (n, 512), np.float32) for l in range(2)]
    with env(KOALA_PARTITION_STREAMS=256):
        eng = kb.BatchKoala(n, model_path=random_model_path, precision="bf16")
    assert (eng.process(pcm) == ref).all()
    assert eng.kernel_launches % 9 == 0                       # three partitions x three launches per chunk
    for l in range(2):
        assert (eng.debug_read(f"h{l}", (n, 512), np.float32) == ref_h[l]).all()
    eng.reset()
    halves = np.concatenate([eng.process(np.ascontiguousarray(pcm[:, :10])), eng.process(np.ascontiguousarray(pcm[:, 10:]))], axis=1)
    assert (halves == ref).all()                               # state carried across calls in every partition
    eng.reset()
    tm = torch.from_numpy(np.ascontiguousarray(pcm.transpose(1, 0, 2))).cuda()
    assert (eng.process(tm, time_major=True).cpu().numpy().transpose(1, 0, 2) == ref).all()
    eng.reset([3, 300, 599])                                   # one stream in each partition restarts, the others carry on
    cont, cont_ref = eng.process(pcm), None
    one.reset([3, 300, 599])
    cont_ref = one.process(pcm)
    assert (cont == cont_ref).all() and (cont[[3, 300, 599]] == ref[[3, 300, 599]]).all()
    with pytest.raises(kb.KoalaInvalidArgumentError):
        eng.reset([600])
    eng.delete()
    one.delete()


def test_long_time_major_host_call_of_a_big_batch(library_path, random_model_path):
    """1100 streams x 75 frames from host memory in the time-major layout (7-frame chunks through the three-buffer input ring and the
    two-buffer output ring, a ragged last chunk): what the same frames give from device memory, bit for bit, state carried into the
    next call."""
    import torch
    n, frames = 1100, 75
    pcm = synth_pcm(n, frames + 9, seed=1100)
    tm = np.ascontiguousarray(pcm.transpose(1, 0, 2))
    eng = kb.BatchKoala(n, model_path=random_model_path, precision="bf16")
    ref = eng.process(torch.from_numpy(tm).cuda(), time_major=True).cpu().numpy()
    eng.reset()
    a = eng.process(tm[:frames], time_major=True)
    b = eng.process(tm[frames:], time_major=True)
    assert (a == ref[:frames]).all() and (b == ref[frames:]).all()
    eng.delete()
