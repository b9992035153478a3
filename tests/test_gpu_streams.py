"""Stream ordering and co-resident engines (run on a B200: pytest -m gpu).

Every step reads and writes the per-stream state rows and the fused kernel's dependency counters, so work enqueued on a new CUDA
stream must run after what was enqueued on the previous one, and the fused launches of DIFFERENT engines on one device must not
run concurrently (each grid spins on counters and assumes it is co-resident).  The reference handle has the same contract in its
single-threaded form: calls on one `pv_koala_t` are serial (/root/reference/include/pv_koala.h:65-90; bindings serialise)."""
import numpy as np
import pytest

import koala_b200 as kb

from conftest import synth_pcm

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]


def test_state_is_ordered_across_streams(library_path, random_model_path):
    """Device-tensor calls return without synchronising.  A call on another torch stream, a per-stream reset and a host-buffer call
    issued right behind them must all see the state the earlier work leaves: same samples as the fully synchronised sequence."""
    import torch
    n, frames = 300, 12
    pcm = synth_pcm(n, 3 * frames, seed=11)
    ref_eng = kb.BatchKoala(n, model_path=random_model_path, precision="bf16")
    a = ref_eng.process(np.ascontiguousarray(pcm[:, :frames]))
    ref_eng.reset([7, 299])
    b = ref_eng.process(np.ascontiguousarray(pcm[:, frames:2 * frames]))
    c = ref_eng.process(np.ascontiguousarray(pcm[:, 2 * frames:]))
    ref_eng.delete()

    eng = kb.BatchKoala(n, model_path=random_model_path, precision="bf16")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    d = [torch.from_numpy(np.ascontiguousarray(pcm[:, i * frames:(i + 1) * frames])).cuda() for i in range(3)]
    torch.cuda.synchronize()
    with torch.cuda.stream(s1):
        out_a = eng.process(d[0])                 # enqueued on s1, not synchronised
    eng.reset([7, 299])                            # engine's own stream: must wait for s1
    with torch.cuda.stream(s2):
        out_b = eng.process(d[1])                 # another stream: must wait for the reset
    out_c = eng.process(np.ascontiguousarray(pcm[:, 2 * frames:]))      # host buffers (engine's own stream): must wait for s2
    torch.cuda.synchronize()
    assert (out_a.cpu().numpy() == a).all()
    assert (out_b.cpu().numpy() == b).all()
    assert (out_c == c).all()
    eng.delete()


def test_two_engines_interleaved_on_two_streams(library_path, random_model_path):
    """Two engines on one GPU, each fed from its own stream without synchronisation in between: their persistent mask-estimator
    grids are serialised by the library (no hang), and each produces what it produces alone."""
    import torch
    n, frames, rounds = 300, 16, 4
    pcm = [synth_pcm(n, frames * rounds, seed=21 + i) for i in range(2)]
    solo = []
    for i in range(2):
        e = kb.BatchKoala(n, model_path=random_model_path, precision="bf16")
        solo.append(e.process(pcm[i]))
        e.delete()
    engines = [kb.BatchKoala(n, model_path=random_model_path, precision="bf16") for _ in range(2)]
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    d_in = [torch.from_numpy(p).cuda() for p in pcm]
    torch.cuda.synchronize()
    outs = [[], []]
    for r in range(rounds):
        for i in range(2):
            with torch.cuda.stream(streams[i]):
                sl = slice(r * frames, (r + 1) * frames)
                outs[i].append(engines[i].process(d_in[i][:, sl].contiguous()))
    torch.cuda.synchronize()
    for i in range(2):
        got = torch.cat(outs[i], dim=1).cpu().numpy()
        assert (got == solo[i]).all(), i
    for e in engines:
        e.delete()
