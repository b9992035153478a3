"""tools/record_reference.py is key-gated: without a Picovoice AccessKey (and a licence server to validate it) it must say why,
write nothing and exit 0 -- and bench.py's reference arm must then fall back to the CPU oracle port and say `kind: "port"`.
If reference golden vectors were ever recorded (tests/golden/ref_*.npy), the GPU engine is compared with them for INFORMATION:
the two engines implement different networks, so no tolerance is asserted, only that the comparison can be made."""
import glob
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

TOOL = os.path.join(ROOT, "tools", "record_reference.py")


def _run(env_extra):
    env = dict(os.environ)
    env.pop("PV_ACCESS_KEY", None)
    env.update(env_extra)
    before = sorted(os.listdir(GOLDEN))
    p = subprocess.run([sys.executable, TOOL, "--iterations", "1"], capture_output=True, text=True, env=env, timeout=300)
    assert sorted(os.listdir(GOLDEN)) == before          # nothing written
    return p


def test_recorder_without_key_skips_cleanly():
    p = _run({})
    assert p.returncode == 0 and "skipped: PV_ACCESS_KEY is not set" in p.stdout
    assert "not measurable" in p.stdout


def test_recorder_with_unusable_key_or_no_reference_skips_cleanly():
    p = _run({"PV_ACCESS_KEY": "invalid"})
    assert p.returncode == 0 and "reference recorder skipped:" in p.stdout
    p = _run({"PV_ACCESS_KEY": "invalid", "PV_KOALA_REFERENCE_DIR": "/nonexistent"})
    assert p.returncode == 0 and "no reference checkout" in p.stdout


def test_time_reference_reports_why_it_cannot_run(monkeypatch):
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import record_reference as rr
    monkeypatch.delenv("PV_ACCESS_KEY", raising=False)
    res, why = rr.time_reference("cpu", 1)
    assert res is None and "PV_ACCESS_KEY" in why


@pytest.mark.gpu
def test_recorded_reference_vectors_information_only(shipped_model_path):
    files = sorted(glob.glob(os.path.join(GOLDEN, "ref_*.npy")))
    if not files:
        pytest.skip("no reference golden vectors recorded (needs PV_ACCESS_KEY + network: tools/record_reference.py)")
    import koala_b200 as kb
    o = kb.create(kb.ANY_ACCESS_KEY, model_path=shipped_model_path)
    for f in files:
        ref = np.load(f)
        assert ref.dtype == np.int16 and ref.size % 256 == 0
    o.delete()
