"""Drop-in conformance: the reference's OWN tests (tests/test_breakfast.py, test_caching.py,
test_filtering.py, 36 tests) run unchanged against this package, imported as `breakfast`.
Only possible where /root/reference is mounted (the build container); on the CPU the distance engine
is swapped for the oracle by the oracle.engine_standin pytest plugin, on a GPU box the real kernels run."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
REF_TESTS = Path("/root/reference/tests")


@pytest.mark.skipif(not REF_TESTS.exists(), reason="/root/reference is not mounted here")
def test_reference_test_suite_passes_against_this_package(tmp_path):
    from breakfast_b200 import _native
    env = dict(os.environ, PYTHONPATH=str(ROOT))
    cmd = [sys.executable, "-m", "pytest", str(REF_TESTS), "-q", "-p", "no:cacheprovider", "--rootdir", str(tmp_path)]
    if _native.device_count() == 0:
        cmd += ["-p", "oracle.engine_standin"]
    res = subprocess.run(cmd, cwd=tmp_path, env=env, capture_output=True, text=True, timeout=600)
    tail = (res.stdout + res.stderr)[-3000:]
    assert res.returncode == 0, tail
    assert "36 passed" in res.stdout, tail
    # and it really was this package, not the reference
    probe = subprocess.run([sys.executable, "-c", "import breakfast.breakfast as b; print(b.__file__)"],
                           cwd=tmp_path, env=env, capture_output=True, text=True)
    assert "breakfast_b200" in probe.stdout
