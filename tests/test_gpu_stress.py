"""Pipelined resident loop under load (four contexts, frames in flight, searches chained by events): the configuration in
which a race on the round flags of the claim-resolution cluster (k_resolve) used to deadlock the GPU about once in 50k
frames. Runs tools/pipeline_stress.py in a subprocess with its own watchdog so that a hang fails instead of stalling."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("trial", [0, 1, 2])
def test_pipelined_loop_makes_progress(trial):
    env = dict(os.environ, FT_STRESS_WATCHDOG="60")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "pipeline_stress.py"), "12000", "4"], env=env, capture_output=True,
                       text=True, timeout=180)
    assert r.returncode == 0 and "ok 12000 steps" in r.stdout, (r.stdout[-300:], r.stderr[-1500:])
