"""The check build (libmchammer_b200_check.so, -DMCH_RACECHECK: device asserts on the step descriptors and the
race detector for the thread-block-cluster protocols) runs the cluster cases without a report and returns the
same bits as the product build.  compute-sanitizer is not available on the GPU pool; this is its stand-in
(csrc/mch_device.cuh, profiles/check_build_case.py; the negative controls are recorded in profiles/)."""

from __future__ import annotations

import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHECK_LIB = os.path.join(REPO, "mchammer_b200", "libmchammer_b200_check.so")


def run_case(lib: str | None) -> dict:
    env = dict(os.environ)
    env.pop("MCHAMMER_B200_LIB", None)
    if lib:
        env["MCHAMMER_B200_LIB"] = lib
    done = subprocess.run([sys.executable, os.path.join(REPO, "profiles", "check_build_case.py")], cwd=REPO, env=env,
                          capture_output=True, text=True, timeout=600)
    assert done.returncode == 0, "check_build_case.py failed:\n" + done.stdout[-3000:] + done.stderr[-3000:]
    return json.loads(done.stdout.strip().splitlines()[-1])


def test_cluster_protocols_under_the_race_detector():
    if not os.path.exists(CHECK_LIB):
        pytest.skip("libmchammer_b200_check.so not built (python -m mchammer_b200.build)")
    checked = run_case(CHECK_LIB)  # a race or a failed assert traps: non-zero exit
    plain = run_case(None)
    assert checked.keys() == plain.keys() and len(checked) >= 15
    for tag in plain:
        assert checked[tag] == plain[tag], tag  # the instrumentation adds no arithmetic: the same bits
