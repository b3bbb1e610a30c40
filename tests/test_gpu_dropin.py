"""The C-ABI library as a drop-in under the reference's own Python layer.

The reference is imported from ``baseline/_ref/warp_src`` (the unmodified tree, built once with its own
``build_lib.py``); the stub of INTEGRATION.md is applied verbatim by ``tests/dropin_driver.py`` in a subprocess, which
then runs the LBVH cases of ``warp/tests/geometry/test_mesh.py:111-357`` and ``test_bvh.py:186-262`` with unmodified
``@wp.kernel`` code going through ``mesh.id`` / ``bvh.id``.  Skipped (with the reason) only where that build is absent.
"""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref", "warp_src")


@pytest.mark.gpu
def test_reference_python_layer_runs_on_the_b200_library():
    if not os.path.exists(os.path.join(REF, "warp", "bin", "warp.so")):
        pytest.skip("baseline/_ref/warp_src (the reference built from /root/reference) is not present on this box")
    from warp_b200 import _lib

    env = dict(os.environ)
    env.setdefault("WARP_CACHE_PATH", os.path.join(ROOT, "gpurun_out", "warp_cache"))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "dropin_driver.py"), REF, _lib.LIB_PATH],
                       capture_output=True, text=True, timeout=1500, env=env)
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("DROPIN_REPORT ")]
    assert lines, f"driver produced no report (rc {r.returncode}):\n{r.stdout[-3000:]}\n{r.stderr[-3000:]}"
    report = json.loads(lines[-1][len("DROPIN_REPORT "):])
    failed = {k: v for k, v in report["cases"].items() if v != "ok"}
    assert not failed, f"{failed}\n{r.stderr[-2000:]}"
    assert report["routed"].get("mesh") and report["routed"].get("bvh")
    assert r.returncode == 0
