"""Drop-in check: the REFERENCE's own test-suite, read in place from ``/root/reference/tests``, run against
``mcframework_b200`` with the oracle-backed stand-in device of ``tests/refsuite/fakedev.py`` (no GPU needed).

What it pins: everything above the C ABI -- class/attribute names the reference tests import and patch, signatures,
validation messages, the error taxonomy of the stats engine, stream bookkeeping (re-seeding, un-reseeded continuation,
spawn side effects), percentile merging, metadata keys.  The kernels below the C ABI are pinned against the same oracle
by the ``-m gpu`` tests.  Skipped where the reference checkout does not exist (the GPU box)."""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_TESTS = "/root/reference/tests"


@pytest.mark.skipif(not os.path.isdir(REF_TESTS), reason="reference checkout not present")
def test_reference_test_suite_passes_against_the_drop_in(tmp_path):
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(ROOT, "tests", "refsuite"), ROOT, env.get("PYTHONPATH", "")])
    cmd = [sys.executable, "-m", "pytest", REF_TESTS, "-o", "addopts=", "-p", "no:cacheprovider", "-p", "refsuite_plugin",
           "-q", "--tb=line", "--rootdir", str(tmp_path)]
    run = subprocess.run(cmd, cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=900)
    tail = run.stdout[-3000:]
    m = re.search(r"(\d+) passed", run.stdout)
    assert run.returncode == 0 and m, tail
    assert "failed" not in run.stdout.splitlines()[-1], tail
    assert int(m.group(1)) >= 231, tail  # the reference ships 231 tests


@pytest.mark.skipif(not os.path.isdir("/root/reference/src/mcframework"), reason="reference checkout not present")
def test_differential_run_against_the_unmodified_reference(tmp_path):
    """``tests/refsuite/differential.py``: the reference itself and the drop-in (stand-in device) driven through the same
    randomised public calls -- engine metrics under every context option and taint, run()/run_simulation/compare_results/
    result_to_string for the four built-in simulations (sequential, block-parallel, un-reseeded continuation), user hooks
    on thread pools, Greeks, paths + LSM, validation errors -- with every output compared (numbers to 1e-9, exception
    types and messages exactly)."""
    script = os.path.join(ROOT, "tests", "refsuite", "differential.py")
    run = subprocess.run([sys.executable, script, "80", "11"], cwd=str(tmp_path), capture_output=True, text=True, timeout=900)
    last = run.stdout.strip().splitlines()[-1] if run.stdout.strip() else ""
    assert run.returncode == 0 and re.fullmatch(r"differential: \d+ comparisons, 0 mismatches", last), run.stdout[-3000:]
    assert int(last.split()[1]) > 2000
