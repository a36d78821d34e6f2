"""pytest plugin (``-p refsuite_plugin``): makes ``import mcframework`` resolve to ``mcframework_b200`` with the
oracle-backed stand-in device installed, BEFORE the reference's conftest and test modules are imported."""
import importlib
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
for _p in (_ROOT, os.path.join(_ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

from refsuite import fakedev  # noqa: E402

fakedev.install()
import mcframework_b200 as _pkg  # noqa: E402

for _name in ("core", "stats_engine", "utils", "sims", "sims.pi", "sims.portfolio", "sims.black_scholes"):
    sys.modules["mcframework." + _name] = importlib.import_module("mcframework_b200." + _name)
sys.modules["mcframework"] = _pkg
