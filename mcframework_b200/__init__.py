"""mcframework_b200 -- the replication loop and StatsEngine reductions of mcframework on B200 GPUs.

Exposes the public names of the reference package (``mcframework/__init__.py:13-28``).  Importing the package
needs neither a GPU nor the CUDA library; running a built-in simulation or a built-in metric does, and fails
loudly without them (no CPU fallback).  Names are resolved lazily (PEP 562) from the table below.
"""
from importlib import import_module as _import_module

__version__ = "0.1.0"

#: public name -> submodule that defines it
_EXPORTS = {
    "core": ("SimulationResult", "MonteCarloSimulation", "MonteCarloFramework"),
    "sims": ("PiEstimationSimulation", "PortfolioSimulation", "BlackScholesSimulation", "BlackScholesPathSimulation"),
    "stats_engine": ("StatsEngine", "StatsContext", "FnMetric", "DEFAULT_ENGINE"),
    "utils": ("z_crit", "t_crit", "autocrit"),
}
_WHERE = {name: mod for mod, names in _EXPORTS.items() for name in names}
__all__ = [name for names in _EXPORTS.values() for name in names]


def __getattr__(name):
    if name in _WHERE:
        value = getattr(_import_module(f"{__name__}.{_WHERE[name]}"), name)
        globals()[name] = value
        return value
    if name in _EXPORTS or name in ("_device", "_native", "_stats_device", "_streams", "_programs", "_multi", "build",
                                    "programs"):  # `programs`: declared draw programs for user simulations (additive)
        return _import_module(f"{__name__}.{name}")
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")


def __dir__():
    return sorted(set(globals()) | set(__all__))
